"""An OUT-OF-TREE (dynamics, cost) pair through the plugin path (include/mppi_b200.h: mppib_load_plugin):
plugins_example/pendulum_pair.cu is compiled into its own shared library from csrc/engine_internal.cuh — libmppi_b200.so is
neither edited nor rebuilt — registered under ids 1000 / 1000 and solved through the ordinary C ABI. The reference lets users
instantiate its templates with their own classes (dynamics.cuh:67-76, cost.cuh:34-35); this is the same contract across
the C-ABI boundary. The check is an independent numpy float32 rollout of the pendulum on the device's own noise."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import mppi_generic_b200 as m

H = m.host
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "plugins_example", "libmppi_plugin_pendulum.so")


class PendulumDynParams(C.Structure):
    _fields_ = [("lim", H.ControlLimits), ("mass", C.c_float), ("length", C.c_float), ("damping", C.c_float), ("gravity", C.c_float)]


class PendulumCostParams(C.Structure):
    _fields_ = [("control_cost_coeff", C.c_float * H.MAX_C), ("discount", C.c_float), ("angle_coeff", C.c_float),
                ("rate_coeff", C.c_float), ("goal_angle", C.c_float), ("terminal_coeff", C.c_float)]


def _plugin():
    """Load the example plugin; build it first if it is missing, rebuild once if the library's layout fingerprint refuses it
    (a plugin must come from the same source revision as libmppi_b200.so)."""
    build = ["bash", os.path.join(ROOT, "plugins_example", "build.sh")]
    if not os.path.exists(PLUGIN):
        subprocess.check_call(build)
    try:
        H.load_plugin(PLUGIN)
    except H.MppibError as ex:
        if "another revision" not in str(ex):
            raise
        subprocess.check_call(build)
        H.load_plugin(PLUGIN)


def test_unknown_user_pair_is_refused_before_the_plugin_is_loaded():
    """CPU-side contract: without the plugin, ids 1000 / 1000 name no kernel (and without a GPU nothing is created either)."""
    dp, cp = PendulumDynParams(), PendulumCostParams()
    dyn = H.UserDynamics(H.USER_ID_BASE + 7, 2, 1, 2, dp)
    cost = H.UserCost(H.USER_ID_BASE + 7, cp)
    with pytest.raises(H.MppibError):
        H.Engine(dyn, cost, H.GaussianDistribution(1, [1.0]), 256, 20, 1)


def test_plugin_library_loads_and_registers_without_a_gpu():
    """Run LAST among the CPU tests of this file (the registration is process-wide): dlopen resolves every symbol the plugin
    needs from libmppi_b200.so (mppib_register_pair, the error channel), mppib_plugin_init registers the pair, loading twice
    is a no-op; creating an engine then gets past the kernel lookup (and fails on the missing
    device here, not on the ids)."""
    _plugin()
    H.load_plugin(PLUGIN)  # loading the same library again is a no-op
    with pytest.raises(H.MppibError):
        H.load_plugin(os.path.join(ROOT, "mppi-generic_b200", "libmppi_b200.so"))  # no mppib_plugin_init in there
    import torch
    if not torch.cuda.is_available():
        dp, cp = PendulumDynParams(), PendulumCostParams()
        with pytest.raises(H.MppibError) as ei:
            H.Engine(H.UserDynamics(H.USER_ID_BASE, 2, 1, 2, dp), H.UserCost(H.USER_ID_BASE, cp),
                     H.GaussianDistribution(1, [1.0]), 256, 20, 1)
        assert "no kernel registered" not in str(ei.value)


@pytest.mark.gpu
@pytest.mark.parametrize("N,T", [(4096, 100), (1000, 37)])
def test_out_of_tree_pendulum_pair_matches_numpy_rollout(N, T):
    _plugin()
    dp, cp = PendulumDynParams(), PendulumCostParams()
    dp.mass, dp.length, dp.damping, dp.gravity = 1.0, 0.8, 0.1, 9.81
    cp.discount, cp.angle_coeff, cp.rate_coeff, cp.goal_angle, cp.terminal_coeff = 1.0, 10.0, 0.5, np.pi, 3.0
    dyn = H.UserDynamics(H.USER_ID_BASE, 2, 1, 2, dp)
    dyn.setControlRanges([(-2.0, 2.0)])
    cost = H.UserCost(H.USER_ID_BASE, cp)
    sampler = H.GaussianDistribution(1, [1.5])
    dt, lam, alpha = 0.02, 1.0, 0.0
    e = H.Engine(dyn, cost, sampler, N, T, 1, flags=H.FLAG_WRITEBACK_CONTROLS)
    e.set_solver(dt, lam, alpha)
    e.seed(42, 0)
    x0 = np.array([[0.3, -0.2]], np.float32)
    U0 = (0.5 * np.sin(np.arange(T) * 0.1)).astype(np.float32).reshape(1, T, 1)
    U, stats = e.solve(x0, U0)
    eps = e.get_noise()[:, :, 0]                                   # [N][T]
    f32 = np.float32
    # the device evaluates mean + sigma eps as ONE fma (nvcc contracts the reference's expression the same way): the
    # float64 product is exact, so rounding the float64 sum reproduces it up to (rare) double rounding
    u = (U0[0, :, 0][None, :].astype(np.float64) + 1.5 * eps.astype(np.float64)).astype(f32)
    u[0] = U0[0, :, 0]                                             # sample 0 is noise-free (gaussian.cu:101)
    tail = np.arange(N) >= np.float32((1.0 - sampler.params.pure_noise_trajectories_percentage) * N)
    u[tail] = (f32(1.5) * eps[tail]).astype(f32)                   # pure-noise tail (gaussian.cu:108)
    u[:, :1] = U0[0, :1, 0]                                        # t < optimization_stride (= 1) keeps the mean (gaussian.cu:101)
    u = np.clip(u, f32(-2.0), f32(2.0))
    dev_u = e.get_samples()[0][:, :, 0]                            # the engine's constrained controls
    np.testing.assert_allclose(dev_u, u, rtol=2e-7, atol=1e-7)
    assert (dev_u != u).mean() < 1e-3
    u = dev_u
    th, om = np.full(N, x0[0, 0], f32), np.full(N, x0[0, 1], f32)
    running = np.zeros(N, f32)
    for t in range(T):
        acc = ((u[:, t] - f32(0.1) * om - f32(1.0 * 9.81 * 0.8) * np.sin(th).astype(f32)) / f32(1.0 * 0.8 * 0.8)).astype(f32)
        th, om = (th + om * f32(dt)).astype(f32), (om + acc * f32(dt)).astype(f32)
        running += (f32(10.0) * (th - f32(np.pi)) ** 2 + f32(0.5) * om ** 2).astype(f32)
    total = (running + f32(3.0) * (f32(10.0) * (th - f32(np.pi)) ** 2 + f32(0.5) * om ** 2)) / f32(T)
    c = e.get_costs()[0]
    rel = np.abs(c - total) / np.maximum(np.abs(total), 1.0)
    assert rel.max() < 1e-4, rel.max()
    w = np.exp(-(c.astype(np.float64) - c.min()) / lam)
    np.testing.assert_allclose(U[0, :, 0], (w[:, None] * u).sum(0) / w.sum(), atol=2e-5)
    assert stats[0][0] == c.min()
    e.close()
