"""the host tail of Controller::computeControl on the device (mppib_nominal_trajectory) — Savitzky-Golay
smoothing (controllers/controller.cuh:557-586) and the nominal state / output roll-forward (:643-663) — against the library's
host twins (mppib_host_smooth_controls, mppib_host_output_trajectory[_lstm]; themselves pinned to the oracle in
tests/test_host_logic.py). The device bodies use the kernels' arithmetic (FFMA contraction, tanh_fast / the mma.sync network,
sincos_cw), the host twins the reference's host arithmetic: the bar is the reference's own CPU≡GPU bar for a trajectory of
step() calls, 1e-4 relative to the trajectory's scale (tests/include/kernel_tests/core/rollout_kernel_test.cu, :258)."""
import numpy as np
import pytest

import mppi_generic_b200 as m
from mppi_generic_b200 import workloads as W

H = m.host
pytestmark = pytest.mark.gpu

CASES = {
    "cartpole": lambda: W.cartpole(1024, 100),
    "double_integrator_tube": lambda: W.double_integrator_tube(1024, 150),
    "autorally": lambda: W.autorally(1024, 100),
    "racer_lstm": lambda: W.racer_lstm(1024, 150),
    "racer_lstm_h32": lambda: W.racer_lstm_h32(1024, 60),
    "quadrotor": lambda: W.quadrotor(1024, 100),
}


def _host_tail(w, x0, U, hist):
    Us = U.copy()
    states = np.zeros((w.D, w.T, w.dyn.STATE_DIM), np.float32)
    outputs = np.zeros((w.D, w.T, w.dyn.OUTPUT_DIM), np.float32)
    for d in range(w.D):
        if hist is not None:
            H.lib().mppib_host_smooth_controls(Us[d].ctypes.data, hist.ctypes.data, w.T, w.dyn.CONTROL_DIM)
        w.dyn.output_trajectory(x0[d], Us[d], w.T, w.dt, states[d], outputs[d])
    return Us, states, outputs


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_tail_matches_host_twins(name):
    w = CASES[name]()
    e = w.make_engine()
    x0 = np.ascontiguousarray(w.x0, np.float32)
    U, _ = e.solve(x0, w.U0)
    rng = np.random.default_rng(3)
    hist = (0.1 * rng.standard_normal((2, w.dyn.CONTROL_DIM))).astype(np.float32)
    for h in (hist, None):
        Us_h, st_h, out_h = _host_tail(w, x0, U, h)
        Us_d, st_d, out_d = e.nominal_trajectory(x0, U, h)
        np.testing.assert_allclose(Us_d, Us_h, rtol=0, atol=2e-6 * max(1.0, float(np.abs(U).max())))
        for a, b in ((st_d, st_h), (out_d, out_h)):
            nan = np.isnan(b)  # the RACER model reports its wheel forces as NaN without an elevation map (both twins)
            assert np.array_equal(np.isnan(a), nan) and np.isfinite(a[~nan]).all()
            a, b = np.where(nan, 0.0, a), np.where(nan, 0.0, b)
            # per state component: 1e-4 of that component's range over the trajectory (the reference's CPU==GPU bar)
            scale = np.maximum(np.abs(b).max(axis=1, keepdims=True), 1.0)
            assert (np.abs(a - b) / scale).max() < 1e-4, (name, (np.abs(a - b) / scale).max())
        assert np.array_equal(st_d[:, 0], x0)
    e.close()


@pytest.mark.parametrize("name", ["cartpole", "double_integrator_tube", "autorally"])
def test_device_tail_chained_behind_an_async_solve(name):
    """U = NULL: the kernel reads the optimised sequence from the result record on the device, ordered behind the solve on
    its stream — identical to passing the solve's U explicitly, and mppib_solve_wait afterwards still hands out the result."""
    w = CASES[name]()
    e = w.make_engine()
    x0 = np.ascontiguousarray(w.x0, np.float32)
    U0 = np.ascontiguousarray(w.U0, np.float32)
    hist = np.zeros((2, w.dyn.CONTROL_DIM), np.float32)
    e.seed(w.seed, 0)
    e.solve_async(x0, U0)
    chained = e.nominal_trajectory(x0, None, hist)
    U, stats = e.solve_wait()
    explicit = e.nominal_trajectory(x0, U, hist)
    for a, b in zip(chained, explicit):
        assert np.array_equal(a, b)
    with pytest.raises(H.MppibError):
        H.Engine.nominal_trajectory(w.make_engine(), x0, None, hist)  # no solve yet: nothing to roll out
    e.close()


def test_controller_mirror_with_the_device_side_tail():
    """VanillaMPPIController.setDeviceSideTail(True): the same closed loop, tail on the device; controls / states agree with
    the default (host-twin) tail to the bar above after several computeControl + slide rounds."""
    w = W.cartpole(2048, 100)

    def run(device_tail):
        ctrl = H.VanillaMPPIController(w.dyn, w.cost, None, w.sampler, w.dt, 1, w.lambda_, w.alpha, w.T, w.N, seed=w.seed)
        ctrl.setDeviceSideTail(device_tail)
        x = w.x0[0].copy()
        for _ in range(5):
            ctrl.computeControl(x, 1)
            x = ctrl.getTargetStateSeq()[1].copy()
            ctrl.slideControlSequence(1)
        return ctrl.getControlSeq().copy(), ctrl.getTargetStateSeq().copy()

    u_h, s_h = run(False)
    u_d, s_d = run(True)
    np.testing.assert_allclose(u_d, u_h, atol=2e-4 * max(1.0, float(np.abs(u_h).max())))
    np.testing.assert_allclose(s_d, s_h, atol=2e-4 * max(1.0, float(np.abs(s_h).max())))


@pytest.mark.parametrize("name", ["cartpole", "double_integrator_tube", "autorally", "racer_lstm", "quadrotor"])
def test_compute_control_is_solve_plus_host_tail(name):
    """mppib_compute_control = mppib_solve followed by the library's host twins on its result (SG smoothing, nominal
    roll-forward): bit-identical to doing the three calls separately, two closed-loop rounds."""
    w = CASES[name]()
    if name == "racer_lstm":  # with an elevation map, so that the engine-held host copy of the map is exercised too
        j, i = np.meshgrid(np.arange(64), np.arange(64))
        w.dyn.setElevationMap((0.3 * np.sin(0.2 * j) * np.cos(0.15 * i)).astype(np.float32), 0.5, (-8.0, -16.0, 0.0))
    a, b = w.make_engine(), w.make_engine()
    x0 = np.ascontiguousarray(w.x0, np.float32)
    hist = np.zeros((2, w.dyn.CONTROL_DIM), np.float32)
    Ua = Ub = np.ascontiguousarray(w.U0, np.float32)
    for it in range(2):
        Ua, st_a, out_a, stats_a = a.compute_control(x0, Ua, hist)
        Us, stats_b = b.solve(x0, Ub)
        Ub, st_b, out_b = _host_tail(w, x0, Us, hist)
        assert np.array_equal(Ua, Ub), it
        assert np.array_equal(st_a, st_b) and np.array_equal(np.nan_to_num(out_a), np.nan_to_num(out_b)), it
        assert stats_a == stats_b
        hist = Ua[0, :2].copy()
    # without the roll-forward / without smoothing
    U2, st2, out2, _ = a.compute_control(x0, w.U0, None, roll_forward=False)
    assert st2 is None and out2 is None and np.isfinite(U2).all()
    a.close()
    b.close()
