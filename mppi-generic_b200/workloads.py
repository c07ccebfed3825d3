"""Synthetic workloads for the BASELINE.json configurations (SURVEY.md §8d "Synthetic inputs (fixed, seeded)").

Each builder returns a ``Workload`` holding the plugin objects (host mirrors of the reference classes), the solver
scalars and the initial conditions, so that tests, ``__graft_entry__.smoke()`` and ``bench.py`` all run the very same
configuration. Nothing here reads /root/reference or any dataset: weights and maps are generated from fixed seeds.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import host as H


@dataclass
class Workload:
    name: str
    controller: str  # "vanilla" | "tube"
    dyn: object
    cost: object
    sampler: object
    N: int
    T: int
    D: int
    dt: float
    lambda_: float
    alpha: float
    x0: np.ndarray  # [D][S]
    U0: np.ndarray  # [D][T][C]
    seed: int = 42
    optimization_stride: int = 1
    extra: dict = field(default_factory=dict)

    @property
    def noise_bytes(self) -> int:
        """Algorithmic HBM bytes of one rollout launch: one read of the unique noise buffer."""
        return self.N * self.T * self.dyn.CONTROL_DIM * 4

    def make_engine(self, **kw) -> "H.Engine":
        e = H.Engine(self.dyn, self.cost, self.sampler, self.N, self.T, self.D, **kw)
        e.set_solver(self.dt, self.lambda_, self.alpha)
        e.seed(self.seed, 0)
        return e


def cartpole(N: int = 8192, T: int = 100) -> Workload:
    """C1/C2: Cartpole + quadratic cost, VanillaMPPI (tests/controllers/vanilla_mppi_test.cu:18-28,81-93,
    examples/cartpole_example.cu:12-13)."""
    dyn = H.CartpoleDynamics(1.0, 1.0, 1.0)
    dyn.setControlRanges([(-5.0, 5.0)])
    cost = H.CartpoleQuadraticCost()
    p = cost.params
    p.cart_position_coeff, p.cart_velocity_coeff = 100.0, 10.0
    p.pole_angle_coeff, p.pole_angular_velocity_coeff = 200.0, 20.0
    p.control_cost_coeff[0] = 1.0
    p.terminal_cost_coeff = 0.0
    p.desired_terminal_state[:] = [-20.0, 0.0, math.pi, 0.0]
    sampler = H.GaussianDistribution(1, [5.0])
    sampler.setControlCostCoeff([1.0])
    x0 = np.zeros((1, 4), np.float32)
    U0 = np.zeros((1, T, 1), np.float32)
    return Workload(f"cartpole_vanilla_N{N}_T{T}", "vanilla", dyn, cost, sampler, N, T, 1, 0.01, 0.25, 0.01, x0, U0)


def double_integrator_tube(N: int = 16384, T: int = 150) -> Workload:
    """C3: DoubleIntegrator circular track (CORL2020), Tube-MPPI (examples/double_integrator_CORL2020.cu:29-39,316-352)."""
    dyn = H.DoubleIntegratorDynamics(1.0)
    cost = H.DoubleIntegratorCircleCost()
    sampler = H.GaussianDistribution(2, [1.0, 1.0])
    x0 = np.tile(np.array([2.0, 0.0, 0.0, 1.0], np.float32), (2, 1))
    U0 = np.zeros((2, T, 2), np.float32)
    return Workload(f"double_integrator_tube_N{N}_T{T}", "tube", dyn, cost, sampler, N, T, 2, 0.02, 2.0, 0.0, x0, U0,
                    extra={"nominal_threshold": 20.0})


def double_integrator_vanilla(N: int = 4096, T: int = 100) -> Workload:
    """DoubleIntegrator with a single distribution (examples/double_integrator_example.cu) — parity-test case."""
    w = double_integrator_tube(N, T)
    w.name, w.controller, w.D = f"double_integrator_vanilla_N{N}_T{T}", "vanilla", 1
    w.x0, w.U0 = w.x0[:1].copy(), w.U0[:1].copy()
    return w


def synthetic_nn_weights(seed: int = 1) -> np.ndarray:
    """theta_i ~ U(-1,1)/sqrt(fan_in), packed W (row-major out x in) then b per layer (fnn_helper.cu:176-183).
    The real Autorally network is a git-LFS stub in the reference tree, so weights are synthetic."""
    rng = np.random.RandomState(seed)
    layers = (6, 32, 32, 4)
    out = []
    for i in range(3):
        fan_in = layers[i]
        out.append((rng.uniform(-1, 1, layers[i + 1] * fan_in) / math.sqrt(fan_in)).astype(np.float32))
        out.append((rng.uniform(-1, 1, layers[i + 1]) / math.sqrt(fan_in)).astype(np.float32))
    th = np.concatenate(out)
    assert th.size == H.AR_NN_NUM_PARAMS
    return th


def track_map_standard() -> tuple:
    """In-memory replica of `track_map_standard.npz` (scripts/autorally/test/generateTestMaps.py:45-75):
    600 x 600 @ 20 ppm, channel0[i][j] = |15 - y| + x/30 with x = j/ppm, y = i/ppm; bounds x in [-13,17], y in [-10,20]."""
    ppm, width, height = 20, 30, 30
    i = np.arange(width * ppm, dtype=np.float64)[:, None]
    j = np.arange(height * ppm, dtype=np.float64)[None, :]
    x, y = j / ppm, i / ppm
    ch0 = (np.abs(height / 2.0 - y) + x / width).astype(np.float32)
    return ch0, (-13.0, 17.0), (-10.0, 20.0), float(ppm)


def autorally(N: int = 32768, T: int = 100) -> Workload:
    """C4: NeuralNetModel<7,2,3> + ARStandardCost on the generated test map (ranges from
    tests/dynamics/ar_dynamics_nn_test.cu:52-58)."""
    dyn = H.NeuralNetModel([(-1.0, 1.0), (-2.0, 2.0)])
    dyn.updateModel([6, 32, 32, 4], synthetic_nn_weights(1))
    cost = H.ARStandardCost()
    ch0, xb, yb, ppm = track_map_standard()
    cost.loadTrackData(ch0, xb[0], xb[1], yb[0], yb[1], ppm)
    sampler = H.GaussianDistribution(2, [0.3, 0.3])
    x0 = np.array([[0.0, 0.0, 0.0, 0.0, 4.0, 0.0, 0.0]], np.float32)
    U0 = np.zeros((1, T, 2), np.float32)
    return Workload(f"autorally_nn_N{N}_T{T}", "vanilla", dyn, cost, sampler, N, T, 1, 0.02, 6.67, 0.0, x0, U0)


def synthetic_lstm_weights(hidden_dim: int = 4, head_hidden: int = 20, seed: int = 2) -> tuple:
    """(lstm, head) ~ U(-1,1)/sqrt(fan_in) in the reference's packed layouts (lstm_helper.cu:72-88, fnn_helper.cu:176-183),
    initial hidden / cell state zero. The RACER networks are not in the reference tree."""
    rng = np.random.RandomState(seed)
    Hd, I = hidden_dim, H.RACER_LSTM_INPUT_DIM
    parts = [rng.uniform(-1, 1, 4 * Hd * Hd) / math.sqrt(Hd + I), rng.uniform(-1, 1, 4 * Hd * I) / math.sqrt(Hd + I),
             rng.uniform(-1, 1, 4 * Hd) / math.sqrt(Hd + I), np.zeros(2 * Hd)]
    lstm = np.concatenate(parts).astype(np.float32)
    IN = Hd + I
    head = np.concatenate([rng.uniform(-1, 1, head_hidden * IN) / math.sqrt(IN), rng.uniform(-1, 1, head_hidden) / math.sqrt(IN),
                           rng.uniform(-1, 1, head_hidden) / math.sqrt(head_hidden),
                           rng.uniform(-1, 1, 1) / math.sqrt(head_hidden)]).astype(np.float32)
    return lstm, head


def racer_lstm(N: int = 65536, T: int = 150, hidden_dim: int = 4, head_hidden: int = 20, colored: bool = True) -> Workload:
    """C5: RacerDubinsElevationLSTMSteering (the in-tree LSTM vehicle model, S19 C2 O28; constructor
    (3, 20, {23, 100, 2H}, 4, H, {H+4, 20, 1}, 11), tests/dynamics/racer_dubins_elevation_lstm_steering_model_test.cu:26-32)
    on flat terrain + ColoredNoise sampler (exponents (1, 1), offset_decay_rate 0.97, colored_noise.cuh:47-49) + our quadratic
    tracking cost."""
    dyn = H.RacerDubinsElevationLSTMSteering(3, 20, (23, 100, 2 * hidden_dim), 4, hidden_dim,
                                             (hidden_dim + 4, head_hidden, 1), 11)
    dyn.setControlRanges([(-1.0, 1.0), (-1.0, 1.0)])  # throttle/brake, steering command
    dyn.setAllValues(*synthetic_lstm_weights(hidden_dim, head_hidden, 2))
    cost = H.RacerQuadraticCost()
    # with the reference's default coefficients (racer_dubins.cuh:78-82) full throttle saturates near 1.6 m/s
    cost.params.desired_speed = 1.2
    if colored:
        sampler = H.ColoredNoiseDistribution(2, [0.3, 0.3], [1.0, 1.0])
    else:
        sampler = H.GaussianDistribution(2, [0.3, 0.3])
    x0 = np.zeros((1, 19), np.float32)
    x0[0, 0] = 3.0  # VEL_X
    x0[0, 9:13] = 1e-6  # covariance diagonal floor (racer_dubins_elevation.cu stateFromMap)
    U0 = np.zeros((1, T, 2), np.float32)
    tag = "colored" if colored else "gaussian"
    return Workload(f"racer_lstm_H{hidden_dim}_{tag}_N{N}_T{T}", "vanilla", dyn, cost, sampler, N, T, 1, 0.02, 1.0, 0.0,
                    x0, U0)


def racer_lstm_gaussian(N: int = 4096, T: int = 100) -> Workload:
    return racer_lstm(N, T, colored=False)


def racer_lstm_h32(N: int = 65536, T: int = 150) -> Workload:
    """C5 at the tensor-core-relevant size of SURVEY §8d: hidden_dim 32 (gate matrix [36 x 128] per step), head {36, 20, 1}."""
    return racer_lstm(N, T, hidden_dim=32, head_hidden=20)


def quadrotor(N: int = 8192, T: int = 100) -> Workload:
    """Quadrotor + quadratic cost, VanillaMPPI (instantiations/quadrotor_mppi/quadrotor_mppi.cuh): fly from the origin
    to a goal 4 m away and 2 m up, hovering there. The only CONTROL_DIM = 4 pair (one 16-byte noise group per step)."""
    dyn = H.QuadrotorDynamics()
    dyn.setControlRanges([(-3.0, 3.0), (-3.0, 3.0), (-3.0, 3.0), (0.0, 36.0)])
    cost = H.QuadrotorQuadraticCost()
    p = cost.params
    p.s_goal[0], p.s_goal[1], p.s_goal[2] = 4.0, 1.0, 2.0
    p.x_coeff, p.v_coeff, p.w_coeff = 10.0, 1.0, 0.5
    p.roll_coeff = p.pitch_coeff = p.yaw_coeff = 5.0
    sampler = H.GaussianDistribution(4, [0.5, 0.5, 0.5, 2.0])
    sampler.setControlCostCoeff([0.1, 0.1, 0.1, 0.01])
    x0 = dyn.getZeroState()[None, :].copy()
    U0 = np.zeros((1, T, 4), np.float32)
    U0[..., 3] = dyn.GRAVITY  # hover thrust (zero_control_[3])
    return Workload(f"quadrotor_vanilla_N{N}_T{T}", "vanilla", dyn, cost, sampler, N, T, 1, 0.02, 1.0, 0.0, x0, U0)


BUILDERS = {
    "racer_lstm": racer_lstm,
    "racer_lstm_gaussian": racer_lstm_gaussian,
    "racer_lstm_h32": racer_lstm_h32,
    "cartpole": cartpole,
    "double_integrator_tube": double_integrator_tube,
    "double_integrator_vanilla": double_integrator_vanilla,
    "autorally": autorally,
    "quadrotor": quadrotor,
}


def by_name(name: str, N: Optional[int] = None, T: Optional[int] = None) -> Workload:
    kw = {}
    if N is not None:
        kw["N"] = N
    if T is not None:
        kw["T"] = T
    return BUILDERS[name](**kw)
