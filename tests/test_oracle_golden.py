"""Pins the CPU oracle (oracle/mppi_oracle.cpp) against the known-answer values the reference's own tests hold for the
hot path. No GPU, no product code: if these fail the oracle is wrong and no parity claim stands."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle
import mppi_generic_b200 as m
from mppi_generic_b200 import workloads as W

H = m.host


def test_fnn_all_ones_gives_33():
    # tests/dynamics/ar_dynamics_nn_test.cu:445-481: theta = 1, s = 0, u = (1,-1)  =>  s_der[3..6] = 33, kinematics 0
    dyn = H.NeuralNetModel()
    theta = np.ones(1412, np.float32)
    x = np.zeros(7, np.float32)
    u = np.array([1.0, -1.0], np.float32)
    xn, xd, y = oracle.dyn_step(H.DYN_AUTORALLY_NN, dyn.params, theta, x, u, 0.1)
    np.testing.assert_array_equal(xd[:3], 0.0)
    np.testing.assert_allclose(xd[3:], 33.0, rtol=4e-7)  # EXPECT_FLOAT_EQ == 4 ULP
    out = oracle.fnn_forward(theta, [6, 32, 32, 4], [0, 0, 0, 0, 1, -1])
    np.testing.assert_allclose(out, 33.0, rtol=4e-7)
    # product host twin agrees (Dynamics::step host method)
    dyn.updateModel([6, 32, 32, 4], theta)
    xn2, xd2, _ = dyn.step(x, u, 0.1)
    np.testing.assert_allclose(xd2, xd, rtol=4e-7)


def test_fnn_weight_layout_matches_reference_packing():
    # tests/nn_helpers/fnn_helper_test.cu:199-253 / fnn_helper.cu:176-183: W (row-major out x in) then b, layer by layer
    rng = np.random.RandomState(0)
    theta = rng.randn(1412).astype(np.float32)
    W1, b1 = theta[:192].reshape(32, 6), theta[192:224]
    W2, b2 = theta[224:1248].reshape(32, 32), theta[1248:1280]
    W3, b3 = theta[1280:1408].reshape(4, 32), theta[1408:]
    x = rng.randn(6).astype(np.float32)
    ref = W3.astype(np.float64) @ np.tanh(W2.astype(np.float64) @ np.tanh(W1.astype(np.float64) @ x + b1) + b2) + b3
    out = oracle.fnn_forward(theta, [6, 32, 32, 4], x)
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-4)  # fnn_helper_test.cu:497-546 tolerance 1e-4


def _ar_cost():
    cost = H.ARStandardCost()
    ch0, xb, yb, ppm = W.track_map_standard()
    cost.loadTrackData(ch0, xb[0], xb[1], yb[0], yb[1], ppm)
    return cost


def test_ar_standard_cost_individual_terms():
    # tests/cost_functions/autorally_standard_cost_test.cu:897-982 on track_map_standard
    cost = _ar_cost()
    p = cost.params
    p.discount = 0.9
    s = np.array([3.0, 0.0, math.pi / 2, 0.0, 2.0, 1.0, 0.1, 0.0], np.float32)
    terms = oracle.ar_cost_terms(p, cost.costmap, s, crash=0)
    assert terms["speed"] == pytest.approx(4.0 ** 2 * 4.25, rel=4e-7)           # 68.0
    assert terms["stabilizing"] == pytest.approx(math.atan(0.5) ** 2 * 10, rel=1e-6)
    # Track term: the state puts both lookups EXACTLY on texel boundaries (u*600 = 320.00002, v*600 = 210.00001 and
    # 189.99999). The reference's GPU texture unit resolves them to texels (319,209)/(319,189) => 1116.3333, the value
    # its test hard-codes for the DEVICE path (autorally_standard_cost_test.cu:958-960); the reference's HOST branch
    # (ar_standard_cost.cu:236-242: float v*600 rounds to 190.0, then -0.5, clamp, std::round) lands on (320,210)/(320,190)
    # => 1106.6666. The oracle
    # restates the host branch; the device value is pinned on the GPU in tests/test_gpu_parity.py.
    host_track = 200.0 * ((4.5 + 320 / 600.0) + (5.5 + 320 / 600.0)) / 2.0
    assert terms["track"] == pytest.approx(host_track, rel=1e-6)
    assert abs(terms["track"] - 1116.3333) < 200 * 0.05  # within one texel of the device golden value
    assert terms["crash_status"] == 1 and terms["crash"] == pytest.approx(10000.0)
    total, _, crash = oracle.state_cost(H.COST_AR_STANDARD, p, cost.costmap, s, t=1, crash=0)
    expect = 68.0 + math.atan(0.5) ** 2 * 10 + host_track + 9000.0
    assert total == pytest.approx(expect, rel=1e-6) and crash == 1
    total4, _, _ = oracle.state_cost(H.COST_AR_STANDARD, p, cost.costmap, s, t=4, crash=0)
    assert total4 == pytest.approx(68.0 + math.atan(0.5) ** 2 * 10 + host_track + 0.9 ** 4 * 10000, rel=1e-6)


def test_ar_cost_texture_lookup_is_point_sampled_and_clamped():
    # texture emulation of ar_standard_cost.cu:225-243: normalised coords, -0.5, clamp, round
    cost = _ar_cost()
    p = cost.params
    ch0 = cost.costmap[..., 0]
    # world (x,y) -> texel (j,i): j = (x+13)*20 - 0.5, i = (y+10)*20 - 0.5
    for (x, y) in [(0.0, 0.0), (3.0, 0.5), (-12.98, 19.9), (16.99, -9.99)]:
        j = int(round(min(max((x + 13) * 20 - 0.5, 0), 599)))
        i = int(round(min(max((y + 10) * 20 - 0.5, 0), 599)))
        assert oracle.ar_query_texture(p, cost.costmap, x, y) == pytest.approx(float(ch0[i, j]), rel=1e-6)
    # out of bounds clamps to the edge texel
    assert oracle.ar_query_texture(p, cost.costmap, -100.0, -100.0) == pytest.approx(float(ch0[0, 0]))
    assert oracle.ar_query_texture(p, cost.costmap, 100.0, 100.0) == pytest.approx(float(ch0[599, 599]))


def test_enforce_constraints_known_answers():
    # tests/dynamics/dynamics_generic_tests.cu:259-283 and :285-357
    lim = H.ControlLimits()
    lim.rng_lo[0], lim.rng_hi[0] = -2, 5
    for u, expect in [(100, 5), (-42178, -2), (2, 2), (-1.5, -1.5)]:
        assert oracle.enforce_constraints(lim, [u])[0] == expect
    lim.rng_lo[1], lim.rng_hi[1] = -6, 8
    lim.rng_lo[2], lim.rng_hi[2] = -11, 16
    np.testing.assert_array_equal(oracle.enforce_constraints(lim, [48, 48, 48]), [5, 8, 16])
    np.testing.assert_array_equal(oracle.enforce_constraints(lim, [-51, -51, -51]), [-2, -6, -11])
    np.testing.assert_array_equal(oracle.enforce_constraints(lim, [-1.5, -1.5, -1.5]), [-1.5, -1.5, -1.5])
    # deadband semantics (dynamics.cuh:254-261)
    lim = H.ControlLimits()
    lim.deadband[0] = 0.5
    assert oracle.enforce_constraints(lim, [0.3])[0] == 0.0
    assert oracle.enforce_constraints(lim, [1.0])[0] == 0.5
    assert oracle.enforce_constraints(lim, [-1.0])[0] == -0.5


def test_update_state_euler():
    # tests/dynamics/dynamics_generic_tests.cu:359-417: s=(0,1,2,3), s_der=(0,1,2,3), dt=0.1 -> (0,1.1,2.2,3.3)
    dyn = H.DoubleIntegratorDynamics()
    x = np.array([0, 1, 2, 3], np.float32)
    u = np.array([2, 3], np.float32)  # xdot = (x2, x3, u0, u1) = (2,3,2,3)
    xn, xd, y = oracle.dyn_step(H.DYN_DOUBLE_INTEGRATOR, dyn.params, None, x, u, 0.1)
    np.testing.assert_allclose(xd, [2, 3, 2, 3])
    np.testing.assert_allclose(xn, [0.2, 1.3, 2.2, 3.3], rtol=4e-7)
    np.testing.assert_array_equal(y, xn)


def test_cartpole_dynamics_formula():
    # dynamics/cartpole/cartpole_dynamics.cu:48-69 evaluated in float64 (tests/dynamics/cartpole_dynamics_tests.cu:153-199
    # only compares CPU with GPU; the formula itself is the known answer)
    dyn = H.CartpoleDynamics(2.0, 3.0, 4.0)
    x = np.array([0.1, 0.3, 0.23, 0.334], np.float32)
    u = np.array([0.654], np.float32)
    _, xd, _ = oracle.dyn_step(H.DYN_CARTPOLE, dyn.params, None, x, u, 0.01)
    th, thd, f, mc, mp_, lp, g = 0.23, 0.334, 0.654, 2.0, 3.0, 4.0, 9.81
    s, c = math.sin(th), math.cos(th)
    ref1 = 1.0 / (mc + mp_ * s * s) * (f + mp_ * s * (lp * thd * thd + g * c))
    ref3 = 1.0 / (lp * (mc + mp_ * s * s)) * (-f * c - mp_ * lp * thd * thd * c * s - (mc + mp_) * g * s)
    np.testing.assert_allclose(xd, [0.3, ref1, 0.334, ref3], rtol=2e-6)


def test_cartpole_quadratic_cost_known_answer():
    # tests/cost_functions/cartpole_quadratic_cost_test.cu:109-177 (defaults 1000,100,2000,100; goal (0,0,pi,0))
    cost = H.CartpoleQuadraticCost()
    s = np.array([1, 2, 3, 4], np.float32)
    c, term, _ = oracle.state_cost(H.COST_CARTPOLE_QUADRATIC, cost.params, None, s)
    ref = 1 * 1000 + 4 * 100 + (3 - math.pi) ** 2 * 2000 + 16 * 100
    assert c == pytest.approx(ref, rel=1e-6) and term == 0.0
    cost.params.terminal_cost_coeff = 2.5
    _, term, _ = oracle.state_cost(H.COST_CARTPOLE_QUADRATIC, cost.params, None, s)
    assert term == pytest.approx(2.5 * ref, rel=1e-6)


def test_double_integrator_circle_cost():
    # cost_functions/double_integrator/double_integrator_circle_cost.cu:34-59
    cost = H.DoubleIntegratorCircleCost()
    c, _, _ = oracle.state_cost(H.COST_DI_CIRCLE, cost.params, None, np.array([2, 0, 0, 2], np.float32))
    assert c == pytest.approx(0.0, abs=1e-6)  # on the circle, |v| = 2, L = 4
    c, _, _ = oracle.state_cost(H.COST_DI_CIRCLE, cost.params, None, np.array([3, 0, 0, 1], np.float32), t=5)
    assert c == pytest.approx(1000.0 + 1.0 + 1.0, rel=1e-6)  # crash + |1-2| + |3-4|


def test_savitzky_golay_smoothing_known_answers():
    # tests/controllers/controller_generic_tests.cu:214-239
    hist = np.zeros((2, 3), np.float32)
    u = np.ones((1, 3), np.float32)
    out = oracle.smooth(u, hist)
    np.testing.assert_allclose(out[0], (17 + 12 - 3) / 35.0, rtol=4e-7)
    u = np.array([[1, 1, 1], [2, 2, 2]], np.float32)
    out = oracle.smooth(u, hist)
    np.testing.assert_allclose(out[0], (1 * 17 + 2 * 12 + 2 * -3) / 35.0, rtol=4e-7)
    np.testing.assert_allclose(out[1], (1 * 12 + 2 * 17 + 2 * 12 + 2 * -3) / 35.0, rtol=4e-7)
    # product host twin == oracle
    L = H.lib()
    v = u.copy()
    L.mppib_host_smooth_controls(v.ctypes.data, hist.ctypes.data, 2, 3)
    np.testing.assert_array_equal(v, out)


def test_slide_control_sequence_known_answers():
    # tests/controllers/controller_generic_tests.cu:241-282 (T=100, scale 0 -> tail becomes zero_control)
    T = 100
    u = np.repeat(np.arange(T, dtype=np.float32)[:, None], 2, axis=1)
    z, sc = np.zeros(2, np.float32), np.zeros(2, np.float32)
    s1 = oracle.slide(u, 1, z, sc)
    for i in range(T):
        assert s1[i, 0] == (0 if i + 1 > T - 1 else min(i + 1, T - 1))
    s2 = oracle.slide(s1, 10, z, sc)
    for i in range(T):
        assert s2[i, 0] == (0 if i + 10 > T - 2 else min(i + 11, T - 1))
    v = u.copy()
    H.lib().mppib_host_slide_controls(v.ctypes.data, 1, T, 2, z.ctypes.data, sc.ctypes.data)
    np.testing.assert_array_equal(v, s1)


def test_norm_exp_baseline_normalizer_identities():
    # tests/mppi_core/normexp_kernel_tests.cu:79-178: N=555 / 28754 / 6048, gamma = 0.3, costs ~ N(100, 2)
    rng = np.random.RandomState(7)
    for n in (555, 6048, 28754):
        costs = (100 + 2 * rng.randn(n)).astype(np.float32)
        base = oracle.baseline(costs)
        assert base == costs.min()
        w = oracle.norm_exp(costs, 0.3, base)
        ref = np.exp(np.float32(-0.3) * (costs - np.float32(base))).astype(np.float32)
        np.testing.assert_allclose(w, ref, rtol=4e-7)
        assert oracle.normalizer(w) == pytest.approx(float(w.astype(np.float64).sum()), rel=1e-7)
    # first minimum wins; value identical either way (mppi_common.cu:885-900)
    assert oracle.baseline(np.array([3, 1, 2, 1], np.float32)) == 1.0


def test_free_energy_formula():
    # core/mppi_common.cu:1065-1081
    rng = np.random.RandomState(3)
    w = rng.rand(1000).astype(np.float32)
    fe = oracle.free_energy(w, 5.0, 2.0)
    norm = w.astype(np.float64).sum() / 1000
    var = (w.astype(np.float64) ** 2).sum()
    assert fe[0] == pytest.approx(-2.0 * math.log(norm) + 5.0, rel=1e-5)
    assert fe[1] == pytest.approx(2.0 * (var / 1000 - norm ** 2), rel=1e-4)


def test_weighted_reduction_matches_triple_loop():
    # tests/mppi_core/weightedreduction_kernel_tests.cu:20-133: N=1024, C=6, T=100, stride 64
    rng = np.random.RandomState(11)
    N, T, Cd = 1024, 100, 6
    w = rng.rand(N).astype(np.float32)
    du = rng.randn(N, T, Cd).astype(np.float32)
    eta = float(w.astype(np.float64).sum())
    out = oracle.weighted_reduction(w, du, eta, T, N, Cd, 64)
    ref = np.einsum("n,ntc->tc", w.astype(np.float64) / eta, du.astype(np.float64))
    np.testing.assert_allclose(out, ref, rtol=2e-5, atol=2e-6)


def test_set_gaussian_controls_three_cases():
    # sampling_distributions/gaussian/gaussian.cu:101-121
    N, T, Cd = 200, 6, 2
    sp = H.GaussianDistribution(2, [0.5, 2.0]).params
    eps = np.random.RandomState(5).randn(1, N, T, Cd).astype(np.float32)
    mean = np.arange(T * Cd, dtype=np.float32).reshape(1, T, Cd)
    s = eps.copy()
    oracle.set_gaussian_controls(mean, sp, s, Cd, T, N, 1, optimization_stride=2)
    np.testing.assert_array_equal(s[0, 0], mean[0])                    # sample 0 == mean
    np.testing.assert_array_equal(s[0, :, :2], np.broadcast_to(mean[0, :2], (N, 2, Cd)))  # t < stride == mean
    sd = np.array([0.5, 2.0], np.float32)
    first_pure = int(math.ceil((1.0 - 0.01) * N))  # n >= 0.99*N  -> 198
    np.testing.assert_array_equal(s[0, first_pure:, 2:], sd * eps[0, first_pure:, 2:])
    np.testing.assert_allclose(s[0, 1:first_pure, 2:], mean[0, 2:] + sd * eps[0, 1:first_pure, 2:], rtol=1e-6)


def test_curand_host_stream_properties():
    # same generator type / seed / offset semantics as controllers/controller.cu:192-207. Values are third-party
    # (libcurand): pinned here only structurally — determinism, continuation, offsets honoured at multiples of 8192.
    a = oracle.curand_normal(42, 0, 32768)
    b = oracle.curand_normal(42, 0, 32768)
    np.testing.assert_array_equal(a, b)
    c = oracle.curand_normal(42, 16384, 16384)
    np.testing.assert_array_equal(a[16384:], c)
    assert abs(float(a.mean())) < 0.03 and abs(float(a.std()) - 1.0) < 0.03
    assert not np.array_equal(a, oracle.curand_normal(43, 0, 32768))


def test_rollout_actual_equals_nominal_when_inputs_equal():
    # tests/mppi_core/rollout_kernel_tests.cu:181-198: same x0 / mean for both systems -> identical costs
    w = W.double_integrator_tube(N=256, T=20)
    eps = oracle.curand_normal(7, 0, 256 * 20 * 2).reshape(256, 20, 2)
    r = oracle.solve(H.DYN_DOUBLE_INTEGRATOR, H.COST_DI_CIRCLE, w.dyn.params, w.cost.params, w.sampler.params, None,
                     None, 256, 20, 2, 2, w.dt, w.lambda_, w.alpha, w.x0, w.U0, eps)
    np.testing.assert_array_equal(r["costs"][0], r["costs"][1])
    np.testing.assert_array_equal(r["U"][0], r["U"][1])


def test_output_trajectory_host_twin_matches_oracle():
    # controllers/controller.cuh:643-663 — product host twin (libmppi_b200) vs oracle restatement
    for w in (W.cartpole(64, 30), W.autorally(64, 30)):
        rng = np.random.RandomState(2)
        u = rng.randn(w.T, w.dyn.CONTROL_DIM).astype(np.float32)
        st, out = oracle.output_trajectory(w.dyn.DYN_ID, w.dyn.params, w.dyn.nn_theta, w.x0[0], u, w.dt)
        st2 = np.zeros_like(st)
        out2 = np.zeros_like(out)
        rc = H.lib().mppib_host_output_trajectory(w.dyn.DYN_ID, C.byref(w.dyn.params),
                                                  None if w.dyn.nn_theta is None else w.dyn.nn_theta.ctypes.data,
                                                  w.x0[0].ctypes.data, u.ctypes.data, w.T, C.c_float(w.dt),
                                                  st2.ctypes.data, out2.ctypes.data)
        assert rc == 0
        np.testing.assert_allclose(st2, st, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(out2, out, rtol=1e-5, atol=1e-6)


# ---- LSTM / RACER model / ColoredNoise -------------------------------------------------
def test_lstm_forward_all_ones_known_answers():
    # tests/nn_helpers/lstm_helper_test.cu:677-718 (forwardCPU): LSTMHelper(8, 20, {28, 3}), every weight, bias and the
    # initial hidden / cell state = 1, input = 1  =>  28.28055, 28.901096, 28.986588, 28.998184, 28.999756
    I, Hd = 8, 20
    w = np.ones(4 * Hd * Hd + 4 * Hd * I + 4 * Hd, np.float32)
    head = np.ones(28 * 3 + 3, np.float32)
    h, c = np.ones(Hd, np.float32), np.ones(Hd, np.float32)
    for expect in (28.28055, 28.901096, 28.986588, 28.998184, 28.999756):
        out, h, c = oracle.lstm_forward(w, I, Hd, head, [28, 3], np.ones(I, np.float32), h, c)
        np.testing.assert_allclose(out, expect, rtol=4e-7)  # EXPECT_FLOAT_EQ


def test_racer_lstm_host_twin_matches_oracle():
    """The product's host twin (host_twins.cpp, used by the controller tail) and the oracle restate the same host path
    (racer_dubins_elevation_lstm_steering.cu:90-118) independently: they must agree over a closed-loop trajectory."""
    w = W.racer_lstm(N=64, T=10)
    dyn = w.dyn
    oracle.set_lstm(dyn.lstm_theta, dyn.hidden_dim, dyn.head_hidden)
    rng = np.random.RandomState(0)
    x = w.x0[0].copy()
    x[1], x[4] = 0.3, 0.1
    x[9:19] = rng.uniform(0, 0.01, 10)
    h, c = dyn.initial_hidden_cell()
    xo, ho, co = x.copy(), h.copy(), c.copy()
    for t in range(60):
        u = rng.uniform(-1, 1, 2).astype(np.float32)
        xn, xd, y, h, c = dyn.step(x, u, 0.02, h, c)
        xn2, xd2, y2, ho, co = oracle.racer_step(dyn.params, xo, u, 0.02, ho, co)
        # the host twin evaluates the activations with its libm-free exp (|abs error| < 2e-7), the oracle with expf/tanhf
        np.testing.assert_allclose(xn, xn2, rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(xd, xd2, rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(y, y2, rtol=2e-5, atol=2e-5)  # NaN == NaN here (WHEEL_FORCE_* outputs)
        x, xo = xn, xn2
    assert np.all(np.isfinite(x))
    # known structure of one step (racer_dubins_elevation.cu:69-227, racer_dubins.cu:427-432): flat terrain, outputs
    assert y[4] == 0.0 and y[6] == 0.0 and y[7] == 0.0 and np.isnan(y[10:13]).all()
    assert y[0] == xn[0] and y[2] == xn[2] and y[3] == xn[3] and y[16] == abs(xn[0])


def test_racer_parametric_known_answers():
    """Hand-evaluated values of the parametric part (racer_dubins.cu:306-319, racer_dubins_elevation.cu:32-67) with the
    LSTM silenced (all weights zero => head output 0)."""
    dyn = H.RacerDubinsElevationLSTMSteering()
    dyn.setControlRanges([(-1.0, 1.0), (-1.0, 1.0)])
    oracle.set_lstm(dyn.lstm_theta, dyn.hidden_dim, dyn.head_hidden)
    x = np.zeros(19, np.float32)
    h, c = dyn.initial_hidden_cell()
    # at rest, full throttle: index 0, |vx| <= 0.2 => throttle = c_t[0] * (1 - 0.13); xdot_vx = that + c_0, clamped to 5.5
    xn, xd, y, _, _ = oracle.racer_step(dyn.params, x, np.array([1.0, 0.0], np.float32), 0.1, h, c)
    assert xd[0] == pytest.approx(min(1.3 * (1.0 - 0.13) + 4.9, 5.5), rel=1e-6)
    assert xn[0] == pytest.approx(xd[0] * 0.1, rel=1e-6)
    assert xd[5] == 0.0 and xd[1] == 0.0
    # moving at 4 m/s (index 2), braking: brake state rises at min(1 * 6.6, 0.33); drag c_v[2] * 4
    x[0] = 4.0
    xn, xd, y, _, _ = oracle.racer_step(dyn.params, x, np.array([-1.0, 0.5], np.float32), 0.1, h, c)
    assert xd[5] == pytest.approx(0.33)
    assert xd[0] == pytest.approx(max(-5.7 * 4.0 + 4.9, -5.5))
    # steering: rate derivative = clamp(((0.5*5 - 0) * 0.6 - 0) * 12.1 - 0, +-5) = 5, angle derivative = current rate = 0
    assert xd[8] == pytest.approx(5.0) and xd[4] == 0.0
    assert xn[8] == pytest.approx(0.5)
    assert xd[2] == pytest.approx(4.0) and xd[3] == pytest.approx(0.0)


def test_colored_noise_matches_the_numpy_algorithm():
    """colored_noise.cu:286-372 mirrors scripts/colored_noise.py:12-104 (Timmer & Koenig via numpy irfft) plus the offset
    subtraction of rearrangeNoise (:39-56); the oracle's restatement must reproduce that algorithm."""
    N, Cd, T = 64, 2, 50
    betas = [1.0, 2.0]
    s = H.ColoredNoiseDistribution(Cd, [0.3, 0.3], betas)
    normals = oracle.curand_normal(42, 0, 2 * N * Cd * (T + 1))
    for offset_t in (1, 4):
        eps = oracle.colored_noise(normals, s.params, N, Cd, T, offset_t=offset_t)
        z = normals.reshape(N, Cd, T + 1, 2).astype(np.float64)
        ref = np.zeros((N, T, Cd))
        for c, beta in enumerate(betas):
            samples = 2 * T
            f = np.fft.rfftfreq(samples)
            fmin = max(0.0, 1.0 / samples)
            ix = int(np.sum(f < fmin))
            sc = f.copy()
            if ix and ix < len(sc):
                sc[:ix] = sc[ix]
            sc = sc ** (-beta / 2.0)
            wv = sc[1:].copy()
            wv[-1] *= (1 + (samples % 2)) / 2.0
            sigma = 2 * np.sqrt(np.sum(wv ** 2)) / samples
            sr, si = z[:, c, :, 0] * sc, z[:, c, :, 1] * sc
            si[:, -1] = 0
            si[:, 0] = 0
            yv = np.fft.irfft(sr + 1j * si, n=samples, axis=-1)[:, :T] / sigma
            ref[:, :, c] = yv - yv[:, offset_t:offset_t + 1] * (0.97 ** np.arange(T))
        np.testing.assert_allclose(eps, ref, atol=5e-6)
        # unit variance before the offset is subtracted (Timmer & Koenig normalisation); pooled over N*T samples
    tab, sigma = oracle.colored_tables(s.params, Cd, T)
    assert tab.shape == (Cd, T + 1) and np.all(sigma > 0)


# ---- RMPPI host logic ----------------------------------------------------------------------------------
def test_rmppi_line_search_weights_strides_and_best_index():
    """controllers/R-MPPI/robust_mppi_controller.cu:472-537 — hand-evaluated values + product host twin == oracle."""
    K = 9
    w = oracle.rmppi_line_search_weights(K)
    # first half interpolates nominal_x_k -> nominal_x_kp1, second half nominal_x_kp1 -> real_x_kp1 (:476-490)
    np.testing.assert_allclose(w[0], [1, .75, .5, .25, 0, 0, 0, 0, 0])
    np.testing.assert_allclose(w[1], [0, .25, .5, .75, 1, .75, .5, .25, 0])
    np.testing.assert_allclose(w[2], [0, 0, 0, 0, 0, .25, .5, .75, 1])
    np.testing.assert_allclose(w.sum(axis=0), 1.0)
    w2 = np.zeros((3, K), np.float32)
    H.lib().mppib_host_rmppi_line_search_weights(K, w2.ctypes.data)
    np.testing.assert_array_equal(w, w2)
    # strides = round([0, s, s] . w) (:493-503)
    np.testing.assert_array_equal(oracle.rmppi_strides(K, 1), [0, 0, 1, 1, 1, 1, 1, 1, 1])  # .25 -> 0, .5 -> 1 (Eigen round: half away from zero)
    np.testing.assert_array_equal(oracle.rmppi_strides(K, 4), [0, 1, 2, 3, 4, 4, 4, 4, 4])
    rng = np.random.RandomState(3)
    xk, xk1, xr = rng.randn(3, 4).astype(np.float32)
    cand = np.zeros((K, 4), np.float32)
    st = np.zeros(K, np.int32)
    H.lib().mppib_host_rmppi_candidates(K, 4, xk.ctypes.data, xk1.ctypes.data, xr.ctypes.data, 4, cand.ctypes.data,
                                        st.ctypes.data)
    np.testing.assert_array_equal(st, oracle.rmppi_strides(K, 4))
    np.testing.assert_allclose(cand, (np.stack([xk, xk1, xr], 1) @ w).T, rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(cand[0], xk)
    np.testing.assert_array_equal(cand[K // 2], xk1)
    np.testing.assert_array_equal(cand[-1], xr)
    # best index: the LAST candidate whose free energy is under the threshold; none => previous value kept
    spc = 16
    costs = (10.0 + rng.rand(K * spc)).astype(np.float32)
    costs[2 * spc:3 * spc] -= 5.0
    costs[6 * spc:7 * spc] -= 4.0
    best, fe = oracle.rmppi_best_index(costs, K, spc, 1.0, 8.0)
    assert best == 6 and fe[2] < 8.0 and fe[6] < 8.0 and fe[0] > 8.0
    fe2 = np.zeros(K, np.float32)
    assert H.lib().mppib_host_rmppi_best_index(costs.ctypes.data, K, spc, C.c_float(1.0), C.c_float(8.0), 3,
                                               fe2.ctypes.data) == 6
    np.testing.assert_allclose(fe2, fe, rtol=1e-6)
    assert H.lib().mppib_host_rmppi_best_index(costs.ctypes.data, K, spc, C.c_float(1.0), C.c_float(1.0), 3,
                                               fe2.ctypes.data) == 3
    assert oracle.rmppi_best_index(costs, K, spc, 1.0, 1.0)[0] == -1
    # free energy formula: -lambda log(mean exp(-(c - b)/lambda)) + b
    b = costs.min()
    ref = -1.0 * np.log(np.mean(np.exp(-(costs[:spc].astype(np.float64) - b)))) + b
    assert fe[0] == pytest.approx(ref, rel=1e-5)


def test_rmppi_oracle_reduces_to_plain_rollout_without_feedback():
    """With zero gains, an infinite value-function threshold floor... the real system's RMPPI cost is the plain rollout
    cost (running + LR + terminal)/T, and the nominal cost is 0.5 c + 0.5 max(min(tracking_real, thr), c) + LR."""
    w = W.double_integrator_tube(256, 40)
    sp = w.sampler.params
    sp.control_cost_coeff[0] = sp.control_cost_coeff[1] = 0.5
    N, T, Cd = w.N, w.T, 2
    eps = oracle.curand_normal(5, 0, N * T * Cd).reshape(N, T, Cd)
    U = np.zeros((2, T, Cd), np.float32)
    U[:, :, 0] = 0.3
    x0 = np.array([[2.0, 0.0, 0.0, 1.0], [2.05, 0.02, 0.0, 1.0]], np.float32)  # [nominal, real]
    samples = np.stack([eps, eps]).copy()
    oracle.set_gaussian_controls(U, sp, samples, Cd, T, N, 2, 1, 0)
    plain = samples.copy()
    c_plain = oracle.rollout(w.dyn.DYN_ID, w.cost.COST_ID, w.dyn.params, w.cost.params, sp, None, None, N, T, 2, w.dt,
                             w.lambda_, w.alpha, x0, U, plain)
    thr = 10.0
    c = oracle.rmppi_rollout(w.dyn.DYN_ID, w.cost.COST_ID, w.dyn.params, w.cost.params, sp, None, None, N, T, w.dt,
                             w.lambda_, w.alpha, thr, x0, U, None, samples)
    np.testing.assert_allclose(c[1], c_plain[1], rtol=2e-6)
    # nominal: plain cost = (state + LR)/T; split it with a run without LR
    sp0 = type(sp).from_buffer_copy(bytes(sp))
    sp0.control_cost_coeff[0] = sp0.control_cost_coeff[1] = 0.0
    c_state = oracle.rollout(w.dyn.DYN_ID, w.cost.COST_ID, w.dyn.params, w.cost.params, sp0, None, None, N, T, 2, w.dt,
                             w.lambda_, w.alpha, x0, U, plain.copy())
    lr_nom = c_plain[0] - c_state[0]
    tracking_real = c_state[1]  # zero gains: tracking cost = state cost + 0 feedback cost
    expect = 0.5 * c_state[0] + 0.5 * np.maximum(np.minimum(tracking_real, thr), c_state[0]) + lr_nom
    np.testing.assert_allclose(c[0], expect, rtol=2e-5, atol=1e-5)


# ---- Quadrotor (the CONTROL_DIM = 4 pair) -----------------------------------------------------------
def _rand_quad_state(rng):
    s = rng.randn(13).astype(np.float32)
    s[6:10] /= np.linalg.norm(s[6:10])
    return s


def test_quadrotor_dynamics_known_answers_and_independent_formula():
    from scipy.spatial.transform import Rotation
    dyn = m.QuadrotorDynamics()
    # hover: identity attitude, thrust = m g -> no acceleration; quaternion stays (1, 0, 0, 0)
    x = dyn.getZeroState()
    xn, xd, y = oracle.dyn_step(H.DYN_QUADROTOR, dyn.params, None, x, [0, 0, 0, 9.81], 0.01)
    np.testing.assert_array_equal(xd, np.zeros(13, np.float32))
    np.testing.assert_array_equal(xn, x)
    np.testing.assert_array_equal(y, x)
    # free fall
    xn, xd, _ = oracle.dyn_step(H.DYN_QUADROTOR, dyn.params, None, x, [0, 0, 0, 0], 0.01)
    assert xd[5] == np.float32(-9.81) and xn[5] == np.float32(-9.81) * np.float32(0.01)
    # body-rate lag: w_dot = (u - w) / tau
    xn, xd, _ = oracle.dyn_step(H.DYN_QUADROTOR, dyn.params, None, x, [1, -2, 0.5, 9.81], 0.01)
    np.testing.assert_allclose(xd[10:], [4.0, -8.0, 2.0], rtol=1e-6)
    # random states against an independent float64 formulation (scipy rotation matrix, quaternion kinematics
    # q_dot = 0.5 q (x) (0, w)), quadrotor_dynamics_tests.cu:25-75 compares the same quantities CPU vs GPU
    rng = np.random.RandomState(0)
    dyn2 = m.QuadrotorDynamics(mass=2.5)
    dt = 0.01
    for _ in range(20):
        s, u = _rand_quad_state(rng), rng.randn(4).astype(np.float32)
        xn, xd, y = oracle.dyn_step(H.DYN_QUADROTOR, dyn2.params, None, s, u, dt)
        w, xq, yq, zq = s[6:10].astype(np.float64)
        R = Rotation.from_quat([xq, yq, zq, w]).as_matrix()
        ref = np.zeros(13)
        ref[0:3] = s[3:6]
        ref[3:6] = (u[3] / 2.5) * R[:, 2] - np.array([0, 0, 9.81])
        om = s[10:13].astype(np.float64)
        ref[6] = -0.5 * np.dot([xq, yq, zq], om)
        ref[7:10] = 0.5 * (w * om + np.cross([xq, yq, zq], om))
        ref[10:13] = (u[:3] - s[10:13]) / 0.25
        np.testing.assert_allclose(xd, ref, rtol=2e-5, atol=2e-6)
        nx = s.astype(np.float64) + ref * dt
        nx[6:10] /= np.linalg.norm(nx[6:10]) * np.copysign(1.0, nx[6])
        np.testing.assert_allclose(xn, nx, rtol=2e-5, atol=2e-6)
        assert xn[6] >= 0 and abs(np.linalg.norm(xn[6:10]) - 1) < 1e-6  # updateState: unit norm, w >= 0
        np.testing.assert_array_equal(y, xn)
    # the default constructor's thrust range and zero control (quadrotor_dynamics.cu:11-19)
    assert (dyn.params.lim.rng_lo[3], dyn.params.lim.rng_hi[3]) == (0.0, 36.0)
    assert dyn.zero_control_[3] == np.float32(9.81)


def test_quadrotor_cost_known_answers():
    from scipy.spatial.transform import Rotation
    cost = m.QuadrotorQuadraticCost()
    p = cost.params
    goal = np.array(list(p.s_goal), np.float32)
    c, term, _ = oracle.state_cost(H.COST_QUADROTOR_QUADRATIC, p, None, goal)
    assert c == 0.0 and term == 0.0
    # position / velocity / body-rate terms are coefficient * squared error
    p.x_coeff, p.v_coeff, p.w_coeff = 3.0, 5.0, 7.0
    s = goal.copy()
    s[0], s[4], s[12] = 2.0, -1.0, 0.5
    c, _, _ = oracle.state_cost(H.COST_QUADROTOR_QUADRATIC, p, None, s)
    assert c == pytest.approx(3.0 * 4.0 + 5.0 * 1.0 + 7.0 * 0.25, rel=1e-6)
    # attitude term: Euler angles of q_goal * q^-1 (quadrotor_quadratic_cost.cu:27-41); a pure yaw of +0.3 rad away from
    # the identity goal gives yaw_coeff * 0.3^2, and roll / pitch likewise
    p.roll_coeff, p.pitch_coeff, p.yaw_coeff = 2.0, 3.0, 4.0
    for axis, coeff in (("x", 2.0), ("y", 3.0), ("z", 4.0)):
        xq, yq, zq, w = Rotation.from_euler(axis, 0.3).as_quat()
        s = goal.copy()
        s[6:10] = [w, xq, yq, zq]
        c, _, _ = oracle.state_cost(H.COST_QUADROTOR_QUADRATIC, p, None, s)
        assert c == pytest.approx(coeff * 0.09, rel=1e-5)
    p.terminal_cost_coeff = 0.5
    c, term, _ = oracle.state_cost(H.COST_QUADROTOR_QUADRATIC, p, None, s)
    assert term == pytest.approx(0.5 * c, rel=1e-6)
    assert list(cost.getDesiredState()) == list(goal)


def test_quadrotor_host_twin_matches_oracle():
    w = W.quadrotor(64, 40)
    rng = np.random.RandomState(3)
    u = (rng.randn(w.T, 4) * [1, 1, 1, 3] + [0, 0, 0, 9.81]).astype(np.float32)
    st, out = oracle.output_trajectory(w.dyn.DYN_ID, w.dyn.params, None, w.x0[0], u, w.dt)
    st2, out2 = np.zeros_like(st), np.zeros_like(out)
    w.dyn.output_trajectory(w.x0[0], u, w.T, w.dt, st2, out2)
    np.testing.assert_array_equal(st2, st)
    np.testing.assert_array_equal(out2, out)
    assert np.all(np.abs(np.linalg.norm(st[:, 6:10], axis=1) - 1) < 1e-5)
    xn, xd, y = w.dyn.step(st[5], u[5], w.dt)
    u5 = oracle.enforce_constraints(w.dyn.params.lim, u[5])
    xn2, xd2, y2 = oracle.dyn_step(w.dyn.DYN_ID, w.dyn.params, None, st[5], u[5], w.dt)
    np.testing.assert_array_equal(xn, xn2)
    np.testing.assert_array_equal(xd, xd2)
    del u5


def test_quadrotor_oracle_solve_flies_towards_the_goal():
    # closed loop on the oracle alone: a few MPPI iterations must cut the distance to the goal
    w = W.quadrotor(512, 50)
    x = w.x0.copy()
    U = w.U0.copy()
    goal = np.array(list(w.cost.params.s_goal[:3]), np.float32)
    d0 = np.linalg.norm(x[0, :3] - goal)
    for it in range(40):
        eps = oracle.curand_normal(42, it * 512 * 50 * 4, 512 * 50 * 4).reshape(512, 50, 4)
        r = oracle.solve(w.dyn.DYN_ID, w.cost.COST_ID, w.dyn.params, w.cost.params, w.sampler.params, None, None, 512, 50,
                         1, 4, w.dt, w.lambda_, w.alpha, x, U, eps)
        U = r["U"]
        u0 = oracle.enforce_constraints(w.dyn.params.lim, U[0, 0])
        x[0], _, _ = oracle.dyn_step(w.dyn.DYN_ID, w.dyn.params, None, x[0], u0, w.dt)
        U[0, :-1] = U[0, 1:]
    assert np.linalg.norm(x[0, :3] - goal) < 0.85 * d0  # 0.8 s of flight: 4.58 m -> 3.67 m, climbing and pitching over


# ---- sampled (visualisation) trajectories ---------------------------------------------------------------
def test_pick_sampled_indices_semantics():
    # controller.cu:55-179: slot 0 = optimised sequence, distinct random picks from the first 98 %, then top-n by weight
    rng = np.random.RandomState(0)
    costs = rng.rand(1000).astype(np.float32)
    idx = H.pick_sampled_indices(50, 5, costs, 0.05, np.random.RandomState(1))
    assert len(idx) == 55 and idx[0] == -1
    body = idx[1:50]
    assert len(set(body.tolist())) == 49 and body.min() >= 0 and body.max() < 980
    np.testing.assert_array_equal(idx[-5:], np.argsort(costs, kind="stable")[:5])
    # above 98 %: everything in order (controller.cu:66-70)
    idx = H.pick_sampled_indices(990, 0, costs, 0.99, np.random.RandomState(1))
    np.testing.assert_array_equal(idx, np.concatenate([[-1], np.arange(1, 990)]))
    assert H.pick_sampled_indices(0, 0, costs, 0.0, rng).size == 0
    np.testing.assert_array_equal(H.pick_sampled_indices(0, 2, costs, 0.0, rng), np.argsort(costs, kind="stable")[:2])


def test_oracle_sampled_trajectory_rows_sum_to_the_rollout_cost():
    w = W.cartpole(128, 40)
    w.sampler.setControlCostCoeff([0.7])
    w.alpha = 0.2
    eps = oracle.curand_normal(3, 0, w.N * w.T).reshape(w.N, w.T, 1)
    r = oracle.solve(w.dyn.DYN_ID, w.cost.COST_ID, w.dyn.params, w.cost.params, w.sampler.params, None, None, w.N, w.T, 1,
                     1, w.dt, w.lambda_, w.alpha, w.x0, w.U0, eps, want_samples=True)
    for n in (0, 17, 127):  # 127: pure-noise sample (likelihood-ratio term with mean 0)
        out, costs, crash = oracle.sampled_trajectory(w.dyn.DYN_ID, w.cost.COST_ID, w.dyn.params, w.cost.params,
                                                      w.sampler.params, None, None, w.N, w.T, 0, n, False, w.dt,
                                                      w.lambda_, w.alpha, w.x0[0], w.U0[0], r["samples"][0, n])
        assert costs.astype(np.float64).sum() == pytest.approx(float(r["costs"][0, n]), rel=1e-5)
        assert out.shape == (w.T, 4) and not crash.any()
        # outputs are the states after each step of the host twin driven by the same controls
        st, _ = oracle.output_trajectory(w.dyn.DYN_ID, w.dyn.params, None, w.x0[0],
                                         np.vstack([r["samples"][0, n], np.zeros((1, 1), np.float32)]), w.dt)
        np.testing.assert_allclose(out, st[1:], rtol=1e-6, atol=1e-6)


# ---- NLN sampler ------------------------------------------------------------------------------------------
def test_nln_noise_oracle_follows_the_reference_call_sequence():
    """nln.cu:114-128: C log-normal planes of N*T (mean 0, std dev sigma_c), then N*T*C normals, multiplied element-wise
    with plane index [c][n][t]. The normals of the k-th draw therefore sit at stream offset (2k - 1) * N*T*C, and the
    quotient noise / normal is log-normal(0, sigma_c): mean exp(sigma^2 / 2) (log_noise_mean_, nln.cu:100)."""
    N, T, Cd, sd = 1024, 64, 2, [0.5, 0.3]  # N*T = 8 * 8192: cuRAND honours these absolute offsets
    for k in (1, 2):
        a = oracle.nln_noise(42, k, N, T, Cd, sd)
        n = oracle.curand_normal(42, (2 * k - 1) * N * T * Cd, N * T * Cd).reshape(N, T, Cd)
        ratio = (a / n).astype(np.float64)
        assert (ratio > 0).all()
        np.testing.assert_allclose(np.log(ratio).std(axis=(0, 1)), sd, rtol=0.01)
        np.testing.assert_allclose(np.log(ratio).mean(axis=(0, 1)), 0.0, atol=0.005)
        np.testing.assert_allclose(ratio.mean(axis=(0, 1)), np.exp(0.5 * np.square(sd)), rtol=0.01)
        # plane c of the log-normal factors is its own cuRAND call: exp(sigma_c * z) of the normals at offset (2k-2)NTC + c NT
        z = oracle.curand_normal(42, (2 * k - 2) * N * T * Cd, N * T * Cd).reshape(Cd, N, T)
        for c in range(Cd):
            np.testing.assert_allclose(ratio[..., c], np.exp(np.float64(sd[c]) * z[c]), rtol=2e-5)
    s = m.NLNDistribution(2, sd)
    mean, std = s.log_noise_mean_and_std_dev()
    np.testing.assert_allclose(mean, np.exp(0.5 * np.square(sd)), rtol=1e-6)
    assert s.SAMPLER_ID == 2
