"""The RACER models' elevation map (texture helpers): TwoDTextureHelper<float> map 0 sampled by
RACER::computeStaticSettling (racer_dubins.cu:359-434) inside RacerDubinsElevationLSTMSteering::step
(racer_dubins_elevation_lstm_steering.cu:105-112).
CPU: the oracle's restatement of the texture query against the reference's OWN known answers
(tests/texture_helpers/two_d_texture_helper_test.cu:368-541, QueryTextureAtMapPose / QueryTextureAtWorldPose: 10 x 20 texture,
resolution 10, swapped axes, origin (1, 2, 3)); static settling against the closed form on a plane; host twin == oracle.
GPU: rollouts / solves over a hilly map against the oracle; the device tail against the host twin."""
import numpy as np
import pytest

import mppi_generic_b200 as m
import oracle
from mppi_generic_b200 import workloads as W

H = m.host

# (normalised x, normalised y) -> expected value, two_d_texture_helper_test.cu:386-447 / :479-541
REF_QUERIES = [((0.0, 0.0), 0.0), ((0.05, 0.0), 0.0), ((0.95, 0.0), 9.0), ((1.0, 0.0), 9.0), ((0.45, 0.0), 4.0),
               ((0.5, 0.0), 4.5), ((0.55, 0.0), 5.0), ((0.0, 0.0), 0.0), ((0.0, 0.025), 0.0), ((0.0, 0.05), 5.0),
               ((0.0, 0.075), 10.0), ((0.0, 0.975), 190.0), ((0.0, 1.0), 190.0), ((0.0, 0.475), 90.0), ((0.0, 0.5), 95.0),
               ((0.0, 0.525), 100.0)]


def _ref_helper(world: bool) -> "H.TwoDTextureHelper":
    t = H.TwoDTextureHelper()
    t.setExtent(0, 10, 20)
    t.updateTexture(0, np.arange(200, dtype=np.float32))  # the .x channel of the reference's float4 texels
    if world:
        t.updateRotation(0, [[0, 1, 0], [1, 0, 0], [0, 0, 1]])
        t.updateOrigin(0, (1, 2, 3))
    t.updateResolution(0, 10)
    t.enableTexture(0)
    return t


@pytest.mark.parametrize("world", [False, True])
def test_texture_query_reproduces_the_reference_known_answers(world):
    t = _ref_helper(world)
    for (nx, ny), expect in REF_QUERIES:
        # resolution * normalised handling exactly as the reference test builds its query points
        q = (ny * 10.0 * 20.0 + 1, nx * 10.0 * 10.0 + 2, 3.0) if world else (nx * 10 * 10, ny * 10 * 20, 0.0)
        assert oracle.elevation_at_world_pose(t.blob(), *q) == pytest.approx(expect, abs=1e-4), (nx, ny)
        assert t.queryTextureAtWorldPose(0, q) == pytest.approx(expect, abs=1e-4), (nx, ny)  # product host twin


def _plane(w=64, h=48, res=0.5, ax=0.08, ay=-0.05, origin=(-3.0, -2.0, 0.0)):
    """z = ax * x + ay * y sampled at the cell centres of a map whose cell (i, j) covers x in [j, j+1) * res + origin."""
    j, i = np.meshgrid(np.arange(w), np.arange(h))
    x = (j + 0.5) * res + origin[0]
    y = (i + 0.5) * res + origin[1]
    return (ax * x + ay * y).astype(np.float32), res, origin


def test_static_settling_on_a_plane_matches_the_closed_form():
    """On z = a x + b y bilinear interpolation is exact, so the four wheel heights are known: with roll = pitch = 0 and yaw
    psi the settled angles are asin of the height differences over track / wheel base (racer_dubins.cu:395-409)."""
    vals, res, origin = _plane()
    dyn = H.RacerDubinsElevationLSTMSteering()
    dyn.setElevationMap(vals, res, origin)
    a, b = 0.08, -0.05
    for yaw in (0.0, 0.7, -2.1):
        x, y = 5.0, 4.0
        c, s = np.cos(yaw), np.sin(yaw)
        wheel = lambda ox, oy: a * (x + c * ox - s * oy) + b * (y + s * ox + c * oy)  # noqa: E731
        fl, fr, rl, rr = wheel(2.981, 0.737), wheel(2.981, -0.737), wheel(0.0, 0.737), wheel(0.0, -0.737)
        roll = 0.5 * (np.arcsin((fl - fr) / 1.474) + np.arcsin((rl - rr) / 1.474))
        pitch = 0.5 * (np.arcsin((rl - fl) / 2.981) + np.arcsin((rr - fr) / 2.981))
        for got in (oracle.static_settling(dyn.getTextureHelper().blob(), yaw, x, y), dyn.staticSettling(yaw, x, y)):
            assert got[0] == pytest.approx(roll, abs=2e-5)
            assert got[1] == pytest.approx(pitch, abs=2e-5)
            assert got[2] == pytest.approx(0.5 * (rl + rr), abs=2e-5)
    # no map / disabled map: flat ground (racer_dubins.cu:427-432)
    dyn.getTextureHelper().disableTexture(0)
    assert dyn.staticSettling(0.3, 1.0, 2.0, 0.2, 0.1) == (0.0, 0.0, 0.0)
    assert oracle.static_settling(None, 0.3, 1.0, 2.0, 0.2, 0.1) == (0.0, 0.0, 0.0)


def _hills(w=96, h=80, res=0.5, origin=(-6.0, -20.0, 0.0), seed=5):
    rng = np.random.default_rng(seed)
    j, i = np.meshgrid(np.arange(w), np.arange(h))
    x, y = (j + 0.5) * res + origin[0], (i + 0.5) * res + origin[1]
    z = 0.6 * np.sin(0.21 * x) * np.cos(0.17 * y) + 0.25 * np.sin(0.05 * x * y * 0.1) + 0.02 * rng.standard_normal(x.shape)
    return z.astype(np.float32), res, origin


def test_host_twin_step_with_a_map_matches_the_oracle():
    w = W.racer_lstm_gaussian(64, 40)
    vals, res, origin = _hills()
    w.dyn.setElevationMap(vals, res, origin, rotation=[[0.96, 0.28, 0], [-0.28, 0.96, 0], [0, 0, 1]])
    oracle.set_lstm(w.dyn.lstm_theta, w.dyn.hidden_dim, w.dyn.head_hidden)
    oracle.set_elevation_map(w.dyn.getTextureHelper().blob())
    try:
        rng = np.random.default_rng(1)
        x = np.ascontiguousarray(w.x0[0], np.float32).copy()
        h, c = w.dyn.initial_hidden_cell()
        ho, co = h.copy(), c.copy()
        xo = x.copy()
        saw_slope = False
        for t in range(40):
            u = np.array([0.6 * rng.uniform(-1, 1), rng.uniform(-1, 1)], np.float32)
            xn, xd, y, h, c = w.dyn.step(x, u, w.dt, h, c)
            on, od, oy, ho, co = oracle.racer_step(w.dyn.params, xo, u, w.dt, ho, co)
            np.testing.assert_allclose(xn, on, rtol=2e-5, atol=2e-5)
            np.testing.assert_allclose(np.nan_to_num(y), np.nan_to_num(oy), rtol=2e-5, atol=2e-5)
            saw_slope = saw_slope or abs(xn[7]) > 1e-3  # PITCH
            x, xo = xn, on
        assert saw_slope, "the trajectory must actually drive over sloped ground"
    finally:
        oracle.set_elevation_map(None)


@pytest.mark.gpu
@pytest.mark.parametrize("hidden", [4, 32])
def test_racer_solve_over_an_elevation_map_matches_the_oracle(hidden):
    """K1 with the map (per-sample costs, the reference's 1e-4 bar; H = 32 runs the LSTM on mma.sync) and the solve's U."""
    w = W.racer_lstm_gaussian(2048, 60) if hidden == 4 else W.racer_lstm(2048, 60, hidden_dim=32, head_hidden=20, colored=False)
    vals, res, origin = _hills()
    w.dyn.setElevationMap(vals, res, origin, rotation=[[0.96, 0.28, 0], [-0.28, 0.96, 0], [0, 0, 1]])
    oracle.set_lstm(w.dyn.lstm_theta, w.dyn.hidden_dim, w.dyn.head_hidden)
    oracle.set_elevation_map(w.dyn.getTextureHelper().blob())
    try:
        e = w.make_engine(flags=H.FLAG_WRITEBACK_CONTROLS)
        U, stats = e.solve(w.x0, w.U0)
        costs = e.get_costs()
        ref = oracle.solve(w.dyn.DYN_ID, w.cost.COST_ID, w.dyn.params, w.cost.params, w.sampler.params, None, None, w.N, w.T,
                           w.D, w.dyn.CONTROL_DIM, w.dt, w.lambda_, w.alpha, w.x0, w.U0, e.get_noise(), nthreads=8)
        rel = np.abs(costs - ref["costs"]) / np.maximum(np.abs(ref["costs"]), 1.0)
        assert rel.max() < 2e-4, rel.max()  # C5's bar (150 LSTM steps with tanh_fast; tests/test_gpu_parity.py)
        assert stats[0][0] == pytest.approx(float(ref["baseline"][0]), rel=2e-4)
        np.testing.assert_allclose(U, ref["U"], atol=2e-3 * max(1.0, float(np.abs(ref["U"]).max())))
        # and the map matters: the same solve on flat ground has different costs
        w.dyn.getTextureHelper().disableTexture(0)
        e.push_params()
        e.seed(w.seed, 0)
        e.solve(w.x0, w.U0)
        assert np.abs(e.get_costs() - costs).max() > 1e-3 * np.abs(costs).max()
        e.close()
    finally:
        oracle.set_elevation_map(None)


@pytest.mark.gpu
def test_device_tail_over_an_elevation_map_matches_the_host_twin():
    w = W.racer_lstm_gaussian(1024, 80)
    vals, res, origin = _hills()
    w.dyn.setElevationMap(vals, res, origin)
    e = w.make_engine()
    x0 = np.ascontiguousarray(w.x0, np.float32)
    U, _ = e.solve(x0, w.U0)
    _, st_d, out_d = e.nominal_trajectory(x0, U, None)
    st_h = np.zeros((w.T, 19), np.float32)
    out_h = np.zeros((w.T, 28), np.float32)
    w.dyn.output_trajectory(x0[0], U[0], w.T, w.dt, st_h, out_h)
    assert np.abs(st_h[:, 7]).max() > 1e-3  # pitch moves
    scale = np.maximum(np.abs(st_h).max(axis=0, keepdims=True), 1.0)
    assert (np.abs(st_d[0] - st_h) / scale).max() < 1e-4
    np.testing.assert_allclose(np.nan_to_num(out_d[0][:, 4]), np.nan_to_num(out_h[:, 4]), atol=1e-4)  # BASELINK_POS_I_Z
    e.close()


def test_texture_query_against_an_independent_interpolator():
    """The oracle's (and the product host twin's) queryTextureAtWorldPose against scipy's order-1 `map_coordinates` with edge
    extension (= clamp addressing) on a random map with a rotated, shifted frame: an independent formulation in float64."""
    from scipy.ndimage import map_coordinates
    rng = np.random.default_rng(11)
    w, h, res = 37, 29, 0.4
    vals = rng.standard_normal((h, w)).astype(np.float32)
    c, s = np.cos(0.6), np.sin(0.6)
    rot = np.array([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]])
    origin = np.array([1.5, -2.0, 0.3])
    t = H.TwoDTextureHelper()
    t.setExtent(0, w, h)
    t.updateTexture(0, vals)
    t.updateRotation(0, rot)
    t.updateOrigin(0, origin)
    t.updateResolution(0, res)
    t.enableTexture(0)
    pts = rng.uniform(-6.0, 18.0, size=(400, 3))
    m = (rot @ (pts - origin).T).T                         # world -> map frame
    qx, qy = m[:, 0] / res - 0.5, m[:, 1] / res - 0.5       # cell-centre convention (two_d_texture_helper.cu:157-160)
    ref = map_coordinates(vals.astype(np.float64), [np.clip(qy, 0, h - 1), np.clip(qx, 0, w - 1)], order=1, mode="nearest")
    for p, r in zip(pts, ref):
        assert oracle.elevation_at_world_pose(t.blob(), *p) == pytest.approx(r, abs=2e-5)
        assert t.queryTextureAtWorldPose(0, p) == pytest.approx(r, abs=2e-5)
