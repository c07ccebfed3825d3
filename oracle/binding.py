"""ctypes binding of oracle/libmppi_oracle.so (the CPU restatement of the reference path). Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmppi_oracle.so")
_lib = None


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_baseline.restype = C.c_float
        _lib.orc_normalizer.restype = C.c_float
        _lib.orc_ar_query_texture.restype = C.c_float
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def curand_normal(seed: int, offset: int, n: int) -> np.ndarray:
    """Host XORWOW normals (same generator type/seed/offset semantics as controller.cu:192-207 + gaussian.cu:380)."""
    out = np.empty(n, np.float32)
    rc = lib().orc_curand_normal(C.c_ulonglong(seed), C.c_ulonglong(offset), C.c_size_t(n), _p(out))
    if rc:
        raise RuntimeError(f"orc_curand_normal failed: {rc}")
    return out


def set_gaussian_controls(means, sp, samples, C_, T, N, D, optimization_stride=1, iteration_num=0):
    lib().orc_set_gaussian_controls(_p(_f32(means)), C.byref(sp), _p(samples), C_, T, N, D, optimization_stride,
                                    iteration_num)


def rollout(dyn_id, cost_id, dyn_params, cost_params, sp, nn_theta, costmap, N, T, D, dt, lam, alpha, x0, means,
            samples, nthreads=1):
    """samples [D][N][T][C] (controls, constrained in place). Returns costs [D][N]."""
    costs = np.empty((D, N), np.float32)
    rc = lib().orc_rollout(dyn_id, cost_id, C.byref(dyn_params), C.byref(cost_params), C.byref(sp), _p(nn_theta),
                           _p(costmap), N, T, D, C.c_float(dt), C.c_float(lam), C.c_float(alpha), _p(_f32(x0)),
                           _p(_f32(means)), _p(samples), _p(costs), nthreads)
    if rc:
        raise RuntimeError(f"orc_rollout failed: {rc}")
    return costs


def solve(dyn_id, cost_id, dyn_params, cost_params, sp, nn_theta, costmap, N, T, D, C_, dt, lam, alpha, x0, U_in, eps,
          optimization_stride=1, iteration_num=0, sum_stride=32, nthreads=1, want_samples=False):
    """One optimisation iteration (see orc_solve). Returns dict(U, baseline, normalizer, free_energy, costs[, samples])."""
    U = np.empty((D, T, C_), np.float32)
    base = np.empty(D, np.float32)
    norm = np.empty(D, np.float32)
    fe = np.empty((D, 3), np.float32)
    costs = np.empty((D, N), np.float32)
    samples = np.empty((D, N, T, C_), np.float32) if want_samples else None
    rc = lib().orc_solve(dyn_id, cost_id, C.byref(dyn_params), C.byref(cost_params), C.byref(sp), _p(nn_theta),
                         _p(costmap), N, T, D, C.c_float(dt), C.c_float(lam), C.c_float(alpha), _p(_f32(x0)),
                         _p(_f32(U_in)), _p(_f32(eps)), optimization_stride, iteration_num, sum_stride, nthreads,
                         _p(U), _p(base), _p(norm), _p(fe), _p(costs), _p(samples))
    if rc:
        raise RuntimeError(f"orc_solve failed: {rc}")
    out = {"U": U, "baseline": base, "normalizer": norm, "free_energy": fe, "costs": costs}
    if want_samples:
        out["samples"] = samples
    return out


def baseline(costs) -> float:
    c = _f32(costs)
    return float(lib().orc_baseline(_p(c), c.size))


def norm_exp(costs, lambda_inv, base) -> np.ndarray:
    c = _f32(costs).copy()
    lib().orc_norm_exp(_p(c), c.size, C.c_float(lambda_inv), C.c_float(base))
    return c


def normalizer(w) -> float:
    w = _f32(w)
    return float(lib().orc_normalizer(_p(w), w.size))


def free_energy(w, base, lam) -> np.ndarray:
    w = _f32(w)
    out = np.empty(3, np.float32)
    lib().orc_free_energy(_p(w), w.size, C.c_float(base), C.c_float(lam), _p(out))
    return out


def weighted_reduction(w, du, normalizer_, T, N, C_, sum_stride=32) -> np.ndarray:
    out = np.empty((T, C_), np.float32)
    lib().orc_weighted_reduction(_p(_f32(w)), _p(_f32(du)), _p(out), C.c_float(normalizer_), T, N, C_, sum_stride)
    return out


def smooth(u, history) -> np.ndarray:
    u = _f32(u).copy()
    T, C_ = u.shape
    lib().orc_smooth(_p(u), _p(_f32(history)), T, C_)
    return u


def slide(u, steps, zero_control, scale) -> np.ndarray:
    u = _f32(u).copy()
    T, C_ = u.shape
    lib().orc_slide(_p(u), steps, T, C_, _p(_f32(zero_control)), _p(_f32(scale)))
    return u


def enforce_constraints(lim, u) -> np.ndarray:
    u = _f32(u).copy()
    lib().orc_enforce_constraints(C.byref(lim), _p(u), u.size)
    return u


def dyn_step(dyn_id, dyn_params, nn_theta, x, u, dt):
    S, C_, O = C.c_int(), C.c_int(), C.c_int()
    lib().orc_dims(dyn_id, C.byref(S), C.byref(C_), C.byref(O))
    xn, xd, y = np.zeros(S.value, np.float32), np.zeros(S.value, np.float32), np.zeros(O.value, np.float32)
    rc = lib().orc_dyn_step(dyn_id, C.byref(dyn_params), _p(nn_theta), _p(_f32(x)), _p(_f32(u)), C.c_float(dt), _p(xn),
                            _p(xd), _p(y))
    if rc:
        raise RuntimeError("orc_dyn_step failed")
    return xn, xd, y


def state_cost(cost_id, cost_params, costmap, y, t=0, crash=0):
    cr = C.c_int(crash)
    c, term = C.c_float(), C.c_float()
    rc = lib().orc_state_cost(cost_id, C.byref(cost_params), _p(costmap), _p(_f32(y)), t, C.byref(cr), C.byref(c),
                              C.byref(term))
    if rc:
        raise RuntimeError("orc_state_cost failed")
    return c.value, term.value, cr.value


def ar_cost_terms(cost_params, costmap, s, crash=0):
    cr = C.c_int(crash)
    sp, st, tr, cc = C.c_float(), C.c_float(), C.c_float(), C.c_float()
    lib().orc_ar_cost_terms(C.byref(cost_params), _p(costmap), _p(_f32(s)), C.byref(cr), C.byref(sp), C.byref(st),
                            C.byref(tr), C.byref(cc))
    return {"speed": sp.value, "stabilizing": st.value, "track": tr.value, "crash": cc.value, "crash_status": cr.value}


def ar_query_texture(cost_params, costmap, x, y) -> float:
    return float(lib().orc_ar_query_texture(C.byref(cost_params), _p(costmap), C.c_float(x), C.c_float(y)))


def fnn_forward(theta, layers, inp) -> np.ndarray:
    layers = np.ascontiguousarray(layers, dtype=np.int32)
    out = np.zeros(int(layers[-1]), np.float32)
    lib().orc_fnn_forward(_p(_f32(theta)), layers.ctypes.data_as(C.c_void_p), len(layers), _p(_f32(inp)), _p(out))
    return out


def output_trajectory(dyn_id, dyn_params, nn_theta, x0, u, dt):
    S, C_, O = C.c_int(), C.c_int(), C.c_int()
    lib().orc_dims(dyn_id, C.byref(S), C.byref(C_), C.byref(O))
    u = _f32(u)
    T = u.shape[0]
    states, outputs = np.zeros((T, S.value), np.float32), np.zeros((T, O.value), np.float32)
    rc = lib().orc_output_trajectory(dyn_id, C.byref(dyn_params), _p(nn_theta), _p(_f32(x0)), _p(u), T, C.c_float(dt),
                                     _p(states), _p(outputs))
    if rc:
        raise RuntimeError("orc_output_trajectory failed")
    return states, outputs


_lstm_keepalive = None


def set_lstm(theta, hidden_dim: int, head_hidden: int) -> None:
    """LSTM weights/architecture for MPPIB_DYN_RACER_LSTM (kept by pointer inside the oracle)."""
    global _lstm_keepalive
    _lstm_keepalive = _f32(theta).copy()
    lib().orc_set_lstm(_p(_lstm_keepalive), hidden_dim, head_hidden)


_elev_keepalive = None


def set_elevation_map(blob) -> None:
    """Elevation map of the RACER model: mppib_elevation_map_header + [height][width] floats as one byte array (what
    host.TwoDTextureHelper.blob() returns), or None = flat ground. Kept by pointer inside the oracle."""
    global _elev_keepalive
    if blob is None:
        _elev_keepalive = None
        lib().orc_set_elevation_map(None)
        return
    _elev_keepalive = np.ascontiguousarray(blob, dtype=np.uint8).copy()
    lib().orc_set_elevation_map(C.c_void_p(_elev_keepalive.ctypes.data))


def elevation_at_world_pose(blob, x: float, y: float, z: float = 0.0) -> float:
    """TwoDTextureHelper<float>::queryTextureAtWorldPose on the host (texture_helper.cu:94-134, two_d_texture_helper.cu:151-243)."""
    b = np.ascontiguousarray(blob, dtype=np.uint8)
    L = lib()
    L.orc_elevation_at_world_pose.restype = C.c_float
    L.orc_elevation_at_world_pose.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
    return float(L.orc_elevation_at_world_pose(b.ctypes.data, x, y, z))


def static_settling(blob, yaw: float, x: float, y: float, roll: float = 0.0, pitch: float = 0.0):
    """RACER::computeStaticSettling (racer_dubins.cu:359-434): returns (roll, pitch, height)."""
    L = lib()
    L.orc_static_settling.restype = C.c_float
    L.orc_static_settling.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    b = None if blob is None else np.ascontiguousarray(blob, dtype=np.uint8)
    r, p = C.c_float(roll), C.c_float(pitch)
    h = L.orc_static_settling(None if b is None else b.ctypes.data, yaw, x, y, C.byref(r), C.byref(p))
    return r.value, p.value, float(h)


def lstm_forward(lstm_w, input_dim, hidden_dim, head_theta, head_layers, inp, h, c):
    """LSTMHelper::forward(input, output), host path. Returns (output, h_next, c_next)."""
    head_layers = np.ascontiguousarray(head_layers, dtype=np.int32)
    h, c = _f32(h).copy(), _f32(c).copy()
    out = np.zeros(int(head_layers[-1]), np.float32)
    lib().orc_lstm_forward(_p(_f32(lstm_w)), input_dim, hidden_dim, _p(_f32(head_theta)),
                           head_layers.ctypes.data_as(C.c_void_p), len(head_layers), _p(_f32(inp)), _p(h), _p(c), _p(out))
    return out, h, c


def lstm_initialize(init_w, input_dim, hidden_dim, head_theta, head_layers, init_len, buffer):
    """LSTMLSTMHelper::initializeLSTM (lstm_lstm_helper.cu:50-73). buffer [cols][input_dim]. Returns the head output
    (first half = hidden, second half = cell of the prediction LSTM)."""
    head_layers = np.ascontiguousarray(head_layers, dtype=np.int32)
    buffer = _f32(buffer)
    out = np.zeros(int(head_layers[-1]), np.float32)
    lib().orc_lstm_initialize(_p(_f32(init_w)), input_dim, hidden_dim, _p(_f32(head_theta)), _p(head_layers), len(head_layers),
                              init_len, _p(buffer), buffer.shape[0], _p(out))
    return out


def racer_step(dyn_params, x, u, dt, h, c):
    """One host step of RacerDubinsElevationLSTMSteering. Returns (x_next, xdot, y, h_next, c_next)."""
    h, c = _f32(h).copy(), _f32(c).copy()
    xn, xd, y = np.zeros(19, np.float32), np.zeros(19, np.float32), np.zeros(28, np.float32)
    rc = lib().orc_racer_step(C.byref(dyn_params), _p(_f32(x)), _p(_f32(u)), C.c_float(dt), _p(h), _p(c), _p(xn), _p(xd),
                              _p(y))
    if rc:
        raise RuntimeError("orc_racer_step failed (set_lstm not called?)")
    return xn, xd, y, h, c


def colored_noise(normals, sp, N, C_, T, offset_t=1, nthreads=1) -> np.ndarray:
    """ColoredNoise block [N][T][C] (before setGaussianControls) from the 2*N*C*(T+1) raw normals of one draw."""
    normals = _f32(normals)
    assert normals.size == 2 * N * C_ * (T + 1)
    eps = np.empty((N, T, C_), np.float32)
    lib().orc_colored_noise(_p(normals), C.byref(sp), N, C_, T, offset_t, _p(eps), nthreads)
    return eps


def colored_tables(sp, C_, T):
    coeffs = np.empty((C_, T + 1), np.float32)
    sigma = np.empty(C_, np.float32)
    lib().orc_colored_tables(C.byref(sp), C_, T, _p(coeffs), _p(sigma))
    return coeffs, sigma


def rmppi_rollout(dyn_id, cost_id, dyn_params, cost_params, sp, nn_theta, costmap, N, T, dt, lam, alpha,
                  value_func_threshold, x0, means, gains, samples, nominal_idx=0, nthreads=1):
    """launchCPURMPPIRolloutKernel: samples [2][N][T][C] sampled controls (constrained in place, feedback added to the
    real system's). gains [T][S][C] or None. Returns costs [2][N]."""
    costs = np.empty((2, N), np.float32)
    g = None if gains is None else _f32(gains)
    rc = lib().orc_rmppi_rollout(dyn_id, cost_id, C.byref(dyn_params), C.byref(cost_params), C.byref(sp), _p(nn_theta),
                                 _p(costmap), N, T, C.c_float(dt), C.c_float(lam), C.c_float(alpha),
                                 C.c_float(value_func_threshold), nominal_idx, _p(_f32(x0)), _p(_f32(means)), _p(g),
                                 _p(samples), _p(costs), nthreads)
    if rc:
        raise RuntimeError(f"orc_rmppi_rollout failed: {rc}")
    return costs


def init_eval(dyn_id, cost_id, dyn_params, cost_params, sp, nn_theta, costmap, N_sampler, T, dt, lam, alpha, candidates,
              strides, samples_per_candidate, means, controls):
    """launchCPUInitEvalKernel: controls [num_samples][T][C] = the sampler's buffer (distribution 0). costs [K * spc]."""
    cand = _f32(candidates)
    st = np.ascontiguousarray(strides, dtype=np.int32)
    K = cand.shape[0]
    costs = np.zeros(K * samples_per_candidate, np.float32)
    rc = lib().orc_init_eval(dyn_id, cost_id, C.byref(dyn_params), C.byref(cost_params), C.byref(sp), _p(nn_theta),
                             _p(costmap), N_sampler, T, C.c_float(dt), C.c_float(lam), C.c_float(alpha), K,
                             samples_per_candidate, _p(cand), st.ctypes.data_as(C.c_void_p), _p(_f32(means)),
                             _p(_f32(controls)), _p(costs))
    if rc:
        raise RuntimeError(f"orc_init_eval failed: {rc}")
    return costs


def rmppi_line_search_weights(K) -> np.ndarray:
    out = np.zeros((3, K), np.float32)
    lib().orc_rmppi_line_search_weights(K, _p(out))
    return out


def rmppi_strides(K, stride) -> np.ndarray:
    out = np.zeros(K, np.int32)
    lib().orc_rmppi_strides(K, stride, out.ctypes.data_as(C.c_void_p))
    return out


def rmppi_best_index(costs, K, spc, lam, threshold):
    fe = np.zeros(K, np.float32)
    best = lib().orc_rmppi_best_index(_p(_f32(costs)), K, spc, C.c_float(lam), C.c_float(threshold), _p(fe))
    return best, fe


def tsallis(costs, gamma, r, base) -> np.ndarray:
    """TsallisTransform (core/mppi_common.cu:968-985): weights (1 - (c - beta)/gamma)^(1/(r-1)) below gamma, else 0."""
    c = _f32(costs).copy()
    lib().orc_tsallis(_p(c), c.size, C.c_float(gamma), C.c_float(r), C.c_float(base))
    return c


def sampled_trajectory(dyn_id, cost_id, dyn_params, cost_params, sp, nn_theta, costmap, N, T, d, sample_index,
                       apply_constraints, dt, lam, alpha, x0, means, controls):
    """One rollout with every step dumped (the visualisation pass): returns outputs [T][O], costs [T + 1], crash [T]."""
    S, C_, O = C.c_int(), C.c_int(), C.c_int()
    lib().orc_dims(dyn_id, C.byref(S), C.byref(C_), C.byref(O))
    outputs = np.zeros((T, O.value), np.float32)
    costs = np.zeros(T + 1, np.float32)
    crash = np.zeros(T, np.int32)
    rc = lib().orc_sampled_trajectory(dyn_id, cost_id, C.byref(dyn_params), C.byref(cost_params), C.byref(sp),
                                      _p(nn_theta), _p(costmap), N, T, d, sample_index, int(apply_constraints),
                                      C.c_float(dt), C.c_float(lam), C.c_float(alpha), _p(_f32(x0)), _p(_f32(means)),
                                      _p(_f32(controls)), _p(outputs), _p(costs), crash.ctypes.data_as(C.c_void_p))
    if rc:
        raise RuntimeError(f"orc_sampled_trajectory failed: {rc}")
    return outputs, costs, crash


def nln_noise(seed: int, draws: int, N: int, T: int, C_: int, std_dev) -> np.ndarray:
    """Raw noise of the `draws`-th NLNDistribution::generateSamples call since seeding (host cuRAND, the reference's call
    sequence): normal * log-normal, [N][T][C]."""
    out = np.empty((N, T, C_), np.float32)
    rc = lib().orc_nln_noise(C.c_ulonglong(seed), draws, N, T, C_, _p(_f32(std_dev)), _p(out))
    if rc:
        raise RuntimeError(f"orc_nln_noise failed: {rc}")
    return out
