"""CPU pin of the arithmetic scheme of csrc/plugins/nn_mma.cuh (the Autorally network on mma.sync): FP16 hi / lo operands
(unscaled residual), three products in one FP32 accumulator, activations handed on as r = 1 / (exp2(z) + 1) with the affine
map of tanh folded into the consuming layer's weights — emulated with numpy float16 / float32, on the synthetic 6-32-32-4
network, against float64. Two things are pinned: (1) the fragment bookkeeping (which lane register holds which matrix element, how a C
fragment becomes the next layer's A fragment, the weight-fragment order of load_weights) reproduces the plain matrix
products exactly; (2) the split reaches FP32-class accuracy where a single FP16 product does not (the reason for it).
The device itself is checked against the oracle and the FFMA2 form in tests/test_gpu_parity.py; this file needs no GPU."""
import numpy as np

from mppi_generic_b200 import workloads as W

SC = 2.8853900817779268  # kTanhScale: tanh(x) = 1 - 2 / (exp2(SC x) + 1)
LANES = np.arange(32)
G, T = LANES >> 2, LANES & 3


def _frag_a16(a):  # a[4][32][2] -> A[16][16]   (PTX m16n8k16 .f16 A layout)
    A = np.zeros((16, 16))
    for l in range(32):
        g, t = G[l], T[l]
        A[g, 2 * t:2 * t + 2], A[g + 8, 2 * t:2 * t + 2] = a[0][l], a[1][l]
        A[g, 2 * t + 8:2 * t + 10], A[g + 8, 2 * t + 8:2 * t + 10] = a[2][l], a[3][l]
    return A


def _frag_b16(b0, b1):  # [32][2] each -> B[16][8]
    B = np.zeros((16, 8))
    for l in range(32):
        g, t = G[l], T[l]
        B[2 * t:2 * t + 2, g], B[2 * t + 8:2 * t + 10, g] = b0[l], b1[l]
    return B


def _c_frag(Cm):  # C[16][8] -> c[32][4]
    return np.array([[Cm[G[l], 2 * T[l]], Cm[G[l], 2 * T[l] + 1], Cm[G[l] + 8, 2 * T[l]], Cm[G[l] + 8, 2 * T[l] + 1]]
                     for l in range(32)])


def _c_mat(c):
    Cm = np.zeros((16, 8))
    for l in range(32):
        g, t = G[l], T[l]
        Cm[g, 2 * t], Cm[g, 2 * t + 1], Cm[g + 8, 2 * t], Cm[g + 8, 2 * t + 1] = c[l]
    return Cm


def test_fragment_bookkeeping_reproduces_the_matrix_products():
    """forward() / load_weights() of nn_mma.cuh restated lane by lane (exact arithmetic): layer 2 and 3 of the network."""
    theta = W.synthetic_nn_weights(1).astype(np.float64)
    W2, b2 = theta[224:1248].reshape(32, 32), theta[1248:1280]
    W3, b3 = theta[1280:1408].reshape(4, 32), theta[1408:1412]
    rng = np.random.RandomState(0)
    h = np.tanh(rng.randn(16, 32))  # one m-tile of layer-1 activations, [row][neuron]
    # layer-1 C fragments of n-tile i, re-used as layer 2's A fragments: k16-tile j <- n-tiles 2j, 2j+1
    cf = [_c_frag(h[:, 8 * i:8 * i + 8]) for i in range(4)]
    a2 = [[cf[2 * j][:, 0:2], cf[2 * j][:, 2:4], cf[2 * j + 1][:, 0:2], cf[2 * j + 1][:, 2:4]] for j in range(2)]
    out2 = np.zeros((16, 32))
    for i in range(4):  # n-tile
        c = np.array([[b2[8 * i + 2 * T[l]], b2[8 * i + 2 * T[l] + 1]] * 2 for l in range(32)])
        for j in range(2):
            b0 = np.array([[W2[8 * i + G[l], 16 * j + 2 * T[l]], W2[8 * i + G[l], 16 * j + 2 * T[l] + 1]] for l in range(32)])
            b1 = np.array([[W2[8 * i + G[l], 16 * j + 8 + 2 * T[l]], W2[8 * i + G[l], 16 * j + 9 + 2 * T[l]]] for l in range(32)])
            c = _c_frag(_c_mat(c) + _frag_a16(a2[j]) @ _frag_b16(b0, b1))
        out2[:, 8 * i:8 * i + 8] = _c_mat(c)
    np.testing.assert_allclose(out2, h @ W2.T + b2, rtol=0, atol=1e-13)
    q = np.tanh(out2)
    cf = [_c_frag(q[:, 8 * i:8 * i + 8]) for i in range(4)]
    a3 = [[cf[2 * j][:, 0:2], cf[2 * j][:, 2:4], cf[2 * j + 1][:, 0:2], cf[2 * j + 1][:, 2:4]] for j in range(2)]
    b3p = np.concatenate([b3, np.zeros(4)])
    c = np.array([[b3p[2 * T[l]], b3p[2 * T[l] + 1]] * 2 for l in range(32)])
    for j in range(2):
        w = lambda g, k: W3[g, k] if g < 4 else 0.0  # noqa: E731  (outputs 4..7 of the tile are padding)
        b0 = np.array([[w(G[l], 16 * j + 2 * T[l]), w(G[l], 16 * j + 2 * T[l] + 1)] for l in range(32)])
        b1 = np.array([[w(G[l], 16 * j + 8 + 2 * T[l]), w(G[l], 16 * j + 9 + 2 * T[l])] for l in range(32)])
        c = _c_frag(_c_mat(c) + _frag_a16(a3[j]) @ _frag_b16(b0, b1))
    np.testing.assert_allclose(_c_mat(c)[:, :4], q @ W3.T + b3, rtol=0, atol=1e-13)


def _split(v):
    """split2 of nn_mma.cuh: hi = half(v), lo = half(v - hi) — the residual is NOT rescaled; below 2^-14 it is an FP16
    subnormal (numpy's float16 keeps subnormals like the tensor cores do)."""
    hi = v.astype(np.float16)
    lo = (v - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def _layer_split(a, Wm, b):
    """bias + hi*hi + lo*hi + hi*lo, all in ONE FP32 accumulator (forward_frag: the accumulator starts at the bias)."""
    ah, al = _split(a.astype(np.float32))
    wh, wl = _split(Wm.astype(np.float32))
    c = b.astype(np.float32) + (ah @ wh.T).astype(np.float32)
    c = (c + (al @ wh.T).astype(np.float32)).astype(np.float32)
    return (c + (ah @ wl.T).astype(np.float32)).astype(np.float32)


def _sigmoid_pre(z):
    """sigmoid2_prescaled: r = 1 / (exp2(z) + 1); tanh(x) = 1 - 2 r with z = 2 log2(e) x."""
    return np.float32(1.0) / (np.exp2(z).astype(np.float32) + np.float32(1.0))


def _folded(theta):
    """load_weights: layers feeding a tanh carry kTanhScale; layers consuming r = (1 - tanh) / 2 carry W' = -2 W and
    b' = b + row sum of W (taken in double, rounded once)."""
    W1, b1 = theta[:192].reshape(32, 6), theta[192:224]
    W2, b2 = theta[224:1248].reshape(32, 32), theta[1248:1280]
    W3, b3 = theta[1280:1408].reshape(4, 32), theta[1408:1412]
    f32 = np.float32
    return ((W1 * SC).astype(f32), (b1 * SC).astype(f32),
            (W2.astype(f32) * f32(-2.0 * SC)).astype(f32), ((b2 + W2.sum(1)).astype(f32) * f32(SC)).astype(f32),
            (W3.astype(f32) * f32(-2.0)).astype(f32), (b3 + W3.sum(1)).astype(f32))


def test_three_product_fp16_split_reaches_fp32_class_accuracy():
    theta = W.synthetic_nn_weights(1).astype(np.float64)
    W1, b1 = theta[:192].reshape(32, 6), theta[192:224]
    W2, b2 = theta[224:1248].reshape(32, 32), theta[1248:1280]
    W3, b3 = theta[1280:1408].reshape(4, 32), theta[1408:1412]
    rng = np.random.RandomState(1)
    x = (rng.randn(16384, 6) * [0.3, 3.0, 1.0, 1.0, 0.5, 0.5]).astype(np.float32)  # roll, vx, vy, yaw rate, steering, throttle
    ref = np.tanh(np.tanh(x.astype(np.float64) @ W1.T + b1) @ W2.T + b2) @ W3.T + b3
    w1, c1, w2, c2, w3, c3 = _folded(theta)
    r1 = _sigmoid_pre(_layer_split(x, w1, c1))
    r2 = _sigmoid_pre(_layer_split(r1, w2, c2))
    out = _layer_split(r2, w3, c3)
    err_split = np.abs(out - ref).max()
    # the same network as plain FP32 FMA chains (what the reference's FNNHelper::forward computes)
    f32 = np.float32

    def chain(a, Wm, b):
        acc = np.zeros((a.shape[0], Wm.shape[0]), f32)
        for k in range(Wm.shape[1]):
            acc = (acc + a[:, k:k + 1].astype(f32) * Wm[:, k].astype(f32)[None, :]).astype(f32)
        return acc + b.astype(f32)
    err_fp32 = np.abs(chain(np.tanh(chain(np.tanh(chain(x, W1, b1)), W2, b2)), W3, b3) - ref).max()
    # and with ONE FP16 product per term (what a plain FP16 / TF32 MMA would do)
    f16 = lambda v: v.astype(np.float16).astype(np.float32)  # noqa: E731
    h1 = np.tanh((f16(x) @ f16(W1.astype(np.float32)).T) + b1.astype(np.float32))
    q1 = np.tanh((f16(h1) @ f16(W2.astype(np.float32)).T) + b2.astype(np.float32))
    out1 = (f16(q1) @ f16(W3.astype(np.float32)).T) + b3.astype(np.float32)
    err_single = np.abs(out1 - ref).max()
    assert err_split < 1.0e-6, err_split          # FP32 class: ~4e-7 here, plain FP32 chains ~2e-7
    assert err_split < 4 * err_fp32 + 2e-7, (err_split, err_fp32)
    assert err_single > 1e-4, err_single          # fails the 1e-4 parity bar of a 100-step recurrence on its own
    assert err_single > 100 * err_split


def test_layer1_k16_concatenation_and_sigmoid_fold_are_exact_identities():
    """Two bookkeeping identities of forward_frag, in exact arithmetic: (1) layer 1 as [a_hi | a_lo] x [w_hi ; w_hi] (one k16
    MMA) + a_hi x w_lo (one k8 MMA) is a_hi w_hi + a_lo w_hi + a_hi w_lo; (2) feeding r = (1 - tanh) / 2 into W' = -2 W,
    b' = b + sum W reproduces W tanh + b."""
    rng = np.random.RandomState(3)
    a_hi, a_lo = rng.randn(16, 8), rng.randn(16, 8) * 1e-4
    w_hi, w_lo = rng.randn(8, 8), rng.randn(8, 8) * 1e-4   # [n][k]
    A16 = np.concatenate([a_hi, a_lo], axis=1)               # [16][16]
    B16 = np.concatenate([w_hi, w_hi], axis=1)               # [n][16]: k 0..7 and k 8..15 both multiply w_hi
    got = A16 @ B16.T + a_hi @ w_lo.T
    np.testing.assert_allclose(got, a_hi @ w_hi.T + a_lo @ w_hi.T + a_hi @ w_lo.T, rtol=0, atol=1e-12)
    Wm, b, t = rng.randn(32, 32), rng.randn(32), np.tanh(rng.randn(16, 32))
    r = (1.0 - t) / 2.0
    np.testing.assert_allclose(r @ (-2.0 * Wm).T + (b + Wm.sum(1)), t @ Wm.T + b, rtol=0, atol=1e-12)


# ---- plugins/lstm_mma.cuh: the steering LSTM (hidden_dim 32) on the same scheme -----------------------------------------------
def test_lstm_fragment_bookkeeping_keeps_the_recurrence_lane_local():
    """Index identities lstm_mma.cuh rests on. Gate tile (q, k) = gate q, hidden units 8k .. 8k+7: lane (g, t) of its C
    fragment holds units 8k+2t, 8k+2t+1 of rows g, g+8 — the same units for all four gates, so c' and h' are lane-local;
    and those two units, packed, are register (k & 1) * 2 + {0: row g, 1: row g+8} of k16-tile k >> 1 of the next step's
    A fragment (m16n8k16 A layout: a0 (g, 2t..), a1 (g+8, 2t..), a2 (g, 2t+8..), a3 (g+8, 2t+8..))."""
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for k in range(4):
            units_c = {8 * k + 2 * t, 8 * k + 2 * t + 1}                     # C fragment columns 2t, 2t+1 of tile (q, k), any q
            j, reg = k >> 1, (k & 1) * 2                                      # where forward() stores h'
            col0 = 2 * t + (8 if reg >= 2 else 0)                             # a0/a1 -> cols 2t.., a2/a3 -> cols 2t+8..
            units_a = {16 * j + col0, 16 * j + col0 + 1}
            assert units_c == units_a, (lane, k)


def test_lstm_step_with_folded_scales_matches_float64():
    """One LSTM step + head for a batch, emulated the way forward() computes it — gate weights pre-scaled by -log2(e)
    (sigmoid gates) / 2 log2(e) (cell candidate), r = 1 / (1 + exp2(z)), tanh = 1 - 2 r, head tanh handed on as r with
    W2' = -2 W2, b2' = b2 + sum W2, every product split hi / lo in FP16 — against the float64 LSTM of lstm_helper.cu:341-463."""
    Hd, I, L1 = 32, 4, 20
    lstm, head = W.synthetic_lstm_weights(Hd, L1, 2)
    lstm, head = lstm.astype(np.float64), head.astype(np.float64)
    HH, IH = Hd * Hd, Hd * I
    Wm = [lstm[q * HH:(q + 1) * HH].reshape(Hd, Hd) for q in range(4)]           # W_im W_fm W_om W_cm
    Wi = [lstm[4 * HH + q * IH:4 * HH + (q + 1) * IH].reshape(Hd, I) for q in range(4)]
    b = [lstm[4 * HH + 4 * IH + q * Hd:4 * HH + 4 * IH + (q + 1) * Hd] for q in range(4)]
    IN = Hd + I
    W1, b1 = head[:L1 * IN].reshape(L1, IN), head[L1 * IN:L1 * IN + L1]
    W2, b2 = head[L1 * IN + L1:L1 * IN + 2 * L1], head[L1 * IN + 2 * L1]
    rng = np.random.RandomState(7)
    n = 4096
    h, c = rng.uniform(-0.9, 0.9, (n, Hd)), rng.uniform(-2, 2, (n, Hd))
    x = rng.uniform(-0.3, 0.3, (n, I))
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))  # noqa: E731
    gi, gf, go = (sig(h @ Wm[q].T + x @ Wi[q].T + b[q]) for q in range(3))
    gc = np.tanh(h @ Wm[3].T + x @ Wi[3].T + b[3])
    c_ref = gi * gc + gf * c
    h_ref = np.tanh(c_ref) * go
    out_ref = np.tanh(np.concatenate([h_ref, x], 1) @ W1.T + b1) @ W2 + b2
    # emulation
    f32, LOG2E = np.float32, 1.4426950408889634
    a = np.concatenate([h, x], 1).astype(f32)
    acts = []
    for q in range(4):
        sc = 2.0 * LOG2E if q == 3 else -LOG2E
        Wq = (np.concatenate([Wm[q], Wi[q]], 1) * sc).astype(f32)
        acts.append(_sigmoid_pre(_layer_split(a, Wq, (b[q] * sc).astype(f32))))
    g_ = f32(1.0) - f32(2.0) * acts[3]
    c_new = (acts[0] * g_ + acts[1] * c.astype(f32)).astype(f32)
    rt = _sigmoid_pre((c_new * f32(2.0 * LOG2E)).astype(f32))
    h_new = (acts[2] * (f32(1.0) - f32(2.0) * rt)).astype(f32)
    a2 = np.concatenate([h_new, x.astype(f32)], 1)
    r1 = _sigmoid_pre(_layer_split(a2, (W1 * 2.0 * LOG2E).astype(f32), (b1 * 2.0 * LOG2E).astype(f32)))
    out = _layer_split(r1, (-2.0 * W2).astype(f32)[None, :], np.array([b2 + W2.sum()], f32))[:, 0]
    assert np.abs(c_new - c_ref).max() < 2e-6 and np.abs(h_new - h_ref).max() < 1e-6, (np.abs(c_new - c_ref).max(),
                                                                                       np.abs(h_new - h_ref).max())
    assert np.abs(out - out_ref).max() < 2e-6, np.abs(out - out_ref).max()
