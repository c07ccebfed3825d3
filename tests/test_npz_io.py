"""npz input files: the library's own reader (csrc/npz_reader.cpp) against numpy on the two file kinds the
reference loads through cnpy — network weights (FNNHelper::loadParams, fnn_helper.cu:44-127) and track maps
(ARStandardCostImpl::loadTrackData, ar_standard_cost.cu:85-142) — stored and deflated, plus the error paths."""
import os
import subprocess

import numpy as np
import pytest

import mppi_generic_b200 as m
from mppi_generic_b200 import workloads as W

H = m.host
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_model(path, compressed=False, seed=1):
    theta = W.synthetic_nn_weights(seed).astype(np.float64)
    layers = (6, 32, 32, 4)
    arrays, at = {}, 0
    for i in range(3):
        n_in, n_out = layers[i], layers[i + 1]
        arrays[f"dynamics_W{i + 1}"] = theta[at:at + n_in * n_out].reshape(n_out, n_in)
        at += n_in * n_out
        arrays[f"dynamics_b{i + 1}"] = theta[at:at + n_out]
        at += n_out
    (np.savez_compressed if compressed else np.savez)(path, **arrays)
    return theta.astype(np.float32)


@pytest.mark.parametrize("compressed", [False, True])
def test_npz_reader_matches_numpy(tmp_path, compressed):
    path = str(tmp_path / "arrays.npz")
    rng = np.random.RandomState(0)
    arrays = {"f8": rng.randn(7, 5), "f4": rng.randn(3, 4, 2).astype(np.float32), "i4": np.arange(12, dtype=np.int32),
              "i8": np.arange(6, dtype=np.int64).reshape(2, 3), "scalar": np.array([2.5], np.float32),
              "big": rng.randn(300, 257).astype(np.float32)}
    (np.savez_compressed if compressed else np.savez)(path, **arrays)
    for k, v in arrays.items():
        got = H.npz_read(path, k)
        assert got.shape == v.shape and got.dtype == np.float32
        np.testing.assert_array_equal(got, v.astype(np.float32))
    with pytest.raises(m.MppibError, match="no array named"):
        H.npz_read(path, "missing")
    with pytest.raises(m.MppibError, match="cannot open"):
        H.npz_read(str(tmp_path / "nope.npz"), "f8")


def test_npz_reader_rejects_what_it_cannot_represent(tmp_path):
    path = str(tmp_path / "odd.npz")
    np.savez(path, fortran=np.asfortranarray(np.ones((3, 2))), text=np.array(["a", "b"]), c=np.ones(3, np.complex64))
    for k in ("fortran", "text", "c"):
        with pytest.raises(m.MppibError):
            H.npz_read(path, k)
    bad = tmp_path / "bad.npz"
    bad.write_bytes(b"this is not a zip archive" * 10)
    with pytest.raises(m.MppibError, match="not a zip"):
        H.npz_read(str(bad), "x")


@pytest.mark.parametrize("compressed", [False, True])
def test_neural_net_model_load_params(tmp_path, compressed):
    path = str(tmp_path / "model.npz")
    theta = _write_model(path, compressed)
    dyn = m.NeuralNetModel([(-1.0, 1.0), (-2.0, 2.0)])
    dyn.loadParams(path)
    np.testing.assert_array_equal(dyn.nn_theta, theta)  # float64 file -> float32 exactly like the reference's assignment
    # a network of another shape is refused like FNNHelper::updateModel does
    np.savez(path, dynamics_W1=np.zeros((8, 6)), dynamics_b1=np.zeros(8), dynamics_W2=np.zeros((4, 8)), dynamics_b2=np.zeros(4))
    with pytest.raises(ValueError):
        dyn.loadParams(path)


def test_ar_standard_cost_load_track_data_from_file(tmp_path):
    ch0, xb, yb, ppm = W.track_map_standard()
    h, w = ch0.shape
    rng = np.random.RandomState(3)
    chans = [ch0.astype(np.float32)] + [rng.rand(h, w).astype(np.float32) for _ in range(3)]
    path = str(tmp_path / "track.npz")
    np.savez(path, xBounds=np.array(xb, np.float32), yBounds=np.array(yb, np.float32),
             pixelsPerMeter=np.array([ppm], np.float32), **{f"channel{c}": chans[c].ravel() for c in range(4)})
    cost = m.ARStandardCost()
    tex = cost.loadTrackDataFromFile(path)
    assert tex.shape == (h, w, 4)
    for c in range(4):
        np.testing.assert_array_equal(tex[..., c], chans[c])
    ref = m.ARStandardCost()
    ref.loadTrackData(ch0, xb[0], xb[1], yb[0], yb[1], ppm)
    assert bytes(cost.params) == bytes(ref.params)  # same transform and map size as the in-memory path
    np.testing.assert_array_equal(cost.costmap[..., 0].reshape(h, w), ch0.astype(np.float32))


def test_cpp_host_layer_loads_the_same_files(tmp_path):
    """NeuralNetModel::loadParams / ARStandardCost::loadTrackData(path) of the C++ mirror (tests/cpp/npz_load_test.cpp)."""
    model = str(tmp_path / "model.npz")
    theta = _write_model(model, compressed=True, seed=4)
    ch0, xb, yb, ppm = W.track_map_standard()
    track = str(tmp_path / "track.npz")
    h, w = ch0.shape
    np.savez(track, xBounds=np.array(xb, np.float32), yBounds=np.array(yb, np.float32),
             pixelsPerMeter=np.array([ppm], np.float32), channel0=ch0.astype(np.float32).ravel(),
             channel1=np.zeros(h * w, np.float32), channel2=np.ones(h * w, np.float32), channel3=np.zeros(h * w, np.float32))
    exe = os.path.join(ROOT, "tests", "cpp", "npz_load_test.bin")
    lib_dir = os.path.join(ROOT, "mppi-generic_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unused-variable", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "npz_load_test.cpp"), "-o", exe, "-L", lib_dir, "-lmppi_b200",
                           "-Wl,-rpath," + lib_dir])
    p = subprocess.run([exe, model, track], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    vals = dict(line.split("=") for line in p.stdout.split())
    assert float(vals["theta_sum"]) == pytest.approx(float(theta.astype(np.float64).sum()), rel=1e-6)
    assert float(vals["theta_abs"]) == pytest.approx(float(np.abs(theta).astype(np.float64).sum()), rel=1e-6)
    assert int(vals["width"]) == w and int(vals["height"]) == h
    assert float(vals["ch0_sum"]) == pytest.approx(float(ch0.astype(np.float64).sum()), rel=1e-6)
    assert float(vals["ch2_sum"]) == pytest.approx(float(h * w), rel=1e-6)
    assert int(vals["missing_rc"]) != 0


def test_racer_lstm_load_params_from_pytorch_layout(tmp_path):
    """LSTMHelper::loadParams (lstm_helper.cu:496-585): PyTorch gate order i, f, g (cell), o in the file, the reference's
    packed i, f, o, c in memory; biases summed; head from output/dynamics_*; "model/" prefix tried first."""
    H_, I, L1 = 4, 4, 20
    lstm_w, head = W.synthetic_lstm_weights(H_, L1, seed=7)
    HH, IH = H_ * H_, H_ * I
    blocks_m = [lstm_w[k * HH:(k + 1) * HH].reshape(H_, H_) for k in range(4)]          # packed: i, f, o, c
    blocks_i = [lstm_w[4 * HH + k * IH:4 * HH + (k + 1) * IH].reshape(H_, I) for k in range(4)]
    bias = [lstm_w[4 * HH + 4 * IH + k * H_:4 * HH + 4 * IH + (k + 1) * H_] for k in range(4)]
    to_file = (0, 1, 3, 2)  # file order i, f, c, o
    rng = np.random.RandomState(0)
    split = [rng.randn(H_).astype(np.float32) * 0.1 for _ in range(4)]
    arrays = {
        "model/lstm/weight_hh_l0": np.concatenate([blocks_m[k] for k in to_file]).astype(np.float64),
        "model/lstm/weight_ih_l0": np.concatenate([blocks_i[k] for k in to_file]).astype(np.float64),
        "model/lstm/bias_hh_l0": np.concatenate([bias[k] - split[i] for i, k in enumerate(to_file)]).astype(np.float64),
        "model/lstm/bias_ih_l0": np.concatenate(split).astype(np.float64),
        "model/output/dynamics_W1": head[:L1 * (H_ + I)].reshape(L1, H_ + I).astype(np.float64),
        "model/output/dynamics_b1": head[L1 * (H_ + I):L1 * (H_ + I) + L1].astype(np.float64),
        "model/output/dynamics_W2": head[L1 * (H_ + I) + L1:L1 * (H_ + I) + 2 * L1].reshape(1, L1).astype(np.float64),
        "model/output/dynamics_b2": head[-1:].astype(np.float64),
    }
    path = str(tmp_path / "lstm.npz")
    np.savez(path, **arrays)
    # the C++ mirror reads the same file to the same packed vector (initial state zero there)
    exe = os.path.join(ROOT, "tests", "cpp", "npz_load_test.bin")
    if os.path.exists(exe):
        model = str(tmp_path / "m.npz")
        _write_model(model)
        ch0, xb, yb, ppm = W.track_map_standard()
        track = str(tmp_path / "t.npz")
        np.savez(track, xBounds=np.array(xb, np.float32), yBounds=np.array(yb, np.float32),
                 pixelsPerMeter=np.array([ppm], np.float32), **{f"channel{c}": ch0.astype(np.float32).ravel() for c in range(4)})
        p = subprocess.run([exe, model, track, path], capture_output=True, text=True, timeout=120)
        assert p.returncode == 0, p.stdout + p.stderr
        vals = dict(line.split("=") for line in p.stdout.split())
        expect = np.concatenate([lstm_w, head]).astype(np.float64)
        assert float(vals["lstm_sum"]) == pytest.approx(float(expect.sum()), rel=1e-5, abs=1e-5)
        assert float(vals["lstm_abs"]) == pytest.approx(float(np.abs(expect).sum()), rel=1e-6)
    dyn = m.RacerDubinsElevationLSTMSteering(hidden_dim=H_, output_layers=(H_ + I, L1, 1))
    h0 = np.linspace(-0.2, 0.2, H_).astype(np.float32)
    dyn.setInitialHiddenCell(h0, -h0)
    dyn.loadParamsLSTM(path)
    block = 4 * HH + 4 * IH + 4 * H_
    np.testing.assert_array_equal(dyn.lstm_theta[:4 * HH + 4 * IH], lstm_w[:4 * HH + 4 * IH])
    np.testing.assert_allclose(dyn.lstm_theta[4 * HH + 4 * IH:block], lstm_w[4 * HH + 4 * IH:block], rtol=0, atol=2e-8)
    np.testing.assert_array_equal(dyn.lstm_theta[block + 2 * H_:], head)
    hh, cc = dyn.initial_hidden_cell()
    np.testing.assert_array_equal(hh, h0)  # the init network's state is not part of this file
    np.testing.assert_array_equal(cc, -h0)
    with pytest.raises(ValueError):
        m.RacerDubinsElevationLSTMSteering(hidden_dim=8, output_layers=(12, L1, 1)).loadParamsLSTM(path)


def test_npz_reader_survives_corrupt_archives(tmp_path):
    """Truncated / bit-flipped archives must come back as an error status, never as a crash or an exception across the C ABI."""
    path = str(tmp_path / "ok.npz")
    np.savez(path, a=np.arange(1000, dtype=np.float32))
    raw = open(path, "rb").read()
    rng = np.random.RandomState(0)
    for trial in range(60):
        b = bytearray(raw)
        if trial % 3 == 0:
            b = b[:rng.randint(10, len(b) - 1)]
        else:
            for _ in range(rng.randint(1, 6)):
                b[rng.randint(0, len(b))] = rng.randint(0, 256)
        bad = tmp_path / f"bad{trial}.npz"
        bad.write_bytes(bytes(b))
        try:
            got = H.npz_read(str(bad), "a")
            assert got.size == 1000  # a flip in the payload or in an unused field still parses
        except m.MppibError:
            pass
