"""The init network of the RACER LSTM model: LSTMLSTMHelper::initializeLSTM
(utils/nn_helpers/lstm_lstm_helper.cu:50-73) and RacerDubinsElevationLSTMSteering::updateFromBuffer
(racer_dubins_elevation_lstm_steering.cu:215-233). Host-only in the reference as well; here it lives in the library's host
twins (mppib_host_lstm_initialize) behind both mirrors. Pinned to the reference's own known answer
(tests/nn_helpers/lstm_lstm_helper_test.cu:161-180: all values 1, a buffer of ones -> hidden = cell = 101), then compared
with the oracle's restatement (built from its LSTMHelper::forward, itself pinned to the reference's LSTM goldens in
tests/test_oracle_golden.py) on random weights, buffers longer than init_len, and through the npz loader."""
import numpy as np
import pytest

import mppi_generic_b200 as m
import oracle

H = m.host


def test_initialize_lstm_reference_known_answer():
    d = H.RacerDubinsElevationLSTMSteering(8, 60, (68, 100, 20), 4, 10, (14, 20, 1), 6)  # the reference test's dimensions
    d.setAllValuesInit(np.ones(d.init_lstm_theta.size), np.ones(d.init_head_theta.size))
    d.initializeLSTM(np.ones((8, 10), np.float32))
    h, c = d.initial_hidden_cell()
    assert np.all(h == 101.0) and np.all(c == 101.0)
    got = oracle.lstm_initialize(np.ones(d.init_lstm_theta.size), 8, 60, np.ones(d.init_head_theta.size), (68, 100, 20), 6,
                                 np.ones((10, 8), np.float32))
    assert np.all(got == 101.0)


def _random_model(seed=0, init_len=11):
    rng = np.random.default_rng(seed)
    d = H.RacerDubinsElevationLSTMSteering(3, 20, (23, 100, 8), 4, 4, (8, 20, 1), init_len)  # the model test's architecture
    lstm = (0.3 * rng.standard_normal(d.init_lstm_theta.size)).astype(np.float32)
    head = (0.2 * rng.standard_normal(d.init_head_theta.size)).astype(np.float32)
    d.setAllValuesInit(lstm, head)
    return d, lstm, head, rng


@pytest.mark.parametrize("cols", [11, 25])
def test_initialize_lstm_matches_the_oracle_on_random_weights(cols):
    d, lstm, head, rng = _random_model()
    buf = rng.standard_normal((3, cols)).astype(np.float32)
    d.initializeLSTM(buf)
    h, c = d.initial_hidden_cell()
    ref = oracle.lstm_initialize(lstm, 3, 20, head, (23, 100, 8), 11, np.ascontiguousarray(buf.T))
    np.testing.assert_allclose(np.concatenate([h, c]), ref, rtol=2e-6, atol=2e-6)
    # only the last init_len columns matter (lstm_lstm_helper.cu:57-61)
    buf2 = buf.copy()
    buf2[:, :cols - 11] = 7.0
    d.initializeLSTM(buf2)
    h2, c2 = d.initial_hidden_cell()
    assert np.array_equal(h2, h) and np.array_equal(c2, c)
    with pytest.raises(ValueError):
        d.initializeLSTM(buf[:, :10])  # shorter than init_len


def test_update_from_buffer_scales_the_steering_rows_and_needs_all_keys():
    d, lstm, head, rng = _random_model(3)
    sa, sr, cmd = (rng.standard_normal(15).astype(np.float32) for _ in range(3))
    assert d.updateFromBuffer({"STEER_ANGLE": sa, "STEER_ANGLE_RATE": sr}) is False  # racer_dubins_elevation_lstm_steering.cu:220-224
    assert d.updateFromBuffer({"STEER_ANGLE": sa, "STEER_ANGLE_RATE": sr, "CAN_STEER_CMD": cmd}) is True
    h, c = d.initial_hidden_cell()
    cols = np.stack([sa * np.float32(0.2), sr * np.float32(0.2), cmd], axis=1)  # [cols][3]
    ref = oracle.lstm_initialize(lstm, 3, 20, head, (23, 100, 8), 11, cols)
    np.testing.assert_allclose(np.concatenate([h, c]), ref, rtol=2e-6, atol=2e-6)


def test_load_params_init_from_an_npz_file(tmp_path):
    """PyTorch layout (gate blocks i, f, g, o; separate ih / hh biases) under "<prefix>init_...", init_length + 1."""
    rng = np.random.default_rng(4)
    Hi, Ii = 20, 3
    whh, wih = rng.standard_normal((4 * Hi, Hi)), rng.standard_normal((4 * Hi, Ii))
    bhh, bih = rng.standard_normal(4 * Hi), rng.standard_normal(4 * Hi)
    W1, b1 = rng.standard_normal((100, 23)), rng.standard_normal(100)
    W2, b2 = rng.standard_normal((8, 100)), rng.standard_normal(8)
    path = str(tmp_path / "model.npz")
    np.savez(path, **{"steer/init_lstm/weight_hh_l0": whh, "steer/init_lstm/weight_ih_l0": wih,
                      "steer/init_lstm/bias_hh_l0": bhh, "steer/init_lstm/bias_ih_l0": bih,
                      "steer/init_output/dynamics_W1": W1, "steer/init_output/dynamics_b1": b1,
                      "steer/init_output/dynamics_W2": W2, "steer/init_output/dynamics_b2": b2,
                      "init_length": np.array([10.0])})
    d = H.RacerDubinsElevationLSTMSteering(3, 20, (23, 100, 8), 4, 4, (8, 20, 1), 5)
    d.loadParamsInit(path, "steer")
    assert d.init_len == 11
    order = (0, 1, 3, 2)
    blk = lambda a: np.concatenate([a[k * Hi:(k + 1) * Hi].ravel() for k in order])  # noqa: E731
    f = lambda a: a.astype(np.float32).astype(np.float64)  # the reader hands out floats; the two biases are summed in double  # noqa: E731
    expect = np.concatenate([blk(whh), blk(wih), blk(f(bhh) + f(bih)), np.zeros(2 * Hi)]).astype(np.float32)
    np.testing.assert_array_equal(d.init_lstm_theta, expect)
    np.testing.assert_array_equal(d.init_head_theta, np.concatenate([W1.ravel(), b1, W2.ravel(), b2]).astype(np.float32))
    buf = rng.standard_normal((3, 11)).astype(np.float32)
    d.initializeLSTM(buf)
    ref = oracle.lstm_initialize(d.init_lstm_theta, 3, 20, d.init_head_theta, (23, 100, 8), 11, np.ascontiguousarray(buf.T))
    np.testing.assert_allclose(np.concatenate(d.initial_hidden_cell()), ref, rtol=2e-6, atol=2e-6)
