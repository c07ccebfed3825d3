"""CPU checks of host-side logic in the mirrors that needs no GPU (run under `-m "not gpu"`)."""
import math

import numpy as np

import mppi_generic_b200 as m

H = m.host


def test_base_enforce_leash_is_componentwise():
    """Dynamics::enforceLeash (dynamics.cuh:448-466)."""
    d = H.CartpoleDynamics(1.0, 1.0, 1.0)
    t = np.array([0.0, 1.0, -2.0, 0.3], np.float32)
    n = np.array([0.4, 3.0, -2.05, -0.3], np.float32)
    leash = np.array([0.5, 0.5, 0.01, 0.2], np.float32)
    out = d.enforceLeash(t, n, leash)
    np.testing.assert_allclose(out, [0.4, 1.5, -2.01, 0.1], rtol=0, atol=1e-6)


def test_racer_dubins_enforce_leash_body_frame_and_yaw_wrap():
    """RacerDubinsImpl::enforceLeash (racer_dubins.cu:177-230): x / y are leashed in the body frame of the TRUE state, yaw by
    the shortest angular distance (3.0 -> -3.0 rad is 0.283 rad apart, not 6), everything else component-wise."""
    d = H.RacerDubinsElevationLSTMSteering()
    t = np.zeros(19, np.float32)
    n = np.zeros(19, np.float32)
    leash = np.full(19, 0.5, np.float32)
    t[1], n[1] = 3.0, -3.0          # YAW
    n[2], n[3] = 2.0, 0.1           # POS_X, POS_Y
    t[0], n[0] = 1.0, 2.0           # VEL_X: 1.0 apart, leash 0.5
    out = d.enforceLeash(t, n, leash)
    assert out[1] == np.float32(-3.0)  # within the leash across the wrap: the nominal yaw is taken
    assert out[0] == np.float32(1.5)
    c, s = math.cos(3.0), math.sin(3.0)
    dxb = np.clip(2.0 * c + 0.1 * s, -0.5, 0.5)
    dyb = np.clip(-2.0 * s + 0.1 * c, -0.5, 0.5)
    np.testing.assert_allclose(out[2:4], [dxb * c - dyb * s, dxb * s + dyb * c], atol=1e-6)
    # the base formula would have moved yaw by the full 0.5 the wrong way and clipped x / y in the map frame
    base = H._Dynamics.enforceLeash(d, t, n, leash)
    assert base[1] == np.float32(2.5) and base[2] == np.float32(0.5)
    # a yaw further away than the leash: leashed along the short way and re-normalised into (-pi, pi]
    t[1], n[1] = 3.0, -2.0
    out = d.enforceLeash(t, n, leash)
    assert out[1] == np.float32(np.float32(3.5) - np.float32(2 * math.pi)) or abs(out[1] - (3.5 - 2 * math.pi)) < 1e-6
