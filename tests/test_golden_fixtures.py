"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle must keep reproducing
them (CPU), and the CUDA path fed with the SAME noise block must match them through the C ABI (GPU)."""
import glob
import json
import os

import numpy as np
import pytest

import oracle
import mppi_generic_b200 as m
from mppi_generic_b200 import workloads as W
from tests.golden import make_golden as G

H = m.host
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name, N, T):
    return np.load(os.path.join(HERE, "golden", f"{name}_N{N}_T{T}.npz"))


def test_fixture_files_exist_and_known_answers_are_cited():
    assert len(glob.glob(os.path.join(HERE, "golden", "*.npz"))) == len(G.CASES)
    ka = json.load(open(os.path.join(HERE, "golden", "reference_known_answers.json")))
    for k, v in ka.items():
        if not k.startswith("_"):
            assert "source" in v and ":" in v["source"], k


@pytest.mark.parametrize("name,N,T,seed", G.CASES)
def test_oracle_reproduces_the_committed_vectors(name, N, T, seed):
    g = _load(name, N, T)
    w, eps, ref = G.solve_case(name, N, T, seed)
    np.testing.assert_array_equal(eps, g["eps"])  # host XORWOW stream is deterministic
    np.testing.assert_allclose(ref["costs"], g["costs"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ref["U"], g["U"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ref["baseline"], g["baseline"], rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name,N,T,seed", G.CASES)
def test_engine_matches_the_committed_vectors(name, N, T, seed):
    g = _load(name, N, T)
    w = W.by_name(name, N, T)
    w.x0, w.U0 = g["x0"], g["U0"]
    e = w.make_engine()
    e.set_noise(g["eps"])
    e.rollout_only(w.x0, w.U0, 1, 0)
    U, stats = e.reduce_only()
    tol = 2e-4 if name.startswith("racer") else 1e-4  # device intrinsics vs host libm, see test_racer_lstm_solve_parity
    np.testing.assert_allclose(e.get_costs(), g["costs"], rtol=tol, atol=1e-5)
    scale = max(1.0, float(np.abs(g["U"]).max()))
    np.testing.assert_allclose(U, g["U"], atol=2e-3 * scale)
    for d in range(w.D):
        assert stats[d][0] == pytest.approx(float(g["baseline"][d]), rel=tol)
    e.close()
