"""Regenerates the golden fixtures of this directory:  python tests/golden/make_golden.py

The reference is C++/CUDA and cannot be built in the development container (Eigen is absent), so the
vectors come from the CPU oracle (oracle/mppi_oracle.cpp), which is itself pinned against the known-answer values of
the reference's own tests (tests/test_oracle_golden.py, values transcribed in reference_known_answers.json with their
file:line). Each .npz holds one small seeded solve: inputs (x0, U0, the raw N(0,1) block from the host XORWOW
generator) and the oracle's outputs (per-sample costs, new mean, baseline, normaliser). Tests: tests/test_golden_fixtures.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from mppi_generic_b200 import workloads as W  # noqa: E402

CASES = [("cartpole", 256, 40, 7), ("double_integrator_tube", 256, 30, 8), ("autorally", 256, 25, 9),
         ("racer_lstm_gaussian", 256, 30, 10)]


def solve_case(name, N, T, seed):
    w = W.by_name(name, N, T)
    rng = np.random.RandomState(seed)
    w.U0 = rng.uniform(-0.2, 0.2, w.U0.shape).astype(np.float32)
    if name == "autorally":
        w.x0[0, :2] = [0.013, 0.021]  # off the texel boundaries of the test map
    if hasattr(w.dyn, "lstm_theta"):
        oracle.set_lstm(w.dyn.lstm_theta, w.dyn.hidden_dim, w.dyn.head_hidden)
    Cd = w.dyn.CONTROL_DIM
    eps = oracle.curand_normal(seed, 0, N * T * Cd).reshape(N, T, Cd)
    ref = oracle.solve(w.dyn.DYN_ID, w.cost.COST_ID, w.dyn.params, w.cost.params, w.sampler.params, w.dyn.nn_theta,
                       getattr(w.cost, "costmap", None), N, T, w.D, Cd, w.dt, w.lambda_, w.alpha, w.x0, w.U0, eps)
    return w, eps, ref


def main():
    for name, N, T, seed in CASES:
        w, eps, ref = solve_case(name, N, T, seed)
        np.savez_compressed(os.path.join(HERE, f"{name}_N{N}_T{T}.npz"), x0=w.x0, U0=w.U0, eps=eps, costs=ref["costs"],
                            U=ref["U"], baseline=ref["baseline"], normalizer=ref["normalizer"], seed=seed)
        print(name, "baseline", ref["baseline"], "normalizer", ref["normalizer"])


if __name__ == "__main__":
    main()
