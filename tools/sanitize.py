"""Small solves of every kernel variant, meant to run under compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mppi_generic_b200 as m  # noqa: E402
from mppi_generic_b200 import workloads as W  # noqa: E402

H = m.host


def run(w, flags=0, solves=2, **kw):
    e = H.Engine(w.dyn, w.cost, w.sampler, w.N, w.T, w.D, flags=flags, **kw)
    e.set_solver(w.dt, w.lambda_, w.alpha)
    e.seed(3, 0)
    for s in range(solves):
        U, st = e.solve(w.x0, w.U0, 1 + (s % 2), 0)
    assert np.all(np.isfinite(U)), w.name
    info = e.launch_info()
    print("ok", w.name, "flags", flags, info["grid"], info["block"], info["smem_bytes"], flush=True)
    return e


for name, N, T in (("cartpole", 512, 64), ("double_integrator_tube", 512, 64), ("autorally", 512, 32),
                   ("racer_lstm", 512, 64), ("racer_lstm_gaussian", 512, 64)):
    run(W.by_name(name, N, T)).close()
os.environ["MPPIB_STREAM"] = "1"
for name, N, T in (("cartpole", 512, 64), ("double_integrator_tube", 512, 64), ("racer_lstm", 512, 64)):
    run(W.by_name(name, N, T)).close()
del os.environ["MPPIB_STREAM"]
w = W.racer_lstm(512, 64, hidden_dim=32)
run(w).close()
w = W.cartpole(512, 64)
e = run(w, flags=H.FLAG_WRITEBACK_CONTROLS)
e.set_tsallis(50.0, 2.0)
e.solve(w.x0, w.U0)
e.close()
w = W.double_integrator_tube(512, 40)
e = H.Engine(w.dyn, w.cost, w.sampler, w.N, w.T, 2, flags=H.FLAG_RMPPI)
e.set_solver(w.dt, w.lambda_, 0.1)
e.seed(1, 0)
g = (np.random.RandomState(0).randn(w.T, 4, 2) * 0.2).astype(np.float32)
e.set_rmppi(5.0, g)
x0 = np.array([[2, 0, 0, 1], [2.05, 0, 0, 1]], np.float32)
U = np.zeros((2, w.T, 2), np.float32)
e.solve(x0, U)
cand = np.tile(x0[0], (5, 1)).astype(np.float32)
e.init_eval(cand, np.array([0, 0, 1, 1, 1], np.int32), 32, U[0], 1)
e.close()
os.environ["MPPIB_NN_TENSOR"] = "1"
run(W.autorally(512, 32)).close()
print("sanitize run complete")
