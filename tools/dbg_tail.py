import numpy as np, sys
sys.path.insert(0, "/root/repo")
import mppi_generic_b200 as m
from mppi_generic_b200 import workloads as W
H = m.host
for name, w in (("racer", W.racer_lstm(1024, 150)), ("racer32", W.racer_lstm_h32(1024, 60))):
    e = w.make_engine()
    x0 = np.ascontiguousarray(w.x0, np.float32)
    U, _ = e.solve(x0, w.U0)
    Us, st, out = e.nominal_trajectory(x0, U, None)
    bad = np.argwhere(~np.isfinite(out))
    print(name, "nonfinite outputs:", len(bad), "first", bad[:5].tolist(), "cols", sorted(set(bad[:, 2].tolist())))
    bad = np.argwhere(~np.isfinite(st))
    print(name, "nonfinite states:", len(bad), "first", bad[:5].tolist())
    states = np.zeros((w.T, 19), np.float32); outputs = np.zeros((w.T, 28), np.float32)
    w.dyn.output_trajectory(x0[0], U[0], w.T, w.dt, states, outputs)
    d = np.abs(st[0] - states); print(name, "state maxdiff per col", np.nanmax(d, axis=0).round(6).tolist())
    d = np.abs(out[0] - outputs); print(name, "output maxdiff per col", np.nanmax(d, axis=0).round(6).tolist())
    print(name, "host out row1", outputs[1].round(4).tolist()); print(name, "dev out row1", out[0, 1].round(4).tolist())
    e.close()
