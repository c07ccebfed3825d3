"""Host-visible cost of the individual stages (GPU box): K1-only, K2-only and full solves, 2000 calls each."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import mppi_generic_b200 as m
from mppi_generic_b200 import workloads as W

name = sys.argv[1] if len(sys.argv) > 1 else "cartpole"
w = W.by_name(name)
e = w.make_engine()
x0, U0 = np.ascontiguousarray(w.x0), np.ascontiguousarray(w.U0)
for _ in range(20):
    e.solve(x0, U0)
def bench(f, n=2000):
    f(); t = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t) / n * 1e6
print(name, e.launch_info(), e.rng_info())
print("solve      us", bench(lambda: e.solve(x0, U0)))
print("rollout    us", bench(lambda: e.rollout_only(x0, U0)))
print("reduce     us", bench(lambda: e.reduce_only()))
print("draw       us", bench(lambda: e.draw_noise()))
U_out = np.empty_like(U0); stats = (m.host.SolveStats * w.D)()
print("solve_into us", bench(lambda: e.solve_into(x0, U0, U_out, stats)))
