import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import mppi_generic_b200 as m
from mppi_generic_b200 import workloads as W
H = m.host
w = W.cartpole(8192, 100)
e = w.make_engine()
print(e.rng_info())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
flush = len(sys.argv) > 2
if flush:
    e.set_option(H.OPT_L2_FLUSH_BYTES, 256 << 20)
    e.enable_timing(True)
for i in range(n):
    U, st = e.solve(w.x0, w.U0)
print("done", st)
if flush:
    print("timing", e.timing())
    e.set_option(H.OPT_L2_FLUSH_BYTES, 0)
    e.enable_timing(True)
    for i in range(5):
        U, st = e.solve(w.x0, w.U0)
    print("after free ok", e.timing())
