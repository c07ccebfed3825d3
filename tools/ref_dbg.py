import sys
sys.path.insert(0, ".")
from mppi_generic_b200 import workloads as W
from oracle import ref_gpu as RG
w = W.autorally(4096, 100)
print("creating", flush=True)
r = RG.autorally(w, 42, small=True)
print("created", r.kernel_choice(), flush=True)
c = r.rollout_costs(w.x0[0])
print("costs", c[:4], flush=True)
U, b, n = r.compute_control(w.x0[0])
print("solve", b, n, flush=True)
