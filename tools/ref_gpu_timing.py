"""GPU-box script: computeControl Hz of the UNMODIFIED reference GPU build (oracle/_ref/libmppi_ref_gpu.so, "reference
kernels, shimmed host") at C2 and C4 over a few rollout block shapes, best shape reported. Every shape runs in its own
process: the reference exit()s on a shape it rejects (mppi_common.cu:1274).   python tools/ref_gpu_timing.py"""
import json
import subprocess
import sys

CHILD = r'''
import json, sys
sys.path.insert(0, ".")
from mppi_generic_b200 import workloads as W
from oracle import ref_gpu as RG
name, b = sys.argv[1], tuple(int(v) for v in sys.argv[2].split(","))
w = W.by_name(name)
r = (RG.cartpole if name == "cartpole" else RG.autorally)(w, 42, block=b)
s = r.time_compute_control(w.x0[0], 1, warmup=5, iters=30 if name == "autorally" else 200)
print("RESULT " + json.dumps({"block": list(b), "kernel": r.kernel_choice(), "ms": s * 1e3, "hz": 1.0 / s}))
'''
SHAPES = {"cartpole": ["64,4", "32,4", "64,1", "32,1", "64,4,64,1", "32,4,100,1", "64,2,100,1"],
          "autorally": ["64,8", "32,8", "64,4", "32,16", "64,8,100,1", "32,8,100,1", "64,8,64,1", "64,8,64,4", "32,8,64,2"]}
out = {}
for name, shapes in SHAPES.items():
    rows = []
    for b in shapes:
        p = subprocess.run([sys.executable, "-c", CHILD, name, b], capture_output=True, text=True, timeout=300)
        res = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        rows.append(json.loads(res[0][7:]) if res else {"block": b, "error": (p.stdout + p.stderr)[-200:]})
        print(name, rows[-1], flush=True)
    ok = [x for x in rows if "hz" in x]
    out[name] = {"best": max(ok, key=lambda x: x["hz"]) if ok else None, "all": rows}
print(json.dumps(out))
