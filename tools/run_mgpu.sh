#!/bin/bash
# GPU-box script: the default bench line (C4 headline + other_configs C2/C3/C5 summaries) on N GPUs, and the sharded==single check.
#   bash tools/run_mgpu.sh 8     -> gpurun_out/r02_bench_default_8gpu.json
N=${1:-2}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps 50 --warmup 5 2>gpurun_out/mgpu_$N.err | tail -1 > gpurun_out/r02_bench_default_${N}gpu.json
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 2>gpurun_out/mgpu_$N.err | tail -1 > gpurun_out/r02_bench_default_${N}gpu.json
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/mgpu_check.py 2>&1 | grep MGPU_OK
fi
python - <<P
import json
d=json.loads(open("gpurun_out/r02_bench_default_${N}gpu.json").read())
def row(w,v,e,k1,p): print("gpus $N", w, "value", round(v,1), "e2e", round(e,1), "K1 us", round(k1*1000,1), "parity_ok", p)
for o in d.get("other_configs",[]): row(o["workload"],o["value"],o["e2e"],o["k1_ms_l2_flushed"],o["parity_ok"])
row(d["config"]["workload"],d["value"],d["e2e"]["value"],d["roofline"]["kernel_ms_l2_flushed"],d.get("parity_ok"))
print("device_tail_value", d["e2e"].get("device_tail_value"), "solve_only", d["e2e"]["solve_only_value"], d["roofline"]["stage_ms_l2_warm"])
P
