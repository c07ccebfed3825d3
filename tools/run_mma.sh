#!/bin/bash
# GPU-box script: mma.sync network — probe, parity of the K1 variant, bench next to the FFMA2 default.
mkdir -p gpurun_out
timeout 120 tools/mma_probe 2>&1 | grep -v scalar
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "autorally and (mma or agree)" 2>&1 | tail -15
MPPIB_NN_MMA=1 timeout 300 python bench.py --no-cpu-baseline --steps 200 --warmup 20 2>/dev/null | tail -1 > gpurun_out/bench_autorally_mma.json
python -c "import json; d=json.load(open('gpurun_out/bench_autorally_mma.json')); print('MMA', d['engine']['k1_launch'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline']['stage_ms_l2_warm'])"
timeout 300 python bench.py --no-cpu-baseline --steps 200 --warmup 20 2>/dev/null | tail -1 > gpurun_out/bench_autorally_ffma2.json
python -c "import json; d=json.load(open('gpurun_out/bench_autorally_ffma2.json')); print('FFMA2', d['engine']['k1_launch'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline']['stage_ms_l2_warm'])"
