#!/bin/bash
# GPU-box script: whole GPU suite + smoke + the default bench line (sanity after a batch of changes).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_autorally_check.json
python -c "import json; d=json.load(open('gpurun_out/bench_autorally_check.json')); print(d['config']['workload'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline']['value'], d['roofline']['stage_ms_l2_warm'], d['clocks'])"
