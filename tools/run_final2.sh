#!/bin/bash
# GPU-box script (end of round 1, after the mma.sync network became the Autorally default): whole GPU suite, smoke, bench
# lines, ncu launch list and one full capture of the new K1.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_autorally.json
python -c "import json; d=json.load(open('gpurun_out/bench_autorally.json')); print(d['config']['workload'], d['engine']['k1_launch'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'solve_only', round(d['e2e']['solve_only_value'],1), 'cpu', d['cpu_baseline']['value'], 'frac', round(d['roofline']['frac'],4), d['roofline']['stage_ms_l2_warm'], d['clocks'])"
MPPIB_NN_FFMA2=1 timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_autorally_ffma2.json
python -c "import json; d=json.load(open('gpurun_out/bench_autorally_ffma2.json')); print('FFMA2 form', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline']['stage_ms_l2_warm'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_autorally_mma.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -o gpurun_out/final_autorally_k1_mma -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -8
