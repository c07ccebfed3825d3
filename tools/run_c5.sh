#!/bin/bash
# GPU-box script: C5 (racer LSTM + colored noise) K1, streaming epilogue recompute (default) vs read-back (round-1 form),
# then ncu DRAM traffic of K1 for the default.
for v in RECOMPUTE READBACK; do
  unset MPPIB_STREAM_READBACK; [ "$v" = "READBACK" ] && export MPPIB_STREAM_READBACK=1
  timeout 300 python bench.py --workload racer_lstm --steps 30 --warmup 5 --no-cpu-baseline --no-reference-gpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$v]', 'K1 us', round(d['roofline']['kernel_ms_l2_warm']*1000,1), 'cold', round(d['roofline']['kernel_ms_l2_flushed']*1000,1), 'value', round(d['value'],1), d['roofline']['stage_ms_l2_warm'], d['engine']['k1_launch'])"
done
unset MPPIB_STREAM_READBACK
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -o gpurun_out/r02_racer_k1 -f python bench.py --workload racer_lstm --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > /dev/null 2>&1
