#!/bin/bash
# GPU-box script: ablations of the mma.sync K1 (libraries built by hand with -DMPPIB_EXP_*; see plugins/dynamics.cuh, costs.cuh)
for v in ${VARIANTS:-"" NEWTON FAST_SINCOS NO_COST}; do
  for n in 32768 8192; do
    if [ "$v" = "DEFAULT" ] || [ -z "$v" ]; then L=""; else L="/root/repo/tools/libexp_$v.so"; fi
    MPPIB_LIB=$L timeout 200 python bench.py --rollouts $n --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$v]', $n, 'K1 us', round(d['roofline']['stage_ms_l2_warm']['rollout_ms']*1000,1), 'value', round(d['value'],1))"
  done
done
