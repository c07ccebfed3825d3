#!/bin/bash
# GPU-box script: is the consumer warp (kinematics + map cost) the limiter of a lone group? K1 of the Autorally pair with the
# map lookups (NO_TEX) or the whole state cost (NO_COST) compiled out (tools/build_exp.sh), over rollout counts.
for lib in "" /root/repo/tools/libexp_NO_TEX.so /root/repo/tools/libexp_NO_COST.so; do
  for n in 4096 8192 32768; do
    MPPIB_LIB=$lib timeout 200 python bench.py --workload autorally --rollouts $n --steps 100 --warmup 10 --no-cpu-baseline --no-reference-gpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[${lib:-DEFAULT}]', $n, 'K1 us', round(d['roofline']['stage_ms_l2_warm']['rollout_ms']*1000,1), 'value', round(d['value'],1))"
  done
done
