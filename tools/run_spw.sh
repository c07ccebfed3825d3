#!/bin/bash
# GPU-box script: K1 of the Autorally pair for every samples-per-warp width (plugins/nn_mma.cuh) and rollout count,
# then the Autorally parity tests with each width forced.   bash tools/run_spw.sh
for spw in ${SPWS:-32 16 8}; do
  for n in ${SIZES:-32768 16384 8192 4096}; do
    MPPIB_SPW=$spw timeout 200 python bench.py --workload autorally --rollouts $n --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[spw $spw]', $n, 'K1 us', round(d['roofline']['stage_ms_l2_warm']['rollout_ms']*1000,1), 'cold', round(d['roofline']['kernel_ms_l2_flushed']*1000,1), 'value', round(d['value'],1), 'grid', d['engine'].get('k1_launch'))"
  done
done
if [ -z "${NO_PARITY:-}" ]; then
for spw in ${SPWS:-32 16 8}; do
  echo "parity spw=$spw"; MPPIB_SPW=$spw timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "autorally" 2>&1 | tail -4
done
fi
