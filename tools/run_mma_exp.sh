#!/bin/bash
# GPU-box script: ablations / experiments on K1. Build the variants first (CPU box): bash tools/build_exp.sh NAME [MACROS..].
#   VARIANTS="DEFAULT DEFER_COST PACKED" PARITY=DEFER_COST bash tools/run_mma_exp.sh                      (Autorally, mma.sync K1)
#   WORKLOAD=racer_lstm SIZES="65536" VARIANTS="DEFAULT LSTM_FFMA2" PARITY=LSTM_FFMA2 PARITY_K=racer bash tools/run_mma_exp.sh
WORKLOAD=${WORKLOAD:-autorally}
for v in ${VARIANTS:-DEFAULT NEWTON FAST_SINCOS NO_COST DEFER_COST PAIR_RCP PACKED}; do
  if [ "$v" = "DEFAULT" ]; then L=""; else L="/root/repo/tools/libexp_$v.so"; [ -f "$L" ] || { echo "[$v] not built"; continue; }; fi
  for n in ${SIZES:-32768 8192}; do
    MPPIB_LIB=$L timeout 200 python bench.py --workload $WORKLOAD --rollouts $n --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$v]', '$WORKLOAD', $n, 'K1 us', round(d['roofline']['stage_ms_l2_warm']['rollout_ms']*1000,1), 'value', round(d['value'],1))"
  done
done
# a variant that changes arithmetic or ordering must stay parity-green before its number means anything
if [ -n "${PARITY:-}" ]; then
  MPPIB_LIB=/root/repo/tools/libexp_$PARITY.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${PARITY_K:-autorally}" 2>&1 | tail -5
fi
