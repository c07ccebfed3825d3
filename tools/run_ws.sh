#!/bin/bash
# GPU-box script: K1 of the Autorally pair over rollout counts — warp-specialised with 16 / 32 samples per producer warp
# (WS16 = default, WS32) and the generic one-thread-per-sample kernel (GENERIC); then the Autorally parity tests.
for v in ${VARIANTS:-WS16 WS32 GENERIC}; do
  for n in ${SIZES:-32768 16384 8192 4096}; do
    unset MPPIB_NO_WS MPPIB_WS_PSPW
    if [ "$v" = "GENERIC" ]; then export MPPIB_NO_WS=1; fi
    if [ "$v" = "WS32" ]; then export MPPIB_WS_PSPW=32; fi
    if [ "$v" = "WS16" ]; then export MPPIB_WS_PSPW=16; fi
    if [ "$v" = "WS8" ]; then export MPPIB_WS_PSPW=8; fi
    timeout 200 python bench.py --workload autorally --rollouts $n --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$v]', $n, 'K1 us', round(d['roofline']['stage_ms_l2_warm']['rollout_ms']*1000,1), 'cold', round(d['roofline']['kernel_ms_l2_flushed']*1000,1), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['engine'].get('k1_launch'))"
  done
done
unset MPPIB_NO_WS MPPIB_WS_PSPW
if [ -z "${NO_PARITY:-}" ]; then
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${PARITY_K:-autorally}" 2>&1 | tail -15
fi
