#!/bin/bash
# GPU-box script: A/B of two library builds on the Autorally K1 over rollout counts (default producers), interleaved twice.
for rep in 1 2; do
for lib in "" /root/repo/tools/libexp_OLDLAYOUT.so; do
  for n in 32768 16384 8192 4096; do
    MPPIB_LIB=$lib timeout 200 python bench.py --workload autorally --rollouts $n --steps 100 --warmup 10 --no-cpu-baseline --no-reference-gpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[${lib:+OLD}${lib:-NEW}]', $n, 'K1 us', round(d['roofline']['stage_ms_l2_warm']['rollout_ms']*1000,1), 'cold', round(d['roofline']['kernel_ms_l2_flushed']*1000,1), 'value', round(d['value'],1), 'smem', d['engine']['k1_launch']['smem_bytes'], d['engine']['k1_launch']['grid'], d['engine']['k1_launch']['block'])"
  done
done
done
MPPIB_NO_WS=1 timeout 200 python bench.py --workload autorally --steps 100 --warmup 10 --no-cpu-baseline --no-reference-gpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[NEW generic]', 'K1 us', round(d['roofline']['stage_ms_l2_warm']['rollout_ms']*1000,1))"
MPPIB_LIB=/root/repo/tools/libexp_OLDLAYOUT.so MPPIB_NO_WS=1 timeout 200 python bench.py --workload autorally --steps 100 --warmup 10 --no-cpu-baseline --no-reference-gpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[OLD generic]', 'K1 us', round(d['roofline']['stage_ms_l2_warm']['rollout_ms']*1000,1))"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py -m gpu -x -q -k "autorally or reference" 2>&1 | tail -4
