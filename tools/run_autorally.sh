python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "autorally or cartpole_parity or tube" 2>&1 | tail -4
for spt in 1 2; do for n in 32768 16384; do
MPPIB_SPT=$spt python bench.py --rollouts $n --steps 200 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('spt $spt', d['config']['workload'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'solve_only', round(d['e2e']['solve_only_value'],1), d['roofline']['stage_ms_l2_warm'], d['config']['k1_launch']['grid'], d['config']['k1_launch']['block'])"
done; done
