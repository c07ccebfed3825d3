python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py -q -m gpu -k "autorally" 2>&1 | tail -3
for n in 32768 8192; do
python bench.py --rollouts $n --steps 200 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'solve_only', round(d['e2e']['solve_only_value'],1), d['roofline']['stage_ms_l2_warm'], d['engine']['k1_launch']['grid'], d['engine']['k1_launch']['block'])"
done
