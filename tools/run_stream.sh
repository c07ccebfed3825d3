MPPIB_STREAM=1 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -6
for wl in racer_lstm double_integrator_tube; do for st in 0 1; do
MPPIB_STREAM=$st python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('stream $st', d['config']['workload'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline']['stage_ms_l2_warm'], d['engine']['k1_launch'])"
done; done
python bench.py --workload autorally --rollouts 65536 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('auto', d['config']['workload'], 'value', round(d['value'],1), d['roofline']['stage_ms_l2_warm'], d['engine']['k1_launch'])"
python bench.py --workload racer_lstm --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('auto', d['config']['workload'], 'value', round(d['value'],1), d['roofline']['stage_ms_l2_warm'], d['engine']['k1_launch'])"
