# usage: bash tools/run_mgpu3.sh N   (under gpurun --gpus N): sharded-solve check + one autorally bench line, short
N=$1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 2>&1 | grep -v "^W0\|^\[W\|^\*\*\*\|OMP_NUM" | tail -4
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('gpus', d['n_gpus'], d['config']['workload'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'solve_only', round(d['e2e']['solve_only_value'],1), d['roofline']['stage_ms_l2_warm'])"
