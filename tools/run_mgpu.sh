# usage: bash tools/run_mgpu.sh N   (under gpurun --gpus N): sharded == single-GPU check, the multi-GPU test, then the bench
N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 2>&1 | grep -v "^W0\|^\[W\|^\*\*\*\|OMP_NUM" | tail -6
for wl in ${WLS:-autorally racer_lstm cartpole}; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload $wl --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/mgpu_$N_$wl.err | tail -1 > gpurun_out/r02_bench_${wl}_${N}gpu.json
python -c "import sys,json; d=json.load(open('gpurun_out/r02_bench_${wl}_${N}gpu.json')); print('gpus', d['n_gpus'], d['config']['workload'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'solve_only', round(d['e2e']['solve_only_value'],1), 'parity_ok', d.get('parity_ok'), d.get('max_abs_dU_vs_single_gpu'), d['roofline']['stage_ms_l2_warm'], d['engine']['k1_launch'])" || tail -5 gpurun_out/mgpu_$N_$wl.err
done
