mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3
for wl in autorally cartpole double_integrator_tube racer_lstm; do
python bench.py --workload $wl --steps 200 --warmup 20 2>/dev/null | tail -1 > gpurun_out/bench_$wl.json
python -c "import sys,json; d=json.load(open('gpurun_out/bench_$wl.json')); print(d['config']['workload'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'solve_only', round(d['e2e']['solve_only_value'],1), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],2), d['cpu_baseline'] and d['cpu_baseline']['cores'], 'frac', round(d['roofline']['frac'],4), d['roofline']['stage_ms_l2_warm'], d['clocks'])"
done
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference_autorally.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_autorally.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -o gpurun_out/final_autorally_k1 -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rollout_kernel|combine_kernel|xorwow_normal|colored_rearrange|regular_fft" -s 12 -c 5 -o gpurun_out/final_racer -f python bench.py --workload racer_lstm --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -12
