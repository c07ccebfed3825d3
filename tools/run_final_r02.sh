#!/bin/bash
# GPU-box script, end of round 2: the full -m gpu suite + smoke, the default bench line and one line per config, the ncu
# launch list of the bench command and one full capture each of the C4 (warp-specialised) and C5 rollout kernels.
#   bash tools/run_final_r02.sh   -> gpurun_out/r02_final_*
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/r02_final_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2 >> $O/r02_final_tests.txt
timeout 600 python bench.py --steps 50 --warmup 5 2>$O/r02_final_bench.err | tail -1 > $O/r02_final_bench_default_1gpu.json
timeout 600 python bench.py --all-configs --steps 50 --warmup 5 2>>$O/r02_final_bench.err | grep '^{' > $O/r02_final_bench_all_configs_1gpu.jsonl
timeout 300 python bench.py --workload racer_lstm_h32 --steps 30 --warmup 5 --no-cpu-baseline --no-reference-gpu 2>>$O/r02_final_bench.err | tail -1 > $O/r02_final_bench_racer_h32_1gpu.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>>$O/r02_final_bench.err | tail -1 > $O/r02_final_bench_reference_arm.json
# launch list of the bench command (per-launch times are cold-cache and serialised: the kernel's SHARE of the step is what counts)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_final_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-other-configs --no-cpu-baseline --no-reference-gpu > $O/r02_final_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel_ar_ws -s 3 -c 1 -o $O/r02_final_autorally_k1 -f \
  python bench.py --steps 5 --warmup 3 --no-other-configs --no-cpu-baseline --no-reference-gpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -o $O/r02_final_racer_k1 -f \
  python bench.py --workload racer_lstm --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > /dev/null 2>&1
for k in autorally racer; do
  python tools/summarize_ncu.py $O/r02_final_${k}_k1.ncu-rep $O/r02_final_${k}_k1 > /dev/null 2>&1
done
cat $O/r02_final_tests.txt; tail -3 $O/r02_final_bench.err
python - <<P
import json
for f in ("$O/r02_final_bench_default_1gpu.json",):
    d=json.loads(open(f).read()); print("default", d["value"], d["e2e"]["value"], d["roofline"]["stage_ms_l2_warm"], d.get("vs_reference_gpu"), [ (o["workload"], round(o["value"]), round(o["e2e"])) for o in d["other_configs"]])
P
