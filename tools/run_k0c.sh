mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "colored or racer" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_racer.csv python bench.py --workload racer_lstm --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/racer_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_racer.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
from collections import defaultdict
d=defaultdict(list)
for r in rows[1:]:
    try: d[r[ki][:60]].append(float(r[vi].replace(',','')))
    except: pass
for k,v in d.items(): print(k, len(v), 'avg us', round(sum(v)/len(v)/1000,1) if max(v)>1000 else round(sum(v)/len(v),1))
PY
python bench.py --workload racer_lstm --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('auto', d['config']['workload'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline']['stage_ms_l2_warm'], d['engine']['k1_launch'])"
