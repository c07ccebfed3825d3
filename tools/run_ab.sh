#!/bin/bash
# GPU-box script: A/B of experimental library builds (tools/build_exp.sh NAME) on one workload.  WORKLOAD=racer_lstm VARIANTS="DEFAULT NO_TEX" bash tools/run_ab.sh
for v in ${VARIANTS:-DEFAULT}; do
  if [ "$v" = "DEFAULT" ]; then L=""; else L="/root/repo/tools/libexp_$v.so"; fi
  for wl in ${WORKLOADS:-racer_lstm}; do
  MPPIB_LIB=$L timeout 300 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline --no-reference-gpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$v]', '$wl', 'K1 us', round(d['roofline']['kernel_ms_l2_warm']*1000,1), 'cold', round(d['roofline']['kernel_ms_l2_flushed']*1000,1), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
  done
done
