for v in "" NO_NN NO_COST NO_TEX; do
  for n in 32768 8192; do
    if [ -z "$v" ]; then L=""; else L="/root/repo/tools/libexp_$v.so"; fi
    MPPIB_LIB=$L python bench.py --rollouts $n --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', $n, 'K1 us', round(d['roofline']['stage_ms_l2_warm']['rollout_ms']*1000,1))"
  done
done
