for v in "" skew1000 skew2000; do
  if [ -z "$v" ]; then L=""; else L="/root/repo/tools/libexp_$v.so"; fi
  MPPIB_LIB=$L MPPIB_NN_TENSOR=1 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tc variant [$v]', 'K1 us', round(d['roofline']['stage_ms_l2_warm']['rollout_ms']*1000,1), d['engine']['k1_launch']['grid'], d['engine']['k1_launch']['block'])"
done
MPPIB_NN_TENSOR=1 python bench.py --rollouts 18944 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tc one CTA per SM', 'K1 us', round(d['roofline']['stage_ms_l2_warm']['rollout_ms']*1000,1), d['engine']['k1_launch']['grid'], d['engine']['k1_launch']['block'])"
python bench.py --rollouts 18944 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ffma2 N=18944', 'K1 us', round(d['roofline']['stage_ms_l2_warm']['rollout_ms']*1000,1), d['engine']['k1_launch']['grid'], d['engine']['k1_launch']['block'])"
