for cfg in "2 256 32768" "2 0 65536" "1 0 65536"; do set -- $cfg
MPPIB_SPT=$1 MPPIB_BX=$2 python bench.py --rollouts $3 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('spt $1 bx $2 N $3', 'K1 us', round(d['roofline']['stage_ms_l2_warm']['rollout_ms']*1000,1), d['engine']['k1_launch']['grid'], d['engine']['k1_launch']['block'])"
done
