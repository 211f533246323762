# usage: bash tools/sweep_reserve.sh   (diagnostic: option-stream SM reserve)
for r in 0 8 16 24; do
  echo "reserve=$r $(VD_OPT_RESERVE_SMS=$r timeout 300 python bench.py --no-cpu --no-resident 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3), round(d['roofline']['avg_launch_ms'],4))")"
done
