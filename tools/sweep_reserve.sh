# usage: bash tools/sweep_reserve.sh   (diagnostic A/B of scheduling knobs; prints ms/step, e2e ms/step, SM MHz)
for f in 16 0 4 8 16 4; do
  echo "fwd_reserve=$f $(VD_OPT_RESERVE_FWD=$f timeout 300 python bench.py --no-cpu --no-resident 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'])")"
done
