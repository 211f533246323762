# usage: bash tools/sweep_reserve.sh   (diagnostic A/B of scheduling knobs; prints ms/step, e2e ms/step, SM MHz)
for t in 0 1 0 1; do
  echo "small_ew4=$t $(VD_SMALL_EW4=$t timeout 300 python bench.py --no-cpu --no-resident 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'])")"
done
for t in 0 1; do VD_SMALL_EW4=$t timeout 250 python tools/phase_times.py 0 2>&1 | tail -1; done
