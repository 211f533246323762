for r in 0 12 20 28 40 56; do
  echo "reserve=$r $(VD_OPT_RESERVE_SMS=$r timeout 300 python bench.py --no-cpu --no-resident 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'])")"
done
VD_OPT_RESERVE_SMS=20 timeout 250 python tools/phase_times.py 1 2>&1 | tail -1
