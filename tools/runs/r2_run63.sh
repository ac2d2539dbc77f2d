#!/bin/bash
# round 2, call 63: final state — full GPU suite, smoke, fast / accurate bench lines, ncu launch list of the fast step
O=gpurun_out/r02_final
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 ${O}_pytest.log
cp gpurun_out/parity_report.json ${O}_parity_report.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee ${O}_smoke.log
summ() { python - "$1" <<'PY'
import json,sys
d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')][0]
print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'launches',d.get('gpu_launches'),'roof',round(d['roofline']['frac'],3),'clk',d.get('clocks',{}).get('sm_mhz'))
if 'stages' in d: print({k:round(v['ms_per_step'],4) for k,v in d['stages'].items()})
PY
}
timeout 300 python bench.py > ${O}_bench_fast.json 2> ${O}_bench_fast.err; echo "bench fast rc=$?"; summ ${O}_bench_fast.json
timeout 300 python bench.py --method accurate > ${O}_bench_accurate.json 2> ${O}_bench_accurate.err; echo "bench accurate rc=$?"; summ ${O}_bench_accurate.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file ${O}_launches_fast.csv \
    python bench.py --steps 2 --warmup 3 > ${O}_ncu_launch_fast.log 2>&1; echo "launch list rc=$?"
