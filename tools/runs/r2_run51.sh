#!/bin/bash
# round 2, call 51: final refresh: full GPU suite, smoke, bench fast / accurate, HBM exhibit
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_51_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2_51_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_51_smoke.log 2>&1; tail -1 gpurun_out/r2_51_smoke.log
cp gpurun_out/parity_report.json gpurun_out/r2_51_parity_report.json
timeout 600 python bench.py > gpurun_out/r2_51_bench_fast.json 2>/dev/null; timeout 600 python bench.py --method accurate > gpurun_out/r2_51_bench_accurate.json 2>/dev/null
python - <<'PY'
import json
for m in ('fast','accurate'):
    d=[json.loads(l) for l in open(f'gpurun_out/r2_51_bench_{m}.json') if l.startswith('{')][0]
    print(m,'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'roof',round(d['roofline']['frac'],3))
    if m=='fast': print({k:round(v['ms_per_step'],4) for k,v in d['stages'].items()})
PY
timeout 600 python tools/bench_hbm_kernels.py > gpurun_out/r2_51_hbm.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/r2_51_hbm.json'))
print({k:(round(v['ms'],4), round(v['frac_of_hbm_peak'],3)) for k,v in d.items()})"
