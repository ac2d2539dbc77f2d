#!/bin/bash
# round 2, call 58: final full GPU suite + smoke on the committed state
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_58_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2_58_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('fast value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'],'roof',round(d['roofline']['frac'],3))"
