#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_beam_gpu.py tests/test_api_gpu.py -m gpu -q -x > gpurun_out/pytest_beam.log 2>&1
echo "== beam rc=$?"; grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_beam.log | tail -8; tail -30 gpurun_out/pytest_beam.log | head -40
python - <<'PY'
import json
d=json.load(open('gpurun_out/parity_report.json'))
print({k:v for k,v in d.items() if k.startswith('beam') or k.startswith('ctc_align')})
PY
