#!/bin/bash
# round 2, call 47: decoder tests with the streaming near-tie rule; full suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/parity_report.json'))
for k,v in d.items():
    if k.startswith('decoder_stream') or k.startswith('encoder_ctc') or k.startswith('config') or k.startswith('wide_fast'): print(k, json.dumps(v)[:300])
PY
