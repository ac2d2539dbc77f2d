#!/bin/bash
# round 2, call 46: conv1 on the tensor pipe as the default: full GPU suite + bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for m in fast accurate; do
timeout 600 python bench.py --method $m 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
print('$m value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'conv1',round(d['stages']['conv1']['ms_per_step'],4))"
done
