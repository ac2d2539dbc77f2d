#!/bin/bash
# round 2, call 54: 8 x B200, final build, default bench line
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29708 bench.py --gpus 8 2> gpurun_out/r2_54_n8.err | grep '^{' > gpurun_out/r2_54_lines_n8.json
timeout 300 python bench.py 2>/dev/null | grep '^{' > gpurun_out/r2_54_lines_n1.json
python - <<'PY'
import json
a=json.loads(open('gpurun_out/r2_54_lines_n8.json').readline()); b=json.loads(open('gpurun_out/r2_54_lines_n1.json').readline())
print('N 8 value',round(a['value']),'ms',round(a['ms_per_step'],4),'e2e',round(a['e2e']['value']),'| N 1 value',round(b['value']),'e2e',round(b['e2e']['value']),'| eff',round(a['value']/8/b['value'],4))
PY
