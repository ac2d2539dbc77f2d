#!/bin/bash
# round 2, call 48: 8 x B200 with the final build: lines (weak), accurate, pages (strong)
mkdir -p gpurun_out
run() { N=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N "$@" 2> gpurun_out/r2_48_n${N}.err | grep '^{' > gpurun_out/r2_48_tmp.json
  python - "$N" "$*" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2_48_tmp.json').readline())
    print('N',sys.argv[1],sys.argv[2],'value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'clk',d.get('clocks',{}).get('sm_mhz'))
except Exception as e:
    print('N',sys.argv[1],'failed',e)
PY
}
run 8; cp gpurun_out/r2_48_tmp.json gpurun_out/r2_48_lines_n8.json
timeout 600 python bench.py 2>/dev/null | grep '^{' > gpurun_out/r2_48_lines_n1.json; python -c "
import json; d=json.loads(open('gpurun_out/r2_48_lines_n1.json').readline()); print('N 1 value',round(d['value']),'e2e',round(d['e2e']['value']))"
run 8 --workload pages; cp gpurun_out/r2_48_tmp.json gpurun_out/r2_48_pages_n8.json
run 8 --method accurate; cp gpurun_out/r2_48_tmp.json gpurun_out/r2_48_accurate_n8.json
