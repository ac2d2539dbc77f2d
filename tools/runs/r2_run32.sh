#!/bin/bash
# round 2, call 32: 8 x B200: weak-scaling lines (N = 8, 4, 2), strong-scaling pages (N = 8)
mkdir -p gpurun_out
run() { # $1 = N, rest = bench args
  N=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N "$@" 2> gpurun_out/r2_32_n${N}.err | grep '^{' > gpurun_out/r2_32_tmp.json
  python - "$N" "$*" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2_32_tmp.json').readline())
    print('N',sys.argv[1],sys.argv[2],'value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'clk',d.get('clocks',{}).get('sm_mhz'), d.get('config',{}).get('ordered_equal_to_single_rank'))
except Exception as e:
    print('N',sys.argv[1],'failed',e)
PY
}
run 8; cp gpurun_out/r2_32_tmp.json gpurun_out/r2_32_lines_n8.json
run 4; cp gpurun_out/r2_32_tmp.json gpurun_out/r2_32_lines_n4.json
run 2; cp gpurun_out/r2_32_tmp.json gpurun_out/r2_32_lines_n2.json
timeout 600 python bench.py 2>/dev/null | grep '^{' > gpurun_out/r2_32_lines_n1.json; python -c "
import json; d=json.loads(open('gpurun_out/r2_32_lines_n1.json').readline()); print('N 1 value',round(d['value']),'e2e',round(d['e2e']['value']))"
run 8 --workload pages; cp gpurun_out/r2_32_tmp.json gpurun_out/r2_32_pages_n8.json
run 8 --method accurate; cp gpurun_out/r2_32_tmp.json gpurun_out/r2_32_accurate_n8.json
tail -3 gpurun_out/r2_32_n8.err
