#!/bin/bash
mkdir -p gpurun_out
for n in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 40 --warmup 3 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err; echo "== ${n}gpu rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${n}gpu.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'n',d['n_gpus'])
PY
done
