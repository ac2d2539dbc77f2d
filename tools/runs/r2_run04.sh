#!/bin/bash
# round 2, call 5 (2 GPUs): NCCL paths — weak-scaling lines with the side-stream exchange, the sharded page pipeline
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2_05_lines_2gpu.json 2> gpurun_out/r2_05_lines_2gpu.err; echo "== lines rc=$?"; tail -3 gpurun_out/r2_05_lines_2gpu.err
timeout 600 $TR bench.py --gpus 2 --method accurate > gpurun_out/r2_05_acc_2gpu.json 2> gpurun_out/r2_05_acc_2gpu.err; echo "== acc rc=$?"; tail -3 gpurun_out/r2_05_acc_2gpu.err
timeout 900 $TR bench.py --gpus 2 --workload pages > gpurun_out/r2_05_pages_2gpu.json 2> gpurun_out/r2_05_pages_2gpu.err; echo "== pages rc=$?"; tail -3 gpurun_out/r2_05_pages_2gpu.err
python - <<PY
import json
for f in ('lines','acc','pages'):
    try:
        d=json.load(open(f'gpurun_out/r2_05_{f}_2gpu.json'))
        print(f,'value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'], d.get('ordered_equal_to_single_gpu'), d.get('identical_on_all_ranks'))
    except Exception as e: print(f,'parse failed',e)
PY
