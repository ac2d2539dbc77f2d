#!/bin/bash
# round 2, call 20: encoder tail: LayerNorm affines folded into the weights, hidden accumulator released before the GELU arithmetic
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "encoder_block" > gpurun_out/r2_20_k.log 2>&1; echo "== kernel tests rc=$?"; tail -3 gpurun_out/r2_20_k.log
timeout 200 python tools/eb_timing.py 2>&1 | grep -v sub-phases | cut -c1-420
timeout 600 python bench.py > gpurun_out/r2_20_bench_fast.json 2> gpurun_out/r2_20_bench_fast.err; echo "== bench rc=$?"; tail -3 gpurun_out/r2_20_bench_fast.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r2_20_bench_fast.json') if l.startswith('{')][0]
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'],'roof',round(d['roofline']['frac'],3),'other',round(d['other_method']['value']))
print({k:round(v['ms_per_step'],4) for k,v in d['stages'].items()})
PY
timeout 1500 python -m pytest tests/test_engine_gpu.py tests/test_wide_gpu.py tests/test_baseline_gpu.py -m gpu -q -x > gpurun_out/r2_20_pytest.log 2>&1; echo "== pytest rc=$?"; tail -6 gpurun_out/r2_20_pytest.log
touch kiri-ocr_b200/csrc/encoder_block.cu
make -C kiri-ocr_b200/csrc EXTRA=-DKIRI_EB_SUBPHASE > gpurun_out/r2_20_make.log 2>&1; echo "make rc=$?"
timeout 200 python tools/eb_timing.py > gpurun_out/r2_20_eb_subphase.txt 2>&1; echo rc=$?; grep -A2 "M=26080" gpurun_out/r2_20_eb_subphase.txt | cut -c1-420
