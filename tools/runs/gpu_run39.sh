#!/bin/bash
# round-1 evidence run of the current build: tests, smoke, bench lines, ncu launch list + full-set capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo "== pytest rc=$?"; tail -1 gpurun_out/pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$?"; tail -1 gpurun_out/smoke.log
for cfg in "fast bucketed" "fast parity" "accurate bucketed"; do set -- $cfg
timeout 600 python bench.py --steps 40 --warmup 3 --method $1 --width-mode $2 > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err; echo "== bench $cfg rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$1_$2.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'],'other',round(d['other_method']['value']), 'roof', d['roofline']['kernel'][:30], round(d['roofline']['frac'],3), 'whole', round(d['whole_step_tensor_frac'],3))
print({k:round(v['ms_per_step'],3) for k,v in d['stages'].items()})
PY
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "== reference rc=$?"; cut -c1-200 gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 230 -c 100 --csv --log-file gpurun_out/launches_fast_v8.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"encoder_block_kernel|gemm_tc_kernel|encoder_attention_kernel|conv1_pair_kernel|preprocess_pack_kernel|ctc_greedy|pool_pos_ln" -s 23 -c 23 -f -o gpurun_out/ncu_step_v8 \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/*.ncu-rep
