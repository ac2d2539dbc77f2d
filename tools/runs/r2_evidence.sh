#!/bin/bash
# round-2 evidence run of the current build: tests, smoke, bench lines (fast / accurate / beam / pages / reference arm),
# ncu launch lists and full-set captures.  Usage: bash tools/runs/r2_evidence.sh <tag>   (files: gpurun_out/r02_<tag>_*)
T=${1:-a}
O=gpurun_out/r02_$T
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > ${O}_pytest_all.log 2>&1; echo "== pytest rc=$?"; tail -1 ${O}_pytest_all.log
cp gpurun_out/parity_report.json ${O}_parity_report.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "== smoke rc=$?"; tail -1 ${O}_smoke.log
summ() { python - "$1" <<'PY'
import json,sys
d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')][0]
print('value',round(d['value']),d['unit'],'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d.get('gpu_launches'),'roof',d['roofline'].get('kernel','')[:28],round(d['roofline']['frac'],3),'clk',d.get('clocks',{}).get('sm_mhz'))
if 'stages' in d: print({k:round(v['ms_per_step'],4) for k,v in d['stages'].items()})
PY
}
timeout 600 python bench.py > ${O}_bench_fast.json 2> ${O}_bench_fast.err; echo "== bench fast rc=$?"; summ ${O}_bench_fast.json
timeout 600 python bench.py --method accurate > ${O}_bench_accurate.json 2> ${O}_bench_accurate.err; echo "== bench accurate rc=$?"; summ ${O}_bench_accurate.json
timeout 600 python bench.py --method beam > ${O}_bench_beam.json 2> ${O}_bench_beam.err; echo "== bench beam rc=$?"; summ ${O}_bench_beam.json
timeout 900 python bench.py --workload pages > ${O}_bench_pages.json 2> ${O}_bench_pages.err; echo "== bench pages rc=$?"; summ ${O}_bench_pages.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_reference.json 2> ${O}_bench_reference.err; echo "== reference rc=$?"; cut -c1-300 ${O}_bench_reference.json
K="encoder_block_kernel|gemm_tc_kernel|conv2_swap_kernel|encoder_attention_kernel|conv1_pair_kernel|conv1_tc_kernel|preprocess_pack_kernel|crop_sum_kernel|ctc_collapse|pool_pos_ln|ln_chain|pack_records|dec_fused|crosskv"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file ${O}_launches_fast.csv \
    python bench.py --steps 2 --warmup 3 > ${O}_ncu_launch_fast.log 2>&1; echo "launch list fast rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file ${O}_launches_accurate.csv \
    python bench.py --steps 2 --warmup 3 --method accurate > ${O}_ncu_launch_acc.log 2>&1; echo "launch list accurate rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 24 -c 24 -f -o ${O}_ncu_step_fast \
    python bench.py --steps 2 --warmup 3 > ${O}_ncu_full_fast.log 2>&1; echo "full fast rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"dec_fused|crosskv" -s 2 -c 2 -f -o ${O}_ncu_dec \
    python bench.py --steps 2 --warmup 3 --method accurate > ${O}_ncu_full_acc.log 2>&1; echo "full accurate rc=$?"
ls -la gpurun_out/*${T}*.ncu-rep
