#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$?"; tail -3 gpurun_out/smoke.log
CMD="python bench.py --steps 2 --warmup 3 --width-mode parity"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 260 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "== launch list rc=$?"
$CMD > gpurun_out/plain2.log 2> gpurun_out/plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 65 -c 3 -o gpurun_out/prof_conv_r1 $CMD > gpurun_out/ncu2.log 2>&1
echo "== full conv rc=$?"
$CMD > gpurun_out/plain3.log 2> gpurun_out/plain3.err && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 113 -c 4 -o gpurun_out/prof_gemm_r1 $CMD > gpurun_out/ncu3.log 2>&1
echo "== full gemm rc=$?"
$CMD > gpurun_out/plain4.log 2> gpurun_out/plain4.err && \
ncu --set full --clock-control none --import-source on -k regex:"conv1|attention|preprocess|ctc_greedy" -s 24 -c 4 -o gpurun_out/prof_misc_r1 $CMD > gpurun_out/ncu4.log 2>&1
echo "== full misc rc=$?"
ls -la gpurun_out/
