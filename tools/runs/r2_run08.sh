timeout 900 python -m pytest tests/test_wide_gpu.py -m gpu -q -x -k "live or validation or ocr_class" > gpurun_out/r2_08_live.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r2_08_live.log
