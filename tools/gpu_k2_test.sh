cd /root/repo
python -m pytest tests/test_gpu_align.py -x -q 2>&1 | tail -3 > gpurun_out/r2_k2long.log
for r in 16 8 4; do PG_K2_LONG_ROWS=$r python tools/run_configs.py c5b >> gpurun_out/r2_k2long.log 2>&1; done
python tools/run_configs.py c5b >> gpurun_out/r2_k2long.log 2>&1
