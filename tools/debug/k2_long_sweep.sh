cd /root/repo
timeout 300 python -m pytest tests/test_gpu_align.py -x -q 2>&1 | tail -2
for r in 16 8 4; do echo "rows $r"; PG_K2_LONG_ROWS=$r timeout 120 python tools/run_configs.py c5b 2>&1 | cut -c1-260; done
