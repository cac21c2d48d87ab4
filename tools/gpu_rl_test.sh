cd /root/repo
python -m pytest tests/test_gpu_groups.py -x -q 2>&1 | tail -2 > gpurun_out/r2_rl_test.log
for rep in 16 1; do for rl in 1 0; do
PG_K3_RL=$rl python tools/bench_groups.py --pairs 24 --replicate $rep --steps 3 --cache gpurun_out/gb.pkl 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rl=$rl rep=$rep kernel_ms %.2f call_ms %.2f gcups %.2f e2e %.2f mism %d'%(d['kernel_ms'],d['call_ms'],d['value'],d['e2e']['value'],d['parity_mismatches']))" >> gpurun_out/r2_rl_test.log
done; done
