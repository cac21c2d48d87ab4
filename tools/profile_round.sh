#!/bin/bash
# Round profile run on the GPU box: tests, bench line, launch list, ncu captures, config runs.  usage: tools/profile_round.sh r2
R=${1:-r2}
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/${R}_gputest_final.log
python bench.py --steps 20 --warmup 5 > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference_arm.json 2>/dev/null
# launch list of the bench command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-prrn --no-scaleout > /dev/null 2>&1
# full captures: the headline fill kernel, the group kernels, the striped long-pair kernel
ncu --set full --clock-control none --import-source on -k regex:k1p_score -s 2 -c 1 -o $O/${R}_k1p python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-prrn --no-scaleout --no-groups > /dev/null 2>&1
python tools/ncu_summary.py $O/${R}_k1p.ncu-rep $O/${R}_k1p_score_ncu_full_summary.csv
ncu --set full --clock-control none --import-source on -k "regex:k4_contract|k3_fill" -c 3 -o $O/${R}_k3 python tools/bench_groups.py --pairs 24 --replicate 16 --steps 1 > /dev/null 2>&1
python tools/ncu_summary.py $O/${R}_k3.ncu-rep $O/${R}_k4_k3_ncu_full_summary.csv
ncu --set full --clock-control none --import-source on -k regex:k2_fill_long -c 1 -o $O/${R}_k2long python tools/run_configs.py c5b > /dev/null 2>&1
python tools/ncu_summary.py $O/${R}_k2long.ncu-rep $O/${R}_k2_long_ncu_full_summary.csv
python tools/bench_k1f.py > $O/${R}_k1f_bench_c2_pam.jsonl 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k1f_score_kernel<float" -c 1 -o $O/${R}_k1f python tools/bench_k1f.py 300 > /dev/null 2>&1
python tools/ncu_summary.py $O/${R}_k1f.ncu-rep $O/${R}_k1f_ncu_full_summary.csv
python tools/run_configs.py c5a c5b c4 > $O/${R}_configs_c4_c5a_c5b.jsonl 2>&1
python tools/run_prrn.py --arm both small mid c3t4 c4n20 c4n50 c4n60 > $O/${R}_prrn_end_to_end.jsonl 2>&1
python tools/run_prrn.py --arm gpu c4n100 >> $O/${R}_prrn_end_to_end.jsonl 2>&1
rm -f $O/${R}_k1f.ncu-rep $O/${R}_k1p.ncu-rep $O/${R}_k3.ncu-rep $O/${R}_k2long.ncu-rep
