cd /root/repo
ncu --set full --clock-control none --import-source on -k regex:k3_fill -c 2 -o gpurun_out/r2_k3lat python tools/bench_groups.py --pairs 3 --replicate 1 --steps 1 --length 900 > /dev/null 2>&1
ncu -i gpurun_out/r2_k3lat.ncu-rep --page source --csv > gpurun_out/r2_k3lat_source.csv 2>/dev/null
ncu -i gpurun_out/r2_k3lat.ncu-rep --page raw --csv > gpurun_out/r2_k3lat_raw.csv 2>/dev/null
rm -f gpurun_out/r2_k3lat.ncu-rep
