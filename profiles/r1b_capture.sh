mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ivf_listmajor" -c 1 -f -o gpurun_out/r1b_prof_ivfflat_listmajor python bench.py --workload ivfflat_1m --steps 1 --warmup 1 --profile --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | grep r1b_
