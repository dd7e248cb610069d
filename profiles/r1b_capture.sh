mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ivf_listmajor_tc|seg_select|dist_tc|lm_" -c 8 -f -o gpurun_out/r1b_prof_ivfflat_listmajor python bench.py --workload ivfflat_1m --steps 1 --warmup 1 --profile --no-cpu-baseline > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ivfpq_scan|rerank" -c 2 -f -o gpurun_out/r1b_prof_ivfpq_scan python bench.py --steps 1 --warmup 1 --profile --no-cpu-baseline > /dev/null 2>&1
for w in ivfflat_1m ivfpq_10m; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r1b_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --profile --no-cpu-baseline > /dev/null 2>&1
done
ls -la gpurun_out | grep r1b_
