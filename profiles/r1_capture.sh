# Round-1 measurement run (one B200): benches of every workload (with the CPU baseline leg), the
# reference arm, ncu launch lists of the same commands and one full ncu capture per dominant kernel.
# Raw artefacts land in gpurun_out/; profiles/summarize.py turns them into the committed summaries.
mkdir -p gpurun_out
for w in flat_100k ivfflat_1m ivfpq_10m ivfflat_768; do
  timeout 900 python bench.py --workload $w 2>/dev/null | tail -1 > gpurun_out/r1_bench_$w.json
done
GB_LISTMAJOR=0 timeout 400 python bench.py --workload ivfflat_1m --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r1_bench_ivfflat_1m_querymajor.json
GB_LISTMAJOR=2 timeout 400 python bench.py --workload ivfflat_1m --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r1_bench_ivfflat_1m_densescores.json
GB_TC_MIRROR=0 timeout 400 python bench.py --workload ivfflat_1m --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r1_bench_ivfflat_1m_regstaged.json
for w in ivfflat_1m ivfpq_10m; do
  timeout 600 python bench.py --workload $w --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r1_bench_reference_$w.json
done
for w in ivfflat_1m ivfpq_10m; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r1_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --profile --no-cpu-baseline > /dev/null 2>&1
done
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ivf_listmajor" -c 1 -f -o gpurun_out/r1_prof_ivfflat_listmajor python bench.py --workload ivfflat_1m --steps 1 --warmup 1 --profile --no-cpu-baseline > /dev/null 2>&1
GB_LISTMAJOR=0 timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ivfflat_scan -c 1 -f -o gpurun_out/r1_prof_ivfflat_scan python bench.py --workload ivfflat_1m --steps 1 --warmup 1 --profile --no-cpu-baseline > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ivfpq_scan -c 1 -f -o gpurun_out/r1_prof_ivfpq_scan python bench.py --steps 1 --warmup 1 --profile --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | grep r1_
