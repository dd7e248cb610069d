# Round-2 measurement run (one B200).  Raw artefacts land in gpurun_out/; profiles/summarize.py condenses them.
#   bash profiles/r2_capture.sh [tag]
mkdir -p gpurun_out
T=${1:-r2}
timeout 600 python bench.py 2>/dev/null | tail -1 > gpurun_out/${T}_bench_default.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches_ivfpq_10m.csv python bench.py --steps 2 --warmup 1 --profile --no-cpu-baseline --no-secondary > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:pqtc_scan -c 1 -f -o gpurun_out/${T}_prof_pqtc_scan python bench.py --steps 1 --warmup 1 --profile --no-cpu-baseline --no-secondary > /dev/null 2>&1
ls -la gpurun_out | grep ${T}_
