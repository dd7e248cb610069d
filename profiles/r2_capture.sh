# Round-2 measurement runs (B200 boxes from gpurun; raw artefacts land in gpurun_out/, profiles/summarize.py condenses them).
# Every GPU command ran under its own `timeout`, after a canary (one pqtc parity test) so that a hung kernel cannot eat the call.
#   1 GPU:
mkdir -p gpurun_out
timeout 400 python bench.py                 | tail -1 > gpurun_out/r2_bench_ivfpq_10m.json          # default line (+ C1/C2/C4/100M secondary)
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | tail -1 > gpurun_out/r2_bench_reference_ivfpq_10m.json
timeout 200 python bench.py --dataset hard --nprobe 32 --steps 3 --warmup 2 --no-secondary --no-cpu-baseline | tail -1 > gpurun_out/r2_bench_ivfpq_10m_hard.json
GB_PQTC=0 timeout 200 python bench.py --dataset hard --nprobe 32 --steps 3 --warmup 2 --no-secondary --no-cpu-baseline | tail -1 > gpurun_out/r2_bench_ivfpq_10m_hard_lut_only.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_ivfpq_10m.csv python bench.py --steps 2 --warmup 1 --profile --no-cpu-baseline --no-secondary
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:pqtc_scan -c 1 -f -o gpurun_out/r2_prof_pqtc_scan python bench.py --steps 1 --warmup 1 --profile --no-cpu-baseline --no-secondary
timeout 300 python bench.py --workload ivfpq_100m --sweep 32:400,32:1000,32:2000,64:1000 --steps 3 --warmup 1 --no-cpu-baseline 2> gpurun_out/r2_sweep_ivfpq_100m.jsonl
#   N GPUs (gpurun --gpus N):
#   TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511"
#   $TR bench.py --gpus N --steps 5 --warmup 3                      -> r2_bench_{2,8}gpu_ivfpq_10m.json (weak: 10M per GPU; secondary: 100M split over N)
#   $TR bench.py --gpus N --steps 2 --warmup 1 --impl reference     -> r2_bench_reference_{2,8}gpu_ivfpq_10m.json
#   $TR bench.py --gpus 2 --workload ivfflat_768 --steps 3 --warmup 3 -> r2_bench_2gpu_ivfflat_768.json
# then:  python profiles/summarize.py full gpurun_out/r2_prof_pqtc_scan.ncu-rep > profiles/r2_ncu_pqtc_scan.txt
#        python profiles/summarize.py launches gpurun_out/r2_launches_ivfpq_10m.csv > profiles/r2_launches_ivfpq_10m.txt
#        python profiles/summarize.py ncu_json profiles/r2_ncu.json
