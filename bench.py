#!/usr/bin/env python
"""bench.py -- queries/s of the gamma vector-search hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of nq synthetic queries: coarse quantiser,
inverted-list scan (IVF-Flat) or LUT + ADC scan (+ exact re-rank) (IVF-PQ), top-k merge.
Default workload = the configuration the metric is quoted on: IVF-PQ d=128 M=16 nbits=8
nlist=4096 N=10M nq=10k (BASELINE.json configs[2]); other configs via --workload.

  value  queries/s with queries and results resident in HBM (CUDA events, max over ranks)
  e2e    the same through the public call with HOST buffers: pinned-host queries -> H2D ->
         search -> [NCCL all-gather + merge] -> D2H results, every step
  roofline      dominant scan kernel: algorithmic bytes / CUDA-event kernel time vs measured HBM peak
  cpu_baseline  the CPU oracle (restatement of gamma's CPU path; faiss is not buildable here)
                on a bounded query sample, same index state, all host threads; rank 0, N=1 only

Multi-GPU (N > 1): one partition of N_per_gpu vectors per rank (Vearch partitions, weak scaling),
every query goes to every partition, per-rank top-k are all-gathered over NCCL and merged on
device in the router's order (internal/client/client.go:1530-1609).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (index type, d, N per GPU, nlist, default nprobe, M, metric, data family)
    "flat_100k": dict(type="FLAT", d=128, n=100_000, nlist=0, nprobe=0, M=0, metric="L2", data="sift", nq=1000,
                      desc="FLAT brute-force L2 d=128 N=100k nq=1k (BASELINE configs[0])"),
    "ivfflat_1m": dict(type="IVFFLAT", d=128, n=1_000_000, nlist=1024, nprobe=32, M=0, metric="L2", data="sift",
                       nq=10_000, desc="IVF-Flat d=128 nlist=1024 nprobe=32 N=1M (BASELINE configs[1])"),
    "ivfpq_10m": dict(type="IVFPQ", d=128, n=10_000_000, nlist=4096, nprobe=32, M=16, metric="L2", data="sift",
                      nq=10_000, desc="IVF-PQ d=128 m=16 nbits=8 nlist=4096 N=10M (BASELINE configs[2])"),
    "ivfflat_768": dict(type="IVFFLAT", d=768, n=10_000_000, nlist=4096, nprobe=32, M=0, metric="InnerProduct",
                        data="embed", nq=10_000,
                        desc="IVF-Flat d=768 cosine nlist=4096 N=10M nq=10k (BASELINE configs[3])"),
    "ivfpq_12m": dict(type="IVFPQ", d=128, n=12_500_000, nlist=4096, nprobe=32, M=16, metric="L2", data="sift",
                      nq=10_000, desc="IVF-PQ d=128 N=100M over 8 partitions (BASELINE configs[4]), 12.5M per GPU"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ivfpq_10m", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="override vectors per GPU (parity/dev runs)")
    ap.add_argument("--nq", type=int, default=0)
    ap.add_argument("--nprobe", type=int, default=0)
    ap.add_argument("--recall-num", type=int, default=-1, help="IVF-PQ exact re-rank depth (0 = off)")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="cudaProfilerStart/Stop around the timed device steps (ncu --profile-from-start off)")
    ap.add_argument("--sweep", default="", help="nprobe:recall_num,... -> recall/QPS table on stderr, then exit")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.uuid = uuid
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.uuid, f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nme, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(np.max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak(kind="hbm"):
    """HBM GB/s, or dense bf16 TFLOP/s (the sustained figure: the kernel is timed inside a step)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            if kind == "hbm":
                return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
            return float(j["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
        except Exception:
            pass
    if kind == "hbm":
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    return 2250.0, "fallback (nominal dense bf16 2.25 PFLOP/s)"


def ncu_traffic(workload, kernel):
    """dram bytes per launch of `kernel` on `workload` from the committed ncu --set full capture, or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(workload, {}).get(kernel)
    except Exception:
        return None


def recall_stats(ids, gt):
    """(reference definition: true 1-NN in the first 10; standard 10-recall@10)"""
    k = min(10, ids.shape[1])
    r1 = float(np.mean([(gt[q, 0] in ids[q, :k]) for q in range(ids.shape[0])]))
    rk = float(np.mean([len(set(ids[q, :k]) & set(gt[q, :k])) / k for q in range(ids.shape[0])]))
    return r1, rk


# ------------------------------------------------------------------------------------------------
def build_index(args, wl, rank, device):
    import torch
    from vearch_b200 import index as gidx, synth
    n = args.n or wl["n"]
    d = wl["d"]
    params = {"metric_type": wl["metric"]}
    if wl["type"] != "FLAT":
        params.update(ncentroids=wl["nlist"], nprobe=min(wl["nprobe"], wl["nlist"]))
        if n < wl["nlist"] * 39:  # dev-size runs
            params["ncentroids"] = max(16, n // 100)
            params["nprobe"] = min(params["nprobe"], params["ncentroids"])
        params["training_threshold"] = min(n, params["ncentroids"] * 200)
    if wl["type"] == "IVFPQ":
        params.update(nsubvector=wl["M"], nbits_per_idx=8)
    idx = gidx.GammaIndex(wl["type"], d, params, device=device)
    t0 = time.time()
    gen = synth.sift_like_torch if wl["data"] == "sift" else synth.embed_like_torch
    chunk = 1 << 20
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x = gen(e - s, d, seed=1234 + 1000 * rank + (s >> 20), device=f"cuda:{device}")
        idx.add_vectors(x)
        del x
    torch.cuda.synchronize()
    t1 = time.time()
    idx.train()
    t2 = time.time()
    idx.add_pending()
    t3 = time.time()
    build = {"gen_add_s": round(t1 - t0, 2), "train_s": round(t2 - t1, 2), "index_s": round(t3 - t2, 2)}
    return idx, params, build


def algorithmic_bytes(idx, wl, params, xq_host, nprobe, k, recall_num):
    """per scanned entry (4d+8) B for IVF-Flat, (M+8) B for IVF-PQ, plus the
    precomputed-table row M*256*4 B per (query, list) and the re-rank gathers; FLAT: N*4d per query."""
    d = wl["d"]
    nq = xq_host.shape[0]
    if wl["type"] == "FLAT":
        return float(nq) * idx.ntotal * 4 * d, {"entries": float(nq) * idx.ntotal}
    lens = np.array([idx.list_len(l) for l in range(idx.nlist)], np.int64)
    _, keys = idx.coarse_search(xq_host, nprobe)
    scanned = float(lens[keys[keys >= 0]].sum())
    pairs = float((keys >= 0).sum())
    if wl["type"] == "IVFFLAT":
        return scanned * (4 * d + 8), {"entries": scanned}
    b = scanned * (wl["M"] + 8)
    if wl["metric"] == "L2":
        b += pairs * wl["M"] * 256 * 4
    if recall_num > 0:
        b += nq * max(k, recall_num) * 4 * d
    return b, {"entries": scanned, "query_list_pairs": pairs}


def export_state(idx, wl):
    st = {}
    if wl["type"] == "FLAT":
        return st
    st["cent"] = idx.get_centroids()
    st["off"], st["codes"], st["ids"] = idx.export_lists()
    if wl["type"] == "IVFPQ":
        st["pq"] = idx.get_pq_centroids()
        st["T"] = idx.get_precomputed_table() if wl["metric"] == "L2" else None
    return st


def cpu_search(orc, wl, st, raw, xq, k, nprobe, recall_num):
    """The CPU restatement of gamma's search path (oracle), one call = coarse + scan (+ re-rank)."""
    metric = orc.METRIC_L2 if wl["metric"] == "L2" else orc.METRIC_IP
    if wl["type"] == "FLAT":
        return orc.flat_search(raw, xq, k, metric)
    cd, keys = orc.coarse_search(st["cent"], xq, nprobe, metric)
    if wl["type"] == "IVFFLAT":
        vecs = st["codes"].view(np.float32).reshape(len(st["ids"]), -1)
        return orc.ivfflat_search_preassigned(st["off"], vecs, st["ids"], xq, k, keys, metric)
    return orc.ivfpq_search_preassigned(st["off"], st["codes"], st["ids"], st["cent"], st["pq"], st["T"], xq, k, keys,
                                        cd, metric, recall_num=max(recall_num, 0), raw=raw)


def main():
    args = parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus and world > 1:
        args.gpus = world
    if args.impl == "reference" and rank != 0:
        return 0  # the CPU arm runs on rank 0 alone

    import torch
    import torch.distributed as dist
    from vearch_b200 import _lib, index as gidx, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this engine has no CPU path")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    use_dist = world > 1 and args.impl == "ours"
    if use_dist:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))

    nq = args.nq or wl["nq"]
    k = args.k
    nprobe = args.nprobe or wl["nprobe"]
    recall_num = args.recall_num if args.recall_num >= 0 else (400 if wl["type"] == "IVFPQ" else 0)
    idx, params, build = build_index(args, wl, rank, local)
    if wl["type"] != "FLAT":
        nprobe = min(nprobe, params["ncentroids"])
    sp = {}
    if wl["type"] != "FLAT":
        sp["nprobe"] = nprobe
    if wl["type"] == "IVFPQ" and recall_num > 0:
        sp["recall_num"] = recall_num
    metric_id = gidx.METRIC_L2 if wl["metric"] == "L2" else gidx.METRIC_IP

    gen = synth.sift_like_torch if wl["data"] == "sift" else synth.embed_like_torch
    nbatches = args.warmup + args.steps
    # a distinct query batch per step (same on every rank), held in pinned host memory for e2e
    q_host = [torch.empty((nq, wl["d"]), dtype=torch.float32).pin_memory() for _ in range(nbatches)]
    q_dev = []
    for b in range(nbatches):
        x = gen(nq, wl["d"], seed=4321 + b, device=dev)
        q_dev.append(x)
        q_host[b].copy_(x)
    torch.cuda.synchronize()

    # ---- recall of this configuration against exact ground truth (untimed) --------------------
    ns = min(1000, nq)
    xs = q_host[0][:ns].numpy()
    gt_d, gt_i = idx.search(xs, 10, brute_force=True)
    if args.sweep:
        for item in args.sweep.split(","):
            npb, rn = (int(v) for v in item.split(":"))
            p = {"nprobe": npb}
            if rn > 0:
                p["recall_num"] = rn
            _, ri_ = idx.search(xs, 10, params=p)
            a1, a10 = recall_stats(ri_, gt_i)
            idx.search_device(q_dev[0], k, params=p)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for b in range(1, min(4, nbatches)):
                idx.search_device(q_dev[b], k, params=p)
            torch.cuda.synchronize()
            qps = nq * (min(4, nbatches) - 1) / (time.perf_counter() - t0)
            print(json.dumps({"sweep": {"nprobe": npb, "recall_num": rn, "recall_1nn_top10": a1, "recall10": a10,
                                        "qps_device": qps}}), file=sys.stderr, flush=True)
        return 0
    rd, ri = idx.search(xs, 10, params=sp or None)
    r1, r10 = recall_stats(ri, gt_i)

    # =============================================================================================
    if args.impl == "reference":
        from oracle import oracle as orc
        st = export_state(idx, wl)
        need_raw = wl["type"] == "FLAT" or recall_num > 0
        # the CPU re-rank touches arbitrary vids: keep a full host copy when it is needed
        raw = fetch_raw(idx, wl) if need_raw else None
        # calibrate the sample so that one step is ~8 s of CPU work
        xq0 = q_host[0].numpy()
        t0 = time.time()
        cpu_search(orc, wl, st, raw, xq0[:64], k, nprobe, recall_num)
        per_q = (time.time() - t0) / 64
        nsamp = int(max(64, min(nq, 8.0 / max(per_q, 1e-6))))
        times = []
        for b in range(nbatches):
            xq = q_host[b].numpy()[:nsamp]
            t0 = time.time()
            cpu_search(orc, wl, st, raw, xq, k, nprobe, recall_num)
            dt = time.time() - t0
            if b >= args.warmup:
                times.append(dt)
        ms = 1000 * float(np.mean(times))
        qps = nsamp / (ms / 1000)
        line = {"impl": "reference", "metric": "queries/sec @ recall@10 (d=%d, N=%d, nq=%d)" % (wl["d"], idx.ntotal, nq),
                "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": wl["desc"], "nprobe": nprobe, "k": k,
                                                "recall_num": recall_num, "recall@10_1nn": r1},
                "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": orc.num_threads(), "kind": "port",
                                 "sample": f"{nsamp} of {nq} queries per step, index state built once on the GPU and "
                                           "exported (BASELINE.md section 3)"},
                "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # =============================================================================================
    # ours
    out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    gather_d = gather_i = None
    if use_dist:
        gather_d = torch.empty((world, nq, k), dtype=torch.float32, device=dev)
        gather_i = torch.empty((world, nq, k), dtype=torch.int64, device=dev)

    def step_device(b):
        idx.search_device(q_dev[b], k, params=sp or None, out=(out_d, out_i))
        if use_dist:
            dist.all_gather_into_tensor(gather_d, out_d)
            dist.all_gather_into_tensor(gather_i, out_i)
            return gidx.merge_partitions_device(gather_d, gather_i, metric_id)
        return out_d, out_i

    def step_e2e(b):
        xq = q_host[b].to(dev, non_blocking=True)  # H2D from pinned memory
        idx.search_device(xq, k, params=sp or None, out=(out_d, out_i))
        if use_dist:
            dist.all_gather_into_tensor(gather_d, out_d)
            dist.all_gather_into_tensor(gather_i, out_i)
            rd_, ri_ = gidx.merge_partitions_device(gather_d, gather_i, metric_id)
        else:
            rd_, ri_ = out_d, out_i
        return rd_.cpu(), ri_.cpu()  # D2H of the step's result

    def sync_all():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
            torch.cuda.synchronize()

    uuid = str(torch.cuda.get_device_properties(local).uuid)
    uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
    sampler = ClockSampler(uuid)

    # ---- device-resident timing ----------------------------------------------------------------
    # nvidia-smi needs ~1 s before its first sample: started ahead of the warm-up and stopped after the
    # end-to-end loop, so the samples cover warm-up + both timed regions (same load throughout)
    sampler.start()
    for b in range(args.warmup):
        step_device(b)
    sync_all()
    idx.set_scan_timing(True)
    _ = idx.last_scan_ms
    launches0 = _lib.lib().gb_launch_count()
    step_ms, scan_ms = [], []
    if args.profile:
        torch.cuda.profiler.start()
    for s in range(args.steps):
        b = args.warmup + s
        flush.fill_(s)  # L2 flush between timed iterations (untimed)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_device(b)
        e1.record()
        sync_all()
        step_ms.append(e0.elapsed_time(e1))
        scan_ms.append(idx.last_scan_ms)
    if args.profile:
        torch.cuda.profiler.stop()
    launches = (_lib.lib().gb_launch_count() - launches0) / max(1, args.steps)
    idx.set_scan_timing(False)
    total_ms = float(np.sum(step_ms))
    if use_dist:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    units = nq * world  # partition-queries processed by all ranks per step (weak scaling)
    value = units / (ms_per_step / 1000)

    # ---- end-to-end timing (host buffers, copies inside the timed region) ---------------------
    for b in range(args.warmup):
        step_e2e(b)
    sync_all()
    e2e_ms = []
    for s in range(args.steps):
        b = args.warmup + s
        flush.fill_(s)
        sync_all()
        t0 = time.perf_counter()
        step_e2e(b)
        torch.cuda.synchronize()
        e2e_ms.append((time.perf_counter() - t0) * 1000)
    e2e_total = float(np.sum(e2e_ms))
    if use_dist:
        t = torch.tensor([e2e_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_total = float(t.item())
    e2e_value = units / (e2e_total / args.steps / 1000)
    clocks = sampler.stop()
    clocks["window"] = "warm-up + timed device steps + timed end-to-end steps"

    # C-ABI host call (gb_index_search: pageable/pinned host in, host out) for N=1
    cabi_qps = None
    if world == 1:
        for b in range(args.warmup):
            idx.search(q_host[b].numpy(), k, params=sp or None)
        cabi_s = 0.0
        for s in range(args.steps):
            flush.fill_(s)  # same L2 flush + full sync between timed calls as the device-resident loop
            torch.cuda.synchronize()
            xh = q_host[args.warmup + s].numpy()
            t0 = time.perf_counter()
            idx.search(xh, k, params=sp or None)  # returns after the D2H of the results
            cabi_s += time.perf_counter() - t0
        cabi_qps = nq * args.steps / cabi_s

    if rank != 0:
        if use_dist:
            dist.barrier()
        return 0

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    xq_host = q_host[args.warmup].numpy()
    abytes, detail = algorithmic_bytes(idx, wl, params, xq_host, nprobe, k, recall_num)
    scan_avg = float(np.mean(scan_ms)) if scan_ms else 0.0
    kname = idx.last_scan_kernel
    if kname.startswith("ivf_listmajor"):
        # list-major grouped GEMM on tcgen05: each list is read once per 128 queries, the kernel is
        # bounded by the tensor pipe.  Algorithmic flops = 2*d per (query, entry) pair (the error-
        # compensated 3xTF32 split issues three MMAs per product: overhead, not credited).
        peak, peak_src = measured_peak("tensor")
        aflops = detail["entries"] * 2.0 * wl["d"]
        achieved = aflops / (scan_avg / 1000) / 1e12 if scan_avg > 0 else None
        roofline = {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": ncu_traffic(args.workload if not (args.n or args.nq or args.nprobe) else "", kname),
                    "peak_source": peak_src, "algorithmic_flops_per_launch": aflops, "mma_kind": "tf32 x3 (kind::tf32 peak is "
                    "half the bf16 figure)", "algorithmic_bytes_per_launch": abytes, "kernel_ms": scan_avg,
                    "kernel_share_of_step": scan_avg / ms_per_step if ms_per_step else None}
    else:
        peak, peak_src = measured_peak("hbm")
        achieved = abytes / (scan_avg / 1000) / 1e9 if scan_avg > 0 else None
        roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": ncu_traffic(args.workload if not (args.n or args.nq or args.nprobe) else "", kname), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": abytes, "kernel_ms": scan_avg,
                    "kernel_share_of_step": scan_avg / ms_per_step if ms_per_step else None}
        if achieved and achieved > peak:
            roofline["note"] = ("algorithmic bytes follow SURVEY 8(d) (every query streams its own entries); the kernel "
                                "serves them from L2 / shares one pass between the queries of a block, so HBM is not its bound")
    roofline.update(detail)

    # ---- CPU baseline on this box's host cores (bounded sample) ---------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        st = export_state(idx, wl)
        raw = fetch_raw(idx, wl) if (wl["type"] == "FLAT" or recall_num > 0) else None
        t0 = time.time()
        cpu_search(orc, wl, st, raw, xq_host[:64], k, nprobe, recall_num)
        per_q = (time.time() - t0) / 64
        nsamp = int(max(64, min(nq, 15.0 / max(per_q, 1e-6))))
        t0 = time.time()
        cd_, ci_ = cpu_search(orc, wl, st, raw, xq_host[:nsamp], k, nprobe, recall_num)
        dt = time.time() - t0
        # parity spot-check of the timed configuration against the checker
        gd_, gi_ = idx.search(xq_host[:nsamp], k, params=sp or None)
        agree = float((gi_ == ci_).mean())
        cpu = {"value": nsamp / dt, "unit": "queries/s", "cores": orc.num_threads(), "kind": "port",
               "sample": f"{nsamp} of {nq} queries, same index state (exported from the GPU build), "
                         f"wall clock {dt:.1f} s; id agreement with the GPU result {agree:.4f}"}

    line = {"metric": "queries/sec @ recall@10 (d=%d, N=%d, nq=%d)" % (wl["d"], idx.ntotal * world, nq),
            "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "index": wl["type"], "d": wl["d"], "n_per_gpu": idx.ntotal,
                       "nq_per_step": nq, "k": k, "nprobe": nprobe, "recall_num": recall_num,
                       "index_params": params, "recall@10_1nn_in_top10": r1, "recall@10_standard": r10,
                       "units": "partition-queries/s summed over ranks (each rank scans its own partition for all "
                                "nq queries); global queries/s on the N-partition database = value / n_gpus",
                       "l2": "512 MiB write between timed steps + distinct query batch per step",
                       "build_seconds": build},
            "clocks": clocks,
            # N=1: the reference-facing C-ABI call gb_index_search(host queries -> host results), H2D/D2H inside;
            # N>1: pinned-host H2D -> search -> NCCL all-gather + merge -> D2H through the Python API
            "e2e": {"value": cabi_qps if cabi_qps else e2e_value, "unit": "queries/s",
                    "h2d_bytes_per_step": nq * wl["d"] * 4, "d2h_bytes_per_step": nq * k * 12,
                    "path": "gb_index_search C-ABI, host buffers" if cabi_qps else
                            "pinned H2D + search_device + all_gather/merge + D2H",
                    "pinned_h2d_search_d2h_qps": e2e_value},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu}
    print(json.dumps(line))
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def fetch_raw(idx, wl):
    """Full host copy of the raw vectors for the CPU re-rank / FLAT legs."""
    return idx.get_vectors(0, idx.ntotal)


if __name__ == "__main__":
    sys.exit(main())
