import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc
from vearch_b200 import index as gidx
L2, IP = orc.METRIC_L2, orc.METRIC_IP
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
def run(metric, d, M):
    n, nq, nlist, nprobe = 40000, 900, 16, 6
    rng = np.random.default_rng(197)
    centers = rng.normal(0, 1, (64, d)).astype(np.float32)
    db = (centers[rng.integers(0, 64, n)] + 0.35 * rng.normal(0, 1, (n, d))).astype(np.float32)
    xq = (centers[rng.integers(0, 64, nq)] + 0.35 * rng.normal(0, 1, (nq, d))).astype(np.float32)
    cent, _, _ = orc.kmeans(db[:4000], nlist, niter=5)
    a = orc.assign(cent, db, metric)
    pqc = orc.pq_train(db[:6000] - cent[a[:6000]], M, niter=4)
    idx = gidx.GammaIndex("IVFPQ", d, {"ncentroids": nlist, "nprobe": nprobe, "nsubvector": M,
                                       "metric_type": "L2" if metric == L2 else "InnerProduct"})
    idx.set_centroids(cent); idx.set_pq_centroids(pqc); idx.add_vectors(db); idx.add_pending()
    off, codes, ids = idx.export_lists()
    pos_of = {}
    for l in range(nlist):
        for p, v in enumerate(ids[off[l]:off[l+1]]):
            pos_of[int(v)] = (l, p)
    cd, keys = orc.coarse_search(cent, xq, nprobe, metric)
    os.environ["GB_PQTC"] = "0"
    de, ie = idx.search_preassigned(xq, 10, keys, cd)
    os.environ["GB_PQTC"] = "1"
    os.environ["GB_PQTC_DUMP"] = "/tmp/pqtc_dump.bin"
    for rep in range(reps):
        dg, ig = idx.search_preassigned(xq, 10, keys, cd)
        raw = open("/tmp/pqtc_dump.bin", "rb").read()
        hdr = np.frombuffer(raw[:16], np.int32); nq_, cap, kk, pa = hdr
        cnt = np.frombuffer(raw[16:16 + 4 * nq_], np.int32)
        cand = np.frombuffer(raw[16 + 4 * nq_:16 + 4 * nq_ + 8 * nq_ * cap], np.uint64).reshape(nq_, cap)
        bad = np.where((dg != de).any(1))[0]
        print(f"metric={metric} d={d} M={M} rep={rep}: bad rows {len(bad)}", flush=True)
        for q in bad:
            for v in set(ie[q]) - set(ig[q]):
                l, p = pos_of[int(v)]
                pr = int(np.where(keys[q] == l)[0][0])
                rec = (pr << 32) | p
                c = int(cnt[q])
                found = rec in set(int(x) for x in cand[q, :min(c, cap)])
                print(f"   q {q} vid {v} list {l} probe {pr} pos {p} tile {p//128} e {p%128}: cand_cnt {c} (cap {cap}) in_cand {found}", flush=True)
    idx.close()
run(IP, 96, 12)
