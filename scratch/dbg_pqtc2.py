import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc
from vearch_b200 import index as gidx
L2, IP = orc.METRIC_L2, orc.METRIC_IP
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
    print("list lens", np.diff(off))
    pos_of = {}
    for l in range(nlist):
        for p, v in enumerate(ids[off[l]:off[l+1]]):
            pos_of[int(v)] = (l, p)
    cd, keys = orc.coarse_search(cent, xq, nprobe, metric)
    os.environ["GB_PQTC"] = "0"
    de, ie = idx.search_preassigned(xq, 10, keys, cd)
    os.environ["GB_PQTC"] = "1"
    for rep in range(4):
        dg, ig = idx.search_preassigned(xq, 10, keys, cd)
        bad = np.where((dg != de).any(1))[0]
        miss = {}
        for q in bad:
            for v in set(ie[q]) - set(ig[q]):
                l, p = pos_of[int(v)]
                miss.setdefault((int(v), l, p, p // 128, p % 128, len(ids[off[l]:off[l+1]])), []).append(int(q))
        print(f"metric={metric} d={d} M={M} rep={rep}: bad rows {len(bad)}; missed (vid,list,pos,tile,e,len)->queries:", miss, flush=True)
    idx.close()
run(IP, 96, 12)
run(IP, 64, 8)
run(IP, 128, 8)
