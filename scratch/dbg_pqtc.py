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
    cd, keys = orc.coarse_search(cent, xq, nprobe, metric)
    for kk in (10, 400):
        for eps in ("1", "8", "1000"):
            os.environ["GB_PQTC"] = "1"; os.environ["GB_PQTC_EPS"] = eps; os.environ["GB_PQTC_STATS"] = "1"
            dg, ig = idx.search_preassigned(xq, kk, keys, cd)
            kname = idx.last_scan_kernel
            os.environ["GB_PQTC"] = "0"
            de, ie = idx.search_preassigned(xq, kk, keys, cd)
            bad = np.where((dg != de).any(1))[0]
            print(f"metric={metric} d={d} M={M} kk={kk} eps={eps} kernel={kname}/{idx.last_scan_kernel}: rows differing {len(bad)}", flush=True)
            for q in bad[:3]:
                j = np.where(dg[q] != de[q])[0][0]
                print("   q", q, "first diff rank", j, "tc", dg[q, j:j+3], ig[q, j:j+3], "exact", de[q, j:j+3], ie[q, j:j+3], "keys", keys[q], flush=True)
    idx.close()
for metric in (IP, L2):
    for d, M in ((96, 12), (128, 16)):
        run(metric, d, M)
