import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc
from vearch_b200 import index as gidx
L2, IP = orc.METRIC_L2, orc.METRIC_IP
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
metric, d, M = IP, 96, 12
n, nq, nlist, nprobe = 40000, 900, 16, 6
rng = np.random.default_rng(197)
centers = rng.normal(0, 1, (64, d)).astype(np.float32)
db = (centers[rng.integers(0, 64, n)] + 0.35 * rng.normal(0, 1, (n, d))).astype(np.float32)
xq = (centers[rng.integers(0, 64, nq)] + 0.35 * rng.normal(0, 1, (nq, d))).astype(np.float32)
cent, _, _ = orc.kmeans(db[:4000], nlist, niter=5)
a = orc.assign(cent, db, metric)
pqc = orc.pq_train(db[:6000] - cent[a[:6000]], M, niter=4)
idx = gidx.GammaIndex("IVFPQ", d, {"ncentroids": nlist, "nprobe": nprobe, "nsubvector": M, "metric_type": "InnerProduct"})
idx.set_centroids(cent); idx.set_pq_centroids(pqc); idx.add_vectors(db); idx.add_pending()
off, codes, ids = idx.export_lists()
cd, keys = orc.coarse_search(cent, xq, nprobe, metric)
dsub = d // M
pq3 = pqc.reshape(M, 256, dsub)
recon = np.concatenate([pq3[m][codes[:, m]] for m in range(M)], axis=1)  # list order
def bf16(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint32) << 16).view(np.float32)
reconb = bf16(recon).astype(np.float64); xb = bf16(xq).astype(np.float64)
R = np.sqrt((pq3 ** 2).sum(-1).max(1).sum())
os.environ["GB_PQTC"] = "1"; os.environ["GB_PQTC_DUMP"] = "/tmp/pqtc_dump.bin"
kk = 10
# expected sure candidates: approx score >= B - eps/2 (well inside the margin)
first = True
for rep in range(reps):
    dg, ig = idx.search_preassigned(xq, kk, keys, cd)
    raw = open("/tmp/pqtc_dump.bin", "rb").read()
    hdr = np.frombuffer(raw[:32], np.int32); nq_, cap, kk_, pa, nitems, npairs, isz, nvalid = [int(v) for v in hdr]
    o = 32
    cnt = np.frombuffer(raw[o:o + 4 * nq_], np.int32); o += 4 * nq_
    cand = np.frombuffer(raw[o:o + 8 * nq_ * cap], np.uint64).reshape(nq_, cap); o += 8 * nq_ * cap
    keysA = np.frombuffer(raw[o:o + 8 * nq_ * kk_], np.uint64).reshape(nq_, kk_); o += 8 * nq_ * kk_
    pair_j = np.frombuffer(raw[o:o + 8 * npairs], np.int64); o += 8 * npairs
    items = np.frombuffer(raw[o:o + isz * nitems], np.int32).reshape(nitems, isz // 4)  # list,pair0,npairs,row0,nrows,seg,grp
    # bound per query from keysA: IP key = ~f2ord(score)
    hi = (keysA[:, kk_ - 1] >> np.uint64(32)).astype(np.uint32)
    x = ~hi
    bits = np.where(x & 0x80000000, x & 0x7FFFFFFF, ~x).astype(np.uint32)
    B = bits.view(np.float32)
    nomiss = 0; missing = []
    pair_slot = {}
    for it in range(nitems):
        l, p0, npr, row0, nrows, seg, grp = items[it][:7]
        for r in range(npr):
            pair_slot[(int(pair_j[p0 + r]), int(seg))] = (it, r)
    for q in range(nq):
        if cnt[q] > cap: continue
        got = set(int(v) for v in cand[q, :cnt[q]])
        na = np.linalg.norm(xq[q])
        for p in range(pa, nprobe):
            l = keys[q, p]
            s = cd[q, p] + reconb[off[l]:off[l+1]] @ xb[q]
            eps = 1.05 * 2**-8 * na * R
            sure = np.where(s >= B[q] - 0.5 * eps)[0]
            for pos in sure:
                if ((p << 32) | int(pos)) not in got:
                    seglen = None
                    missing.append((q, p, int(l), int(pos)))
    print(f"rep {rep}: bad rows {(dg != idx_ref[0]).any(1).sum() if not first else 'n/a'} missing sure-candidates {len(missing)}", flush=True)
    if first:
        os.environ["GB_PQTC"] = "0"
        idx_ref = idx.search_preassigned(xq, kk, keys, cd)
        os.environ["GB_PQTC"] = "1"
        first = False
    for (q, p, l, pos) in missing[:40]:
        # locate item/row
        j = q * nprobe + p
        hit = [(it, r) for (jj, sg), (it, r) in pair_slot.items() if jj == j]
        desc = []
        for it, r in hit:
            li, p0, npr, row0, nrows, seg, grp = items[it][:7]
            if row0 <= pos < row0 + nrows:
                desc.append(f"item {it} grp {grp} seg {seg} row {r} (warp {r//32}) tile {(pos-row0)//128} col {(pos-row0)%128} (c0 {((pos-row0)%128)//32*32})")
        print("   missing q", q, "probe", p, "list", l, "pos", pos, desc, flush=True)
idx.close()
