import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc
L2, IP = orc.METRIC_L2, orc.METRIC_IP
metric, d, M = IP, 96, 12
n, nq, nlist, nprobe = 40000, 900, 16, 6
rng = np.random.default_rng(197)
centers = rng.normal(0, 1, (64, d)).astype(np.float32)
db = (centers[rng.integers(0, 64, n)] + 0.35 * rng.normal(0, 1, (n, d))).astype(np.float32)
xq = (centers[rng.integers(0, 64, nq)] + 0.35 * rng.normal(0, 1, (nq, d))).astype(np.float32)
cent, _, _ = orc.kmeans(db[:4000], nlist, niter=5)
a = orc.assign(cent, db, metric)
pqc = orc.pq_train(db[:6000] - cent[a[:6000]], M, niter=4)
off, order = orc.build_lists(a, nlist)
codes = orc.ivfpq_encode(cent, pqc, db, a)
cd, keys = orc.coarse_search(cent, xq, nprobe, metric)
def bf16(x):
    x = np.asarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint32) << 16
    return r.view(np.float32)
dsub = d // M
pq3 = pqc.reshape(M, 256, dsub)
R2 = (pq3 ** 2).sum(-1).max(1).sum(); R = np.sqrt(R2)
print("Rmax", R)
def analyse(q, vid, kk=10):
    l = a[vid]; p = int(np.where(keys[q] == l)[0][0])
    code = codes[vid]
    r = np.concatenate([pq3[m, code[m]] for m in range(M)])
    x = xq[q]
    dis0 = cd[q, p]
    # reference fp32 value
    ip = np.array([np.float32(sum(np.float32(x[m*dsub+j]) * np.float32(pq3[m, code[m], j]) for j in range(dsub))) for m in range(M)], np.float32)
    s = np.float32(dis0)
    for m in range(M): s = np.float32(s + ip[m])
    # approx
    acc = -np.dot(bf16(x).astype(np.float64), bf16(-r).astype(np.float64) * -1.0)
    st = dis0 - acc
    # bound from probe 0
    l0 = keys[q, 0]
    members = order[off[l0]:off[l0+1]]
    rr = np.stack([np.concatenate([pq3[m, codes[v][m]] for m in range(M)]) for v in members])
    s0 = cd[q, 0] + rr @ x
    B = np.sort(s0)[::-1][kk-1]
    na = np.linalg.norm(x)
    Z = abs(dis0) + na * R
    eps = 1.05 * 2**-8 * na * R + 2**-17 * Z
    print(f"q {q} vid {vid} list {l} probe {p}: S_ref {s:.6f} S_approx {st:.6f} err {st - s:+.6f} | B {B:.6f} eps {eps:.6f} -> pass iff S~ >= B-eps: {st >= B - eps} (margin {st - (B - eps):+.5f}); |x| {na:.3f} |r| {np.linalg.norm(r):.3f} sum|x r| {np.abs(x*r).sum():.3f}")
analyse(497, 28021)
for q, v in [(355, 25939), (432, 25939), (641, 25939), (68, 4328), (623, 4328), (122, 29871), (613, 29871), (838, 27515)]:
    analyse(q, v)
