"""CPU tests of the oracle (oracle/gamma_oracle.c) against independent numpy/sklearn maths and
against the only numeric pins the reference's own tests hold for this path:
  * L2 score == sum((x-y)^2)                       test/test_module_vector.py:337-364
  * FLAT recall@1 >= 0.95, recall@10 >= 1.0         test/test_vector_index_flat.py:94-96
  * IVFFLAT r@1 >= 0.8, r@10 >= 0.9 (nprobe > 1)    test/test_vector_index_ivfflat.py:89-94
  * IVFPQ r@10 >= 0.9 (nprobe > 10)                 test/test_vector_index_ivfpq.py:105-111
  * self-query top-1 score ~ 1.0, normalised IP     internal/engine/tests/test.h:554-565
Recall definition: fraction of queries whose true 1-NN is in the first i results
(test/utils/vearch_utils.py:1481-1483).
"""
import numpy as np
import pytest

from oracle import oracle as orc
from vearch_b200 import synth

L2, IP = orc.METRIC_L2, orc.METRIC_IP


def brute(db, xq, k, metric):
    db64, xq64 = db.astype(np.float64), xq.astype(np.float64)
    if metric == L2:
        s = (xq64 ** 2).sum(1)[:, None] + (db64 ** 2).sum(1)[None, :] - 2 * xq64 @ db64.T
        order = np.lexsort((np.broadcast_to(np.arange(db.shape[0]), s.shape), s), axis=1)[:, :k]
    else:
        s = xq64 @ db64.T
        order = np.lexsort((np.broadcast_to(np.arange(db.shape[0]), s.shape), -s), axis=1)[:, :k]
    return np.take_along_axis(s, order, 1), order


def recall_1nn_in_topi(ids, gt1, i):
    return float(np.mean([(gt1[q] in ids[q, :i]) for q in range(ids.shape[0])]))


@pytest.fixture(scope="module")
def sift():
    db = synth.sift_like(6000, 32, seed=1)
    xq = synth.sift_like(64, 32, seed=2)
    return db, xq


def test_mt19937_matches_numpy_legacy_stream():
    n, seed = 1000, 1235
    rs = np.random.RandomState(seed)
    perm = np.arange(n, dtype=np.int64)
    for i in range(n - 1):
        r = int(rs.randint(0, 2 ** 32, dtype=np.uint64))
        i2 = i + r % (n - i)
        perm[i], perm[i2] = perm[i2], perm[i]
    assert np.array_equal(orc.rand_perm(n, seed), perm)


@pytest.mark.parametrize("metric", [L2, IP])
def test_flat_exact_on_integer_data(sift, metric):
    db, xq = sift
    k = 10
    dis, ids = orc.flat_search(db, xq, k, metric)
    gd, gi = brute(db, xq, k, metric)
    # integer-valued data: every partial sum is exact => distances bit-equal to fp64 maths
    assert np.array_equal(dis.astype(np.float64), gd)
    if metric == L2:
        assert np.array_equal(ids, gi)  # (dist, id) ascending == faiss CMax order for an id-ordered scan
    else:
        # CMin reorder lists equal scores with the larger id first; compare as sets per score
        for q in range(xq.shape[0]):
            assert sorted(zip(-dis[q], ids[q])) == sorted(zip(-gd[q].astype(np.float32), gi[q]))
    # reference pins
    assert recall_1nn_in_topi(ids, gi[:, 0], 1) >= 0.95 and recall_1nn_in_topi(ids, gi[:, 0], 10) >= 1.0
    if metric == L2:
        manual = ((xq[:, None, :] - db[ids]) ** 2).sum(-1)
        assert np.abs(manual - dis).max() <= 0.01


def test_flat_tie_rule_smaller_id_kept():
    db = np.zeros((8, 4), np.float32)
    db[:, 0] = [3, 1, 1, 1, 2, 1, 0, 1]
    xq = np.zeros((1, 4), np.float32)
    dis, ids = orc.flat_search(db, xq, 3, L2)
    assert ids.tolist() == [[6, 1, 2]] and dis.tolist() == [[0, 1, 1]]
    dis, ids = orc.flat_search(db, xq, 20, L2)  # k > n: tail is (-1, FLT_MAX)
    assert ids[0, :8].tolist() == [6, 1, 2, 3, 5, 7, 4, 0] and (ids[0, 8:] == -1).all()
    assert (dis[0, 8:] == np.finfo(np.float32).max).all()


def test_flat_filters_and_score_window(sift):
    db, xq = sift
    n = db.shape[0]
    rng = np.random.default_rng(0)
    deleted = rng.random(n) < 0.3
    allowed = rng.random(n) < 0.5
    delb = np.packbits(deleted, bitorder="little")
    filb = np.packbits(allowed, bitorder="little")
    dis, ids = orc.flat_search(db, xq, 10, L2, del_bitmap=delb, filter_bitmap=filb)
    keep = np.flatnonzero(~deleted & allowed)
    gd, gi = brute(db[keep], xq, 10, L2)
    assert np.array_equal(ids, keep[gi]) and np.array_equal(dis.astype(np.float64), gd)
    # score window (SearchCondition::IsSimilarScoreValid, gamma_common_data.h:94-96)
    full, _ = orc.flat_search(db, xq, 50, L2)
    lo, hi = float(full[0, 5]), float(full[0, 30])
    dis, ids = orc.flat_search(db, xq[:1], 50, L2, min_score=lo, max_score=hi)
    got = dis[0][ids[0] >= 0]
    assert got.min() >= lo and got.max() <= hi and len(got) == np.sum((full[0] >= lo) & (full[0] <= hi))


def test_self_query_ip_normalised():
    db = synth.embed_like(2000, 64, seed=3)
    dis, ids = orc.flat_search(db, db[:50], 1, IP)
    assert np.array_equal(ids[:, 0], np.arange(50)) and np.abs(dis[:, 0] - 1.0).max() < 1e-5


@pytest.fixture(scope="module")
def ivf(sift):
    db, xq = sift
    nlist = 32
    cent, _, obj = orc.kmeans(db, nlist, niter=10)
    assert obj[-1] <= obj[0]
    a = orc.assign(cent, db, L2)
    off, order = orc.build_lists(a, nlist)
    return dict(db=db, xq=xq, nlist=nlist, cent=cent, assign=a, off=off, order=order)


def test_kmeans_update_is_mean(ivf):
    cent, h = orc.kmeans_update(ivf["db"], ivf["nlist"], ivf["assign"])
    for c in range(ivf["nlist"]):
        m = ivf["db"][ivf["assign"] == c]
        assert h[c] == len(m)
        if len(m):
            assert np.allclose(cent[c], m.mean(0), rtol=1e-5, atol=1e-4)


def test_kmeans_quality_vs_sklearn(sift):
    from sklearn.cluster import KMeans
    db, _ = sift
    cent, _, obj = orc.kmeans(db, 16, niter=25)
    a = orc.assign(cent, db, L2)
    inertia = float(((db - cent[a]) ** 2).sum())
    sk = KMeans(16, n_init=1, max_iter=25, random_state=0).fit(db)
    assert inertia <= 1.10 * sk.inertia_


@pytest.mark.parametrize("metric", [L2, IP])
def test_ivfflat_matches_bruteforce_over_probed_lists(ivf, metric):
    db, xq, off, order = ivf["db"], ivf["xq"], ivf["off"], ivf["order"]
    nprobe, k = 8, 10
    cdis, keys = orc.coarse_search(ivf["cent"], xq, nprobe, metric)
    dis, ids = orc.ivfflat_search_preassigned(off, db[order], order, xq, k, keys, metric)
    for q in range(xq.shape[0]):
        members = np.concatenate([order[off[l]:off[l + 1]] for l in keys[q]])
        gd, gi = brute(db[members], xq[q:q + 1], k, metric)
        assert np.array_equal(dis[q].astype(np.float64), gd[0])
        assert set(ids[q]) == set(members[gi[0]]) or len(set(gd[0])) < k  # exact unless boundary ties
    if metric == L2:
        _, gt = brute(db, xq, 1, L2)
        assert recall_1nn_in_topi(ids, gt[:, 0], 1) >= 0.8 and recall_1nn_in_topi(ids, gt[:, 0], 10) >= 0.9


def test_ivfflat_tombstone_bit_and_bad_keys(ivf):
    db, xq, off, order = ivf["db"], ivf["xq"], ivf["off"], ivf["order"]
    _, keys = orc.coarse_search(ivf["cent"], xq, 4, L2)
    base_d, base_i = orc.ivfflat_search_preassigned(off, db[order], order, xq, 5, keys, L2)
    ids = order.copy()
    victim = base_i[:, 0]
    pos = np.flatnonzero(np.isin(order, victim))
    ids[pos] |= orc.DEL_MASK  # realtime_mem_data.h:26 tombstone
    d2, i2 = orc.ivfflat_search_preassigned(off, db[order], ids, xq, 5, keys, L2)
    assert not np.isin(i2, victim).any()
    keys2 = keys.copy()
    keys2[:, 1] = -1  # "not enough centroids for multiprobe" (ivfflat.cc:653)
    d3, i3 = orc.ivfflat_search_preassigned(off, db[order], order, xq, 5, keys2, L2)
    assert (i3 >= -1).all()


@pytest.fixture(scope="module")
def pq(ivf):
    db = ivf["db"]
    M = 8
    resid = db - ivf["cent"][ivf["assign"]]
    pqc = orc.pq_train(resid, M, niter=8)
    codes = orc.ivfpq_encode(ivf["cent"], pqc, db, ivf["assign"])
    return dict(M=M, pqc=pqc, codes=codes, resid=resid)


def test_pq_codes_are_per_slice_argmin(ivf, pq):
    M, pqc, resid = pq["M"], pq["pqc"], pq["resid"]
    dsub = resid.shape[1] // M
    r = resid[:500].astype(np.float64)
    for m in range(M):
        dd = ((r[:, None, m * dsub:(m + 1) * dsub] - pqc[m][None].astype(np.float64)) ** 2).sum(-1)
        ref = dd.argmin(1)
        got = pq["codes"][:500, m]
        bad = np.flatnonzero(ref != got)
        for b in bad:  # only fp32-vs-fp64 near ties may differ
            assert abs(dd[b, ref[b]] - dd[b, got[b]]) <= 1e-4 * max(1.0, dd[b, ref[b]])


def test_ivfpq_table_identity_and_rerank(ivf, pq):
    db, xq, off, order = ivf["db"], ivf["xq"], ivf["off"], ivf["order"]
    cent, pqc, codes = ivf["cent"], pq["pqc"], pq["codes"]
    nprobe, k = 16, 10
    cdis, keys = orc.coarse_search(cent, xq, nprobe, L2)
    T = orc.ivfpq_precompute_table(cent, pqc)
    d_tab, i_tab = orc.ivfpq_search_preassigned(off, codes[order], order, cent, pqc, T, xq, k, keys, cdis, L2)
    d_dir, i_dir = orc.ivfpq_search_preassigned(off, codes[order], order, cent, pqc, None, xq, k, keys, cdis, L2)
    # |x-c-r|^2 = |x-c|^2 + (|r|^2 + 2<c,r>) - 2<x,r>   
    assert np.allclose(d_tab, d_dir, rtol=2e-3, atol=2.0)
    # ADC distance == distance to the decoded vector
    dsub = db.shape[1] // pq["M"]
    for q in range(4):
        for j in range(3):
            vid = i_dir[q, j]
            dec = cent[ivf["assign"][vid]] + np.concatenate([pqc[m][codes[vid, m]] for m in range(pq["M"])])
            assert np.isclose(((xq[q] - dec) ** 2).sum(), d_dir[q, j], rtol=1e-3, atol=1.0)
    # exact re-rank (ivfpq.cc:675-726): scores become exact L2 of the returned ids
    d_rr, i_rr = orc.ivfpq_search_preassigned(off, codes[order], order, cent, pqc, T, xq, k, keys, cdis, L2,
                                              recall_num=100, raw=db)
    exact = ((xq[:, None, :] - db[i_rr]) ** 2).sum(-1)
    assert np.array_equal(exact, d_rr)
    _, gt = brute(db, xq, 1, L2)
    assert recall_1nn_in_topi(i_rr, gt[:, 0], 10) >= 0.9
    assert recall_1nn_in_topi(i_rr, gt[:, 0], 10) >= recall_1nn_in_topi(i_tab, gt[:, 0], 10)


def test_ivfpq_inner_product(ivf, pq):
    db, xq, off, order = ivf["db"], ivf["xq"], ivf["off"], ivf["order"]
    cent, pqc, codes = ivf["cent"], pq["pqc"], pq["codes"]
    cdis, keys = orc.coarse_search(cent, xq, 8, IP)
    d, i = orc.ivfpq_search_preassigned(off, codes[order], order, cent, pqc, None, xq, 5, keys, cdis, IP)
    for q in range(4):
        vid = i[q, 0]
        dec = cent[ivf["assign"][vid]] + np.concatenate([pqc[m][codes[vid, m]] for m in range(pq["M"])])
        assert np.isclose(float(xq[q] @ dec), d[q, 0], rtol=1e-3)
    assert (np.diff(d, axis=1) <= 0).all()


def test_merge_partitions_router_order():
    dis = np.array([[[1, 3, 5]], [[1, 2, 6]]], np.float32)
    ids = np.array([[[10, 11, 12]], [[20, 21, -1]]], np.int64)
    od, oi = orc.merge_partitions(dis, ids, L2)
    # equal scores: the later partition first (client.go:1553-1573)
    assert od.tolist() == [[1, 1, 2]]
    assert oi.tolist() == [[(1 << 32) | 20, 10, (1 << 32) | 21]]
