"""The CPU arm of bench.py (baseline/cpu_gamma.c, built for speed: FMA, AVX-512 dispatch, blocked sgemm) must
return what the correctness checker (oracle/, built with -ffp-contract=off) returns: identical ids on integer-valued
SIFT-shaped data where every distance is exact, scores within 1e-5 relative on float data."""
import numpy as np
import pytest

from baseline import cpu_gamma as cg
from oracle import oracle as orc
from vearch_b200 import synth

L2, IP = orc.METRIC_L2, orc.METRIC_IP


@pytest.fixture(scope="module")
def state():
    d, n, nlist, M = 64, 20000, 32, 16
    db = synth.sift_like(n, d, seed=11)
    xq = synth.sift_like(64, d, seed=12)
    cent, _, _ = orc.kmeans(db[:4000], nlist, niter=4)
    return dict(d=d, n=n, nlist=nlist, M=M, db=db, xq=xq, cent=cent)


@pytest.mark.parametrize("metric", [L2, IP])
def test_flat_and_coarse_match_oracle(state, metric):
    s = state
    do, io = orc.flat_search(s["db"][:5000], s["xq"], 10, metric)
    dc, ic = cg.flat_search(s["db"][:5000], s["xq"], 10, metric)
    assert np.array_equal(do, dc) and np.array_equal(io, ic)  # integer data: exact whatever the FMA choices
    cent = np.rint(s["cent"]).astype(np.float32)
    do, io = orc.coarse_search(cent, s["xq"], 8, metric)
    dc, ic = cg.coarse_search(cent, s["xq"], 8, metric)
    assert np.array_equal(do, dc) and np.array_equal(io, ic)
    # float centroids: expanded form |x|^2+|c|^2-2x.c vs direct form, same lists up to near-ties
    do, io = orc.coarse_search(s["cent"], s["xq"], 8, metric)
    dc, ic = cg.coarse_search(s["cent"], s["xq"], 8, metric)
    assert np.allclose(do, dc, rtol=1e-4, atol=1e-2) and (io == ic).mean() > 0.98


@pytest.mark.parametrize("metric", [L2, IP])
def test_ivfflat_matches_oracle(state, metric):
    s = state
    a = orc.assign(s["cent"], s["db"], metric)
    off, order = orc.build_lists(a, s["nlist"])
    vecs = s["db"][order]
    _, keys = orc.coarse_search(s["cent"], s["xq"], 6, metric)
    deleted = np.random.default_rng(3).random(s["n"]) < 0.2
    delb = np.packbits(deleted, bitorder="little")
    for kw in ({}, {"del_bitmap": delb}):
        do, io = orc.ivfflat_search_preassigned(off, vecs, order, s["xq"], 10, keys, metric, **kw)
        dc, ic = cg.ivfflat_search_preassigned(off, vecs, order, s["xq"], 10, keys, metric, **kw)
        assert np.array_equal(do, dc) and np.array_equal(io, ic)


@pytest.mark.parametrize("metric", [L2, IP])
def test_ivfpq_matches_oracle(state, metric):
    s = state
    a = orc.assign(s["cent"], s["db"], metric)
    off, order = orc.build_lists(a, s["nlist"])
    pqc = orc.pq_train(s["db"][:6000] - s["cent"][a[:6000]], s["M"], niter=4)
    codes = orc.ivfpq_encode(s["cent"], pqc, s["db"], a)[order]
    T = orc.ivfpq_precompute_table(s["cent"], pqc) if metric == L2 else None
    cd, keys = orc.coarse_search(s["cent"], s["xq"], 6, metric)
    # ADC stage: float tables, FMA may change the last bit of a LUT entry -> compare scores to 1e-5, ids where scores differ clearly
    do, io = orc.ivfpq_search_preassigned(off, codes, order, s["cent"], pqc, T, s["xq"], 50, keys, cd, metric)
    dc, ic = cg.ivfpq_search_preassigned(off, codes, order, s["cent"], pqc, T, s["xq"], 50, keys, cd, metric)
    assert np.allclose(do, dc, rtol=1e-5, atol=1e-3)
    assert (io == ic).mean() > 0.99
    # with the exact re-rank the final answer is the exact distance of integer vectors: identical
    do, io = orc.ivfpq_search_preassigned(off, codes, order, s["cent"], pqc, T, s["xq"], 10, keys, cd, metric, recall_num=100,
                                          raw=s["db"])
    dc, ic = cg.ivfpq_search_preassigned(off, codes, order, s["cent"], pqc, T, s["xq"], 10, keys, cd, metric, recall_num=100,
                                         raw=s["db"])
    assert np.array_equal(do, dc) and (io == ic).mean() > 0.995


def test_thread_control_and_isa():
    n = cg.set_threads(2)
    assert n == 2 and cg.num_threads() == 2
    assert cg.set_threads() == cg.physical_cores() >= 1
    assert cg.isa() in ("avx512f", "avx2+fma", "scalar")
