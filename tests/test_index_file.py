"""gamma's own index dump files (ivfflat.index / ivfpq.index).

CPU part: the independent struct.pack restatement of the format (tests/gamma_index_file.py) round
trips and matches a committed golden file.  GPU part: the library's Dump is byte-identical to that
restatement for the same index state, and Load turns a file back into an index that answers exactly
like the one that was dumped (index/impl/gamma_index_ivfflat.cc:807-892, gamma_index_ivfpq.cc:1019-1116,
index/index_io.cc:108-194)."""
import os
import struct

import numpy as np
import pytest

import gamma_index_file as gif
from oracle import oracle as orc
from vearch_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
L2, IP = orc.METRIC_L2, orc.METRIC_IP
DEL = np.int64(-1) << np.int64(63)


def faiss_metric(metric):
    return gif.METRIC_L2 if metric == L2 else gif.METRIC_IP


def mt(metric):
    return "L2" if metric == L2 else "InnerProduct"


def tiny_state():
    rng = np.random.default_rng(7)
    d, nlist = 8, 3
    cent = rng.integers(0, 9, (nlist, d)).astype(np.float32)
    off = np.array([0, 2, 2, 5])
    vecs = rng.integers(0, 9, (5, d)).astype(np.float32)
    ids = np.array([0, 3, 1, 2, 4], np.int64)
    ids[3] |= DEL
    return d, nlist, cent, off, vecs, ids


def test_reference_writer_round_trip_and_golden_bytes():
    d, nlist, cent, off, vecs, ids = tiny_state()
    buf = gif.write_ivfflat(d, gif.METRIC_L2, 2, cent, off, vecs, ids, 5)
    with open(os.path.join(HERE, "golden", "ivfflat_tiny.index"), "rb") as f:
        assert f.read() == buf
    # fixed offsets of the header fields (x86-64 widths; index_io.cc:15-23, 41-47)
    assert buf[:4] == b"IvFl" and struct.unpack_from("<i", buf, 4)[0] == d
    assert struct.unpack_from("<q", buf, 8)[0] == 5 and struct.unpack_from("<qq", buf, 16) == (1 << 20, 1 << 20)
    assert buf[32] == 1 and struct.unpack_from("<i", buf, 33)[0] == gif.METRIC_L2
    assert struct.unpack_from("<QQ", buf, 37) == (nlist, 2) and buf[53:57] == b"IxF2"
    r = gif.read_index_file(buf)
    assert r["kind"] == "IvFl" and r["indexed_count"] == 5 and r["code_bytes"] == d * 4
    assert np.array_equal(r["centroids"], cent) and np.array_equal(r["ids"], ids) and np.array_equal(r["list_off"], off)
    assert np.array_equal(r["codes"].view(np.float32).reshape(-1, d), vecs)
    pq = np.arange(4 * 256 * 2, dtype=np.float32).reshape(4, 256, 2)
    codes = (np.arange(20) * 13 % 256).astype(np.uint8).reshape(5, 4)
    r = gif.read_index_file(gif.write_ivfpq(d, gif.METRIC_IP, 2, cent, pq, off, codes, ids, 5))
    assert r["kind"] == "IwPQ" and r["metric"] == gif.METRIC_IP and r["code_bytes"] == 4
    assert np.array_equal(r["pq_centroids"], pq) and np.array_equal(r["codes"], codes)


def gi():
    from vearch_b200 import index as gidx
    return gidx


@pytest.mark.gpu
@pytest.mark.parametrize("metric,d", [(L2, 32), (IP, 32), (L2, 18)])
def test_ivfflat_dump_is_byte_exact_and_loads_back(tmp_path, metric, d):
    n, nlist, nq, nprobe, k = 6000, 12, 40, 5, 10
    db, xq = synth.sift_like(n, d, seed=71), synth.sift_like(nq, d, seed=72)
    cent, _, _ = orc.kmeans(db[:2000], nlist, niter=4)
    params = {"ncentroids": nlist, "nprobe": nprobe, "metric_type": mt(metric)}
    idx = gi().GammaIndex("IVFFLAT", d, params)
    idx.set_centroids(cent)
    idx.add_vectors(db)
    idx.add_pending()
    idx.update_vector(17, db[18])  # leaves a tombstone behind (realtime_mem_data.cc:298-320)
    off, codes, ids = idx.export_lists()
    assert (ids < 0).sum() == 1
    vecs = codes.view(np.float32).reshape(len(ids), -1)[:, :d]
    idx.dump(tmp_path, "emb.000")
    with open(tmp_path / "emb.000" / "ivfflat.index", "rb") as f:
        got = f.read()
    assert got == gif.write_ivfflat(d, faiss_metric(metric), nprobe, cent, off, vecs, ids, n)
    cd, keys = orc.coarse_search(cent, xq, nprobe, metric)
    ref = idx.search_preassigned(xq, k, keys, cd)

    idx2 = gi().GammaIndex("IVFFLAT", d, params)
    with pytest.raises(RuntimeError):  # the file indexes vectors the store does not hold yet
        idx2.load(tmp_path, "emb.000")
    idx2 = gi().GammaIndex("IVFFLAT", d, params)
    db2 = db.copy()
    db2[17] = db[18]
    idx2.add_vectors(db2)
    assert idx2.load(tmp_path, "nothing.000") == 0 and not idx2.is_trained  # no file: train again
    assert idx2.load(tmp_path, "emb.000") == n and idx2.is_trained and idx2.indexed_count == n
    o2, c2, i2 = idx2.export_lists()
    assert np.array_equal(o2, off) and np.array_equal(c2, codes) and np.array_equal(i2, ids)
    got2 = idx2.search_preassigned(xq, k, keys, cd)
    assert np.array_equal(got2[0], ref[0]) and np.array_equal(got2[1], ref[1])
    # a file whose coarse quantizer is an IndexHNSWFlat: the centroids are taken, the graph is skipped
    hdir = tmp_path / "hnsw" / "emb.000"
    hdir.mkdir(parents=True)
    (hdir / "ivfflat.index").write_bytes(gif.write_ivfflat(d, faiss_metric(metric), nprobe, cent, off, vecs, ids, n,
                                                          hnsw_quantizer=True))
    idx4 = gi().GammaIndex("IVFFLAT", d, params)
    idx4.add_vectors(db2)
    assert idx4.load(tmp_path / "hnsw", "emb.000") == n and np.array_equal(idx4.get_centroids(), cent)
    got4 = idx4.search_preassigned(xq, k, keys, cd)
    assert np.array_equal(got4[0], ref[0]) and np.array_equal(got4[1], ref[1])
    idx4.close()
    # the loaded index keeps working: realtime adds and an update that must find vid -> (list, pos)
    more = synth.sift_like(500, d, seed=73)
    idx2.add_vectors(more)
    idx2.add_pending()
    assert idx2.indexed_count == n + 500
    idx2.update_vector(100, more[0])
    _, _, i3 = idx2.export_lists()
    assert (i3 < 0).sum() == 2 and (i3 == 100).sum() == 1
    db3 = np.vstack([db2, more])
    db3[100] = more[0]
    do, io = orc.flat_search(db3, more[:1], 2, metric)
    dg, ig = idx2.search(more[:1], 2, params={"nprobe": nlist})
    assert sorted(ig[0]) == sorted(io[0]) and np.array_equal(dg, do)
    if metric == L2:
        assert sorted(ig[0]) == [100, n]
    idx.close()
    idx2.close()


@pytest.mark.gpu
@pytest.mark.parametrize("metric", [L2, IP])
def test_ivfpq_dump_is_byte_exact_and_loads_back(tmp_path, metric):
    d, n, nlist, M, nq, nprobe, k = 32, 6000, 12, 8, 40, 5, 10
    db, xq = synth.sift_like(n, d, seed=74), synth.sift_like(nq, d, seed=75)
    cent, _, assign = orc.kmeans(db[:2000], nlist, niter=4)
    _, a = orc.coarse_search(cent, db[:3000], 1, L2)
    pqc = orc.pq_train(db[:3000] - cent[a[:, 0]], M, niter=4)
    params = {"ncentroids": nlist, "nprobe": nprobe, "nsubvector": M, "metric_type": mt(metric)}
    idx = gi().GammaIndex("IVFPQ", d, params)
    idx.set_centroids(cent)
    idx.set_pq_centroids(pqc)
    idx.add_vectors(db)
    idx.add_pending()
    off, codes, ids = idx.export_lists()
    idx.dump(tmp_path, "emb.000")
    with open(tmp_path / "emb.000" / "ivfpq.index", "rb") as f:
        got = f.read()
    assert got == gif.write_ivfpq(d, faiss_metric(metric), nprobe, cent, pqc, off, codes, ids, n)
    cd, keys = orc.coarse_search(cent, xq, nprobe, metric)
    ref = idx.search_preassigned(xq, k, keys, cd, params={"recall_num": 50})
    idx2 = gi().GammaIndex("IVFPQ", d, params)
    idx2.add_vectors(db)
    assert idx2.load(tmp_path, "emb.000") == n and idx2.is_trained
    assert np.array_equal(idx2.get_pq_centroids(), pqc)
    if metric == L2:  # "precomputed table not stored. It is cheaper to recompute it" (ivfpq.cc:1091-1095)
        assert np.array_equal(idx2.get_precomputed_table(), idx.get_precomputed_table())
    got2 = idx2.search_preassigned(xq, k, keys, cd, params={"recall_num": 50})
    assert np.array_equal(got2[0], ref[0]) and np.array_equal(got2[1], ref[1])
    # a file for another table layout is refused, not half-loaded
    idx3 = gi().GammaIndex("IVFPQ", d, dict(params, nsubvector=4))
    idx3.add_vectors(db)
    with pytest.raises(RuntimeError):
        idx3.load(tmp_path, "emb.000")
    for i in (idx, idx2, idx3):
        i.close()
