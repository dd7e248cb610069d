"""Independent restatement (struct.pack, no shared code with csrc/index_io.cu) of gamma's index dump
files, used by the tests as the golden writer / reader:

  GammaIVFFlatIndex::Dump   index/impl/gamma_index_ivfflat.cc:807-839   ivfflat.index
  GammaIVFPQIndex::Dump     index/impl/gamma_index_ivfpq.cc:1019-1051   ivfpq.index
  write_ivf_header / write_index_header / write_direct_map / write_product_quantizer /
  WriteInvertedLists        index/index_io.cc:15-141
  faiss::write_index(IndexFlat): fourcc, index header, WRITEXBVECTOR(codes)  (faiss v1.14.1, restated)

Type widths are the C++ ones on x86-64 Linux: int 4, idx_t/long/size_t 8, bool 1, MetricType 4.
"""
import struct

import numpy as np

METRIC_IP, METRIC_L2 = 0, 1  # faiss::MetricType


def fourcc(s):
    return struct.pack("<4s", s.encode())


def _index_header(d, ntotal, metric):
    return struct.pack("<iqqq?i", d, ntotal, 1 << 20, 1 << 20, True, metric)


def _hnsw_graph(n, m=4):
    """faiss::HNSW as write_hnsw stores it (index/index_io.cc:196-212): five vectors, five ints.  A
    syntactically valid single-level graph; readers that only want the centroids skip it."""
    nb = 2 * m
    neigh = np.full(n * nb, -1, np.int32)
    for i in range(n):
        neigh[i * nb] = (i + 1) % n
    out = struct.pack("<Q", 1) + struct.pack("<d", 1.0)                       # assign_probas
    out += struct.pack("<Q", 2) + struct.pack("<ii", 0, nb)                    # cum_nneighbor_per_level
    out += struct.pack("<Q", n) + np.ones(n, np.int32).tobytes()               # levels
    out += struct.pack("<Q", n + 1) + (np.arange(n + 1, dtype=np.uint64) * nb).tobytes()  # offsets
    out += struct.pack("<Q", neigh.size) + neigh.tobytes()                     # neighbors
    out += struct.pack("<iiiii", 0, 0, 40, 16, 1)   # entry_point, max_level, efConstruction, efSearch, upper_beam
    return out


def _ivf_header(d, ntotal, metric, nlist, nprobe, centroids, hnsw_quantizer=False):
    cent = np.ascontiguousarray(centroids, np.float32)
    assert cent.shape == (nlist, d)
    out = _index_header(d, ntotal, metric) + struct.pack("<QQ", nlist, nprobe)
    if hnsw_quantizer:  # faiss::write_index(IndexHNSWFlat): "IHNf", header, graph, then the flat storage
        out += fourcc("IHNf") + _index_header(d, nlist, metric) + _hnsw_graph(nlist)
    out += fourcc("IxF2" if metric == METRIC_L2 else "IxFI") + _index_header(d, nlist, metric)
    out += struct.pack("<Q", cent.size) + cent.tobytes()  # WRITEXBVECTOR: count in 4-byte units
    out += struct.pack("<bQ", 0, 0)                        # DirectMap::NoMap, empty array
    return out


def _inverted_lists(list_off, codes, ids, code_bytes):
    nlist = len(list_off) - 1
    sizes = np.diff(np.asarray(list_off, np.int64)).astype(np.uint64)
    out = fourcc("ilar") + struct.pack("<QQ", nlist, code_bytes) + fourcc("full")
    out += struct.pack("<Q", nlist) + sizes.tobytes()
    codes = np.ascontiguousarray(codes, np.uint8).reshape(-1, code_bytes)
    ids = np.ascontiguousarray(ids, np.int64)
    for l in range(nlist):
        a, b = int(list_off[l]), int(list_off[l + 1])
        if b > a:
            out += codes[a:b].tobytes() + ids[a:b].tobytes()
    return out


def write_ivfflat(d, metric, nprobe, centroids, list_off, vecs, ids, indexed_count, hnsw_quantizer=False):
    """vecs: [n, d] fp32 in list order; ids: int64 with the tombstone top bit."""
    nlist = len(list_off) - 1
    vecs = np.ascontiguousarray(vecs, np.float32).reshape(-1, d)
    out = fourcc("IvFl") + _ivf_header(d, indexed_count, metric, nlist, nprobe, centroids, hnsw_quantizer)
    out += _inverted_lists(list_off, vecs.view(np.uint8), ids, d * 4)
    out += struct.pack("<i", indexed_count)
    return out


def write_ivfpq(d, metric, nprobe, centroids, pq_centroids, list_off, codes, ids, ntotal, opq=None):
    """opq: d x d rotation (row-major, y = A x) written as the LinearTransform block of write_opq
    (index/index_io.cc:230-246) between the product quantizer and the lists."""
    nlist = len(list_off) - 1
    pq = np.ascontiguousarray(pq_centroids, np.float32)
    M = pq.shape[0]
    assert pq.shape == (M, 256, d // M)
    out = fourcc("IwPQ") + _ivf_header(d, ntotal, metric, nlist, nprobe, centroids)
    out += struct.pack("<?QQQQQ", True, M, d, M, 8, pq.size) + pq.tobytes()
    if opq is not None:
        A = np.ascontiguousarray(opq, np.float32)
        assert A.shape == (d, d)
        out += fourcc("LTra") + struct.pack("<?Q", False, A.size) + A.tobytes() + struct.pack("<Qii?", 0, d, d, True)
    out += _inverted_lists(list_off, codes, ids, M)
    return out


class _Reader:
    def __init__(self, buf):
        self.b, self.p = buf, 0

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.p)
        self.p += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def arr(self, dtype, n):
        a = np.frombuffer(self.b, dtype, n, self.p).copy()
        self.p += a.nbytes
        return a


def read_index_file(buf):
    """Parse either file into a dict (the reverse of the writers above)."""
    r = _Reader(buf)
    kind = r.take("4s").decode()
    assert kind in ("IvFl", "IwPQ")
    d, ntotal, _, _, trained, metric = r.take("iqqq?i")
    nlist, nprobe = r.take("QQ")
    q = r.take("4s").decode()
    assert q in ("IxF2", "IxFI")
    qd, qn, _, _, _, qm = r.take("iqqq?i")
    n = r.take("Q")
    assert (qd, qn, n) == (d, nlist, nlist * d)
    cent = r.arr(np.float32, n).reshape(nlist, d)
    assert r.take("bQ") == (0, 0)
    out = dict(kind=kind, d=d, ntotal=ntotal, trained=trained, metric=metric, nlist=nlist, nprobe=nprobe,
               centroids=cent)
    if kind == "IwPQ":
        by_res, code_size, pd, M, nbits, n = r.take("?QQQQQ")
        assert by_res and (pd, nbits, code_size) == (d, 8, M) and n == M * 256 * (d // M)
        out["pq_centroids"] = r.arr(np.float32, n).reshape(M, 256, d // M)
    assert r.take("4s") == b"ilar"
    nb, cb = r.take("QQ")
    assert r.take("4s") == b"full" and nb == nlist
    assert r.take("Q") == nlist
    sizes = r.arr(np.uint64, nlist).astype(np.int64)
    off = np.zeros(nlist + 1, np.int64)
    np.cumsum(sizes, out=off[1:])
    codes = np.empty((int(off[-1]), cb), np.uint8)
    ids = np.empty(int(off[-1]), np.int64)
    for l in range(nlist):
        if sizes[l]:
            codes[off[l]:off[l + 1]] = r.arr(np.uint8, int(sizes[l]) * cb).reshape(-1, cb)
            ids[off[l]:off[l + 1]] = r.arr(np.int64, int(sizes[l]))
    out.update(code_bytes=cb, list_off=off, codes=codes, ids=ids)
    if kind == "IvFl":
        out["indexed_count"] = r.take("i")
    assert r.p == len(buf), "trailing bytes"
    return out
