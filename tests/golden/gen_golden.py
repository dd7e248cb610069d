"""Generates the golden wire fixtures under tests/golden/ with the OFFICIAL protobuf runtime
(google.protobuf, pure python) from the REFERENCE'S OWN descriptors: tests/golden/vearchpb_descriptor_set.binpb holds
the FileDescriptorProtos protoc embedded in internal/proto/vearchpb/*.pb.go (extract_pb_descriptors.py).  The hand
restatement below (router_grpc.proto:19-35,129-136,168-219, data_model.proto:32-37: the messages that cross the
gamma C-ABI) is kept as a cross-check: tests/test_boundary_cpu.py asserts it agrees field by field with the real
descriptors.  Run here (CPU container); the .bin/.json outputs are committed.

    python tests/golden/gen_golden.py
"""
import json
import os
import struct

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

HERE = os.path.dirname(os.path.abspath(__file__))
F = descriptor_pb2.FieldDescriptorProto


def build_pool():
    fd = descriptor_pb2.FileDescriptorProto(name="vearch_subset.proto", package="vearchpb", syntax="proto3")

    def msg(name, fields, nested=None):
        m = fd.message_type.add(name=name)
        for fname, num, ftype, label, tname in fields:
            f = m.field.add(name=fname, number=num, type=ftype, label=label)
            if tname:
                f.type_name = tname
        return m

    OPT, REP = F.LABEL_OPTIONAL, F.LABEL_REPEATED
    head = msg("RequestHead", [("time_out_ms", 1, F.TYPE_INT64, OPT, None), ("user_name", 2, F.TYPE_STRING, OPT, None),
                               ("db_name", 4, F.TYPE_STRING, OPT, None), ("space_name", 5, F.TYPE_STRING, OPT, None),
                               ("params", 7, F.TYPE_MESSAGE, REP, ".vearchpb.RequestHead.ParamsEntry")])
    e = head.nested_type.add(name="ParamsEntry")
    e.field.add(name="key", number=1, type=F.TYPE_STRING, label=OPT)
    e.field.add(name="value", number=2, type=F.TYPE_STRING, label=OPT)
    e.options.map_entry = True
    msg("VectorQuery", [("name", 1, F.TYPE_STRING, OPT, None), ("value", 2, F.TYPE_BYTES, OPT, None),
                        ("min_score", 3, F.TYPE_DOUBLE, OPT, None), ("max_score", 4, F.TYPE_DOUBLE, OPT, None),
                        ("format", 5, F.TYPE_STRING, OPT, None), ("index_type", 6, F.TYPE_STRING, OPT, None)])
    msg("RangeFilter", [("field", 1, F.TYPE_STRING, OPT, None), ("lower_value", 2, F.TYPE_BYTES, OPT, None),
                        ("upper_value", 3, F.TYPE_BYTES, OPT, None), ("include_lower", 4, F.TYPE_BOOL, OPT, None),
                        ("include_upper", 5, F.TYPE_BOOL, OPT, None), ("is_union", 6, F.TYPE_INT32, OPT, None)])
    msg("TermFilter", [("field", 1, F.TYPE_STRING, OPT, None), ("value", 2, F.TYPE_BYTES, OPT, None),
                       ("is_union", 3, F.TYPE_INT32, OPT, None)])
    msg("SearchRequest", [("head", 1, F.TYPE_MESSAGE, OPT, ".vearchpb.RequestHead"), ("req_num", 2, F.TYPE_INT32, OPT, None),
                          ("topN", 3, F.TYPE_INT32, OPT, None), ("is_brute_search", 4, F.TYPE_INT32, OPT, None),
                          ("vec_fields", 5, F.TYPE_MESSAGE, REP, ".vearchpb.VectorQuery"),
                          ("fields", 6, F.TYPE_STRING, REP, None),
                          ("range_filters", 7, F.TYPE_MESSAGE, REP, ".vearchpb.RangeFilter"),
                          ("term_filters", 8, F.TYPE_MESSAGE, REP, ".vearchpb.TermFilter"),
                          ("index_params", 9, F.TYPE_STRING, OPT, None), ("multi_vector_rank", 10, F.TYPE_INT32, OPT, None),
                          ("l2_sqrt", 11, F.TYPE_BOOL, OPT, None), ("ranker", 15, F.TYPE_STRING, OPT, None),
                          ("trace", 16, F.TYPE_BOOL, OPT, None), ("operator", 17, F.TYPE_INT32, OPT, None),
                          ("offset", 20, F.TYPE_INT32, OPT, None),
                          ("partition_names", 22, F.TYPE_STRING, REP, None)])
    msg("QueryRequest", [("head", 1, F.TYPE_MESSAGE, OPT, ".vearchpb.RequestHead"),
                         ("document_ids", 2, F.TYPE_STRING, REP, None), ("partition_id", 3, F.TYPE_INT32, OPT, None),
                         ("next", 4, F.TYPE_BOOL, OPT, None),
                         ("range_filters", 5, F.TYPE_MESSAGE, REP, ".vearchpb.RangeFilter"),
                         ("term_filters", 6, F.TYPE_MESSAGE, REP, ".vearchpb.TermFilter"),
                         ("fields", 7, F.TYPE_STRING, REP, None), ("is_vector_value", 8, F.TYPE_BOOL, OPT, None),
                         ("limit", 9, F.TYPE_INT32, OPT, None), ("operator", 15, F.TYPE_INT32, OPT, None),
                         ("offset", 17, F.TYPE_INT32, OPT, None)])
    msg("Field", [("name", 1, F.TYPE_STRING, OPT, None), ("type", 2, F.TYPE_INT32, OPT, None),
                  ("value", 3, F.TYPE_BYTES, OPT, None)])
    msg("ResultItem", [("score", 1, F.TYPE_DOUBLE, OPT, None), ("fields", 2, F.TYPE_MESSAGE, REP, ".vearchpb.Field"),
                       ("p_key", 3, F.TYPE_STRING, OPT, None)])
    msg("SearchStatus", [("total", 1, F.TYPE_INT32, OPT, None), ("failed", 2, F.TYPE_INT32, OPT, None),
                         ("successful", 3, F.TYPE_INT32, OPT, None), ("msg", 4, F.TYPE_STRING, OPT, None)])
    msg("SearchResult", [("total_hits", 1, F.TYPE_INT32, OPT, None), ("max_score", 2, F.TYPE_DOUBLE, OPT, None),
                         ("status", 5, F.TYPE_MESSAGE, OPT, ".vearchpb.SearchStatus"), ("msg", 6, F.TYPE_STRING, OPT, None),
                         ("result_items", 7, F.TYPE_MESSAGE, REP, ".vearchpb.ResultItem"),
                         ("timeout", 9, F.TYPE_BOOL, OPT, None)])
    msg("SearchResponse", [("results", 2, F.TYPE_MESSAGE, REP, ".vearchpb.SearchResult"),
                           ("timeout", 3, F.TYPE_BOOL, OPT, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return pool


def reference_pool():
    """descriptor pool of the reference's real .proto files (errors, data_model, router_grpc), or None"""
    path = os.path.join(HERE, "vearchpb_descriptor_set.binpb")
    if not os.path.exists(path):
        return None
    fds = descriptor_pb2.FileDescriptorSet()
    fds.ParseFromString(open(path, "rb").read())
    pool = descriptor_pool.DescriptorPool()
    for f in fds.file:
        pool.Add(f)
    return pool


NAMES = ("SearchRequest", "SearchResponse", "VectorQuery", "RequestHead", "QueryRequest")


def classes(restated=False):
    """message classes from the reference's own descriptors (default) or from the hand restatement"""
    pool = None if restated else reference_pool()
    if pool is None:
        pool = build_pool()
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("vearchpb." + n))
    return {n: get(n) for n in NAMES}


def main():
    cls = classes()
    rng = np.random.default_rng(7)
    # ---- request 1: what the router sends (doc_query.go:1220-1248): window = -/+ MaxFloat64 ----
    q = rng.integers(0, 200, size=(3, 8)).astype(np.float32)
    r = cls["SearchRequest"]()
    r.head.params["request_id"] = "req-42"
    r.head.params["partition_id"] = "7"
    r.head.space_name = "ts_space"
    r.req_num, r.topN, r.is_brute_search = 3, 10, 0
    v = r.vec_fields.add()
    v.name, v.value, v.min_score, v.max_score = "emb", q.tobytes(), -1.7976931348623157e308, 1.7976931348623157e308
    r.fields.append("_id")
    r.index_params = json.dumps({"nprobe": 16, "metric_type": "L2", "recall_num": 100})
    r.trace, r.offset, r.l2_sqrt = True, 2, True
    r.partition_names.append("p0")  # unknown to the engine: must be skipped
    open(os.path.join(HERE, "search_request_router.bin"), "wb").write(r.SerializeToString(deterministic=True))  # map entries in key order
    exp1 = {"request_id": "req-42", "partition_id": 7, "req_num": 3, "topn": 10, "brute_force_search": 0,
            "index_params": r.index_params, "trace": True, "offset": 2, "l2_sqrt": True, "fields": ["_id"],
            "n_range_filters": 0,
            "vec_fields": [{"name": "emb", "value_len": q.nbytes, "index_type": "", "min_score": v.min_score,
                            "max_score": v.max_score}]}
    # ---- request 2: defaults dropped by proto3, filters present, brute force ----
    r2 = cls["SearchRequest"]()
    r2.req_num, r2.topN, r2.is_brute_search = 1, 5, 1
    v2 = r2.vec_fields.add()
    v2.name, v2.value, v2.max_score, v2.index_type = "emb", q[:1].tobytes(), 1234.5, "IVFPQ"
    rf = r2.range_filters.add()
    rf.field, rf.lower_value = "price", b"\x01\x02"
    open(os.path.join(HERE, "search_request_minimal.bin"), "wb").write(r2.SerializeToString())
    exp2 = {"request_id": "", "partition_id": 0, "req_num": 1, "topn": 5, "brute_force_search": 1, "index_params": "",
            "trace": False, "offset": 0, "l2_sqrt": False, "fields": [], "n_range_filters": 1,
            "vec_fields": [{"name": "emb", "value_len": 32, "index_type": "IVFPQ", "min_score": 0.0, "max_score": 1234.5}]}
    json.dump({"search_request_router.bin": exp1, "search_request_minimal.bin": exp2, "queries_router": q.tolist()},
              open(os.path.join(HERE, "search_requests.json"), "w"), indent=1)
    # ---- response: what Response::Serialize emits (response.cc:89-162) ----
    resp = cls["SearchResponse"]()
    resp.timeout = False
    scores = [[0.0, 12.5, 100.25], [3.0, 4.0, 5.0]]
    keys = [["a", "bb", "ccc"], ["k0", "k1", "k2"]]
    for i in range(2):
        sr = resp.results.add()
        sr.status.total = 1000
        sr.status.successful = 1000
        sr.status.failed = 0
        sr.status.msg = ""
        for s, kk in zip(scores[i], keys[i]):
            it = sr.result_items.add()
            it.score = s
            f = it.fields.add()
            f.name, f.value = "_id", kk.encode()
        sr.msg = "OK"
        sr.max_score = max(scores[i])
        sr.timeout = False
    open(os.path.join(HERE, "search_response.bin"), "wb").write(resp.SerializeToString())
    json.dump({"scores": scores, "keys": keys, "total": 1000}, open(os.path.join(HERE, "search_response.json"), "w"))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()


def gen_flat_golden():
    """Known-answer vectors for FLAT / IVF-Flat from an INDEPENDENT implementation (numpy fp64
    brute force), stored with their inputs so they do not depend on generator stability."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from vearch_b200 import synth
    db = synth.sift_like(2000, 32, seed=5)
    xq = synth.sift_like(16, 32, seed=6)
    d64, q64 = db.astype(np.float64), xq.astype(np.float64)
    l2 = ((q64[:, None, :] - d64[None, :, :]) ** 2).sum(-1)
    ip = q64 @ d64.T
    ids = np.broadcast_to(np.arange(db.shape[0]), l2.shape)
    o_l2 = np.lexsort((ids, l2), axis=1)[:, :10]
    o_ip = np.lexsort((ids, -ip), axis=1)[:, :10]
    np.savez_compressed(os.path.join(HERE, "flat_small.npz"), db=db.astype(np.uint8), xq=xq.astype(np.uint8),
                        l2_ids=o_l2, l2_dis=np.take_along_axis(l2, o_l2, 1).astype(np.float32),
                        ip_ids=o_ip, ip_dis=np.take_along_axis(ip, o_ip, 1).astype(np.float32))


if __name__ == "__main__":
    gen_flat_golden()


def gen_index_file_golden():
    """ivfflat_tiny.index: gamma's IVF-Flat dump format (index/index_io.cc, gamma_index_ivfflat.cc:807-839)
    written by the independent struct.pack restatement in tests/gamma_index_file.py."""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import gamma_index_file as gif
    import test_index_file as tif
    d, nlist, cent, off, vecs, ids = tif.tiny_state()
    with open(os.path.join(HERE, "ivfflat_tiny.index"), "wb") as f:
        f.write(gif.write_ivfflat(d, gif.METRIC_L2, 2, cent, off, vecs, ids, 5))


if __name__ == "__main__":
    gen_index_file_golden()
