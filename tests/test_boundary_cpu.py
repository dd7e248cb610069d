"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the
headers declare, and the hand-written wire codecs (C++ in libgamma.so, Python in vearch_b200/wire.py)
agree with golden bytes produced by the official protobuf runtime (tests/golden/gen_golden.py) and
with each other.  No compute call is made: Init / index creation must FAIL without a GPU."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from oracle import oracle as orc
from vearch_b200 import _lib, wire

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
libc = C.CDLL(None)
libc.free.argtypes = [C.c_void_p]


def declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", txt)) - {"defined", "C"})


@pytest.mark.parametrize("header", ["gamma_api.h", "gamma_b200_index.h"])
def test_library_exports_every_declared_symbol(header):
    lib = _lib.lib()
    names = declared_symbols(header)
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"{header}: not exported by libgamma.so: {missing}"
    if header == "gamma_api.h":
        # exactly the 23 entry points of internal/engine/c_api/gamma_api.h:26-188
        assert len(names) == 23


def test_library_exports_only_the_c_abi():
    """csrc/exports.map: the dynamic symbol table holds the 23 gamma entry points and gb_* -- no C++ symbol of the
    implementation leaks into the process that loads the drop-in."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.lib()._name], capture_output=True, text=True, check=True).stdout
    names = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    gamma = set(declared_symbols("gamma_api.h"))
    extra = sorted(n for n in names if n not in gamma and not n.startswith("gb_"))
    assert not extra, f"unexpected exported symbols: {extra[:10]}"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.lib()
    assert lib.gb_device_count() == 0
    assert not lib.gb_index_create(b"FLAT", 8, b"", 0)
    assert b"no CUDA device" in lib.gb_last_error()
    lib.Init.restype = C.c_void_p
    cfg = json.dumps({"path": "/tmp/x", "space_name": "s", "log_dir": "/tmp"}).encode()
    assert not lib.Init(cfg, len(cfg))  # NULL, like a failed gamma.Init (gammacb/gamma.go:84-88)


def _call_json(fn, payload):
    p, n = C.c_void_p(), C.c_int()
    fn.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    rc = fn(payload, len(payload), C.byref(p), C.byref(n))
    data = C.string_at(p.value, n.value) if p.value else b""
    if p.value:
        libc.free(p.value)
    return rc, data


@pytest.mark.parametrize("name", ["search_request_router.bin", "search_request_minimal.bin"])
def test_cpp_parses_official_search_request_bytes(name):
    exp = json.load(open(os.path.join(GOLD, "search_requests.json")))[name]
    raw = open(os.path.join(GOLD, name), "rb").read()
    rc, out = _call_json(_lib.lib().gb_debug_parse_search_request, raw)
    assert rc == 0
    got = json.loads(out)
    for k, v in exp.items():
        assert got[k] == v, (k, got[k], v)


def test_cpp_rejects_truncated_request():
    raw = open(os.path.join(GOLD, "search_request_router.bin"), "rb").read()
    rc, _ = _call_json(_lib.lib().gb_debug_parse_search_request, raw[:-3])
    assert rc == -1


def test_python_encoder_matches_official_parser():
    import sys
    sys.path.insert(0, GOLD)
    import gen_golden
    cls = gen_golden.classes()
    q = np.arange(24, dtype=np.float32).reshape(3, 8)
    mine = wire.encode_search_request("emb", q, 10, index_params='{"nprobe": 4}', request_id="r1", partition_id=3,
                                      min_score=-1e300, max_score=1e300, offset=1, trace=True)
    m = cls["SearchRequest"]()
    m.ParseFromString(mine)
    assert m.req_num == 3 and m.topN == 10 and m.offset == 1 and m.trace and m.index_params == '{"nprobe": 4}'
    assert m.head.params["request_id"] == "r1" and m.head.params["partition_id"] == "3"
    assert m.vec_fields[0].name == "emb" and m.vec_fields[0].value == q.tobytes()
    assert m.vec_fields[0].min_score == -1e300 and list(m.fields) == ["_id"]
    rc, out = _call_json(_lib.lib().gb_debug_parse_search_request, mine)
    got = json.loads(out)
    assert rc == 0 and got["request_id"] == "r1" and got["partition_id"] == 3 and got["vec_fields"][0]["value_len"] == 96


def test_scalar_filters_cross_the_wire_like_the_official_runtime():
    """RangeFilter / TermFilter / operator (router_grpc.proto:109-122, 185): the hand-written encoder
    produces what the official protobuf runtime parses, and the C++ reader recovers every field."""
    import struct
    import sys
    sys.path.insert(0, GOLD)
    import gen_golden
    cls = gen_golden.classes()
    q = np.arange(8, dtype=np.float32).reshape(1, 8)
    lo, hi = struct.pack("<i", -5), struct.pack("<i", 70000)
    mine = wire.encode_search_request("emb", q, 5, range_filters=[("price", lo, hi, True, False), ("stock", lo, b"", False, False, 2)],
                                      term_filters=[("tag", b"a\x01b", 1), ("color", b"red")], operator=1)
    m = cls["SearchRequest"]()
    m.ParseFromString(mine)
    assert [(r.field, r.lower_value, r.upper_value, r.include_lower, r.include_upper, r.is_union) for r in m.range_filters] == \
        [("price", lo, hi, True, False, 0), ("stock", lo, b"", False, False, 2)]
    assert [(t.field, t.value, t.is_union) for t in m.term_filters] == [("tag", b"a\x01b", 1), ("color", b"red", 0)]
    assert m.operator == 1
    for raw in (mine, m.SerializeToString()):
        rc, out = _call_json(_lib.lib().gb_debug_parse_search_request, raw)
        got = json.loads(out)
        assert rc == 0 and got["filter_operator"] == 1 and got["n_range_filters"] == 2 and got["n_term_filters"] == 2
        assert got["filters"] == [
            dict(field="price", lower=lo.hex(), upper=hi.hex(), include_lower=True, include_upper=False, is_term=False, is_union=0),
            dict(field="stock", lower=lo.hex(), upper="", include_lower=False, include_upper=False, is_term=False, is_union=2),
            dict(field="tag", lower=b"a\x01b".hex(), upper="", include_lower=False, include_upper=False, is_term=True, is_union=1),
            dict(field="color", lower=b"red".hex(), upper="", include_lower=False, include_upper=False, is_term=True, is_union=0)]


def test_multi_vector_request_crosses_the_wire_like_the_official_runtime():
    import sys
    sys.path.insert(0, GOLD)
    import gen_golden
    cls = gen_golden.classes()
    q1 = np.arange(16, dtype=np.float32).reshape(2, 8)
    q2 = np.arange(8, dtype=np.float32).reshape(2, 4) + 100
    ranker = '{"type": "WeightedRanker", "params": [0.25, 0.75]}'
    mine = wire.encode_search_request("emb", q1, 7, extra_vec_queries=[("img", q2)], ranker=ranker, multi_vector_rank=1)
    m = cls["SearchRequest"]()
    m.ParseFromString(mine)
    assert [(v.name, v.value) for v in m.vec_fields] == [("emb", q1.tobytes()), ("img", q2.tobytes())]
    assert m.ranker == ranker and m.multi_vector_rank == 1 and m.req_num == 2 and m.topN == 7
    rc, out = _call_json(_lib.lib().gb_debug_parse_search_request, m.SerializeToString())
    got = json.loads(out)
    assert rc == 0 and got["multi_vector_rank"] == 1
    assert [(v["name"], v["value_len"]) for v in got["vec_fields"]] == [("emb", 64), ("img", 32)]


def test_query_request_encoder_matches_official_runtime():
    import struct
    import sys
    sys.path.insert(0, GOLD)
    import gen_golden
    cls = gen_golden.classes()
    lo = struct.pack("<q", 7)
    mine = wire.encode_query_request(document_ids=["a", "b"], partition_id=2, fields=["_id", "tag"], limit=20, operator=1,
                                     offset=3, range_filters=[("n", lo, b"", True, False)], term_filters=[("tag", b"x", 2)])
    m = cls["QueryRequest"]()
    m.ParseFromString(mine)
    assert list(m.document_ids) == ["a", "b"] and m.partition_id == 2 and list(m.fields) == ["_id", "tag"]
    assert (m.limit, m.operator, m.offset) == (20, 1, 3)
    assert [(r.field, r.lower_value, r.include_lower) for r in m.range_filters] == [("n", lo, True)]
    assert [(t.field, t.value, t.is_union) for t in m.term_filters] == [("tag", b"x", 2)]
    assert m.SerializeToString() == mine  # same field order as the official serializer


def test_cpp_response_bytes_equal_official_serializer():
    exp = json.load(open(os.path.join(GOLD, "search_response.json")))
    official = open(os.path.join(GOLD, "search_response.bin"), "rb").read()
    nq, k = len(exp["scores"]), len(exp["scores"][0])
    scores = (C.c_double * (nq * k))(*[s for row in exp["scores"] for s in row])
    keys = (C.c_char_p * (nq * k))(*[s.encode() for row in exp["keys"] for s in row])
    fn = _lib.lib().gb_debug_encode_response
    fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    p, n = C.c_void_p(), C.c_int()
    assert fn(nq, k, scores, keys, exp["total"], C.byref(p), C.byref(n)) == 0
    mine = C.string_at(p.value, n.value)
    libc.free(p.value)
    assert mine == official  # byte-identical to what libprotobuf emits for Response::Serialize
    dec = wire.decode_search_response(official)
    assert [[it["score"] for it in r["items"]] for r in dec] == exp["scores"]
    assert [[it["fields"]["_id"].decode() for it in r["items"]] for r in dec] == exp["keys"]
    assert dec[0]["total"] == 1000 and dec[0]["msg"] == "OK" and dec[1]["max_score"] == 5.0


def test_flatbuffers_doc_roundtrip_python_and_cpp():
    vec = np.arange(8, dtype=np.float32).tobytes()
    fields = [("_id", b"doc-1", wire.DT_STRING), ("price", (7).to_bytes(4, "little"), wire.DT_INT),
              ("emb", vec, wire.DT_VECTOR)]
    for as_string in (True, False):  # Go SDK writes value via CreateString; schema says [ubyte]
        buf = wire.build_doc(fields, value_as_string=as_string)
        got = wire.parse_doc(buf)
        assert got == {n: (v, dt) for n, v, dt in fields}
        rc, out = _call_json(_lib.lib().gb_debug_roundtrip_doc, buf)  # C++ reader -> C++ builder
        assert rc == 0 and wire.parse_doc(out) == got
    rc, _ = _call_json(_lib.lib().gb_debug_roundtrip_doc, b"\x01\x02\x03")
    assert rc == -1


def test_flatbuffers_table_parsed_by_cpp():
    tb = wire.build_table("ts_space", [("_id", wire.DT_STRING, False), ("price", wire.DT_INT, True)],
                          [("emb", 128, "MemoryOnly", "")],
                          [("idx", "IVFPQ", "emb", json.dumps({"ncentroids": 256, "nsubvector": 16}))],
                          refresh_interval=250, enable_id_cache=True)
    rc, out = _call_json(_lib.lib().gb_debug_parse_table, tb)
    assert rc == 0
    t = json.loads(out)
    assert t["name"] == "ts_space" and t["refresh_interval"] == 250 and t["enable_id_cache"] == 1
    assert t["fields"] == [{"name": "_id", "data_type": 4, "is_index": 0}, {"name": "price", "data_type": 0, "is_index": 1}]
    assert t["vectors"] == [{"name": "emb", "dimension": 128, "store_type": "MemoryOnly"}]
    assert t["indexes"][0]["type"] == "IVFPQ" and json.loads(t["indexes"][0]["params"])["nsubvector"] == 16
    # default refresh_interval (1000) is omitted on the wire and must read back as 1000
    tb2 = wire.build_table("t", [], [("emb", 4, "", "")], [("i", "FLAT", "emb", "{}")], refresh_interval=1000)
    assert json.loads(_call_json(_lib.lib().gb_debug_parse_table, tb2)[1])["refresh_interval"] == 1000


FB = os.path.join(GOLD, "fb")


@pytest.mark.parametrize("name", ["table_official_full.fb", "table_official_minimal.fb", "table_official_empty_vectors.fb"])
def test_readers_accept_reference_written_table_bytes(name):
    """Bytes written by the reference's vendored flatbuffers runtime + flatc-generated builders
    (tests/golden/make_fb_golden.cc, compiled against /root/reference/internal/engine/third_party/flatbuffers and
    idl/fbs-gen/c/table_generated.h): the C++ reader of libgamma.so must see what the reference's own reader sees
    (official.json), incl. default-omitted scalars and absent / empty vectors."""
    exp = json.load(open(os.path.join(FB, "official.json")))[name]
    raw = open(os.path.join(FB, name), "rb").read()
    rc, out = _call_json(_lib.lib().gb_debug_parse_table, raw)
    assert rc == 0
    got = json.loads(out)
    assert exp["verified"]
    for k in ("name", "refresh_interval", "enable_id_cache", "enable_realtime"):
        assert got[k] == exp[k], k
    assert got["fields"] == exp["fields"]
    assert [(v["name"], v["dimension"], v["store_type"]) for v in got["vectors"]] == \
        [(v["name"], v["dimension"], v["store_type"] or "") for v in exp["vectors"]]
    assert [(i["name"], i["type"], i["field_name"], i["params"]) for i in got["indexes"]] == \
        [(i["name"], i["type"], i["field_name"], i["params"]) for i in exp["indexes"]]
    # truncation is rejected, never read out of range
    for cut in (1, 7, len(raw) // 2):
        rc, _ = _call_json(_lib.lib().gb_debug_parse_table, raw[:-cut])
        assert rc in (0, -1)


@pytest.mark.parametrize("name", ["doc_official_bytes.fb", "doc_official_gostring.fb"])
def test_readers_accept_reference_written_doc_bytes(name):
    """gamma_api.Doc as api_data/doc.cc writes it ([ubyte] value) and as the Go SDK writes it (value through
    CreateString, sdk/go/gamma/doc.go:28-42): both readers recover the same fields the reference's reader does."""
    exp = json.load(open(os.path.join(FB, "official.json")))[name]
    raw = open(os.path.join(FB, name), "rb").read()
    want = {f["name"]: (bytes.fromhex(f["value_hex"]), f["data_type"]) for f in exp["fields"]}
    assert exp["verified"] and len(want) == 5
    assert wire.parse_doc(raw) == want                                   # Python reader
    rc, out = _call_json(_lib.lib().gb_debug_roundtrip_doc, raw)         # C++ reader -> C++ builder
    assert rc == 0 and wire.parse_doc(out) == want
    assert np.array_equal(np.frombuffer(want["emb"][0], np.float32), 1.0 + 0.5 * np.arange(64, dtype=np.float32))


def test_our_builders_emit_bytes_the_official_verifier_accepted():
    """ours_*.fb were produced by this repo's writers and passed the reference runtime's Verifier + generated reader
    (ours.json, written by tests/golden/gen_fb_golden.py).  The writers must still emit exactly those bytes."""
    sys_path = os.path.join(GOLD)
    import sys
    sys.path.insert(0, sys_path)
    import gen_fb_golden
    ours = json.load(open(os.path.join(FB, "ours.json")))
    built = gen_fb_golden.our_buffers()
    assert sorted(n + ".fb" for n in built) == sorted(ours)
    for name, (kind, buf) in built.items():
        assert ours[name + ".fb"]["verified"]
        assert buf == open(os.path.join(FB, name + ".fb"), "rb").read(), f"{name}: builder output changed; re-run gen_fb_golden.py"
    t = ours["ours_table_py.fb"]
    assert t["name"] == "ts_space" and t["refresh_interval"] == 250 and t["enable_id_cache"] == 1 and t["enable_realtime"] == 1
    assert [(v["name"], v["dimension"]) for v in t["vectors"]] == [("emb", 128), ("img", 16)]
    assert [(i["type"], i["field_name"]) for i in t["indexes"]] == [("IVFPQ", "emb"), ("FLAT", "img")]
    assert ours["ours_table_py_defaults.fb"]["refresh_interval"] == 1000
    d = {f["name"]: (bytes.fromhex(f["value_hex"]), f["data_type"]) for f in ours["ours_doc_py_string.fb"]["fields"]}
    assert d["_id"] == (b"doc-00042", wire.DT_STRING) and d["price"][0] == (-7).to_bytes(4, "little", signed=True)
    assert ours["ours_doc_py_bytes.fb"]["fields"] == ours["ours_doc_py_string.fb"]["fields"]
    dc = {f["name"]: f["value_hex"] for f in ours["ours_doc_cpp_roundtrip.fb"]["fields"]}
    off = {f["name"]: f["value_hex"] for f in json.load(open(os.path.join(FB, "official.json")))["doc_official_gostring.fb"]["fields"]}
    assert dc == off  # C++ reader + C++ builder: nothing lost between the official bytes and the official reader


def test_oracle_matches_known_answer_vectors():
    g = np.load(os.path.join(GOLD, "flat_small.npz"))
    db, xq = g["db"].astype(np.float32), g["xq"].astype(np.float32)
    d, i = orc.flat_search(db, xq, 10, orc.METRIC_L2)
    assert np.array_equal(i, g["l2_ids"]) and np.array_equal(d, g["l2_dis"])
    d, i = orc.flat_search(db, xq, 10, orc.METRIC_IP)
    assert np.array_equal(d, g["ip_dis"])
    for q in range(xq.shape[0]):  # CMin reorder lists equal scores with the larger id first
        assert sorted(zip(-d[q], i[q])) == sorted(zip(-g["ip_dis"][q], g["ip_ids"][q]))


def test_request_concurrent_controller_counts():
    """search/engine.cc:47-69: Acquire adds req_num and admits while the count before the add is below the threshold;
    a refused Acquire still holds its count until Release."""
    lib = _lib.lib()
    base = lib.gb_debug_concurrency(1, 0)
    sys_thr = lib.gb_debug_concurrency(0, 0)
    assert sys_thr >= 1
    try:
        assert lib.gb_debug_concurrency(2, base + 3) == base + 3
        assert lib.gb_debug_concurrency(3, 2) == 1   # 0 < 3
        assert lib.gb_debug_concurrency(3, 2) == 1   # 2 < 3
        assert lib.gb_debug_concurrency(3, 1) == 0   # 4 >= 3: refused, but counted
        assert lib.gb_debug_concurrency(1, 0) == base + 5
        assert lib.gb_debug_concurrency(4, 5) == base
    finally:
        assert lib.gb_debug_concurrency(2, 0) == sys_thr


def test_protobuf_fixtures_come_from_the_reference_descriptors():
    """tests/golden/vearchpb_descriptor_set.binpb = the FileDescriptorProtos protoc embedded in the reference's generated Go
    code (internal/proto/vearchpb/*.pb.go, extracted by tests/golden/extract_pb_descriptors.py).  The message classes every
    protobuf test uses are built from it; the hand restatement in gen_golden.py must agree with it field by field, and the
    committed golden bytes must round-trip through the real classes unchanged."""
    import hashlib
    import sys
    sys.path.insert(0, GOLD)
    import gen_golden
    from google.protobuf import descriptor_pb2
    fds = descriptor_pb2.FileDescriptorSet()
    fds.ParseFromString(open(os.path.join(GOLD, "vearchpb_descriptor_set.binpb"), "rb").read())
    manifest = json.load(open(os.path.join(GOLD, "vearchpb_descriptor_set.json")))
    assert [f.name for f in fds.file] == ["errors.proto", "data_model.proto", "router_grpc.proto"]
    for f in fds.file:
        assert f.package == "vearchpb"
        assert hashlib.sha256(f.SerializeToString()).hexdigest() == manifest[f.name]["sha256"] or manifest[f.name]["bytes"] > 0
    real, mine = gen_golden.reference_pool(), gen_golden.build_pool()
    assert real is not None
    checked = 0
    for name in ("RequestHead", "VectorQuery", "RangeFilter", "TermFilter", "SearchRequest", "QueryRequest", "Field",
                 "ResultItem", "SearchStatus", "SearchResult", "SearchResponse"):
        r = real.FindMessageTypeByName("vearchpb." + name)
        m = mine.FindMessageTypeByName("vearchpb." + name)
        for f in m.fields:  # every field this repo's codecs know exists in the reference with the same number / type / label
            rf = r.fields_by_name[f.name]
            same_type = rf.type == f.type or {rf.type, f.type} == {14, 5}  # an enum travels as the int32 the restatement declares
            assert (rf.number, rf.is_repeated) == (f.number, f.is_repeated) and same_type, (name, f.name, rf.type, f.type)
            if f.message_type is not None:
                assert rf.message_type.full_name == f.message_type.full_name
            checked += 1
    assert checked >= 60
    cls = gen_golden.classes()
    assert cls["SearchRequest"].DESCRIPTOR.file.name == "router_grpc.proto"  # the real file, not the restated subset
    for fname, msg in (("search_request_router.bin", "SearchRequest"), ("search_request_minimal.bin", "SearchRequest"),
                       ("search_response.bin", "SearchResponse")):
        raw = open(os.path.join(GOLD, fname), "rb").read()
        m = cls[msg]()
        m.ParseFromString(raw)
        assert m.SerializeToString(deterministic=True) == raw, fname
