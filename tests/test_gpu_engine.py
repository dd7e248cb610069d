"""GPU tests of the gamma C-ABI drop-in (include/gamma_api.h) driven the way the Go partition
server drives it: Init -> CreateTable(flatbuffers) -> AddOrUpdateDoc(flatbuffers) x N ->
BuildIndex / auto-build -> poll GetEngineStatus -> Search(protobuf) -> Dump -> Load
(internal/engine/tests/test.h:1000-1033 TestIndexes, internal/ps/engine/gammacb/*).
Results are checked against the CPU oracle; error codes against the reference's conventions."""
import ctypes as C
import json
import threading

import numpy as np
import pytest

from oracle import oracle as orc
from vearch_b200 import synth, wire

pytestmark = pytest.mark.gpu
L2, IP = orc.METRIC_L2, orc.METRIC_IP
D, N, NQ = 32, 6000, 24


def eng_mod():
    from vearch_b200 import engine
    return engine


@pytest.fixture(scope="module")
def data():
    return synth.sift_like(N, D, seed=81), synth.sift_like(NQ, D, seed=82)


def keys_of(res):
    return [[it["fields"]["_id"].decode() for it in r["items"]] for r in res]


def scores_of(res):
    return [[it["score"] for it in r["items"]] for r in res]


def make_engine(tmp_path, index_type, params, name="ts"):
    e = eng_mod().GammaEngine(str(tmp_path), space_name=name)
    e.create_table(name, D, index_type, params, fields=(("_id", wire.DT_STRING, False), ("tag", wire.DT_STRING, False)))
    return e


def add_all(e, db, start=0):
    for i, v in enumerate(db):
        assert e.add_doc(f"doc{start + i}", v, extra_fields=[("tag", f"t{(start + i) % 7}".encode(), wire.DT_STRING)]) == 0


def test_ivfflat_full_flow_matches_oracle(tmp_path, data):
    db, xq = data
    e = make_engine(tmp_path, "IVFFLAT", {"ncentroids": 32, "nprobe": 8, "metric_type": "L2", "training_threshold": 2000})
    assert e.status() == {"backup_status": 0, "doc_num": 0, "index_status": 0, "max_docid": -1, "min_indexed_num": 0}
    add_all(e, db[:1500])
    # untrained, > 100 docs, brute_force_search == 0  =>  IndexNotTrained (search/engine.cc:285-302)
    with pytest.raises(eng_mod().GammaStatusError) as ei:
        e.search(xq, 10)
    assert ei.value.code == 2 and "index not trained" in ei.value.msg
    # is_brute_search = 1 works before training and is exact (FLAT path, gamma_index_flat.cc)
    res = e.search(xq, 10, is_brute_search=1)
    do, io = orc.flat_search(db[:1500], xq, 10, L2)
    assert keys_of(res) == [[f"doc{i}" for i in row] for row in io]
    assert np.array_equal(np.array(scores_of(res), np.float32), do)
    assert res[0]["total"] == 1500 and res[0]["msg"] == "OK" and res[0]["max_score"] == max(scores_of(res)[0])
    add_all(e, db[1500:], start=1500)  # crossing training_threshold auto-starts indexing (engine.cc:753-761)
    st = e.wait_indexed(N)
    assert st["index_status"] == 2 and st["doc_num"] == N and st["max_docid"] == N - 1
    res = e.search(xq, 10, index_params={"nprobe": 8})
    _, gt = orc.flat_search(db, xq, 10, L2)
    got = keys_of(res)
    hit = np.mean([f"doc{gt[q, 0]}" in got[q] for q in range(NQ)])
    assert hit >= 0.9  # reference CI pin for IVFFLAT r@10 (test/test_vector_index_ivfflat.py:89-94)
    for q in range(NQ):  # returned score == sum (x - y)^2 of the returned doc (test_module_vector.py:337-364)
        for key, s in zip(got[q], scores_of(res)[q]):
            assert s == float(((xq[q] - db[int(key[3:])]) ** 2).sum())
        assert scores_of(res)[q] == sorted(scores_of(res)[q])
    # auto mode (2) on an indexed table == normal search; offset trims the head (vector_manager.cc:1054-1068)
    res2 = e.search(xq, 5, index_params={"nprobe": 8}, is_brute_search=2, offset=3)
    assert keys_of(res2) == [row[3:8] for row in keys_of(e.search(xq, 8, index_params={"nprobe": 8}))]
    # requested fields
    res3 = e.search(xq[:1], 3, fields=("_id", "tag", "emb"))
    it = res3[0]["items"][0]
    docid = int(it["fields"]["_id"].decode()[3:])
    assert it["fields"]["tag"] == f"t{docid % 7}".encode()
    assert np.array_equal(np.frombuffer(it["fields"]["emb"], np.float32), db[docid])
    mi = e.memory_info()
    assert mi["vector_mem"] > 0 and mi["index_mem"] > 0 and set(mi) == {"table_mem", "index_mem", "vector_mem",
                                                                        "field_range_mem", "bitmap_mem"}
    e.close()


def test_error_conventions(tmp_path, data):
    db, xq = data
    E = eng_mod()
    e = E.GammaEngine(str(tmp_path))
    with pytest.raises(E.GammaStatusError) as ei:  # search before CreateTable
        e.vec_name = "emb"
        e.search(xq, 10)
    assert ei.value.code == 4
    with pytest.raises(E.GammaStatusError) as ei:  # unknown index type
        e.create_table("t", D, "HNSW", {})
    assert ei.value.code == 3
    with pytest.raises(E.GammaStatusError) as ei:  # nprobe > ncentroids (gamma_index_ivfflat.cc:95-99)
        e.create_table("t", D, "IVFFLAT", {"ncentroids": 8, "nprobe": 32})
    assert ei.value.code == 4
    with pytest.raises(E.GammaStatusError) as ei:  # d % nsubvector != 0 (gamma_index_ivfpq.cc:125-133)
        e.create_table("t", D, "IVFPQ", {"ncentroids": 8, "nprobe": 4, "nsubvector": 5})
    assert ei.value.code == 4 and "cannot divide by nsubvector" in ei.value.msg
    e.create_table("t", D, "FLAT", {"metric_type": "L2"})
    with pytest.raises(E.GammaStatusError) as ei:
        e.create_table("t", D, "FLAT", {})
    assert ei.value.code == 4
    assert e.add_doc("a", db[0]) == 0
    assert e.add_doc("b", db[1][:-1]) == -3  # CheckDoc: wrong vector length (engine.cc:802-813)
    assert e.add_doc_raw(wire.build_doc([("_id", b"c", wire.DT_STRING)])) == -3  # no vector field
    assert e.delete_doc("nope") == -1
    for kw, code in ((dict(topn=0), 4), (dict(topn=5, req_num=0), 4)):
        with pytest.raises(E.GammaStatusError) as ei:
            e.search(xq[:1], kw.pop("topn"), **kw)
        assert ei.value.code == code
    with pytest.raises(E.GammaStatusError) as ei:
        e.search_raw(wire.encode_search_request("other_field", xq[:1], 5))
    assert ei.value.code == 4
    with pytest.raises(E.GammaStatusError) as ei:
        e.search_raw(b"\x0a\xff\xff")  # malformed protobuf
    assert ei.value.code == 4
    assert e.query()["items"] == []  # Query without ids or filters: one empty result (engine.cc:443-521)
    # FLAT, few docs, not "indexed": allowed because max_docid <= 100 (brute_force_search_threshold)
    res = e.search(xq[:1], 5)
    assert keys_of(res) == [["a"]] or keys_of(res) == [["a"]]
    e.close()


def test_delete_update_getdoc(tmp_path, data):
    db, xq = data
    e = make_engine(tmp_path, "IVFFLAT", {"ncentroids": 16, "nprobe": 16, "metric_type": "L2", "training_threshold": 1000})
    add_all(e, db[:3000])
    e.wait_indexed(3000)
    base = keys_of(e.search(xq, 5, index_params={"nprobe": 16}))
    victim = base[0][0]
    assert e.delete_doc(victim) == 0 and e.delete_doc(victim) == -1  # key mapping removed (table_->Delete)
    assert e.status()["doc_num"] == 2999
    after = keys_of(e.search(xq, 5, index_params={"nprobe": 16}))
    assert victim not in after[0]
    vid = int(victim[3:])
    alive = np.ones(3000, bool)
    alive[vid] = False
    delb = np.packbits(~alive, bitorder="little")
    do, io = orc.flat_search(db[:3000], xq, 5, L2, del_bitmap=delb)
    assert after == [[f"doc{i}" for i in row] for row in io]  # nprobe == nlist => exact
    rc, doc = e.get_doc_by_id(victim)
    assert rc == -1 and doc == {}
    rc, doc = e.get_doc_by_id("doc5")
    assert rc == 0 and doc["_id"][0] == b"doc5" and doc["tag"][0] == b"t5"
    assert np.array_equal(np.frombuffer(doc["emb"][0], np.float32), db[5]) and doc["emb"][1] == wire.DT_VECTOR
    rc, doc = e.get_doc_by_docid(vid - 1, next_=True)  # next undeleted docid after vid-1 skips the victim
    assert rc == 0 and int.from_bytes(doc["_docid"][0], "little") == vid + 1
    assert e.get_doc_by_docid(vid)[0] == -1 and e.get_doc_by_docid(10 ** 6)[0] == -1
    # update: same key, new vector -> old entry tombstoned, new vector searchable (engine.cc:774-850)
    target = xq[3] + 1.0
    assert e.add_doc("doc7", target) == 0
    assert e.status()["doc_num"] == 2999
    res = e.search(xq[3:4], 1, index_params={"nprobe": 16})
    assert keys_of(res) == [["doc7"]] and scores_of(res)[0][0] == float(D)
    rc, doc = e.get_doc_by_id("doc7")
    assert np.array_equal(np.frombuffer(doc["emb"][0], np.float32), target)
    # re-adding a deleted key creates a fresh doc
    assert e.add_doc(victim, db[vid]) == 0
    e.wait_indexed(3001)
    assert e.status()["doc_num"] == 3000 and e.status()["max_docid"] == 3000
    e.close()


@pytest.mark.parametrize("index_type,params", [
    ("FLAT", {"metric_type": "InnerProduct"}),
    ("IVFPQ", {"ncentroids": 16, "nprobe": 16, "nsubvector": 8, "metric_type": "L2", "training_threshold": 1000}),
])
def test_flat_ip_and_ivfpq_tables(tmp_path, data, index_type, params):
    db, xq = data
    e = make_engine(tmp_path, index_type, params)
    add_all(e, db[:2500])
    assert e.build_index() == 0 and e.build_index() == 0  # idempotent while running (engine.cc:953-976)
    e.wait_indexed(2500)
    if index_type == "FLAT":
        res = e.search(xq, 10)
        do, io = orc.flat_search(db[:2500], xq, 10, IP)
        assert np.array_equal(np.array(scores_of(res), np.float32), do)
        assert all(sorted(a) == sorted(f"doc{i}" for i in b) for a, b in zip(keys_of(res), io))
    else:
        res = e.search(xq, 10, index_params={"nprobe": 16, "recall_num": 100})
        _, gt = orc.flat_search(db[:2500], xq, 1, L2)
        got = keys_of(res)
        assert np.mean([f"doc{gt[q, 0]}" in got[q] for q in range(NQ)]) >= 0.9
        for q in range(NQ):  # re-ranked scores are exact
            assert scores_of(res)[q][0] == float(((xq[q] - db[int(got[q][0][3:])]) ** 2).sum())
    e.close()


def test_dump_load_roundtrip(tmp_path, data):
    db, xq = data
    params = {"ncentroids": 16, "nprobe": 4, "metric_type": "L2", "training_threshold": 1000}
    e = make_engine(tmp_path, "IVFFLAT", params)
    add_all(e, db[:2000])
    e.wait_indexed(2000)
    e.delete_doc("doc11")
    before = e.search(xq, 10, index_params={"nprobe": 4})
    assert e.dump() == 0
    e.close()
    # the index travels in gamma's own file format (gamma_index_ivfflat.cc:807-839), readable by the
    # independent restatement of that format
    import gamma_index_file as gif
    with open(tmp_path / "retrieval_model_index" / "emb.000" / "ivfflat.index", "rb") as fh:
        dumped = gif.read_index_file(fh.read())
    assert dumped["kind"] == "IvFl" and dumped["nlist"] == 16 and dumped["indexed_count"] == 2000
    assert len(dumped["ids"]) == 2000 and sorted(dumped["ids"].tolist()) == list(range(2000))
    e2 = make_engine(tmp_path, "IVFFLAT", params)  # gammacb.New: CreateTable then Load (gamma.go:104-130)
    assert e2.load() == 0
    st = e2.status()  # lists came from the file: indexed before the indexing thread has done anything
    assert st["min_indexed_num"] == 2000 and st["index_status"] == 2
    e2.wait_indexed(2000)
    assert e2.status()["doc_num"] == 1999
    after = e2.search(xq, 10, index_params={"nprobe": 4})
    assert keys_of(after) == keys_of(before) and scores_of(after) == scores_of(before)
    assert e2.get_doc_by_id("doc11")[0] == -1 and e2.get_doc_by_id("doc12")[0] == 0
    e2.close()


def test_kill_switch_and_config(tmp_path, data):
    db, xq = data
    E = eng_mod()
    e = make_engine(tmp_path, "FLAT", {"metric_type": "L2"})
    add_all(e, db[:50])
    api = E._api()
    api.SetKillStatus(b"rq-1", 9, 1)
    with pytest.raises(E.GammaStatusError) as ei:  # killed request -> MemoryExceeded, code 8 (reader.go:170-174)
        e.search(xq[:1], 3, request_id="rq-1", partition_id=9)
    assert ei.value.code == 8
    assert len(e.search(xq[:1], 3, request_id="rq-2", partition_id=9)[0]["items"]) == 3
    api.DeleteKillStatus(b"rq-1", 9)
    assert len(e.search(xq[:1], 3, request_id="rq-1", partition_id=9)[0]["items"]) == 3
    assert e.set_config({"refresh_interval": 50, "slow_search_time": 7}) == 0
    cfg = e.get_config()
    assert cfg["refresh_interval"] == 50 and cfg["slow_search_time"] == 7
    e.close()


def test_reference_written_flatbuffers_drive_the_engine(tmp_path):
    """Boundary pin: gamma_api.Table / gamma_api.Doc buffers written by the REFERENCE's own flatc-generated
    builders (tests/golden/make_fb_golden.cc, compiled against /root/reference's vendored flatbuffers + idl/fbs-gen/c) go
    through CreateTable / AddOrUpdateDoc / GetDocByID / Search exactly as the Go partition server would send them --
    including the Go SDK's habit of writing `value` with CreateString (sdk/go/gamma/doc.go:28-42)."""
    import json as _json
    import os
    E = eng_mod()
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fb")
    meta = _json.load(open(os.path.join(gold, "official.json")))
    e = E.GammaEngine(str(tmp_path / "fbref"), space_name="ts_space")
    e.create_table_raw(open(os.path.join(gold, "table_official_full.fb"), "rb").read(), "emb", 64)
    # a second CreateTable must be refused like the reference does (engine.cc: "table is created")
    with pytest.raises(E.GammaStatusError):
        e.create_table_raw(open(os.path.join(gold, "table_official_minimal.fb"), "rb").read(), "emb", 64)
    exp = {f["name"]: bytes.fromhex(f["value_hex"]) for f in meta["doc_official_bytes.fb"]["fields"]}
    assert e.add_doc_raw(open(os.path.join(gold, "doc_official_bytes.fb"), "rb").read()) == 0
    assert e.status()["doc_num"] == 1
    # the same document again, `value` written as a flatbuffers string: an update of the same _id, not a second doc
    assert e.add_doc_raw(open(os.path.join(gold, "doc_official_gostring.fb"), "rb").read()) == 0
    assert e.status()["doc_num"] == 1
    rc, doc = e.get_doc_by_id("doc-00042")
    assert rc == 0
    for name in ("_id", "price", "tag", "emb", "img"):
        assert doc[name][0] == exp[name], name
    emb = np.frombuffer(exp["emb"], np.float32)
    res = e.search(emb[None, :], 1, is_brute_search=1)
    assert res[0]["items"][0]["fields"]["_id"] == b"doc-00042" and res[0]["items"][0]["score"] == 0.0
    # the second vector field of the table (FLAT, inner product) is searchable too
    img = np.frombuffer(exp["img"], np.float32)
    req = wire.encode_search_request("img", img[None, :], 1, is_brute_search=1)
    r2 = wire.decode_search_response(e.search_raw(req))
    assert r2[0]["items"][0]["fields"]["_id"] == b"doc-00042" and abs(r2[0]["items"][0]["score"] - float(img @ img)) < 1e-3
    e.close()


def test_admission_control_refuses_with_resource_exhausted(tmp_path, data):
    """RequestConcurrentController (search/engine.cc:47-119, :252-260): a Search arriving while the in-flight request
    count is at the threshold fails with Status::ResourceExhausted() = code kBusy (6), "Resource busy: Resource
    temporarily unavailable", and gives its count back; below the threshold it is admitted."""
    db, xq = data
    E = eng_mod()
    e = make_engine(tmp_path, "FLAT", {"metric_type": "L2"})
    add_all(e, db[:50])
    api = E._api()
    base = api.gb_debug_concurrency(1, 0)
    assert api.gb_debug_concurrency(0, 0) >= 1  # system-derived threshold
    try:
        assert api.gb_debug_concurrency(2, base + 4) == base + 4
        assert api.gb_debug_concurrency(3, 4) == 1  # four requests in flight elsewhere: admitted, threshold reached
        with pytest.raises(E.GammaStatusError) as ei:
            e.search(xq[:1], 3)
        assert ei.value.code == 6 and ei.value.msg == "Resource busy: Resource temporarily unavailable"
        assert api.gb_debug_concurrency(1, 0) == base + 4  # the refused request released its count
        api.gb_debug_concurrency(4, 4)
        assert len(e.search(xq[:2], 3)) == 2
        assert api.gb_debug_concurrency(1, 0) == base
    finally:
        api.gb_debug_concurrency(2, 0)
    e.close()


def test_enable_realtime_searches_unindexed_tail(tmp_path, data):
    """table.enable_realtime: documents that are stored but not yet picked up by the indexing thread
    are searched brute-force and merged (vector_manager.cc:854-889, 971-1053)."""
    db, xq = data
    e = eng_mod().GammaEngine(str(tmp_path), space_name="rt")
    e.create_table("rt", D, "IVFFLAT", {"ncentroids": 16, "nprobe": 16, "metric_type": "L2", "training_threshold": 2000},
                   refresh_interval=600000, enable_realtime=True)  # the indexing loop sleeps 10 min between passes
    add_all(e, db[:2000])
    e.wait_indexed(2000)
    add_all(e, db[2000:2300], start=2000)  # these stay un-indexed during the test
    assert e.status()["min_indexed_num"] == 2000
    res = e.search(db[2100:2104], 1, index_params={"nprobe": 16})
    assert keys_of(res) == [[f"doc{i}"] for i in range(2100, 2104)] and all(s[0] == 0.0 for s in scores_of(res))
    res = e.search(xq, 10, index_params={"nprobe": 16})
    do, io = orc.flat_search(db[:2300], xq, 10, L2)
    assert keys_of(res) == [[f"doc{i}" for i in row] for row in io]  # nprobe == nlist + tail => exact
    e.close()
    # without enable_realtime the tail is invisible until the next indexing pass (reference behaviour)
    e2 = eng_mod().GammaEngine(str(tmp_path / "b"), space_name="nrt")
    e2.create_table("nrt", D, "IVFFLAT", {"ncentroids": 16, "nprobe": 16, "metric_type": "L2", "training_threshold": 2000},
                    refresh_interval=600000)
    add_all(e2, db[:2000])
    e2.wait_indexed(2000)
    add_all(e2, db[2000:2300], start=2000)
    res = e2.search(db[2100:2101], 1, index_params={"nprobe": 16})
    assert keys_of(res) != [["doc2100"]]
    e2.close()


def test_concurrent_search_while_adding(tmp_path, data):
    db, xq = data
    e = make_engine(tmp_path, "IVFFLAT", {"ncentroids": 16, "nprobe": 8, "metric_type": "L2", "training_threshold": 1000})
    add_all(e, db[:2000])
    e.wait_indexed(2000)
    errors = []

    def searcher():
        try:
            for _ in range(30):
                res = e.search(xq[:4], 5, index_params={"nprobe": 8})
                assert all(len(r["items"]) == 5 for r in res)
        except Exception as ex:  # noqa: BLE001
            errors.append(ex)

    ts = [threading.Thread(target=searcher) for _ in range(4)]
    for t in ts:
        t.start()
    add_all(e, db[2000:4000], start=2000)  # raft apply thread keeps writing (raft_state_machine.go:128)
    for t in ts:
        t.join()
    assert not errors
    e.wait_indexed(4000)
    res = keys_of(e.search(db[3999:4000], 1, index_params={"nprobe": 16}))
    assert res == [["doc3999"]]
    e.close()


def test_scalar_filters_restrict_the_candidate_set(tmp_path, data):
    """range / term filters (search/engine.cc:349-366, table/scalar_index_manager.cc:294-345, 588-651):
    the engine turns them into the dense docid bitmap every scan kernel applies before selection, so a
    filtered search is the exact search over the surviving documents."""
    import struct
    db, xq = data
    n = 3000
    e = eng_mod().GammaEngine(str(tmp_path), space_name="ts")
    e.create_table("ts", D, "IVFFLAT", {"ncentroids": 16, "nprobe": 16, "metric_type": "L2", "training_threshold": 1000},
                   fields=(("_id", wire.DT_STRING, False), ("price", wire.DT_INT, True), ("tag", wire.DT_STRING, True),
                           ("cats", wire.DT_STRINGARRAY, True), ("plain", wire.DT_INT, False)))
    price = (np.arange(n) * 37) % 1000
    tags = np.array([f"t{i % 7}" for i in range(n)])
    for i in range(n):
        cats = f"c{i % 5}\x01c{(i // 5) % 11}".encode()
        assert e.add_doc(f"doc{i}", db[i], extra_fields=[("price", struct.pack("<i", int(price[i])), wire.DT_INT),
                                                          ("tag", tags[i].encode(), wire.DT_STRING),
                                                          ("cats", cats, wire.DT_STRINGARRAY),
                                                          ("plain", struct.pack("<i", i), wire.DT_INT)]) == 0
    e.wait_indexed(n)
    e.delete_doc("doc7")
    alive = np.ones(n, bool)
    alive[7] = False

    def check(mask, **kw):
        res = e.search(xq, 5, index_params={"nprobe": 16}, **kw)
        fb = np.packbits(mask & alive, bitorder="little")
        do, io = orc.flat_search(db[:n], xq, 5, L2, filter_bitmap=fb)
        assert keys_of(res) == [[f"doc{i}" for i in row if i >= 0] for row in io]
        assert scores_of(res) == [[float(s) for s, i in zip(drow, irow) if i >= 0] for drow, irow in zip(do, io)]

    i32 = lambda v: struct.pack("<i", v)
    check((price >= 100) & (price < 300), range_filters=[("price", i32(100), i32(300), True, False)])
    check(price > 900, range_filters=[("price", i32(900), b"", False, False)])
    check(price <= 36, range_filters=[("price", b"", i32(36), False, True)])
    check(price == 370, range_filters=[("price", i32(370), i32(370), True, True)])
    check(price != 370, range_filters=[("price", i32(370), i32(370), True, True, 2)])
    check((tags == "t1") | (tags == "t4"), term_filters=[("tag", b"t1\x01t4")])
    check(~((tags == "t1") | (tags == "t4")), term_filters=[("tag", b"t1\x01t4", 2)])
    cats_hit = np.array([(i % 5 == 2) or ((i // 5) % 11 in (2, 9)) for i in range(n)])  # any element in {c2, c9}
    check(cats_hit, term_filters=[("cats", b"c2\x01c9")])
    both = (price < 500) & (tags == "t3")
    check(both, range_filters=[("price", b"", i32(500), False, False)], term_filters=[("tag", b"t3")])
    check((price < 50) | (tags == "t3"), range_filters=[("price", b"", i32(50), False, False)], term_filters=[("tag", b"t3")],
          operator=1)
    # nothing survives / the field has no scalar index: req_num empty results, not an error
    for kw in (dict(range_filters=[("price", i32(5000), b"", True, False)]),
               dict(range_filters=[("plain", i32(5), b"", True, False)]),
               dict(term_filters=[("nofield", b"x")])):
        res = e.search(xq, 5, **kw)
        assert len(res) == NQ and all(r["items"] == [] and "no result" in r["msg"] for r in res)
    # AddFieldIndexWithParams / RemoveFieldIndex switch a field's filterability (engine.cc:1471-1600)
    e.add_field_index("plain")
    res = e.search(xq, 5, index_params={"nprobe": 16}, range_filters=[("plain", i32(2990), b"", True, False)])
    assert all(int(k[3:]) >= 2990 for row in keys_of(res) for k in row) and len(keys_of(res)[0]) == 5
    e.remove_field_index("plain")
    assert e.search(xq, 5, range_filters=[("plain", i32(2990), b"", True, False)])[0]["items"] == []
    with pytest.raises(eng_mod().GammaStatusError) as ei:
        e.add_field_index("emb")
    assert ei.value.code == 3
    # brute-force path (FLAT kernel) honours the same bitmap
    res = e.search(xq, 5, is_brute_search=1, range_filters=[("price", i32(100), i32(300), True, False)])
    fb = np.packbits((price >= 100) & (price < 300) & alive, bitorder="little")
    _, io = orc.flat_search(db[:n], xq, 5, L2, filter_bitmap=fb)
    assert keys_of(res) == [[f"doc{i}" for i in row] for row in io]
    e.close()


def test_process_exit_without_close_is_clean(tmp_path):
    """A partition server that is killed never calls Close: the library's own threads (indexing loop,
    request coalescer) must be parked before the CUDA runtime unloads, or the process dies in a
    signal instead of exiting."""
    import subprocess
    import sys
    code = f"""
import numpy as np
from vearch_b200 import engine, synth, wire
e = engine.GammaEngine({str(tmp_path)!r}, space_name="ts")
e.create_table("ts", 32, "IVFFLAT", {{"ncentroids": 16, "nprobe": 4, "metric_type": "L2", "training_threshold": 1000}},
               refresh_interval=20)
db = synth.sift_like(1500, 32, seed=5)
for i, v in enumerate(db):
    assert e.add_doc(f"doc{{i}}", v) == 0
e.wait_indexed(1500)
assert len(e.search(db[:1], 3)[0]["items"]) == 3   # nq = 1: goes through the coalescer thread
print("leaving without close", flush=True)
"""
    root = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=300)
    assert "leaving without close" in p.stdout, p.stderr[-2000:]
    assert p.returncode == 0, (p.returncode, p.stderr[-2000:])


def test_query_by_ids_and_by_filters(tmp_path, data):
    """Query (search/engine.cc:404-523): by key, by docid (partition_id > 0), by scalar filters with
    limit / offset; deleted documents never come back."""
    import struct
    db, _ = data
    e = eng_mod().GammaEngine(str(tmp_path), space_name="ts")
    e.create_table("ts", D, "FLAT", {"metric_type": "L2"},
                   fields=(("_id", wire.DT_STRING, False), ("n", wire.DT_LONG, True), ("tag", wire.DT_STRING, True)))
    for i in range(200):
        assert e.add_doc(f"doc{i}", db[i], extra_fields=[("n", struct.pack("<q", i * 3), wire.DT_LONG),
                                                          ("tag", f"t{i % 4}".encode(), wire.DT_STRING)]) == 0
    e.delete_doc("doc9")
    r = e.query(document_ids=["doc5", "nope", "doc9", "doc7"], fields=["_id", "n"])
    assert [it["fields"]["_id"] for it in r["items"]] == [b"doc5", b"doc7"] and r["total"] == 2
    assert struct.unpack("<q", r["items"][1]["fields"]["n"])[0] == 21 and "tag" not in r["items"][0]["fields"]
    r = e.query(document_ids=["3", "9", "4000", "x"], partition_id=1)  # docids; every field comes back
    assert [it["fields"]["_id"] for it in r["items"]] == [b"doc3"] and set(r["items"][0]["fields"]) == {"_id", "n", "tag"}
    q8 = lambda v: struct.pack("<q", v)
    r = e.query(range_filters=[("n", q8(12), q8(60), True, True)], term_filters=[("tag", b"t1\x01t3")], limit=100,
                fields=["_id"])
    want = [i for i in range(200) if 12 <= i * 3 <= 60 and i % 4 in (1, 3) and i != 9]
    assert [it["fields"]["_id"] for it in r["items"]] == [f"doc{i}".encode() for i in want]
    r = e.query(term_filters=[("tag", b"t2")], limit=3, offset=2, fields=["_id"])
    assert [it["fields"]["_id"] for it in r["items"]] == [b"doc10", b"doc14", b"doc18"]
    r = e.query(term_filters=[("tag", b"zzz")], limit=5)
    assert r["items"] == [] and "no result" in r["msg"]
    e.close()


def test_multi_vector_table_join_and_weighted_rank(tmp_path, data):
    """Two vector fields, each with its own index: a request with two vec_fields searches both, keeps the
    documents BOTH returned and scores them with the WeightedRanker (vector_manager.cc:747, 900-964)."""
    db, xq = data
    n, d2, topn = 2500, 16, 60
    img = synth.sift_like(n, d2, seed=91)
    imq = synth.sift_like(NQ, d2, seed=92)
    e = eng_mod().GammaEngine(str(tmp_path), space_name="ts")
    e.create_table("ts", D, "IVFFLAT", {"ncentroids": 16, "nprobe": 16, "metric_type": "L2", "training_threshold": 1000},
                   extra_vectors=[("img", d2, "FLAT", {"metric_type": "L2"})])
    assert e.add_doc("bad", db[0]) == -3  # CheckDoc: every vector field must be present (engine.cc:802-813)
    for i in range(n):
        assert e.add_doc(f"doc{i}", db[i], extra_fields=[("img", img[i].tobytes(), wire.DT_VECTOR)]) == 0
    st = e.wait_indexed(n)
    assert st["min_indexed_num"] == n
    rc, doc = e.get_doc_by_id("doc3")
    assert rc == 0 and np.array_equal(np.frombuffer(doc["img"][0], np.float32), img[3])
    assert np.array_equal(np.frombuffer(doc["emb"][0], np.float32), db[3])
    # one field at a time: either field can be the query
    d1, i1 = orc.flat_search(db[:n], xq, topn, L2)
    d2s, i2 = orc.flat_search(img, imq, topn, L2)
    w = np.array([0.3, 0.7], np.float32)

    def expect(rank):
        out = []
        for q in range(NQ):
            m1 = dict(zip(i1[q].tolist(), d1[q].tolist()))
            m2 = dict(zip(i2[q].tolist(), d2s[q].tolist()))
            both = sorted(set(m1) & set(m2))
            sc = [float(np.float32(float(np.float32(m1[k]) * w[0]) + float(np.float32(m2[k]) * w[1]))) for k in both]
            pairs = list(zip(both, sc))
            if rank:
                pairs.sort(key=lambda t: t[1])  # stable: docid order among equal scores
            out.append(pairs)
        return out

    for rank in (0, 1):
        res = e.search(xq, topn, is_brute_search=1, extra_vec_queries=[("img", imq)],
                       ranker='{"type": "WeightedRanker", "params": [0.3, 0.7]}', multi_vector_rank=rank)
        exp = expect(rank)
        assert keys_of(res) == [[f"doc{k}" for k, _ in row] for row in exp]
        got_sc = scores_of(res)
        for q in range(NQ):
            assert np.allclose(got_sc[q], [s for _, s in exp[q]], rtol=1e-6)
    assert sum(len(r) for r in exp) > NQ  # the join is not trivially empty
    # default weights 1 / vec_num; bad ranker -> InvalidArgument (common_query_data.h:257-300)
    res = e.search(xq[:2], topn, is_brute_search=1, extra_vec_queries=[("img", imq[:2])], multi_vector_rank=1)
    assert len(res) == 2 and all(len(r["items"]) > 0 for r in res)
    with pytest.raises(eng_mod().GammaStatusError) as ei:
        e.search(xq[:2], topn, is_brute_search=1, extra_vec_queries=[("img", imq[:2])], ranker='{"type": "WeightedRanker", "params": [1.0]}')
    assert ei.value.code == 4 and "length don't equal" in ei.value.msg
    # dump / load keeps both fields
    before = keys_of(e.search(xq, topn, is_brute_search=1, extra_vec_queries=[("img", imq)], multi_vector_rank=1))
    assert e.dump() == 0
    e.close()
    e2 = eng_mod().GammaEngine(str(tmp_path), space_name="ts")
    e2.create_table("ts", D, "IVFFLAT", {"ncentroids": 16, "nprobe": 16, "metric_type": "L2", "training_threshold": 1000},
                    extra_vectors=[("img", d2, "FLAT", {"metric_type": "L2"})])
    assert e2.load() == 0
    e2.wait_indexed(n)
    after = keys_of(e2.search(xq, topn, is_brute_search=1, extra_vec_queries=[("img", imq)], multi_vector_rank=1))
    assert after == before
    e2.close()


def test_rebuild_index_retrains_and_reindexes(tmp_path, data):
    """RebuildIndex (search/engine.cc:991-1089): a no-op on an engine that is not indexing; otherwise the
    index of every vector field is dropped, trained again on the current vectors and refilled."""
    db, xq = data
    e = make_engine(tmp_path, "IVFPQ", {"ncentroids": 16, "nprobe": 16, "nsubvector": 8, "metric_type": "L2",
                                        "training_threshold": 1000})
    assert e.rebuild_index() == 0  # "index not running, no need to rebuild!"
    add_all(e, db[:2500])
    e.wait_indexed(2500)
    before = e.search(xq, 10, index_params={"nprobe": 16, "recall_num": 100})
    for i in range(0, 300):
        e.delete_doc(f"doc{i}")
    assert e.rebuild_index() == 0
    st = e.wait_indexed(2500)
    assert st["index_status"] == 2 and st["doc_num"] == 2200
    after = e.search(xq, 10, index_params={"nprobe": 16, "recall_num": 100})
    alive = np.ones(2500, bool)
    alive[:300] = False
    _, gt = orc.flat_search(db[:2500], xq, 1, L2, del_bitmap=np.packbits(~alive, bitorder="little"))
    got = keys_of(after)
    assert np.mean([f"doc{gt[q, 0]}" in got[q] for q in range(NQ)]) >= 0.9
    assert all(int(k[3:]) >= 300 for row in got for k in row) and len(before) == len(after)
    assert e.rebuild_index(describe=1) == 0 and e.status()["index_status"] == 2  # describe: the thread is only stopped
    e.close()
