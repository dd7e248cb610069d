#!/usr/bin/env python
"""Regenerates tests/golden/fb/ (run in the build container, where /root/reference exists):

  * official_*.fb      Table / Doc buffers written by the reference's vendored flatbuffers runtime and its
                       flatc-generated builders (make_fb_golden.cc `emit`), incl. the Go SDK's CreateString quirk
  * official.json      what the reference's generated READER sees in them (make_fb_golden.cc `check`)
  * ours_*.fb          buffers produced by THIS repo's builders (vearch_b200/wire.py and the C++ FbBuilder
                       in vearch_b200/csrc/wire.h through gb_debug_roundtrip_doc)
  * ours.json          the official Verifier's verdict on them + what the generated reader decodes

tests/test_boundary_cpu.py then (a) feeds the official bytes to our C++ and Python readers and (b) checks that
our builders still emit exactly the bytes the official Verifier accepted.  Nothing here reads /root/reference
at test time."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/internal/engine"
OUT = os.path.join(HERE, "fb")
TOOL = "/tmp/make_fb_golden"


def our_buffers():
    """name -> (kind, bytes) built by this repository's writers"""
    from vearch_b200 import _lib, wire
    vec = (np.arange(64, dtype=np.float32) * 0.5 + 1.0).tobytes()
    fields = [("_id", b"doc-00042", wire.DT_STRING), ("price", (-7).to_bytes(4, "little", signed=True), wire.DT_INT),
              ("tag", b"red\x01blue", wire.DT_STRING), ("emb", vec, wire.DT_VECTOR)]
    out = {}
    out["ours_doc_py_string"] = ("doc", wire.build_doc(fields, value_as_string=True))
    out["ours_doc_py_bytes"] = ("doc", wire.build_doc(fields, value_as_string=False))
    # the C++ builder: parse an OFFICIAL doc with the C++ reader and re-emit it with the C++ FbBuilder
    lib = _lib.lib()
    fn = lib.gb_debug_roundtrip_doc
    fn.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    src = open(os.path.join(OUT, "doc_official_gostring.fb"), "rb").read()
    p, n = C.c_void_p(), C.c_int()
    assert fn(src, len(src), C.byref(p), C.byref(n)) == 0
    out["ours_doc_cpp_roundtrip"] = ("doc", C.string_at(p.value, n.value))
    libc.free(p.value)
    out["ours_table_py"] = ("table", wire.build_table(
        "ts_space", [("_id", wire.DT_STRING, False), ("price", wire.DT_INT, True)],
        [("emb", 128, "MemoryOnly", ""), ("img", 16, "", "")],
        [("idx", "IVFPQ", "emb", json.dumps({"ncentroids": 256, "nsubvector": 16})), ("idx2", "FLAT", "img", "{}")],
        refresh_interval=250, enable_id_cache=True, enable_realtime=True))
    out["ours_table_py_defaults"] = ("table", wire.build_table("t", [], [("emb", 4, "", "")], [("i", "FLAT", "emb", "{}")]))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", f"-I{REF}/third_party", f"-I{REF}/idl/fbs-gen/c",
                           os.path.join(HERE, "make_fb_golden.cc"), "-o", TOOL])
    subprocess.check_call([TOOL, "emit", OUT])
    official = {}
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".fb") and "_official_" in f:
            kind = "table" if f.startswith("table") else "doc"
            official[f] = json.loads(subprocess.check_output([TOOL, "check", os.path.join(OUT, f), kind]))
    json.dump(official, open(os.path.join(OUT, "official.json"), "w"), indent=1, sort_keys=True)
    ours = {}
    for name, (kind, buf) in our_buffers().items():
        path = os.path.join(OUT, name + ".fb")
        open(path, "wb").write(buf)
        ours[name + ".fb"] = json.loads(subprocess.check_output([TOOL, "check", path, kind]))
        assert ours[name + ".fb"]["verified"], name
    json.dump(ours, open(os.path.join(OUT, "ours.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(official), "official and", len(ours), "verified buffers to", OUT)


if __name__ == "__main__":
    main()
