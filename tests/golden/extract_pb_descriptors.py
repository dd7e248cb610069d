"""Extracts the serialized FileDescriptorProtos that protoc embedded in the reference's generated Go code
(internal/proto/vearchpb/*.pb.go: `var file_<name>_proto_rawDesc = []byte{...}`) -- i.e. protoc's own output for
the reference's internal/proto/{errors,data_model,router_grpc}.proto -- and writes them as ONE FileDescriptorSet:

    python tests/golden/extract_pb_descriptors.py            # -> tests/golden/vearchpb_descriptor_set.binpb (+ .json manifest)

Run HERE (the GPU box has no /root/reference); the outputs are committed.  tests/golden/gen_golden.py builds its message
classes from this set, so the protobuf fixtures come from the reference's real descriptors, not from a hand restatement
(there is no protoc / grpc_tools in this image, and none is needed)."""
import hashlib
import json
import os
import re

from google.protobuf import descriptor_pb2

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/internal/proto/vearchpb"
FILES = ["errors", "data_model", "router_grpc"]  # dependency order


def raw_desc(go_path, stem):
    src = open(go_path).read()
    m = re.search(r"var file_%s_proto_rawDesc = \[\]byte\{(.*?)\n\}" % stem, src, re.S)
    if not m:
        raise SystemExit(f"{go_path}: no rawDesc literal")
    return bytes(int(tok, 16) for tok in re.findall(r"0x([0-9a-fA-F]{2})", m.group(1)))


def main():
    fds = descriptor_pb2.FileDescriptorSet()
    manifest = {}
    for stem in FILES:
        go = os.path.join(REF, stem + ".pb.go")
        raw = raw_desc(go, stem)
        fd = fds.file.add()
        fd.ParseFromString(raw)
        assert fd.name == stem + ".proto", fd.name
        manifest[fd.name] = {"source": go, "bytes": len(raw), "sha256": hashlib.sha256(raw).hexdigest(),
                             "package": fd.package, "messages": [m.name for m in fd.message_type]}
    out = os.path.join(HERE, "vearchpb_descriptor_set.binpb")
    open(out, "wb").write(fds.SerializeToString())
    json.dump(manifest, open(os.path.join(HERE, "vearchpb_descriptor_set.json"), "w"), indent=1)
    print("wrote", out, {k: v["bytes"] for k, v in manifest.items()})


if __name__ == "__main__":
    main()
