// Golden-fixture generator for the flatbuffers side of the drop-in boundary.
//
// Compiled HERE (the GPU box has no /root/reference) against the reference's own vendored flatbuffers
// runtime and its flatc-generated readers/builders:
//     /root/reference/internal/engine/third_party/flatbuffers/flatbuffers.h
//     /root/reference/internal/engine/idl/fbs-gen/c/{table,doc}_generated.h
// so every byte below comes from the code the reference engine and its SDKs run, not from this
// repository's hand-written codec (vearch_b200/csrc/wire.h, vearch_b200/wire.py).
//
//   make_fb_golden emit  <dir>       write official Table / Doc buffers into <dir>
//   make_fb_golden check <file> table|doc
//                                    run the official Verifier + generated reader over a buffer
//                                    produced by THIS repo's builders and print what it reads as JSON
//
// Recipe (tests/golden/make_fb_golden.sh):
//   g++ -std=c++17 -I/root/reference/internal/engine/third_party -I/root/reference/internal/engine/idl/fbs-gen/c
//       tests/golden/make_fb_golden.cc -o /tmp/make_fb_golden
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "flatbuffers/flatbuffers.h"
#include "doc_generated.h"
#include "table_generated.h"

using namespace gamma_api;

static void write_file(const std::string& path, const uint8_t* p, size_t n) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) {
    perror(path.c_str());
    exit(1);
  }
  fwrite(p, 1, n, f);
  fclose(f);
}

static std::vector<uint8_t> f32_bytes(int d, float base) {
  std::vector<float> v(d);
  for (int i = 0; i < d; i++) v[i] = base + 0.5f * (float)i;
  std::vector<uint8_t> b(d * 4);
  memcpy(b.data(), v.data(), b.size());
  return b;
}

static std::vector<uint8_t> i32_bytes(int32_t v) {
  std::vector<uint8_t> b(4);
  memcpy(b.data(), &v, 4);
  return b;
}

// Table as api_data/table.cc:97-157 (Table::Serialize) builds it
static void emit_table_full(const std::string& dir) {
  flatbuffers::FlatBufferBuilder fbb;
  std::vector<flatbuffers::Offset<FieldInfo>> fields;
  fields.push_back(CreateFieldInfo(fbb, fbb.CreateString("_id"), STRING, false));
  fields.push_back(CreateFieldInfo(fbb, fbb.CreateString("price"), INT, true, 1));
  fields.push_back(CreateFieldInfo(fbb, fbb.CreateString("tag"), STRING, true));
  fields.push_back(CreateFieldInfo(fbb, fbb.CreateString("weight"), DOUBLE, false));
  std::vector<flatbuffers::Offset<VectorInfo>> vecs;
  vecs.push_back(CreateVectorInfo(fbb, fbb.CreateString("emb"), VECTOR, true, 64, fbb.CreateString("MemoryOnly"),
                                  fbb.CreateString("{\"cache_size\": 1024}")));
  vecs.push_back(CreateVectorInfo(fbb, fbb.CreateString("img"), VECTOR, true, 16, fbb.CreateString(""), fbb.CreateString("")));
  std::vector<flatbuffers::Offset<IndexInfo>> idx;
  {
    std::vector<flatbuffers::Offset<flatbuffers::String>> names = {fbb.CreateString("emb")};
    idx.push_back(CreateIndexInfo(fbb, fbb.CreateString("emb_idx"), fbb.CreateString("IVFPQ"), fbb.CreateString("emb"),
                                  fbb.CreateVector(names),
                                  fbb.CreateString("{\"ncentroids\": 256, \"nsubvector\": 16, \"metric_type\": \"L2\", "
                                                   "\"training_threshold\": 10000, \"nprobe\": 20}")));
  }
  {
    std::vector<flatbuffers::Offset<flatbuffers::String>> names = {fbb.CreateString("img")};
    idx.push_back(CreateIndexInfo(fbb, fbb.CreateString("img_idx"), fbb.CreateString("FLAT"), fbb.CreateString("img"),
                                  fbb.CreateVector(names), fbb.CreateString("{\"metric_type\": \"InnerProduct\"}")));
  }
  {
    std::vector<flatbuffers::Offset<flatbuffers::String>> names = {fbb.CreateString("price")};
    idx.push_back(CreateIndexInfo(fbb, fbb.CreateString("price_idx"), fbb.CreateString("SCALAR"), fbb.CreateString("price"),
                                  fbb.CreateVector(names), fbb.CreateString("")));
  }
  auto t = CreateTable(fbb, fbb.CreateString("ts_space"), fbb.CreateVector(fields), fbb.CreateVector(vecs),
                       fbb.CreateString(""), fbb.CreateString(""), 250, true, true, fbb.CreateVector(idx));
  fbb.Finish(t);
  write_file(dir + "/table_official_full.fb", fbb.GetBufferPointer(), fbb.GetSize());
}

// every scalar at its schema default (omitted on the wire), no indexes vector at all
static void emit_table_minimal(const std::string& dir) {
  flatbuffers::FlatBufferBuilder fbb;
  std::vector<flatbuffers::Offset<VectorInfo>> vecs;
  vecs.push_back(CreateVectorInfo(fbb, fbb.CreateString("v"), VECTOR, false, 4, 0, 0));
  auto t = CreateTable(fbb, fbb.CreateString("t"), 0, fbb.CreateVector(vecs));
  fbb.Finish(t);
  write_file(dir + "/table_official_minimal.fb", fbb.GetBufferPointer(), fbb.GetSize());
}

// empty-but-present vectors (what the Go SDK emits for a table without scalar fields / indexes)
static void emit_table_empty_vectors(const std::string& dir) {
  flatbuffers::FlatBufferBuilder fbb;
  std::vector<flatbuffers::Offset<FieldInfo>> fields;
  std::vector<flatbuffers::Offset<VectorInfo>> vecs;
  vecs.push_back(CreateVectorInfo(fbb, fbb.CreateString("v"), VECTOR, true, 8, fbb.CreateString("RocksDB"), 0));
  std::vector<flatbuffers::Offset<IndexInfo>> idx;
  auto t = CreateTable(fbb, fbb.CreateString("empty_vectors"), fbb.CreateVector(fields), fbb.CreateVector(vecs), 0, 0,
                       1000, false, false, fbb.CreateVector(idx));
  fbb.Finish(t);
  write_file(dir + "/table_official_empty_vectors.fb", fbb.GetBufferPointer(), fbb.GetSize());
}

// Doc as api_data/doc.cc:16-50 (Doc::Serialize) builds it: value is a [ubyte] vector
static void emit_doc_bytes(const std::string& dir) {
  flatbuffers::FlatBufferBuilder fbb;
  std::vector<flatbuffers::Offset<Field>> fs;
  const std::string id = "doc-00042";
  std::vector<uint8_t> idb(id.begin(), id.end());
  fs.push_back(CreateField(fbb, fbb.CreateString("_id"), fbb.CreateVector(idb), STRING));
  fs.push_back(CreateField(fbb, fbb.CreateString("price"), fbb.CreateVector(i32_bytes(-7)), INT));
  const std::string tag = "red\001blue";
  std::vector<uint8_t> tagb(tag.begin(), tag.end());
  fs.push_back(CreateField(fbb, fbb.CreateString("tag"), fbb.CreateVector(tagb), STRING));
  fs.push_back(CreateField(fbb, fbb.CreateString("emb"), fbb.CreateVector(f32_bytes(64, 1.0f)), VECTOR));
  fs.push_back(CreateField(fbb, fbb.CreateString("img"), fbb.CreateVector(f32_bytes(16, -3.0f)), VECTOR));
  fbb.Finish(CreateDoc(fbb, fbb.CreateVector(fs)));
  write_file(dir + "/doc_official_bytes.fb", fbb.GetBufferPointer(), fbb.GetSize());
}

// Doc as the Go SDK builds it (sdk/go/gamma/doc.go:28-42): value written with CreateString, i.e. the
// same length-prefixed bytes followed by a NUL, although the schema says [ubyte]
static void emit_doc_gostring(const std::string& dir) {
  flatbuffers::FlatBufferBuilder fbb;
  std::vector<flatbuffers::Offset<flatbuffers::String>> names, values;
  const std::string id = "doc-00042";
  std::vector<uint8_t> emb = f32_bytes(64, 1.0f), img = f32_bytes(16, -3.0f), price = i32_bytes(-7);
  const std::string tag = "red\001blue";
  const char* nm[5] = {"_id", "price", "tag", "emb", "img"};
  const std::string vals[5] = {id, std::string(price.begin(), price.end()), tag, std::string(emb.begin(), emb.end()),
                               std::string(img.begin(), img.end())};
  const DataType dts[5] = {STRING, INT, STRING, VECTOR, VECTOR};
  for (int i = 0; i < 5; i++) {  // the Go code creates all names and values first, then the tables
    names.push_back(fbb.CreateString(nm[i]));
    values.push_back(fbb.CreateString(vals[i].data(), vals[i].size()));
  }
  std::vector<flatbuffers::Offset<Field>> fs;
  for (int i = 0; i < 5; i++) {
    FieldBuilder b(fbb);
    b.add_name(names[i]);
    b.add_value(flatbuffers::Offset<flatbuffers::Vector<uint8_t>>(values[i].o));
    b.add_data_type(dts[i]);
    fs.push_back(b.Finish());
  }
  fbb.Finish(CreateDoc(fbb, fbb.CreateVector(fs)));
  write_file(dir + "/doc_official_gostring.fb", fbb.GetBufferPointer(), fbb.GetSize());
}

static std::string jstr(const flatbuffers::String* s) {
  if (!s) return "null";
  std::string o = "\"";
  for (char c : s->str()) {
    if (c == '"' || c == '\\') o += '\\';
    o += c;
  }
  return o + "\"";
}

static int check(const char* path, const char* kind) {
  FILE* f = fopen(path, "rb");
  if (!f) {
    perror(path);
    return 2;
  }
  std::vector<uint8_t> buf;
  uint8_t tmp[4096];
  size_t n;
  while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
  fclose(f);
  flatbuffers::Verifier v(buf.data(), buf.size());
  if (!strcmp(kind, "table")) {
    if (!VerifyTableBuffer(v)) {
      printf("{\"verified\": false}\n");
      return 1;
    }
    const Table* t = GetTable(buf.data());
    printf("{\"verified\": true, \"name\": %s, \"refresh_interval\": %d, \"enable_id_cache\": %d, \"enable_realtime\": %d, ",
           jstr(t->name()).c_str(), t->refresh_interval(), (int)t->enable_id_cache(), (int)t->enable_realtime());
    printf("\"fields\": [");
    for (unsigned i = 0; t->fields() && i < t->fields()->size(); i++) {
      const FieldInfo* fi = t->fields()->Get(i);
      printf("%s{\"name\": %s, \"data_type\": %d, \"is_index\": %d}", i ? ", " : "", jstr(fi->name()).c_str(),
             (int)fi->data_type(), (int)fi->is_index());
    }
    printf("], \"vectors\": [");
    for (unsigned i = 0; t->vectors_info() && i < t->vectors_info()->size(); i++) {
      const VectorInfo* vi = t->vectors_info()->Get(i);
      printf("%s{\"name\": %s, \"dimension\": %d, \"is_index\": %d, \"store_type\": %s, \"store_param\": %s}", i ? ", " : "",
             jstr(vi->name()).c_str(), vi->dimension(), (int)vi->is_index(), jstr(vi->store_type()).c_str(),
             jstr(vi->store_param()).c_str());
    }
    printf("], \"indexes\": [");
    for (unsigned i = 0; t->indexes() && i < t->indexes()->size(); i++) {
      const IndexInfo* ii = t->indexes()->Get(i);
      printf("%s{\"name\": %s, \"type\": %s, \"field_name\": %s, \"params\": %s}", i ? ", " : "", jstr(ii->name()).c_str(),
             jstr(ii->type()).c_str(), jstr(ii->field_name()).c_str(), jstr(ii->params()).c_str());
    }
    printf("]}\n");
    return 0;
  }
  if (!VerifyDocBuffer(v)) {
    printf("{\"verified\": false}\n");
    return 1;
  }
  const Doc* d = GetDoc(buf.data());
  printf("{\"verified\": true, \"fields\": [");
  for (unsigned i = 0; d->fields() && i < d->fields()->size(); i++) {
    const Field* fl = d->fields()->Get(i);
    printf("%s{\"name\": %s, \"data_type\": %d, \"value_hex\": \"", i ? ", " : "", jstr(fl->name()).c_str(), (int)fl->data_type());
    for (unsigned j = 0; fl->value() && j < fl->value()->size(); j++) printf("%02x", fl->value()->Get(j));
    printf("\"}");
  }
  printf("]}\n");
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 3 && !strcmp(argv[1], "emit")) {
    const std::string dir = argv[2];
    emit_table_full(dir);
    emit_table_minimal(dir);
    emit_table_empty_vectors(dir);
    emit_doc_bytes(dir);
    emit_doc_gostring(dir);
    return 0;
  }
  if (argc >= 4 && !strcmp(argv[1], "check")) return check(argv[2], argv[3]);
  fprintf(stderr, "usage: %s emit <dir> | check <file> table|doc\n", argv[0]);
  return 2;
}
