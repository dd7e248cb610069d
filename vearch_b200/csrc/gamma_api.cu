// The gamma C-ABI (include/gamma_api.h): the symbols the Go partition server binds through cgo
// (internal/engine/sdk/go/gamma/gamma.go), implemented over gb::Engine.
// Mirrors internal/engine/c_api/gamma_api.cc:35-374; every output buffer is malloc()'d.
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>

#include "../../include/gamma_api.h"
#include "engine.h"
#include "json.h"

using gb::Engine;

static void to_cstatus(const gb::Status& s, struct CStatus* c) {
  c->code = s.code;
  c->msg = nullptr;
  if (s.code != 0) {
    std::string m = s.ToString();
    // the reference allocates with new[] while Go frees with C.free (gamma_api.cc:149 vs
    // gamma.go:47-53); the drop-in must malloc
    c->msg = static_cast<char*>(malloc(m.size() + 1));
    if (c->msg) memcpy(c->msg, m.c_str(), m.size() + 1);
  }
}

static void out_buffer(const std::string& s, char** out, int* len) {
  *len = (int)s.size();
  *out = static_cast<char*>(malloc(s.size() ? s.size() : 1));
  if (*out && !s.empty()) memcpy(*out, s.data(), s.size());
}

static std::atomic<int> g_log_dir_flag{0};

extern "C" {

void* Init(const char* config_str, int len) {
  gb::JsonValue j;
  if (!config_str || len <= 0 || !gb::JsonParser::parse(std::string(config_str, (size_t)len), &j) ||
      j.type != gb::JsonValue::Object)
    return nullptr;
  std::string log_dir, path, space = "default";
  int flag = g_log_dir_flag.fetch_add(1);
  if (flag == 0 && !j.get_string("log_dir", &log_dir)) return nullptr;  // mandatory on first call (gamma_api.cc:40-47)
  if (!j.get_string("path", &path)) return nullptr;                     // gamma_api.cc:49-53
  j.get_string("space_name", &space);
  int device = 0;
  j.get_int("device", &device);
  if (const char* env = getenv("GAMMA_B200_DEVICE")) device = atoi(env);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return nullptr;  // no CPU path
  if (device < 0 || device >= ndev) device = device % ndev;
  return static_cast<void*>(new Engine(path, space, device));
}

int Close(void* engine) {
  delete static_cast<Engine*>(engine);
  return 0;
}

struct CStatus CreateTable(void* engine, const char* table_str, int len) {
  struct CStatus c;
  if (!engine) {
    to_cstatus(gb::Status::Make(gb::kInvalidArgument, "null engine"), &c);
    return c;
  }
  to_cstatus(static_cast<Engine*>(engine)->CreateTable(reinterpret_cast<const uint8_t*>(table_str), (size_t)len), &c);
  return c;
}

int AddOrUpdateDoc(void* engine, const char* doc_str, int len) {
  if (!engine) return -1;
  return static_cast<Engine*>(engine)->AddOrUpdate(reinterpret_cast<const uint8_t*>(doc_str), (size_t)len);
}

int DeleteDoc(void* engine, const char* docid, int docid_len) {
  if (!engine) return -1;
  return static_cast<Engine*>(engine)->Delete(std::string(docid, (size_t)docid_len));
}

void GetEngineStatus(void* engine, char** status, int* len) {
  out_buffer(engine ? static_cast<Engine*>(engine)->EngineStatus() : std::string("{}"), status, len);
}

void GetMemoryInfo(void* engine, char** memory_info, int* len) {
  out_buffer(engine ? static_cast<Engine*>(engine)->MemoryInfo() : std::string("{}"), memory_info, len);
}

int GetDocByID(void* engine, const char* docid, int docid_len, char** doc_str, int* len) {
  if (!engine) return -1;
  std::string fb;
  int ret = static_cast<Engine*>(engine)->GetDocByKey(std::string(docid, (size_t)docid_len), &fb);
  out_buffer(fb, doc_str, len);
  return ret;
}

int GetDocByDocID(void* engine, int docid, char next, char** doc_str, int* len) {
  if (!engine) return -1;
  std::string fb;
  int ret = static_cast<Engine*>(engine)->GetDocByDocid(docid, next != 0, &fb);
  out_buffer(fb, doc_str, len);
  return ret;
}

int BuildIndex(void* engine) { return engine ? static_cast<Engine*>(engine)->BuildIndex() : -1; }

int RebuildIndex(void* engine, int, int, int) {
  // RebuildIndex only acts on a running index and re-trains in place (engine.cc:991-1089);
  // not offered by this build: report "nothing to do" exactly like an idle reference engine.
  (void)engine;
  return 0;
}

int Dump(void* engine) { return engine ? static_cast<Engine*>(engine)->Dump() : -1; }
int Load(void* engine) { return engine ? static_cast<Engine*>(engine)->Load() : -1; }

struct CStatus Search(void* engine, const char* request_str, int req_len, char** response_str, int* res_len) {
  struct CStatus c;
  if (!engine) {
    to_cstatus(gb::Status::Make(gb::kInvalidArgument, "null engine"), &c);
    return c;
  }
  gb::SearchRequestPB req;
  if (!req.parse(reinterpret_cast<const uint8_t*>(request_str), (size_t)req_len)) {
    to_cstatus(gb::Status::Make(gb::kInvalidArgument, "parse search request failed"), &c);
    return c;
  }
  std::string resp;
  gb::Status st = static_cast<Engine*>(engine)->Search(req, &resp);
  to_cstatus(st, &c);
  if (st.ok()) out_buffer(resp, response_str, res_len);
  return c;
}

struct CStatus Query(void*, const char*, int, char**, int*) {
  struct CStatus c;
  to_cstatus(gb::Status::Make(gb::kNotSupported, "Query (scalar-only document query) is outside the vector hot path"), &c);
  return c;
}

int SetConfig(void* engine, const char* config_str, int len) {
  return engine ? static_cast<Engine*>(engine)->SetConfig(std::string(config_str, (size_t)len)) : -1;
}

int GetConfig(void* engine, char** config_str, int* len) {
  if (!engine) return -1;
  out_buffer(static_cast<Engine*>(engine)->GetConfig(), config_str, len);
  return 0;
}

struct CStatus Backup(void*, int) {
  struct CStatus c;
  to_cstatus(gb::Status::Make(gb::kNotSupported, "Backup is handled by the storage layer, outside the vector hot path"), &c);
  return c;
}

struct CStatus AddFieldIndexWithParams(void*, const char*, int, const char*, int, const char*, int) {
  struct CStatus c;
  to_cstatus(gb::Status::Make(gb::kNotSupported, "scalar field indexes are outside the vector hot path"), &c);
  return c;
}

struct CStatus RemoveFieldIndex(void*, const char*, int) {
  struct CStatus c;
  to_cstatus(gb::Status::Make(gb::kNotSupported, "scalar field indexes are outside the vector hot path"), &c);
  return c;
}

void SetMemoryLimitConfig(int) {}  // host-RSS watchdog (memory/memoryManager.cc): not applicable, vectors live in HBM

void SetKillStatus(const char* request_id, int partition_id, int reason) {
  if (request_id) Engine::SetKill(request_id, partition_id, reason);
}

void DeleteKillStatus(const char* request_id, int partition_id) {
  if (request_id) Engine::ClearKill(request_id, partition_id);
}

}  // extern "C"
