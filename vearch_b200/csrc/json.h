// Minimal JSON reader/writer for the engine's config strings (Init config, index model params,
// per-request retrieval params, status output).  Plays the role of util/utils.h JsonParser over
// cJSON in the reference (internal/engine/util/utils.{h,cc}).
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

namespace gb {

struct JsonValue {
  enum Type { Null, Bool, Number, String, Array, Object } type = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<JsonValue> arr;
  std::vector<std::pair<std::string, JsonValue>> obj;

  const JsonValue* get(const std::string& key) const {
    if (type != Object) return nullptr;
    for (auto& kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  bool get_int(const std::string& key, int* out) const {
    const JsonValue* v = get(key);
    if (!v || v->type != Number) return false;
    *out = (int)v->num;
    return true;
  }
  bool get_double(const std::string& key, double* out) const {
    const JsonValue* v = get(key);
    if (!v || v->type != Number) return false;
    *out = v->num;
    return true;
  }
  bool get_string(const std::string& key, std::string* out) const {
    const JsonValue* v = get(key);
    if (!v || v->type != String) return false;
    *out = v->str;
    return true;
  }
  bool get_bool(const std::string& key, bool* out) const {
    const JsonValue* v = get(key);
    if (!v) return false;
    if (v->type == Bool) {
      *out = v->b;
      return true;
    }
    if (v->type == Number) {
      *out = v->num != 0;
      return true;
    }
    return false;
  }
};

class JsonParser {
 public:
  // returns true on success
  static bool parse(const std::string& text, JsonValue* out) {
    JsonParser p(text);
    p.skip_ws();
    if (!p.value(out)) return false;
    p.skip_ws();
    return p.pos_ == p.s_.size();
  }

 private:
  explicit JsonParser(const std::string& s) : s_(s) {}
  const std::string& s_;
  size_t pos_ = 0;
  int depth_ = 0;

  void skip_ws() {
    while (pos_ < s_.size() && (s_[pos_] == ' ' || s_[pos_] == '\t' || s_[pos_] == '\n' || s_[pos_] == '\r')) pos_++;
  }
  bool lit(const char* w) {
    size_t n = strlen(w);
    if (s_.compare(pos_, n, w) != 0) return false;
    pos_ += n;
    return true;
  }
  bool string(std::string* out) {
    if (pos_ >= s_.size() || s_[pos_] != '"') return false;
    pos_++;
    out->clear();
    while (pos_ < s_.size()) {
      char c = s_[pos_++];
      if (c == '"') return true;
      if (c == '\\') {
        if (pos_ >= s_.size()) return false;
        char e = s_[pos_++];
        switch (e) {
          case '"': out->push_back('"'); break;
          case '\\': out->push_back('\\'); break;
          case '/': out->push_back('/'); break;
          case 'b': out->push_back('\b'); break;
          case 'f': out->push_back('\f'); break;
          case 'n': out->push_back('\n'); break;
          case 'r': out->push_back('\r'); break;
          case 't': out->push_back('\t'); break;
          case 'u': {
            if (pos_ + 4 > s_.size()) return false;
            unsigned cp = (unsigned)strtoul(s_.substr(pos_, 4).c_str(), nullptr, 16);
            pos_ += 4;
            if (cp < 0x80) {
              out->push_back((char)cp);
            } else if (cp < 0x800) {
              out->push_back((char)(0xC0 | (cp >> 6)));
              out->push_back((char)(0x80 | (cp & 0x3F)));
            } else {
              out->push_back((char)(0xE0 | (cp >> 12)));
              out->push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
              out->push_back((char)(0x80 | (cp & 0x3F)));
            }
            break;
          }
          default: return false;
        }
      } else {
        out->push_back(c);
      }
    }
    return false;
  }
  bool value(JsonValue* v) {
    if (++depth_ > 64) return false;
    skip_ws();
    if (pos_ >= s_.size()) return false;
    bool ok = false;
    char c = s_[pos_];
    if (c == '{') {
      pos_++;
      v->type = JsonValue::Object;
      skip_ws();
      if (pos_ < s_.size() && s_[pos_] == '}') {
        pos_++;
        ok = true;
      } else {
        while (true) {
          skip_ws();
          std::string key;
          if (!string(&key)) break;
          skip_ws();
          if (pos_ >= s_.size() || s_[pos_] != ':') break;
          pos_++;
          JsonValue child;
          if (!value(&child)) break;
          v->obj.emplace_back(std::move(key), std::move(child));
          skip_ws();
          if (pos_ < s_.size() && s_[pos_] == ',') {
            pos_++;
            continue;
          }
          if (pos_ < s_.size() && s_[pos_] == '}') {
            pos_++;
            ok = true;
          }
          break;
        }
      }
    } else if (c == '[') {
      pos_++;
      v->type = JsonValue::Array;
      skip_ws();
      if (pos_ < s_.size() && s_[pos_] == ']') {
        pos_++;
        ok = true;
      } else {
        while (true) {
          JsonValue child;
          if (!value(&child)) break;
          v->arr.push_back(std::move(child));
          skip_ws();
          if (pos_ < s_.size() && s_[pos_] == ',') {
            pos_++;
            continue;
          }
          if (pos_ < s_.size() && s_[pos_] == ']') {
            pos_++;
            ok = true;
          }
          break;
        }
      }
    } else if (c == '"') {
      v->type = JsonValue::String;
      ok = string(&v->str);
    } else if (c == 't') {
      v->type = JsonValue::Bool;
      v->b = true;
      ok = lit("true");
    } else if (c == 'f') {
      v->type = JsonValue::Bool;
      v->b = false;
      ok = lit("false");
    } else if (c == 'n') {
      v->type = JsonValue::Null;
      ok = lit("null");
    } else {
      const char* start = s_.c_str() + pos_;
      char* end = nullptr;
      double d = strtod(start, &end);
      if (end != start) {
        v->type = JsonValue::Number;
        v->num = d;
        pos_ += (size_t)(end - start);
        ok = true;
      }
    }
    depth_--;
    return ok;
  }
};

inline std::string json_escape(const std::string& s) {
  std::string o;
  for (char c : s) {
    switch (c) {
      case '"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\n': o += "\\n"; break;
      case '\r': o += "\\r"; break;
      case '\t': o += "\\t"; break;
      default:
        if ((unsigned char)c < 0x20) {
          char buf[8];
          snprintf(buf, sizeof buf, "\\u%04x", c);
          o += buf;
        } else {
          o.push_back(c);
        }
    }
  }
  return o;
}

}  // namespace gb
