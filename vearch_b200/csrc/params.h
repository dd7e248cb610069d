// JSON -> parameter structs (model params at Init, retrieval params per request).
#pragma once
#include <string>

#include "index.h"

namespace gb {
bool parse_model_params(const std::string& text, ModelParams* mp, std::string* err);
bool parse_retrieval_params(const std::string& text, RetrievalParams* rp, std::string* err);
}  // namespace gb
