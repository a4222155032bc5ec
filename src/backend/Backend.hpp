// Backend plug-in interface — the drop-in boundary on the compiler side.
//
// Mirrors reference src/backend/Backend.hpp:22-36 (same member names, argument
// meaning and error convention): a backend gets the analysed script plus a context
// with output/asset/deps directories and the `-C` configuration, writes its sources
// together with executable `build.sh` and `run.sh` into outputDir, and throws
// BackendError for "model uses a feature I do not support" (driver exit code 2,
// reference src/main.cpp:270-274).
#pragma once

#include <map>
#include <stdexcept>
#include <string>

#include "../Ast.hpp"

namespace abl {

struct BackendError : public std::runtime_error {
  explicit BackendError(const std::string &msg) : std::runtime_error(msg) {}
};

struct ConfigError : public std::runtime_error {
  explicit ConfigError(const std::string &msg) : std::runtime_error(msg) {}
};

// `-C name=value` options, read lazily by backends (reference src/Config.cpp:20-48).
struct Config {
  std::map<std::string, std::string> config;
  bool getBool(const std::string &name, bool defaultValue) const;
  long getInt(const std::string &name, long defaultValue) const;
  std::string getString(const std::string &name, const std::string &defaultValue) const;
};

struct BackendContext {
  const std::string &outputDir;
  const std::string &assetDir;
  const std::string &depsDir;
  const Config &config;
};

struct Backend {
  virtual void generate(Script &script, const BackendContext &ctx) = 0;
  virtual void initEnv(const BackendContext &) {}
  virtual ~Backend() {}
};

// The B200-native backend (`-b cuda`).
struct CudaBackend : public Backend {
  void generate(Script &script, const BackendContext &ctx) override;
};

}  // namespace abl
