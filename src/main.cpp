// OpenABL driver with the B200-native `cuda` backend.
//
// Command line, exit codes and console output follow the reference driver
// (reference src/main.cpp:170-315, src/Cli.cpp:31-95): parse model + asset/lib.abl,
// analyse, `generate` through the backend registry, then optionally ./build.sh and
// ./run.sh inside the output directory, printing `Execution time: <s>s`.
//   exit 0  success          exit 1  usage / parse / analysis / build error
//   exit 2  BackendError (feature unsupported by the chosen backend)
#include <chrono>
#include <cstdio>
#include <iostream>
#include <map>
#include <memory>

#include "FileUtil.hpp"
#include "Parser.hpp"
#include "Sema.hpp"
#include "backend/Backend.hpp"

namespace abl {

struct Options {
  bool help = false, lintOnly = false, build = false, run = false;
  std::string fileName, backend, outputDir, assetDir, depsDir;
  std::map<std::string, std::string> params, config;
};

struct OptionError : public std::runtime_error {
  explicit OptionError(const std::string &m) : std::runtime_error(m) {}
};

static void splitPair(const std::string &arg, const char *what,
                      std::map<std::string, std::string> &into) {
  size_t eq = arg.find('=');
  if (eq == std::string::npos)
    throw OptionError(std::string("Malformed ") + what + ": Missing \"=\"");
  into[arg.substr(0, eq)] = arg.substr(eq + 1);
}

static Options parseOptions(int argc, char **argv) {
  Options o;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a == "-h" || a == "--help") { o.help = true; return o; }
    if (a == "--lint-only") { o.lintOnly = true; continue; }
    if (a == "-B" || a == "--build") { o.build = true; continue; }
    if (a == "-R" || a == "--run") { o.run = true; continue; }
    if (i + 1 == argc) throw OptionError("Missing argument for option \"" + a + "\"");
    std::string v = argv[++i];
    if (a == "-b" || a == "--backend") o.backend = v;
    else if (a == "-i" || a == "--input") o.fileName = v;
    else if (a == "-o" || a == "--output-dir") o.outputDir = v;
    else if (a == "-A" || a == "--asset-dir") o.assetDir = v;
    else if (a == "-D" || a == "--deps-dir") o.depsDir = v;
    else if (a == "-P" || a == "--param") splitPair(v, "parameter", o.params);
    else if (a == "-C" || a == "--config") splitPair(v, "configuration value", o.config);
    else throw OptionError("Unknown option \"" + a + "\"");
  }
  if (o.fileName.empty()) throw OptionError("Missing input file (-i or --input)");
  if (o.assetDir.empty()) {
    // default ./asset like the reference; fall back to the asset dir next to the binary
    o.assetDir = "./asset";
    if (!directoryExists(o.assetDir)) {
      std::string alt = executableDirectory() + "/asset";
      std::string alt2 = executableDirectory() + "/../asset";
      if (directoryExists(alt)) o.assetDir = alt;
      else if (directoryExists(alt2)) o.assetDir = alt2;
    }
  }
  if (o.depsDir.empty()) o.depsDir = "./deps";
  if (o.lintOnly) return o;
  if (o.backend.empty()) throw OptionError("Missing backend (-b or --backend)");
  return o;
}

static void printHelp() {
  std::cout << "Usage: ./OpenABL -i input.abl -o ./output-dir -b backend\n\n"
               "Options:\n"
               "  -A, --asset-dir    Asset directory (default: ./asset)\n"
               "  -b, --backend      Backend\n"
               "  -B, --build        Build the generated code\n"
               "  -C, --config       Specify a configuration value (name=value)\n"
               "  -D, --deps-dir     Deps directory (default: ./deps)\n"
               "  -h, --help         Display this help\n"
               "  -i, --input        Input file\n"
               "  -o, --output-dir   Output directory\n"
               "  -P, --param        Specify a simulation parameter (name=value)\n"
               "  -R, --run          Build and run the generated code\n"
               "      --lint-only    Only parse and analyse the model\n"
               "\n"
               "Available backends:\n"
               " * cuda   (NVIDIA B200, sm_100a)\n"
               "\n"
               "Available configuration options:\n"
               " * bool use_float       (default: false) single precision agent state\n"
               " * int  cuda.block_size (default: 0)     threads per CTA of step kernels (0 = automatic)\n"
               " * int  cuda.gpus       (default: 1)     GPUs of this machine the generated program runs on (slab\n"
               "                                          decomposition; ABL_CUDA_GPUS overrides it at run time)\n"
               " * bool cuda.tile       (default: false) stage neighbour cells in shared memory (ABL_MODE 2)\n"
               " * bool cuda.bulk       (default: true)  print the TMA-staged tile variant of sparse 2-D loops (ABL_MODE 7)\n"
               " * bool cuda.dense      (default: true)  print the single-precision pre-filter variant of dense loops (ABL_MODE 8)\n"
               " * bool cuda.flat       (default: true)  print the flat candidate loop of 2-D loops (ABL_MODE 3)\n"
               " * bool cuda.cull       (default: true)  skip cells a small radius cannot reach\n"
               " * bool cuda.rowcull    (default: false) narrow every row of cells to the agent's reach (dense loops)\n"
               " * bool cuda.sqcmp      (default: true)  compare squared distances instead of taking square roots\n"
               " * bool cuda.nlist      (default: false) cache neighbour lists of step functions whose agents never move\n"
               " * bool cuda.unroll     (default: false) unroll the for-near candidate loop by two\n"
               " * str  cuda.save_format (default: json) save() output: json, flame_xml or flamegpu_xml\n"
               " * int  cuda.dump_state (default: 0)     also write raw binary state on save()\n"
               " * bool visualize       (default: false) write frames/frame_<timestep>.ppm, agents painted by the model's\n"
               "                                          getColor / getSize (every cuda.frame_interval timesteps, default 1)\n"
            << std::flush;
}

static std::map<std::string, std::unique_ptr<Backend>> getBackends() {
  std::map<std::string, std::unique_ptr<Backend>> b;
  b["cuda"] = std::unique_ptr<Backend>(new CudaBackend);
  return b;
}

static std::unique_ptr<Script> parseFile(const std::string &path, const char *what) {
  std::string text;
  if (!readFile(path, text)) {
    std::cerr << what << " \"" << path << "\" could not be opened." << std::endl;
    return nullptr;
  }
  ParseError perr;
  auto s = parseScript(text, perr);
  if (!s) std::cerr << "Parse error: " << perr.msg << " on line " << perr.line << std::endl;
  return s;
}

static int run(int argc, char **argv) {
  Options opt;
  try {
    opt = parseOptions(argc, argv);
  } catch (const OptionError &e) {
    printHelp();
    std::cerr << "\nERROR: " << e.what() << std::endl;
    return 1;
  }
  if (opt.help) { printHelp(); return 0; }

  std::string probe;
  if (!readFile(opt.fileName, probe)) {
    std::cerr << "File \"" << opt.fileName << "\" could not be opened." << std::endl;
    return 1;
  }
  if (!directoryExists(opt.assetDir)) {
    std::cerr << "Asset directory \"" << opt.assetDir << "\" does not exist "
              << "(override with -A or --asset-dir)" << std::endl;
    return 1;
  }
  auto mainScript = parseFile(opt.fileName, "File");
  auto libScript = parseFile(opt.assetDir + "/lib.abl", "Library file");
  if (!mainScript || !libScript) return 1;

  Sema sema(*mainScript, opt.params, opt.backend);
  sema.analyseLibrary(*libScript);
  sema.analyseMain();
  for (const Diagnostic &d : sema.diagnostics())
    std::cerr << d.msg << " on line " << d.line << std::endl;
  if (!sema.diagnostics().empty()) return 1;
  if (opt.lintOnly) return 0;

  if (!opt.outputDir.empty()) {
    createDirectory(opt.outputDir);
  } else if (opt.build || opt.run) {
    opt.outputDir = createTemporaryDirectory();
    std::cout << "Writing to directory " << opt.outputDir << std::endl;
  } else {
    std::cerr << "Missing output directory (-o or --output-dir)" << std::endl;
    return 1;
  }
  opt.depsDir = getAbsolutePath(opt.depsDir);
  opt.assetDir = getAbsolutePath(opt.assetDir);

  auto backends = getBackends();
  auto it = backends.find(opt.backend);
  if (it == backends.end()) {
    std::cerr << "Unknown backend \"" << opt.backend << "\"" << std::endl;
    return 1;
  }
  Config config{opt.config};
  BackendContext ctx{opt.outputDir, opt.assetDir, opt.depsDir, config};
  try {
    it->second->generate(*mainScript, ctx);
  } catch (const BackendError &e) {
    std::cerr << e.what() << std::endl;
    return 2;
  } catch (const std::runtime_error &e) {
    std::cerr << e.what() << std::endl;
    return 1;
  }

  if (opt.build || opt.run) {
    changeWorkingDirectory(opt.outputDir);
    it->second->initEnv(ctx);
    if (!fileExists("./build.sh")) {
      std::cerr << "Build file for this backend not found" << std::endl;
      return 1;
    }
    if (!executeCommand("./build.sh")) {
      std::cerr << "Build failed" << std::endl;
      return 1;
    }
    if (opt.run) {
      if (!fileExists("./run.sh")) {
        std::cerr << "Run file for this backend not found" << std::endl;
        return 1;
      }
      auto t0 = std::chrono::high_resolution_clock::now();
      if (!executeCommand("./run.sh")) {
        std::cerr << "Run failed" << std::endl;
        return 1;
      }
      auto t1 = std::chrono::high_resolution_clock::now();
      auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count();
      std::cout << "Execution time: " << ms / 1000.0 << "s" << std::endl;
    }
  }
  return 0;
}

// ---- Config --------------------------------------------------------------
bool Config::getBool(const std::string &name, bool def) const {
  auto it = config.find(name);
  if (it == config.end()) return def;
  Const v = Sema::parseCliValue(it->second);
  if (v.k != TK::Bool) throw ConfigError("Value of " + name + " must be boolean");
  return v.b;
}
long Config::getInt(const std::string &name, long def) const {
  auto it = config.find(name);
  if (it == config.end()) return def;
  Const v = Sema::parseCliValue(it->second);
  if (v.k != TK::Int) throw ConfigError("Value of " + name + " must be integer");
  return v.i;
}
std::string Config::getString(const std::string &name, const std::string &def) const {
  auto it = config.find(name);
  return it == config.end() ? def : it->second;
}

}  // namespace abl

int main(int argc, char **argv) { return abl::run(argc, argv); }
