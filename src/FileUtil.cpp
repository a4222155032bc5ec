#include "FileUtil.hpp"

#include <climits>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <sys/stat.h>
#include <unistd.h>

namespace abl {

bool fileExists(const std::string &path) {
  struct stat st;
  return stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

bool directoryExists(const std::string &path) {
  struct stat st;
  return stat(path.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

bool createDirectory(const std::string &path) {
  std::string partial;
  for (size_t i = 0; i <= path.size(); i++) {
    if (i == path.size() || path[i] == '/') {
      if (!partial.empty() && !directoryExists(partial) && mkdir(partial.c_str(), 0755) != 0 &&
          !directoryExists(partial))
        return false;
    }
    if (i < path.size()) partial.push_back(path[i]);
  }
  return true;
}

std::string createTemporaryDirectory() {
  char tmpl[] = "/tmp/openabl_XXXXXX";
  char *res = mkdtemp(tmpl);
  if (!res) throw std::runtime_error("Could not create temporary directory");
  return res;
}

std::string getAbsolutePath(const std::string &path) {
  char buf[PATH_MAX];
  if (realpath(path.c_str(), buf)) return buf;
  if (!path.empty() && path[0] == '/') return path;
  if (getcwd(buf, sizeof buf)) return std::string(buf) + "/" + path;
  return path;
}

bool readFile(const std::string &path, std::string &out) {
  std::ifstream f(path.c_str(), std::ios::binary);
  if (!f) return false;
  std::ostringstream ss;
  ss << f.rdbuf();
  out = ss.str();
  return true;
}

void writeToFile(const std::string &path, const std::string &contents) {
  std::ofstream f(path.c_str(), std::ios::binary);
  if (!f) throw std::runtime_error("Could not write \"" + path + "\"");
  f << contents;
}

void copyFile(const std::string &src, const std::string &dst) {
  std::string data;
  if (!readFile(src, data)) throw std::runtime_error("Could not read \"" + src + "\"");
  writeToFile(dst, data);
}

void makeFileExecutable(const std::string &path) { chmod(path.c_str(), 0755); }

void changeWorkingDirectory(const std::string &path) {
  if (chdir(path.c_str()) != 0) throw std::runtime_error("Could not enter \"" + path + "\"");
}

bool executeCommand(const std::string &cmd) { return system(cmd.c_str()) == 0; }

std::string executableDirectory() {
  char buf[PATH_MAX];
  ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
  if (n <= 0) return ".";
  buf[n] = 0;
  char *slash = strrchr(buf, '/');
  if (slash) *slash = 0;
  return buf;
}

}  // namespace abl
