// Small filesystem / process helpers used by the driver and the backends.
#pragma once

#include <string>

namespace abl {

bool fileExists(const std::string &path);
bool directoryExists(const std::string &path);
bool createDirectory(const std::string &path);      // mkdir -p
std::string createTemporaryDirectory();
std::string getAbsolutePath(const std::string &path);
bool readFile(const std::string &path, std::string &out);
void writeToFile(const std::string &path, const std::string &contents);
void copyFile(const std::string &src, const std::string &dst);
void makeFileExecutable(const std::string &path);
void changeWorkingDirectory(const std::string &path);
bool executeCommand(const std::string &cmd);
std::string executableDirectory();

}  // namespace abl
