// Stand-in for the bison-generated location.hh the reference sources include
// (reference src/AST.hpp:21, src/ErrorHandling.hpp:19).  bison is not installed in
// this toolchain; only the members the reference actually touches are provided.
#pragma once
namespace OpenABL {
struct position {
  unsigned line = 1, column = 1;
};
struct location {
  position begin, end;
  location() {}
  explicit location(unsigned line) { begin.line = end.line = line; }
};
}  // namespace OpenABL
