// Stand-in for the bison-generated Parser.hpp (reference src/ParserContext.hpp:17).
#pragma once
#include "AST.hpp"
namespace OpenABL {
struct ParserContext;
class Parser {
public:
  explicit Parser(ParserContext &ctx) : ctx(ctx) {}
  int parse();
private:
  ParserContext &ctx;
};
}  // namespace OpenABL
