// Hand-written lexer + recursive-descent parser for the OpenABL DSL.
//
// The language accepted is the one defined by the reference's flex/bison sources
// (reference src/Lexer.l:49-161 tokens, src/Parser.y:115-134 precedence,
// src/Parser.y:162-364 grammar).  Neither generator exists in this toolchain, so
// this is an independent implementation; the only behaviour deliberately mirrored
// is *observable* behaviour: which programs parse, operator precedence, and the
// line number each construct reports in diagnostics (the 23 golden files under
// reference test/*.exp pin those line numbers, including the quirk that a token
// at column 1 directly after a newline reports the line of the previous token's
// end, because the reference lexer only advances its location start on
// whitespace, not on newlines — Lexer.l:57-60, 158-159).
#pragma once

#include <string>

#include "Ast.hpp"

namespace abl {

struct ParseError {
  std::string msg;
  int line;
};

// Parses `text`; on failure returns nullptr and fills `err`.
std::unique_ptr<Script> parseScript(const std::string &text, ParseError &err);

}  // namespace abl
