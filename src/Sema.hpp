// Semantic analysis for the OpenABL DSL: scoping, typing, int->float promotion,
// constant folding of globals (+ -P overrides), overload resolution and naming,
// environment bounds, for-near metadata, add/remove/reduction bookkeeping.
//
// Behavioural contract = reference src/AnalysisVisitor.cpp (whole file) and
// src/Value.cpp:57-322; diagnostics text/line numbers are pinned by the reference's
// golden files test/*.exp.  The implementation is a single recursive pass over the
// tagged tree in Ast.hpp (no visitor classes).
#pragma once

#include <functional>
#include <map>
#include <string>
#include <vector>

#include "Ast.hpp"

namespace abl {

struct Diagnostic {
  std::string msg;
  int line;
};

struct Signature {
  enum Special : uint8_t { None, Sum, CountMember, LogCsv };
  static const unsigned SEQ_STEP_ONLY = 1u << 2;
  std::string name;       // name used in source
  std::string emitName;   // name emitted by printers (dist_float2, random_4, ...)
  std::vector<Ty> params;
  Ty ret;
  unsigned flags = 0;
  FuncDecl *decl = nullptr;
  Special special = None;
};

class Sema {
public:
  Sema(Script &mainScript, const std::map<std::string, std::string> &cliParams,
       const std::string &backend);

  void analyseLibrary(Script &lib);
  void analyseMain();

  const std::vector<Diagnostic> &diagnostics() const { return diags; }

  // Constant evaluation helpers, also used by backends (reference Value.cpp).
  static Const parseCliValue(const std::string &text);
  Const eval(const Expr &e) const;

private:
  Script &script;  // main script: collects agents/consts/funcs of lib + main
  const std::map<std::string, std::string> &cliParams;
  std::string backend;
  std::vector<Diagnostic> diags;

  std::map<std::string, std::vector<Signature>> functions;
  std::map<std::string, AgentDecl *> agentsByName;
  std::map<std::string, FuncDecl *> funcsByName;
  std::map<std::string, Symbol *> names;
  std::vector<std::map<std::string, Symbol *>> nameStack;
  std::vector<Const> radii;
  FuncDecl *curFunc = nullptr;
  Symbol *nearVar = nullptr;
  int loopDepth = 0;
  bool isLib = false;
  int nextUid = 0;

  void error(const std::string &msg, int line) { diags.push_back({msg, line}); }
  void registerBuiltins();
  void addBuiltin(const std::string &name, const std::string &emit, std::vector<Ty> params, Ty ret,
                  unsigned flags = 0, Signature::Special sp = Signature::None);

  Ty resolveType(const std::string &name, int line);
  Symbol *declare(const std::string &name, int line, Ty type, bool immutable, bool global,
                  const Const &val);
  void pushScope() { nameStack.push_back(names); }
  void popScope() { names = nameStack.back(); nameStack.pop_back(); }

  void script_(Script &s);
  void agent(AgentDecl &a);
  void constant(ConstDecl &c);
  void environment(EnvDecl &e);
  void function(FuncDecl &f);
  void stmt(Stmt &s);
  void expr(ExprP &e);
  void call(ExprP &e);
  void finishMain();

  bool promote(ExprP &e, const Ty &want);
  Ty binaryType(Op op, ExprP &l, ExprP &r);
  bool isImmutableTarget(const Expr &e) const;
  const Signature *findCompatible(const std::vector<Signature> &sigs, const std::vector<Ty> &args) const;
  Signature concretize(const Signature &sig, const std::vector<Ty> &args) const;
};

// Turns a folded constant back into a typed literal / constructor expression.
ExprP constToExpr(const Const &c);

}  // namespace abl
