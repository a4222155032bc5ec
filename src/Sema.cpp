#include "Sema.hpp"

#include <cassert>
#include <cmath>
#include <sstream>
#include <stdexcept>

namespace abl {

// ---------------------------------------------------------------------------
// Small helpers
// ---------------------------------------------------------------------------
const char *opSigil(Op op) {
  switch (op) {
    case Op::Add: return "+"; case Op::Sub: return "-"; case Op::Mul: return "*";
    case Op::Div: return "/"; case Op::Mod: return "%"; case Op::BitAnd: return "&";
    case Op::BitXor: return "^"; case Op::BitOr: return "|"; case Op::Shl: return "<<";
    case Op::Shr: return ">>"; case Op::Eq: return "=="; case Op::Ne: return "!=";
    case Op::Lt: return "<"; case Op::Le: return "<="; case Op::Gt: return ">";
    case Op::Ge: return ">="; case Op::And: return "&&"; case Op::Or: return "||";
    case Op::Range: return ".."; case Op::Neg: return "-"; case Op::Pos: return "+";
    case Op::Not: return "!"; case Op::BitNot: return "~";
  }
  return "?";
}

std::string Ty::str() const {
  switch (k) {
    case TK::Invalid: return "INVALID";
    case TK::Void: return "void";
    case TK::Bool: return "bool";
    case TK::Int: return "int";
    case TK::Float: return "float";
    case TK::String: return "string";
    case TK::Vec2: return "float2";
    case TK::Vec3: return "float3";
    case TK::Agent: return agent ? agent->name : "agent";
    case TK::Array: return elem().str() + "[]";
    case TK::AgentType: return agent ? "agentType{" + agent->name + "}" : "agentType";
    case TK::AgentMember:
      return agent ? "agentMember{" + agent->name + "." + member->name + "}" : "agentMember";
    case TK::Unresolved: return "";
  }
  return "";
}

static TK builtinTypeKind(const std::string &n) {
  if (n == "void") return TK::Void;
  if (n == "bool") return TK::Bool;
  if (n == "int") return TK::Int;
  if (n == "float") return TK::Float;
  if (n == "string") return TK::String;
  if (n == "float2") return TK::Vec2;
  if (n == "float3") return TK::Vec3;
  return TK::Invalid;
}

static std::string argList(const std::vector<Ty> &tys) {
  std::string s = "(";
  for (size_t i = 0; i < tys.size(); i++) {
    if (i) s += ", ";
    s += tys[i].str();
  }
  return s + ")";
}

static ExprP typedLit(Expr::Kind k, Ty t) {
  ExprP e(new Expr(k, 1));
  e->type = t;
  return e;
}

ExprP constToExpr(const Const &c) {
  switch (c.k) {
    case TK::Bool: { ExprP e = typedLit(Expr::BoolLit, TK::Bool); e->bval = c.b; return e; }
    case TK::Int: { ExprP e = typedLit(Expr::IntLit, TK::Int); e->ival = c.i; return e; }
    case TK::Float: { ExprP e = typedLit(Expr::FloatLit, TK::Float); e->fval = c.f; return e; }
    case TK::String: { ExprP e = typedLit(Expr::StrLit, TK::String); e->name = c.s; return e; }
    case TK::Vec2:
    case TK::Vec3: {
      ExprP e = typedLit(Expr::Call, c.k);
      e->name = c.k == TK::Vec2 ? "float2" : "float3";
      e->ckind = Expr::Ctor;
      for (int i = 0; i < c.vecLen(); i++) {
        ExprP f = typedLit(Expr::FloatLit, TK::Float);
        f->fval = c.v[i];
        e->kids.push_back(std::move(f));
      }
      return e;
    }
    default: return nullptr;
  }
}

// ---------------------------------------------------------------------------
// Constant evaluation (reference Value.cpp:57-322, AnalysisVisitor.cpp:277-384)
// ---------------------------------------------------------------------------
Const Sema::parseCliValue(const std::string &text) {
  if (text == "true") return Const::ofBool(true);
  if (text == "false") return Const::ofBool(false);
  try {
    size_t used;
    long v = std::stoi(text, &used);
    if (used == text.size()) return Const::ofInt(v);
  } catch (const std::logic_error &) {}
  try {
    size_t used;
    double v = std::stod(text, &used);
    if (used == text.size()) return Const::ofFloat(v);
  } catch (const std::logic_error &) {}
  return Const();
}

static Const foldUnary(Op op, const Const &v) {
  switch (op) {
    case Op::Pos:
      return (v.isNum() || v.isVec()) ? v : Const();
    case Op::Neg:
      if (v.k == TK::Int) return Const::ofInt(-v.i);
      if (v.k == TK::Float) return Const::ofFloat(-v.f);
      if (v.isVec()) return Const::ofVec(v.vecLen(), -v.v[0], -v.v[1], -v.v[2]);
      return Const();
    case Op::Not:
      return v.k == TK::Bool ? Const::ofBool(!v.b) : Const();
    case Op::BitNot:
      return v.k == TK::Int ? Const::ofInt(~v.i) : Const();
    default: return Const();
  }
}

static Const foldBinary(Op op, const Const &l, const Const &r) {
  const bool ints = l.k == TK::Int && r.k == TK::Int;
  const bool nums = l.isNum() && r.isNum();
  const bool vecs = l.isVec() && l.k == r.k;
  switch (op) {
    case Op::Add:
      if (ints) return Const::ofInt(l.i + r.i);
      if (nums) return Const::ofFloat(l.num() + r.num());
      if (vecs) return Const::ofVec(l.vecLen(), l.v[0] + r.v[0], l.v[1] + r.v[1], l.v[2] + r.v[2]);
      return Const();
    case Op::Sub:
      if (ints) return Const::ofInt(l.i - r.i);
      if (nums) return Const::ofFloat(l.num() - r.num());
      if (vecs) return Const::ofVec(l.vecLen(), l.v[0] - r.v[0], l.v[1] - r.v[1], l.v[2] - r.v[2]);
      return Const();
    case Op::Mul:
      if (ints) return Const::ofInt(l.i * r.i);
      if (nums) return Const::ofFloat(l.num() * r.num());
      if (l.isVec() && r.isNum()) {
        double f = r.num();
        return Const::ofVec(l.vecLen(), l.v[0] * f, l.v[1] * f, l.v[2] * f);
      }
      if (l.isNum() && r.isVec()) {
        double f = l.num();
        return Const::ofVec(r.vecLen(), r.v[0] * f, r.v[1] * f, r.v[2] * f);
      }
      return Const();
    case Op::Div:
      if (ints) return r.i == 0 ? Const() : Const::ofInt(l.i / r.i);
      if (nums) return Const::ofFloat(l.num() / r.num());
      if (l.isVec() && r.isNum()) {
        double f = r.num();
        return Const::ofVec(l.vecLen(), l.v[0] / f, l.v[1] / f, l.v[2] / f);
      }
      return Const();
    case Op::Mod:
      if (ints) return r.i == 0 ? Const() : Const::ofInt(l.i % r.i);
      if (nums) return Const::ofFloat(fmod(l.num(), r.num()));
      return Const();
    case Op::BitOr: return ints ? Const::ofInt(l.i | r.i) : Const();
    case Op::BitAnd: return ints ? Const::ofInt(l.i & r.i) : Const();
    case Op::BitXor: return ints ? Const::ofInt(l.i ^ r.i) : Const();
    case Op::Shl: return ints ? Const::ofInt(l.i << r.i) : Const();
    case Op::Shr: return ints ? Const::ofInt(l.i >> r.i) : Const();
    case Op::Eq: case Op::Ne:
      if (vecs) {
        // The reference compares only x and y, also for float3 (Value.cpp:254-259).
        bool eq = l.v[0] == r.v[0] && l.v[1] == r.v[1];
        return Const::ofBool(op == Op::Eq ? eq : !eq);
      }
      /* fallthrough */
    case Op::Lt: case Op::Le: case Op::Gt: case Op::Ge: {
      if (!nums) return Const();
      int cmp;
      if (ints) cmp = l.i > r.i ? 1 : l.i < r.i ? -1 : 0;
      else { double a = l.num(), b = r.num(); cmp = a > b ? 1 : a < b ? -1 : 0; }
      switch (op) {
        case Op::Eq: return Const::ofBool(cmp == 0);
        case Op::Ne: return Const::ofBool(cmp != 0);
        case Op::Lt: return Const::ofBool(cmp < 0);
        case Op::Le: return Const::ofBool(cmp <= 0);
        case Op::Gt: return Const::ofBool(cmp > 0);
        default: return Const::ofBool(cmp >= 0);
      }
    }
    case Op::Or: case Op::And:
      if (l.k != TK::Bool || r.k != TK::Bool) return Const();
      return Const::ofBool(op == Op::Or ? (l.b || r.b) : (l.b && r.b));
    default: return Const();
  }
}

static Const foldMath(const std::string &fn, const std::vector<Const> &args) {
  if (args.size() == 1) {
    if (!args[0].isNum()) return Const();
    double x = args[0].num();
    if (fn == "sin") return Const::ofFloat(sin(x));
    if (fn == "cos") return Const::ofFloat(cos(x));
    if (fn == "tan") return Const::ofFloat(tan(x));
    if (fn == "sinh") return Const::ofFloat(sinh(x));
    if (fn == "cosh") return Const::ofFloat(cosh(x));
    if (fn == "tanh") return Const::ofFloat(tanh(x));
    if (fn == "asin") return Const::ofFloat(asin(x));
    if (fn == "acos") return Const::ofFloat(acos(x));
    if (fn == "atan") return Const::ofFloat(atan(x));
    if (fn == "exp") return Const::ofFloat(exp(x));
    if (fn == "log") return Const::ofFloat(log(x));
    if (fn == "sqrt") return Const::ofFloat(sqrt(x));
    if (fn == "cbrt") return Const::ofFloat(cbrt(x));
    if (fn == "round") return Const::ofFloat(round(x));
    return Const();
  }
  if (args.size() == 2 && fn == "pow" && args[0].isNum() && args[1].isNum())
    return Const::ofFloat(pow(args[0].num(), args[1].num()));
  return Const();
}

static Const toFloatImplicit(const Const &c) {
  return c.isNum() ? Const::ofFloat(c.num()) : Const();
}

Const Sema::eval(const Expr &e) const {
  switch (e.kind) {
    case Expr::BoolLit: return Const::ofBool(e.bval);
    case Expr::IntLit: return Const::ofInt(e.ival);
    case Expr::FloatLit: return Const::ofFloat(e.fval);
    case Expr::StrLit: return Const::ofStr(e.name);
    case Expr::Var: return e.sym ? e.sym->value : Const();
    case Expr::Unary: {
      Const v = eval(*e.kids[0]);
      return v.valid() ? foldUnary(e.op, v) : Const();
    }
    case Expr::Binary: {
      Const l = eval(*e.kids[0]), r = eval(*e.kids[1]);
      return (l.valid() && r.valid()) ? foldBinary(e.op, l, r) : Const();
    }
    case Expr::Call: {
      if (e.ckind == Expr::Ctor) {
        if (e.kids.empty()) return Const();
        Const a = eval(*e.kids[0]);
        switch (e.type.k) {
          case TK::Bool:
            if (a.k == TK::Bool) return a;
            if (a.k == TK::Int) return Const::ofBool(a.i != 0);
            if (a.k == TK::Float) return Const::ofBool(a.f != 0);
            return Const();
          case TK::Int:
            if (a.k == TK::Int) return a;
            if (a.k == TK::Float) return Const::ofInt((long)a.f);
            if (a.k == TK::Bool) return Const::ofInt(a.b);
            return Const();
          case TK::Float:
            if (a.k == TK::Float) return a;
            if (a.k == TK::Int) return Const::ofFloat((double)a.i);
            if (a.k == TK::Bool) return Const::ofFloat(a.b);
            return Const();
          case TK::Vec2:
          case TK::Vec3: {
            int n = e.type.vecLen();
            double v[3] = {0, 0, 0};
            if (e.kids.size() == 1) {
              Const f = toFloatImplicit(a);
              if (!f.valid()) return Const();
              v[0] = v[1] = v[2] = f.f;
            } else {
              for (int i = 0; i < n; i++) {
                Const f = toFloatImplicit(eval(*e.kids[i]));
                if (!f.valid()) return Const();
                v[i] = f.f;
              }
            }
            return Const::ofVec(n, v[0], v[1], n == 3 ? v[2] : 0);
          }
          default: return Const();
        }
      }
      if (e.ckind == Expr::Builtin) {
        std::vector<Const> args;
        for (const ExprP &k : e.kids) {
          Const a = eval(*k);
          if (!a.valid()) return Const();
          args.push_back(a);
        }
        return foldMath(e.target, args);
      }
      return Const();
    }
    default: return Const();
  }
}

// ---------------------------------------------------------------------------
// Construction / builtin function table (reference main.cpp:26-131)
// ---------------------------------------------------------------------------
Sema::Sema(Script &mainScript, const std::map<std::string, std::string> &cliParams,
           const std::string &backend)
    : script(mainScript), cliParams(cliParams), backend(backend) {
  static FuncDecl globalContext;  // "current function" while analysing global initialisers
  curFunc = &globalContext;
  registerBuiltins();
}

void Sema::addBuiltin(const std::string &name, const std::string &emit, std::vector<Ty> params,
                      Ty ret, unsigned flags, Signature::Special sp) {
  Signature s;
  s.name = name; s.emitName = emit; s.params = std::move(params); s.ret = ret;
  s.flags = flags; s.special = sp;
  functions[name].push_back(s);
}

void Sema::registerBuiltins() {
  const Ty F(TK::Float), I(TK::Int), V2(TK::Vec2), V3(TK::Vec3), V(TK::Void);
  addBuiltin("dot", "dot_float2", {V2, V2}, F);
  addBuiltin("dot", "dot_float3", {V3, V3}, F);
  addBuiltin("length", "length_float2", {V2}, F);
  addBuiltin("length", "length_float3", {V3}, F);
  addBuiltin("dist", "dist_float2", {V2, V2}, F);
  addBuiltin("dist", "dist_float3", {V3, V3}, F);
  addBuiltin("normalize", "normalize_float2", {V2}, V2);
  addBuiltin("normalize", "normalize_float3", {V3}, V3);
  addBuiltin("random", "random_float", {F, F}, F);
  addBuiltin("randomInt", "random_int", {I, I}, I);
  for (const char *fn : {"sin", "cos", "tan", "sinh", "cosh", "tanh", "asin", "acos", "atan",
                         "exp", "log", "sqrt", "cbrt", "round"})
    addBuiltin(fn, fn, {F}, F);
  addBuiltin("pow", "pow", {F, F}, F);
  addBuiltin("min", "min", {F, F}, F);
  addBuiltin("max", "max", {F, F}, F);

  addBuiltin("add", "add", {Ty(TK::Agent)}, V);
  addBuiltin("removeCurrent", "removeCurrent", {}, V);
  addBuiltin("near", "near", {Ty(TK::Agent), F}, Ty::arrayOf(Ty(TK::Agent)));
  addBuiltin("save", "save", {Ty(TK::String)}, V);

  const unsigned SEQ = Signature::SEQ_STEP_ONLY;
  addBuiltin("count", "count", {Ty(TK::AgentType)}, I, SEQ);
  addBuiltin("getLastExecTime", "getLastExecTime", {}, F, SEQ);
  addBuiltin("sum", "sum", {Ty(TK::AgentMember)}, Ty(TK::Unresolved), SEQ, Signature::Sum);
  addBuiltin("count", "count_member", {Ty(TK::AgentMember), Ty(TK::Unresolved)}, I, SEQ,
             Signature::CountMember);
  addBuiltin("log_csv", "log_csv", {}, V, SEQ, Signature::LogCsv);
}

const Signature *Sema::findCompatible(const std::vector<Signature> &sigs,
                                      const std::vector<Ty> &args) const {
  for (const Signature &s : sigs) {
    bool ok;
    switch (s.special) {
      case Signature::Sum:
        ok = args.size() == 1 && args[0].isAgentMember();
        break;
      case Signature::CountMember:
        ok = args.size() == 2 && args[0].isAgentMember() && args[0].member &&
             args[1].fits(args[0].member->type, false);
        break;
      case Signature::LogCsv:
        ok = true;
        for (const Ty &t : args) ok = ok && (t.isInt() || t.isFloat());
        break;
      default:
        ok = args.size() == s.params.size();
        for (size_t i = 0; ok && i < args.size(); i++) ok = args[i].fits(s.params[i], true);
    }
    if (ok) return &s;
  }
  return nullptr;
}

Signature Sema::concretize(const Signature &sig, const std::vector<Ty> &args) const {
  Signature out = sig;
  switch (sig.special) {
    case Signature::Sum: {
      out.params = args;
      Ty mt = args[0].member->type;
      out.ret = mt.isBool() ? Ty(TK::Int) : mt;  // bools sum to int
      return out;
    }
    case Signature::CountMember:
    case Signature::LogCsv:
      out.params = args;
      return out;
    default: break;
  }
  // Replace "any agent" placeholders by the concrete argument type.
  Ty agentTy(TK::Agent);
  for (size_t i = 0; i < sig.params.size(); i++) {
    const Ty &p = sig.params[i];
    if ((p.isAgent() || p.isAgentType() || p.isAgentMember()) && !p.agent) {
      agentTy = args[i];
      out.params[i] = args[i];
    } else if (p.isArray() && p.base == TK::Agent && !p.agent) {
      agentTy = args[i].elem();
      out.params[i] = Ty::arrayOf(agentTy);
    }
  }
  if (sig.ret.isAgent() && !sig.ret.agent) out.ret = agentTy;
  else if (sig.ret.isArray() && sig.ret.base == TK::Agent && !sig.ret.agent)
    out.ret = Ty::arrayOf(agentTy);
  return out;
}

// ---------------------------------------------------------------------------
// Scopes
// ---------------------------------------------------------------------------
Ty Sema::resolveType(const std::string &name, int line) {
  TK k = builtinTypeKind(name);
  if (k != TK::Invalid) return Ty(k);
  auto it = agentsByName.find(name);
  if (it == agentsByName.end()) {
    error("Unknown type \"" + name + "\"", line);
    return Ty();
  }
  return Ty::agentOf(it->second);
}

Symbol *Sema::declare(const std::string &name, int line, Ty type, bool immutable, bool global,
                      const Const &val) {
  auto it = names.find(name);
  if (it != names.end()) {
    // A local may shadow a global, nothing else may be redeclared.
    if (!it->second->global || global) {
      error("Cannot redeclare variable \"" + name + "\"", line);
      return nullptr;
    }
  }
  Symbol *s = new Symbol;
  s->name = name; s->type = type; s->immutable = immutable; s->global = global; s->value = val;
  s->uid = ++nextUid;
  script.symbols.emplace_back(s);
  // Reference quirk (AnalysisVisitor.cpp:63-79 uses map::insert): when a local shadows
  // a global the *name keeps resolving to the global entry*.  Typing decisions that
  // follow (e.g. which vector helper an operator lowers to) must match the reference,
  // so the quirk is preserved.
  if (it == names.end()) names[name] = s;
  return s;
}

// ---------------------------------------------------------------------------
// Promotion and operator typing
// ---------------------------------------------------------------------------
bool Sema::promote(ExprP &e, const Ty &want) {
  if (e->type.same(want)) return true;
  if (e->type.isInt() && want.isFloat()) {
    if (e->kind == Expr::IntLit) {
      ExprP f(new Expr(Expr::FloatLit, e->line));
      f->fval = (double)e->ival;
      f->type = Ty(TK::Float);
      e = std::move(f);
    } else {
      ExprP cast(new Expr(Expr::Call, e->line));
      cast->name = "float";
      cast->ckind = Expr::Ctor;
      cast->type = Ty(TK::Float);
      cast->kids.push_back(std::move(e));
      e = std::move(cast);
    }
    return true;
  }
  return false;
}

Ty Sema::binaryType(Op op, ExprP &l, ExprP &r) {
  Ty lt = l->type, rt = r->type;
  auto promoteEither = [&]() { return promote(l, r->type) || promote(r, l->type); };
  switch (op) {
    case Op::Add: case Op::Sub:
      if (!lt.isNumOrVec() || !rt.isNumOrVec()) return Ty();
      if (!lt.same(rt)) return promoteEither() ? Ty(TK::Float) : Ty();
      return lt;
    case Op::Mul: case Op::Div:
      if (!lt.isNumOrVec() || !rt.isNumOrVec()) return Ty();
      if (lt.isVec() && rt.isVec()) return Ty();
      if (lt.isVec()) { promote(r, Ty(TK::Float)); return lt; }
      if (rt.isVec()) {
        if (op == Op::Div) return Ty();
        promote(l, Ty(TK::Float));
        return rt;
      }
      if (!lt.same(rt)) return promoteEither() ? Ty(TK::Float) : Ty();
      return lt;
    case Op::Mod:
      if (lt.isInt() && rt.isInt()) return Ty(TK::Int);
      return promoteEither() ? Ty(TK::Float) : Ty();
    case Op::BitOr: case Op::BitAnd: case Op::BitXor: case Op::Shl: case Op::Shr:
      return (lt.isInt() && rt.isInt()) ? Ty(TK::Int) : Ty();
    case Op::Eq: case Op::Ne:
      if (lt.isNum() && rt.isNum()) return Ty(TK::Bool);
      if (lt.isVec() && lt.same(rt)) return Ty(TK::Bool);
      return Ty();
    case Op::Lt: case Op::Le: case Op::Gt: case Op::Ge:
      return (lt.isNum() && rt.isNum()) ? Ty(TK::Bool) : Ty();
    case Op::Or: case Op::And:
      return (lt.isBool() && rt.isBool()) ? Ty(TK::Bool) : Ty();
    case Op::Range:
      return (lt.isInt() && rt.isInt()) ? Ty::arrayOf(Ty(TK::Int)) : Ty();
    default: return Ty();
  }
}

bool Sema::isImmutableTarget(const Expr &e) const {
  if (e.kind == Expr::Var) return e.sym ? e.sym->immutable : false;
  if (e.kind == Expr::Member || e.kind == Expr::Index) return isImmutableTarget(*e.kids[0]);
  return false;
}

// ---------------------------------------------------------------------------
// Expressions
// ---------------------------------------------------------------------------
void Sema::expr(ExprP &e) {
  switch (e->kind) {
    case Expr::BoolLit: e->type = Ty(TK::Bool); return;
    case Expr::IntLit: e->type = Ty(TK::Int); return;
    case Expr::FloatLit: e->type = Ty(TK::Float); return;
    case Expr::StrLit: e->type = Ty(TK::String); return;

    case Expr::Var: {
      auto it = names.find(e->name);
      if (it == names.end()) {
        error("Use of undeclared variable " + e->name, e->line);
        return;
      }
      e->sym = it->second;
      e->type = it->second->type;
      return;
    }

    case Expr::Unary: {
      expr(e->kids[0]);
      Ty t = e->kids[0]->type;
      if (t.invalid()) return;
      bool ok = false;
      switch (e->op) {
        case Op::Pos: case Op::Neg: ok = t.isNumOrVec(); break;
        case Op::Not: ok = t.isBool(); break;
        case Op::BitNot: ok = t.isInt(); break;
        default: break;
      }
      if (!ok) {
        error(std::string("Type mismatch: Applying unary operator \"") + opSigil(e->op) +
              "\" to " + t.str(), e->line);
        return;
      }
      e->type = t;
      return;
    }

    case Expr::Binary: {
      expr(e->kids[0]);
      expr(e->kids[1]);
      if (e->kids[0]->type.invalid() || e->kids[1]->type.invalid()) return;
      e->type = binaryType(e->op, e->kids[0], e->kids[1]);
      if (e->type.invalid()) {
        error("Type mismatch (" + e->kids[0]->type.str() + " " + opSigil(e->op) + " " +
              e->kids[1]->type.str() + ")", e->line);
        return;
      }
      // canonical form: vector * scalar
      if (e->op == Op::Mul && e->kids[1]->type.isVec()) std::swap(e->kids[0], e->kids[1]);
      return;
    }

    case Expr::Ternary: {
      expr(e->kids[0]);
      expr(e->kids[1]);
      expr(e->kids[2]);
      Ty a = e->kids[1]->type, b = e->kids[2]->type;
      if (a.invalid() || b.invalid()) return;
      if (promote(e->kids[1], b)) e->type = b;
      else if (promote(e->kids[2], a)) e->type = a;
      else error("Branches of ternary operator have divergent types " + a.str() + " and " +
                 b.str(), e->line);
      return;
    }

    case Expr::Member: {
      expr(e->kids[0]);
      Ty t = e->kids[0]->type;
      if (t.isVec()) {
        bool ok = e->name == "x" || e->name == "y" || (t.k == TK::Vec3 && e->name == "z");
        if (!ok) { error("Vector has no member \"" + e->name + "\"", e->line); return; }
        e->type = Ty(TK::Float);
      } else if (t.isAgent() || t.isAgentType()) {
        AgentMember *m = t.agent ? t.agent->find(e->name) : nullptr;
        if (!m) { error("Agent has no member \"" + e->name + "\"", e->line); return; }
        e->type = t.isAgent() ? m->type : Ty::memberOf(t.agent, m);
      } else {
        error("Can only access members on agent or vector type", e->line);
        return;
      }
      if (e->kids[0]->kind == Expr::Var && nearVar && e->kids[0]->sym == nearVar)
        curFunc->nearMembers.insert(e->name);
      return;
    }

    case Expr::EnvAccess: {
      if (!script.env) {
        error("Cannot access environment prior to its declaration", e->line);
        return;
      }
      ExprP repl;
      if (e->name == "max") repl = constToExpr(script.env->envMax);
      else if (e->name == "min") repl = constToExpr(script.env->envMin);
      else { error("Unknown environment member \"" + e->name + "\"", e->line); return; }
      if (repl) e = std::move(repl);
      return;
    }

    case Expr::Index: {
      expr(e->kids[0]);
      expr(e->kids[1]);
      if (!e->kids[0]->type.isArray()) {
        error("Can only index into arrays", e->kids[0]->line);
        return;
      }
      if (!e->kids[1]->type.isInt()) {
        error("Array offset must be an integer", e->kids[1]->line);
        return;
      }
      e->type = e->kids[0]->type.elem();
      return;
    }

    case Expr::Call:
      call(e);
      return;

    case Expr::AgentCreate: {
      for (ExprP &k : e->kids) expr(k);
      auto it = agentsByName.find(e->name);
      if (it == agentsByName.end()) {
        error("Unknown agent type \"" + e->name + "\"", e->line);
        return;
      }
      AgentDecl *a = it->second;
      std::set<std::string> seen;
      for (size_t i = 0; i < e->kids.size(); i++) {
        AgentMember *m = a->find(e->initNames[i]);
        if (!m) {
          error("Agent has no member \"" + e->initNames[i] + "\"", e->initLines[i]);
          return;
        }
        Ty have = e->kids[i]->type;
        if (have.invalid()) return;
        if (!promote(e->kids[i], m->type)) {
          error("Trying to initialize member of type " + m->type.str() +
                " from expression of type " + have.str(), e->kids[i]->line);
          return;
        }
        seen.insert(m->name);
      }
      for (auto &m : a->members) {
        if (!seen.count(m->name)) {
          error("Agent member \"" + m->name + "\" has not been initialized", e->line);
          return;
        }
      }
      e->type = Ty::agentOf(a);
      return;
    }

    case Expr::ArrayInit:
      for (ExprP &k : e->kids) expr(k);
      return;

    case Expr::NewArray: {
      e->elemTy = resolveType(e->name, e->line);
      expr(e->kids[0]);
      if (e->elemTy.invalid()) return;
      e->type = Ty::arrayOf(e->elemTy);
      return;
    }
  }
}

static bool ctorArgsValid(TK k, const std::vector<Ty> &args) {
  switch (k) {
    case TK::Bool: case TK::Int: case TK::Float:
      return args.size() == 1 && (args[0].isBool() || args[0].isNum());
    case TK::Vec2:
      if (args.size() != 1 && args.size() != 2) return false;
      break;
    case TK::Vec3:
      if (args.size() != 1 && args.size() != 3) return false;
      break;
    default: return false;
  }
  for (const Ty &t : args) if (!t.fits(Ty(TK::Float), true)) return false;
  return true;
}

void Sema::call(ExprP &e) {
  std::vector<Ty> args;
  for (ExprP &k : e->kids) { expr(k); args.push_back(k->type); }
  for (const Ty &t : args) if (t.invalid()) return;

  TK ctor = builtinTypeKind(e->name);
  if (ctor != TK::Invalid) {
    if (!ctorArgsValid(ctor, args)) {
      error("Type constructor called with invalid arguments: " + e->name + argList(args), e->line);
      return;
    }
    e->ckind = Expr::Ctor;
    e->type = Ty(ctor);
    return;
  }

  auto fit = functions.find(e->name);
  if (fit == functions.end()) {
    error("Call to unknown function \"" + e->name + "\"", e->line);
    return;
  }
  const Signature *sig = findCompatible(fit->second, args);
  if (!sig) {
    std::string msg = "Function called with invalid arguments: " + e->name + argList(args) +
                      ", expected ";
    bool first = true;
    for (const Signature &s : fit->second) {
      if (!first) msg += " or ";
      first = false;
      msg += e->name + argList(s.params);
    }
    error(msg, e->line);
    return;
  }
  if (sig->decl && sig->decl->kind != FuncDecl::Normal) {
    error("Cannot directly call step function " + e->name + "()", e->line);
    return;
  }
  if ((sig->flags & Signature::SEQ_STEP_ONLY) && !curFunc->isSeqStep()) {
    error(e->name + "() can only be used inside a sequential step function", e->line);
    return;
  }

  if (e->name == "removeCurrent") {
    if (!curFunc->isStep()) {
      error("removeCurrent() can only be used inside a step function", e->line);
      return;
    }
    curFunc->usesRemoval = true;
    if (AgentDecl *a = curFunc->stepAgent()) a->usesRemoval = true;
    script.usesRemoval = true;
  }

  if (e->name == "add") {
    if (!curFunc->isStep() && !curFunc->isMain()) {
      error("add() can only be used in main() or a step function", e->line);
      return;
    }
    if (curFunc->isStep()) {
      const Expr &arg = *e->kids[0];
      if (arg.kind != Expr::AgentCreate) {
        error("Argument of add() must be an agent creation expression", e->line);
        return;
      }
      auto it = agentsByName.find(arg.name);
      if (it == agentsByName.end()) return;
      if (curFunc->addedAgent) {
        error("Only one agent per step function may be added at runtime", e->line);
        return;
      }
      curFunc->addedAgent = it->second;
      it->second->receivesAdds = true;
      script.usesAddition = true;
    }
  }

  Signature conc = concretize(*sig, args);
  e->ckind = sig->decl ? Expr::User : Expr::Builtin;
  e->target = conc.emitName;
  e->type = conc.ret;
  e->callee = sig->decl;
  e->paramTys = conc.params;

  if (e->name == "count" || e->name == "sum") {
    Reduction r;
    r.kind = conc.emitName == "count" ? Reduction::CountType
           : conc.emitName == "count_member" ? Reduction::CountMember : Reduction::SumMember;
    r.agent = conc.params[0].agent;
    r.member = conc.params[0].member;
    script.reductions.insert(r);
  }
  if (e->name == "log_csv") script.usesLogging = true;
  if (e->name == "getLastExecTime") script.usesTiming = true;
  if (e->name == "random" || e->name == "randomInt") curFunc->usesRng = true;
  if (sig->decl && sig->decl->usesRng) curFunc->usesRng = true;
}

// ---------------------------------------------------------------------------
// Statements
// ---------------------------------------------------------------------------
void Sema::stmt(Stmt &s) {
  switch (s.kind) {
    case Stmt::ExprS:
      expr(s.e[0]);
      return;

    case Stmt::Block:
      pushScope();
      for (StmtP &b : s.body) stmt(*b);
      popScope();
      return;

    case Stmt::VarDecl: {
      s.declTy = resolveType(s.typeName, s.typeLine);
      if (!s.e.empty()) expr(s.e[0]);
      s.sym = declare(s.varName, s.varLine, s.declTy, false, false, Const());
      if (s.e.empty()) {
        error("Variable declaration must have an initializer", s.line);
        return;
      }
      Ty have = s.e[0]->type;
      if (s.declTy.invalid() || have.invalid()) return;
      if (!promote(s.e[0], s.declTy))
        error("Trying to assign value of type " + have.str() + " to variable of type " +
              s.declTy.str(), s.e[0]->line);
      return;
    }

    case Stmt::Assign: {
      expr(s.e[0]);
      expr(s.e[1]);
      if (isImmutableTarget(*s.e[0])) {
        error("Trying to assign to immutable variable", s.e[0]->line);
        return;
      }
      Ty lt = s.e[0]->type, rt = s.e[1]->type;
      if (lt.invalid() || rt.invalid()) return;
      if (!rt.fits(lt, true))
        error("Cannot assign value of type " + rt.str() + " to variable of type " + lt.str(),
              s.e[1]->line);
      return;
    }

    case Stmt::AssignOp:
      expr(s.e[0]);
      expr(s.e[1]);
      if (isImmutableTarget(*s.e[0]))
        error("Trying to assign to immutable variable", s.e[0]->line);
      return;

    case Stmt::If: {
      expr(s.e[0]);
      for (StmtP &b : s.body) stmt(*b);
      Ty t = s.e[0]->type;
      if (!t.invalid() && !t.isBool())
        error("if() condition must be bool, but received " + t.str(), s.e[0]->line);
      return;
    }

    case Stmt::While: {
      loopDepth++;
      expr(s.e[0]);
      stmt(*s.body[0]);
      loopDepth--;
      Ty t = s.e[0]->type;
      if (!t.invalid() && !t.isBool())
        error("while() condition must be bool, but received " + t.str(), s.e[0]->line);
      return;
    }

    case Stmt::For: {
      loopDepth++;
      pushScope();
      Ty declTy = resolveType(s.typeName, s.typeLine);
      s.declTy = declTy;
      s.sym = declare(s.varName, s.varLine, declTy, true, false, Const());
      bool near = false;
      if (s.e[0]->kind == Expr::Call && s.e[0]->name == "near") {
        if (!declTy.isAgent()) {
          error("Type specified in for-near loop is not an agent", s.typeLine);
        } else if (curFunc->nearAgent) {
          error("Multiple for-near loops in a single step function", s.line);
        } else if (!declTy.agent->position()) {
          error("Cannot use for-near loop on agent without position member", s.line);
        } else {
          curFunc->nearAgent = declTy.agent;
          s.forKind = Stmt::ForNear;
          nearVar = s.sym;
          near = true;
        }
      }
      expr(s.e[0]);
      stmt(*s.body[0]);
      loopDepth--;
      popScope();

      if (near) {
        nearVar = nullptr;
        Const r = s.e[0]->kids.size() > 1 ? eval(*s.e[0]->kids[1]) : Const();
        radii.push_back(r);
        curFunc->nearRadius = r;
        return;
      }
      Ty et = s.e[0]->type;
      if (!et.isArray()) {
        error("Can only use for with array type, received " + et.str(), s.e[0]->line);
        return;
      }
      if (!et.elem().fits(declTy, false)) {
        error("For expression type " + et.str() + " not compatible with declared " +
              declTy.str(), s.e[0]->line);
        return;
      }
      s.forKind = (s.e[0]->kind == Expr::Binary && s.e[0]->op == Op::Range) ? Stmt::ForRange
                                                                             : Stmt::ForArray;
      return;
    }

    case Stmt::Simulate: {
      expr(s.e[0]);
      if (script.simulate) {
        error("Script can only contain a single simulate statement", s.line);
        return;
      }
      if (!s.e[0]->type.isInt()) {
        error("Number of timesteps must be an integer, " + s.e[0]->type.str() + " given",
              s.e[0]->line);
        return;
      }
      for (const std::string &name : s.stepNames) {
        auto it = funcsByName.find(name);
        if (it == funcsByName.end()) {
          error("Unknown step function \"" + name + "\"", s.line);
          return;
        }
        FuncDecl *f = it->second;
        if (f->kind == FuncDecl::Normal) {
          error("Function \"" + name + "\" is not a step function", s.line);
          return;
        }
        if (f->isSeqStep()) {
          if (s.seqStep) { error("Can only use single sequential step function", s.line); return; }
          s.seqStep = f;
        } else {
          if (s.seqStep) { error("Sequential step function must be last", s.line); return; }
          s.steps.push_back(f);
        }
      }
      script.simulate = &s;
      return;
    }

    case Stmt::Return: {
      if (!s.e.empty()) expr(s.e[0]);
      Ty want = curFunc->retTy;
      if (want.isVoid()) {
        if (!s.e.empty()) error("Cannot return value from void function", s.e[0]->line);
        return;
      }
      if (s.e.empty()) {
        error("Return from non-void function must specify value", s.line);
        return;
      }
      Ty have = s.e[0]->type;
      if (!promote(s.e[0], want))
        error("Trying to return " + have.str() + " from function with return type " + want.str(),
              s.e[0]->line);
      return;
    }

    case Stmt::Break:
      if (loopDepth == 0) error("Cannot use break outside a loop", s.line);
      return;
    case Stmt::Continue:
      if (loopDepth == 0) error("Cannot use continue outside a loop", s.line);
      return;
  }
}

// ---------------------------------------------------------------------------
// Declarations
// ---------------------------------------------------------------------------
void Sema::agent(AgentDecl &a) {
  if (agentsByName.count(a.name)) {
    error("Redefinition of agent " + a.name, a.line);
  } else {
    agentsByName[a.name] = &a;
    script.agents.push_back(&a);
    // the type name itself is a value, for count(Agent) / sum(Agent.member)
    Symbol *s = new Symbol;
    s->name = a.name; s->type = Ty::agentType(&a); s->immutable = true; s->global = true;
    s->uid = ++nextUid;
    script.symbols.emplace_back(s);
    if (!names.count(a.name)) names[a.name] = s;
  }
  for (auto &m : a.members) m->type = resolveType(m->typeName, m->typeLine);
}

static bool isConstantInitializer(const Expr &e) {
  switch (e.kind) {
    case Expr::BoolLit: case Expr::IntLit: case Expr::FloatLit: case Expr::StrLit:
    case Expr::Unary:
      return true;
    case Expr::ArrayInit:
      for (const ExprP &k : e.kids) if (!isConstantInitializer(*k)) return false;
      return true;
    default: return false;
  }
}

void Sema::constant(ConstDecl &c) {
  script.consts.push_back(&c);
  c.type = resolveType(c.typeName, c.typeLine);
  expr(c.init);

  Const val;
  if (c.isArray) {
    Ty elemTy = c.type;
    c.type = Ty::arrayOf(elemTy);
    if (c.init->kind != Expr::ArrayInit) {
      error("Array must be initialized using array initializer", c.init->line);
      return;
    }
    for (ExprP &k : c.init->kids) {
      if (!promote(k, elemTy)) {
        error("Element of type " + k->type.str() + " inside initializer for array of type " +
              elemTy.str(), k->line);
        return;
      }
    }
    c.init->type = c.type;
    if (!isConstantInitializer(*c.init)) {
      error("Initializer of global constant must be a constant expression", c.init->line);
      return;
    }
  } else {
    val = eval(*c.init);
    if (!val.valid()) {
      error("Initializer of global constant must be a constant expression", c.init->line);
      return;
    }
    int line = c.init->line;
    c.init = constToExpr(val);  // backends only ever see the folded literal
    c.init->line = line;
  }

  if (!promote(c.init, c.type)) {
    error("Trying to assign value of type " + c.init->type.str() + " to global of type " +
          c.type.str(), c.init->line);
    return;
  }

  auto it = cliParams.find(c.name);
  if (it != cliParams.end()) {
    if (!c.isParam) {
      error("Only constants marked with \"param\" can be specified as parameters", c.nameLine);
      return;
    }
    val = parseCliValue(it->second);
    if (!val.valid()) {
      error("Value \"" + it->second + "\" provided for parameter \"" + c.name +
            "\" could not parsed", c.nameLine);
      return;
    }
    Ty vt(val.k);
    if (!vt.fits(c.type, true)) {
      error("Provided parameter \"" + c.name + "\" is of type " + vt.str() + ", but " +
            c.type.str() + " expected", c.nameLine);
      return;
    }
    c.init = constToExpr(val);
  }
  if (c.isParam) script.params.insert(c.name);
  // NB: the symbol keeps the *unpromoted* value (an `param float x = 500` stays the
  // integer 500 for later folding, so `n / x` folds as integer division) — this is
  // what the reference does (AnalysisVisitor.cpp:422-471) and it decides e.g. boids2d's
  // max_pos.
  c.sym = declare(c.name, c.nameLine, c.type, true, true, val);
}

void Sema::environment(EnvDecl &e) {
  for (ExprP &v : e.values) expr(v);
  if (script.env) {
    error("Script can only contain a single environment specification", e.line);
    return;
  }
  for (size_t i = 0; i < e.names.size(); i++) {
    const Expr &x = *e.values[i];
    Const v = eval(x);
    if (!v.valid()) {
      error("Environment member \"" + e.names[i] + "\" must be a constant expression", x.line);
      return;
    }
    if (e.names[i] == "min") {
      if (!x.type.isVec()) { error("Environment min bound must be float2 or float3", x.line); return; }
      e.envMin = v;
    } else if (e.names[i] == "max") {
      if (!x.type.isVec()) { error("Environment max bound must be float2 or float3", x.line); return; }
      e.envMax = v;
    } else if (e.names[i] == "granularity") {
      if (!x.type.isNum()) { error("Environment granularity must be a number", x.line); return; }
      e.granularity = v;
    } else {
      error("Unknown environment member \"" + e.names[i] + "\"", e.lines[i]);
      return;
    }
  }
  if (e.envMax.valid()) {
    e.dim = e.envMax.vecLen();
    if (e.envMin.valid()) {
      if (e.envMin.k != e.envMax.k) {
        error("min and max environment bounds must have the same type", e.line);
        return;
      }
    } else {
      e.envMin = Const::ofVec(e.dim, 0, 0, 0);
    }
    e.envSize = foldBinary(Op::Sub, e.envMax, e.envMin);
    for (int i = 0; i < e.dim; i++) {
      if (e.envSize.v[i] < 0) {
        error("Environment minimum should be smaller or equal than the maximum", e.line);
        return;
      }
    }
  }
  script.env = &e;
}

void Sema::function(FuncDecl &f) {
  pushScope();
  FuncDecl *savedFunc = curFunc;
  bool registered = false;

  f.retTy = resolveType(f.retTypeName, f.retLine);
  bool typesOk = !f.retTy.invalid();
  std::vector<Ty> paramTys;
  if (typesOk) {
    for (Param &p : f.params) {
      p.type = resolveType(p.typeName, p.typeLine);
      if (p.type.invalid()) { typesOk = false; break; }
      paramTys.push_back(p.type);
      if (!f.isStep() && !p.outName.empty())
        error("Out variable (-> " + p.outName + ") can only be used in step functions", p.outLine);
    }
  }

  if (typesOk) {
    bool conflict = false;
    auto fit = functions.find(f.name);
    if (fit != functions.end()) {
      for (const Signature &s : fit->second) {
        if (s.params.size() != paramTys.size()) continue;  // arity overloading is always fine
        bool differs = false, scalarClash = false;
        for (size_t i = 0; i < paramTys.size(); i++) {
          const Ty &a = paramTys[i], &b = s.params[i];
          if (!a.same(b)) {
            differs = true;
            auto scalar = [](const Ty &t) { return t.isBool() || t.isInt() || t.isFloat(); };
            if (scalar(a) && scalar(b)) scalarClash = true;
          }
        }
        if (scalarClash || !differs) {
          error("Declaration " + f.name + argList(paramTys) +
                " conflicts with previous declaration " + f.name + argList(s.params), f.line);
          conflict = true;
          break;
        }
      }
    }
    if (!conflict) {
      f.emitName = f.name;
      if (fit != functions.end()) f.emitName += "_" + std::to_string(fit->second.size());
      curFunc = &f;
      // visualisation hooks are only meaningful for the JVM backends
      if ((f.name == "getColor" || f.name == "getSize") && backend != "mason" &&
          backend != "dmason") {
        f.skipped = true;
      } else {
        Signature s;
        s.name = f.name; s.emitName = f.emitName; s.params = paramTys; s.ret = f.retTy;
        s.decl = &f;
        functions[f.name].push_back(s);
        if (!funcsByName.count(f.name)) funcsByName[f.name] = &f;  // first declaration wins
        script.funcs.push_back(&f);
        if (f.isMain()) script.mainFunc = &f;
        registered = true;
      }
    }
  }
  (void)registered;
  if (curFunc != &f && !f.skipped) {
    // Signature errors: the body is still analysed (against the enclosing function
    // context of the reference, which is the previous function) to surface more errors.
    // We use the function itself as context; it only affects follow-up diagnostics.
    curFunc = &f;
  }

  for (Param &p : f.params) {
    if (p.type.invalid()) p.type = resolveType(p.typeName, p.typeLine);
    bool inIsImmutable = !p.outName.empty();
    p.sym = declare(p.name, p.nameLine, p.type, inIsImmutable, false, Const());
    if (!p.outName.empty()) p.outSym = declare(p.outName, p.outLine, p.type, false, false, Const());
  }
  for (StmtP &s : f.body) stmt(*s);
  popScope();

  if (f.isStep()) {
    bool valid = f.params.size() == 1 && !f.params[0].outName.empty() && f.params[0].type.isAgent();
    if (!valid) error("Step function " + f.name + " does not have a valid signature", f.line);
  }
  curFunc = savedFunc;
}

// ---------------------------------------------------------------------------
// Script level
// ---------------------------------------------------------------------------
void Sema::script_(Script &s) {
  for (Decl &d : s.decls) {
    switch (d.kind) {
      case Decl::Agent: agent(*d.agent); break;
      case Decl::Func: function(*d.func); break;
      case Decl::ConstD: constant(*d.cnst); break;
      case Decl::Env: environment(*d.env); break;
    }
  }
}

void Sema::analyseLibrary(Script &lib) {
  isLib = true;
  script_(lib);
}

void Sema::analyseMain() {
  isLib = false;
  script_(script);
  finishMain();
}

void Sema::finishMain() {
  EnvDecl *env = script.env;
  for (AgentDecl *a : script.agents) {
    AgentMember *pos = a->position();
    if (!pos) continue;
    if (!env || env->dim < 0) {
      error("An environment { } declaration is required to use position members", pos->line);
      return;
    }
    if (!pos->type.isVec() || pos->type.vecLen() != env->dim) {
      error("Dimensionality of position member does not match environment dimension", pos->line);
      return;
    }
  }

  if (env && !env->granularity.valid()) {
    // automatic cell size: the largest statically known interaction radius
    Const best;
    for (const Const &r : radii) {
      if (!r.valid() || !r.isNum()) continue;
      if (!best.valid() || r.num() > best.num()) best = r;
    }
    if (!best.valid()) {
      error("Could not automatically determine partitioning granularity. "
            "Please explicitly specify it in the environment { } declaration", env->line);
      return;
    }
    env->granularity = best;
  }

  if (!script.mainFunc) {
    error("Script must have a main function", 1);
    return;
  }

  if (script.simulate) {
    bool topLevel = false;
    for (const StmtP &s : script.mainFunc->body) if (s.get() == script.simulate) topLevel = true;
    if (!topLevel) {
      error("Simulate statement cannot be used conditionally", script.simulate->line);
      return;
    }
  }

  for (const auto &kv : cliParams) {
    if (!script.params.count(kv.first)) {
      error("Unknown parameter \"" + kv.first + "\" specified through -P", 1);
      return;
    }
  }
}

}  // namespace abl
