// OpenABL-cuda front end: syntax tree.
//
// Design note: unlike the reference (one class per node + double-dispatch visitor,
// reference src/AST.hpp:36-703), this tree is three tagged structs (Expr / Stmt / Decl)
// walked by plain recursive functions.  The *information* carried is what a GPU
// backend needs from the reference's analysis: resolved types,
// folded constants, overload-resolved callee names, for-near metadata, add/remove flags.
#pragma once

#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

namespace abl {

struct AgentDecl;
struct AgentMember;
struct FuncDecl;

// ---------------------------------------------------------------------------
// Types
// ---------------------------------------------------------------------------
enum class TK : uint8_t {
  Invalid, Void, Bool, Int, Float, String, Vec2, Vec3,
  Agent,       // value of an agent type (agent == nullptr: "any agent")
  Array,       // array of `base` (arrays of agents keep `agent`)
  AgentType,   // the agent type itself, e.g. count(Prey)
  AgentMember, // a member of an agent type, e.g. sum(Grass.avail)
  Unresolved,  // builtin return type computed from arguments
};

struct Ty {
  TK k = TK::Invalid;
  TK base = TK::Invalid;          // element kind for Array
  AgentDecl *agent = nullptr;     // Agent / AgentType / AgentMember / Array-of-agent
  AgentMember *member = nullptr;  // AgentMember

  Ty() {}
  Ty(TK k) : k(k) {}
  static Ty agentOf(AgentDecl *a) { Ty t(TK::Agent); t.agent = a; return t; }
  static Ty agentType(AgentDecl *a) { Ty t(TK::AgentType); t.agent = a; return t; }
  static Ty memberOf(AgentDecl *a, AgentMember *m) {
    Ty t(TK::AgentMember); t.agent = a; t.member = m; return t;
  }
  static Ty arrayOf(const Ty &e) {
    Ty t(TK::Array); t.base = e.k; t.agent = e.agent; t.member = e.member; return t;
  }
  Ty elem() const {
    Ty t(base); if (base == TK::Agent) t.agent = agent; return t;
  }

  bool invalid() const { return k == TK::Invalid; }
  bool isVoid() const { return k == TK::Void; }
  bool isBool() const { return k == TK::Bool; }
  bool isInt() const { return k == TK::Int; }
  bool isFloat() const { return k == TK::Float; }
  bool isString() const { return k == TK::String; }
  bool isNum() const { return k == TK::Int || k == TK::Float; }
  bool isVec() const { return k == TK::Vec2 || k == TK::Vec3; }
  bool isNumOrVec() const { return isNum() || isVec(); }
  bool isAgent() const { return k == TK::Agent; }
  bool isArray() const { return k == TK::Array; }
  bool isAgentType() const { return k == TK::AgentType; }
  bool isAgentMember() const { return k == TK::AgentMember; }
  int vecLen() const { return k == TK::Vec2 ? 2 : 3; }

  bool same(const Ty &o) const {
    if (k != o.k) return false;
    if (k == TK::Array) return elem().same(o.elem());
    return agent == o.agent && member == o.member;
  }
  // `promote`: additionally accept int where float is expected.
  bool fits(const Ty &want, bool promote) const {
    if (k != want.k) return promote && k == TK::Int && want.k == TK::Float;
    if (k == TK::Agent) return agent == want.agent || want.agent == nullptr;
    if (k == TK::Array) return elem().fits(want.elem(), false);
    return true;
  }
  std::string str() const;
};

// ---------------------------------------------------------------------------
// Compile-time values (constant folding of globals, -P overrides, env bounds)
// ---------------------------------------------------------------------------
struct Const {
  TK k = TK::Invalid;
  bool b = false;
  long i = 0;
  double f = 0;
  double v[3] = {0, 0, 0};
  std::string s;

  static Const ofBool(bool x) { Const c; c.k = TK::Bool; c.b = x; return c; }
  static Const ofInt(long x) { Const c; c.k = TK::Int; c.i = x; return c; }
  static Const ofFloat(double x) { Const c; c.k = TK::Float; c.f = x; return c; }
  static Const ofStr(const std::string &x) { Const c; c.k = TK::String; c.s = x; return c; }
  static Const ofVec(int n, double x, double y, double z = 0) {
    Const c; c.k = n == 2 ? TK::Vec2 : TK::Vec3; c.v[0] = x; c.v[1] = y; c.v[2] = z; return c;
  }
  bool valid() const { return k != TK::Invalid; }
  bool isNum() const { return k == TK::Int || k == TK::Float; }
  bool isVec() const { return k == TK::Vec2 || k == TK::Vec3; }
  int vecLen() const { return k == TK::Vec2 ? 2 : 3; }
  double num() const { return k == TK::Int ? (double)i : f; }
};

// ---------------------------------------------------------------------------
// Expressions
// ---------------------------------------------------------------------------
enum class Op : uint8_t {
  // binary
  Add, Sub, Mul, Div, Mod, BitAnd, BitXor, BitOr, Shl, Shr,
  Eq, Ne, Lt, Le, Gt, Ge, And, Or, Range,
  // unary
  Neg, Pos, Not, BitNot,
};
const char *opSigil(Op op);

struct Symbol {            // one declared variable
  std::string name;
  Ty type;
  bool immutable = false;
  bool global = false;
  Const value;             // folded value for global constants
  int uid = 0;
};

struct Expr;
using ExprP = std::unique_ptr<Expr>;

struct Expr {
  enum Kind : uint8_t {
    BoolLit, IntLit, FloatLit, StrLit, Var, Unary, Binary, Call, Member,
    EnvAccess, Index, Ternary, AgentCreate, ArrayInit, NewArray,
  };
  enum CallKind : uint8_t { User, Builtin, Ctor };

  Kind kind;
  int line = 1;
  Ty type;

  bool bval = false;
  long ival = 0;
  double fval = 0;
  std::string name;             // Var/Call/AgentCreate name, Member/EnvAccess member, StrLit text, NewArray elem type
  Op op = Op::Add;
  std::vector<ExprP> kids;      // operands / args / init values
  std::vector<std::string> initNames;  // AgentCreate: member names parallel to kids
  std::vector<int> initLines;

  // filled by Sema
  Symbol *sym = nullptr;        // Var
  CallKind ckind = User;
  std::string target;           // Call: overload-resolved emitted name (e.g. random_4, dist_float2)
  FuncDecl *callee = nullptr;   // Call to a user function
  std::vector<Ty> paramTys;     // Call: concrete parameter types
  Ty elemTy;                    // NewArray

  Expr(Kind k, int line) : kind(k), line(line) {}
  const Expr *init(const std::string &member) const {
    for (size_t i = 0; i < initNames.size(); i++)
      if (initNames[i] == member) return kids[i].get();
    return nullptr;
  }
};

// ---------------------------------------------------------------------------
// Statements
// ---------------------------------------------------------------------------
struct Stmt;
using StmtP = std::unique_ptr<Stmt>;

struct Stmt {
  enum Kind : uint8_t {
    ExprS, Block, VarDecl, If, While, For, Return, Break, Continue, Simulate, Assign, AssignOp,
  };
  enum ForKind : uint8_t { ForArray, ForRange, ForNear };

  Kind kind;
  int line = 1;
  std::vector<ExprP> e;         // 0..2 expressions (cond / lhs,rhs / init / range expr)
  std::vector<StmtP> body;      // Block stmts; If: then[,else]; loops: body
  std::string typeName;         // VarDecl / For: declared type
  int typeLine = 1;
  std::string varName;
  int varLine = 1;
  Op op = Op::Add;              // AssignOp
  std::vector<std::string> stepNames;  // Simulate

  // filled by Sema
  Ty declTy;
  Symbol *sym = nullptr;
  ForKind forKind = ForArray;
  std::vector<FuncDecl *> steps;  // Simulate: parallel step functions in order
  FuncDecl *seqStep = nullptr;

  Stmt(Kind k, int line) : kind(k), line(line) {}
};

// ---------------------------------------------------------------------------
// Declarations
// ---------------------------------------------------------------------------
struct AgentMember {
  std::string name;
  std::string typeName;
  bool isPosition = false;
  int line = 1;
  int typeLine = 1;
  Ty type;
};

struct AgentDecl {
  std::string name;
  std::vector<std::unique_ptr<AgentMember>> members;
  int line = 1;
  bool usesRemoval = false;     // some step function removes agents of this type
  bool receivesAdds = false;    // some step function adds agents of this type at run time
  AgentMember *position() const {
    for (auto &m : members) if (m->isPosition) return m.get();
    return nullptr;
  }
  AgentMember *find(const std::string &n) const {
    for (auto &m : members) if (m->name == n) return m.get();
    return nullptr;
  }
  int memberIndex(const std::string &n) const {
    for (size_t i = 0; i < members.size(); i++) if (members[i]->name == n) return (int)i;
    return -1;
  }
};

struct Param {
  std::string typeName;
  std::string name, outName;    // outName empty unless `in -> out`
  int line = 1, typeLine = 1, nameLine = 1, outLine = 1;
  Ty type;
  Symbol *sym = nullptr, *outSym = nullptr;
};

struct FuncDecl {
  enum Kind : uint8_t { Normal, Step, SeqStep };
  Kind kind = Normal;
  std::string name;
  std::string emitName;         // unique name after overload numbering (name, name_1, ...)
  std::string retTypeName;
  int line = 1, retLine = 1;
  std::vector<Param> params;
  std::vector<StmtP> body;
  Ty retTy;
  bool skipped = false;         // getColor/getSize: analysed but never emitted

  // step-function facts (reference AST.hpp:536-544 equivalents)
  AgentDecl *nearAgent = nullptr;           // agent type iterated by the for-near loop
  std::set<std::string> nearMembers;        // members read from the neighbour
  Const nearRadius;                         // folded radius (invalid if dynamic)
  bool usesRemoval = false;
  AgentDecl *addedAgent = nullptr;
  bool usesRng = false;
  std::set<std::string> writtenMembers;     // members of `out` assigned (non-identity)
  bool isMain() const { return name == "main"; }
  bool isStep() const { return kind == Step; }
  bool isSeqStep() const { return kind == SeqStep; }
  AgentDecl *stepAgent() const { return params.empty() ? nullptr : params[0].type.agent; }
};

struct ConstDecl {
  std::string typeName, name;
  int line = 1, typeLine = 1, nameLine = 1;
  bool isArray = false, isParam = false;
  ExprP init;
  Ty type;
  Symbol *sym = nullptr;
};

struct EnvDecl {
  int line = 1;
  std::vector<std::string> names;
  std::vector<int> lines;
  std::vector<ExprP> values;
  Const envMin, envMax, envSize, granularity;
  int dim = -1;
};

struct Decl {
  enum Kind : uint8_t { Agent, Func, ConstD, Env } kind;
  std::unique_ptr<AgentDecl> agent;
  std::unique_ptr<FuncDecl> func;
  std::unique_ptr<ConstDecl> cnst;
  std::unique_ptr<EnvDecl> env;
};

struct Reduction {
  enum Kind : uint8_t { CountType, CountMember, SumMember } kind;
  AgentDecl *agent;
  AgentMember *member;
  bool operator<(const Reduction &o) const {
    if (kind != o.kind) return kind < o.kind;
    if (agent != o.agent) return agent < o.agent;
    return member < o.member;
  }
};

struct Script {
  std::vector<Decl> decls;
  int line = 1;

  // filled by Sema
  std::vector<AgentDecl *> agents;
  std::vector<ConstDecl *> consts;
  std::vector<FuncDecl *> funcs;
  std::set<Reduction> reductions;
  std::set<std::string> params;
  Stmt *simulate = nullptr;
  FuncDecl *mainFunc = nullptr;
  EnvDecl *env = nullptr;
  bool usesRemoval = false, usesAddition = false, usesLogging = false, usesTiming = false;
  std::vector<std::unique_ptr<Symbol>> symbols;  // owns all symbols
};

}  // namespace abl
