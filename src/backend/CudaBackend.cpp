// The B200-native `cuda` backend: code printer + build script generation.
//
// Replaces reference src/backend/CBackend.cpp + CPrinter.cpp for `-b cuda`.  From the
// analysed script it writes into outputDir:
//   model_host.c     C99 host program: agent records + type tables, folded constants,
//                    library/user functions, main() whose `simulate` hands the population
//                    to the device runtime (include/abl_cuda.h).  Everything outside
//                    `simulate` is lowered with the same expression shapes and literal
//                    formatting as the reference C printer (GenericPrinter.cpp:39-58 float
//                    literals with 6 significant digits, GenericCPrinter.cpp:20-75 vector
//                    operator lowering, CPrinter.cpp:64-173), because initial state and
//                    arithmetic must match the `c` backend bit for bit.
//   model_kernels.cu one __global__ kernel per `step` function (sm_100a) + launcher +
//                    registration of pools and steps with the runtime.
//   build.sh/run.sh  nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false + gcc -O2 -std=c99.
#include "Backend.hpp"

#include <cmath>
#include <functional>
#include <map>
#include <sstream>

#include "../FileUtil.hpp"

namespace abl {

namespace {

enum class Target { Host, Device };

// Float literal text: what `std::ostream << double` prints (6 significant digits), with
// ".0" appended when the text has no '.', as the reference does (GenericPrinter.cpp:42-50).
// Deviation: the reference appends ".0" also to exponent forms ("1e+06.0", which is not
// valid C); exponent forms are left untouched here.
std::string floatText(double v) {
  std::ostringstream s;
  s << v;
  std::string t = s.str();
  if (t.find('.') == std::string::npos && t.find('e') == std::string::npos && std::isfinite(v))
    t += ".0";
  return t;
}

std::string quoted(const std::string &s) {
  std::string out = "\"";
  for (char c : s) {
    if (c == '"' || c == '\\') out.push_back('\\');
    out.push_back(c);
  }
  return out + "\"";
}

class Writer {
public:
  Writer &operator<<(const std::string &s) { buf << s; return *this; }
  Writer &operator<<(const char *s) { buf << s; return *this; }
  Writer &operator<<(char c) { buf << c; return *this; }
  Writer &operator<<(long v) { buf << v; return *this; }
  Writer &operator<<(int v) { buf << v; return *this; }
  Writer &operator<<(unsigned v) { buf << v; return *this; }
  Writer &operator<<(size_t v) { buf << v; return *this; }
  void nl() { buf << "\n" << std::string(4 * depth, ' '); }
  void indent() { depth++; }
  void outdent() { depth--; }
  std::string str() const { return buf.str(); }
private:
  std::ostringstream buf;
  int depth = 0;
};

struct StepInfo {
  FuncDecl *fn;
  AgentDecl *self;
  std::set<std::string> reads;   // members of `in` read
  std::set<std::string> writes;  // members of `out` assigned (identity copies excluded)
  bool wholeIn = false, wholeOut = false;
};

class CudaPrinter {
public:
  CudaPrinter(Script &script, bool useFloat, const Config &config)
      : script(script), useFloat(useFloat), config(config) {}

  std::string hostSource();
  std::string kernelSource();

private:
  Script &script;
  bool useFloat;
  const Config &config;
  Writer w;
  Target target = Target::Host;
  int anon = 0;
  // device-function context
  const FuncDecl *curFn = nullptr;
  const StepInfo *curStep = nullptr;
  bool curStepHasLimit = false;
  bool curStepTile = false;     // the step's for-near loop can run from a shared-memory tile
  bool curStepFlat = false;     // `-C cuda.flat=true` and a 2-D for-near loop: ABL_MODE 3 is printed
  bool curStepList = false;     // `-C cuda.nlist=true` and a static neighbourhood: list kernels (ABL_MODE 4/5/6) are printed
  bool curStepDense = false;    // for-near loop with a host-evaluable radius: ABL_MODE 8 (single-precision shadow pre-filter, `-C cuda.dense=false` omits it)
  bool curStepSplit = false;    // ... whose loop ranges over the stepped agent itself: the pre-filter can run as a kernel of its own (ABL_MODE 9)
  bool curStepBulk = false;     // tileable 2-D flat loop: ABL_MODE 7 (rows staged by cp.async.bulk, `-C cuda.bulk=false` omits it)
  bool stepListEligible(const StepInfo &si) const;
  // one neighbour column staged in shared memory by a tiled kernel
  struct TileCol {
    int member;          // member of the neighbour agent
    int comp;            // component of a float3 member (its columns are scalar), else 0
    int column;          // SoA column index in the pool
    std::string ctype;   // element type in shared memory
    std::string bytes;   // sizeof expression of one element
  };
  std::vector<TileCol> tileColumns(const FuncDecl &f, const AgentDecl &nbr) const;
  // byte offset (per tile entry) of staged column q; and of the whole entry for q == size
  static std::string tileOffset(const std::vector<TileCol> &cols, size_t q) {
    std::string r = "(0";
    for (size_t i = 0; i < q && i < cols.size(); i++) r += " + " + cols[i].bytes;
    return r + ")";
  }
  const Stmt *findNearStmt(const std::vector<StmtP> &body) const;
  // inside a for-near body: dist(in.pos, nx.pos) / length(in.pos - nx.pos) reuse the squared
  // distance of the radius filter ((a-b)^2 == (b-a)^2 bit for bit)
  const Symbol *nearVar = nullptr, *nearSelf = nullptr;
  std::string nearPos, nearSelfPos, nearD2;
  bool isNearPair(const Expr &a, const Expr &b) const;
  void setNearContext(const Stmt &loop, const Expr &agentExpr, const std::string &pos, const std::string &selfPos,
                      const std::string &d2) {
    nearVar = loop.sym;
    nearSelf = agentExpr.kind == Expr::Var ? agentExpr.sym : nullptr;
    nearPos = pos; nearSelfPos = selfPos;
    nearD2 = nearSelf ? d2 : std::string();
    nearBody = loop.body.empty() ? nullptr : loop.body[0].get();
    sqAlias.clear();
  }
  void clearNearContext() { nearVar = nearSelf = nullptr; nearD2.clear(); nearBody = nullptr; sqAlias.clear(); }
  // `-C cuda.sqcmp=true`: comparisons of the near pair's distance with a constant become
  // comparisons of the squared distance with a host-computed bound (abl_sq_cmp_limit).
  // sqAlias: local float variables of the loop body that hold dist(in.pos, nx.pos) and are never
  // assigned again -> name of the squared-distance variable; sqLimits: the bounds of the step
  // kernel being printed, (operator, constant text), deduplicated across the loop variants.
  const Stmt *nearBody = nullptr;
  std::map<const Symbol *, std::string> sqAlias;
  std::vector<std::pair<int, std::string>> sqLimits;
  bool sqcmpOn() const { return config.getBool("cuda.sqcmp", true); }
  const std::string *nearDistSquare(const Expr &e) const;
  bool isNearDistCall(const Expr &e) const;
  bool sqCompare(const Expr &e);
  static bool assignsTo(const Stmt &s, const Symbol *sym);
  // chunked near loop: `break` of the DSL body must leave two nested C++ loops
  std::string arrayElemName(const Ty &base) const;
  std::string nearBreakLabel;
  std::string nearContinueLabel;   // non-empty: `continue` of the for-near body jumps there (unrolled loop)
  int innerLoopDepth = 0;
  std::vector<StepInfo> steps;
  std::set<const FuncDecl *> listSteps;   // steps whose kernels were printed with the list modes
  std::set<const FuncDecl *> denseSteps;  // steps whose kernels were printed with the shadow pre-filter (ABL_MODE 8)

  std::string label() { return "_var" + std::to_string(anon++); }
  bool dev() const { return target == Target::Device; }

  int agentIndex(const AgentDecl *a) const {
    for (size_t i = 0; i < script.agents.size(); i++) if (script.agents[i] == a) return (int)i;
    return -1;
  }
  // first SoA column of member m (float3 members take three columns)
  static int columnOf(const AgentDecl &a, int member) {
    int c = 0;
    for (int i = 0; i < member; i++) c += a.members[i]->type.k == TK::Vec3 ? 3 : 1;
    return c;
  }
  static int columnCount(const AgentDecl &a) { return columnOf(a, (int)a.members.size()); }

  std::string typeName(const Ty &t) const;
  std::string storageName(const Ty &t) const;
  static const char *typeTag(const Ty &t);

  void expr(const Expr &e);
  std::string exprText(const Expr &e) {
    Writer tmp;
    std::swap(w, tmp);
    expr(e);
    std::string text = w.str();
    std::swap(w, tmp);
    return text;
  }
  void args(const Expr &call);
  void vecBinary(Op op, const Expr &l, const Expr &r);
  void callExpr(const Expr &e);
  void stmt(const Stmt &s);
  void stmts(const std::vector<StmtP> &v);
  void forStmt(const Stmt &s);
  void nearLoop(const Stmt &s);
  // what the per-variant emitters of a for-near loop share
  struct NearLoop {
    const Stmt &s;
    const Expr &agentExpr, &radius;
    AgentDecl *nbr;
    AgentMember *pos, *selfPos;
    int dim, posIndex;
    std::string it, sdim, selfPosText;
    std::vector<int> others;      // neighbour members the body reads (besides the position)
    bool prefetchOthers;          // ... fetched one candidate ahead / with the position
    std::string ptypeS;
  };
  void nearLoadOthers(const NearLoop &L, const std::string &idx);
  void nearLoopBody(const NearLoop &L, const std::string &d2, const std::string &breakLabel, const std::string &continueLabel);
  void nearListLoops(const NearLoop &L);
  void nearTileLoop(const NearLoop &L);
  void nearBulkFlatLoop(const NearLoop &L);
  void nearShadowLoop(const NearLoop &L);
  void nearChunkedLoop(const NearLoop &L);
  void nearFlatLoop(const NearLoop &L);
  void nearCursorLoop(const NearLoop &L);
  void constDecl(const ConstDecl &c);
  void function(const FuncDecl &f);
  void agentStruct(const AgentDecl &a);

  void analyseSteps();
  void scanStepExpr(const Expr &e, StepInfo &si, const FuncDecl &f);
  void scanStepStmt(const Stmt &s, StepInfo &si, const FuncDecl &f);
  void reachable(const FuncDecl *f, std::set<const FuncDecl *> &seen);
  void reachableExpr(const Expr &e, std::set<const FuncDecl *> &seen);
  void reachableStmt(const Stmt &s, std::set<const FuncDecl *> &seen);

  void stepKernel(const StepInfo &si, int index);
  struct StepKernelCtx {
    const StepInfo &si;
    const Expr *radius;
    const std::vector<TileCol> &tcols;
    int tdim;
    const std::string &trows;
    bool sql;
  };
  void stepKernelWrapper(const StepKernelCtx &C);
  void stepPrefilterKernel(const StepKernelCtx &C);
  void stepLauncher(const StepKernelCtx &C);
  static const Expr *findNearRadius(const std::vector<StmtP> &body);
  static bool hostEvaluable(const Expr &e);
  void loadMember(const AgentDecl &a, int m, const std::string &dst, const std::string &view,
                  const std::string &idx);
  void storeMember(const AgentDecl &a, int m, const std::string &cols, const std::string &idx,
                   const std::string &value);
  void hostSimulate();
  void hostVisualize();
  bool visualize() const { return config.getBool("visualize", false); }
  void hostSeqSupport();
};

// ---------------------------------------------------------------------------------------
// types
// ---------------------------------------------------------------------------------------
std::string CudaPrinter::typeName(const Ty &t) const {
  switch (t.k) {
    case TK::Void: return "void";
    case TK::Bool: return "bool";
    case TK::Int: return "int";
    case TK::Float: return "abl_real";
    case TK::String: return "const char*";
    case TK::Vec2: return "abl_float2";
    case TK::Vec3: return "abl_float3";
    case TK::Agent: return dev() ? t.agent->name : t.agent->name + "*";
    case TK::Array: return dev() ? storageName(t.elem()) : "abl_array*";
    default: throw BackendError("cuda backend: unsupported type " + t.str());
  }
}

std::string CudaPrinter::storageName(const Ty &t) const {
  if (t.isAgent()) return t.agent->name;
  return typeName(t);
}

const char *CudaPrinter::typeTag(const Ty &t) {
  switch (t.k) {
    case TK::Bool: return "ABL_TYPE_BOOL";
    case TK::Int: return "ABL_TYPE_INT";
    case TK::Float: return "ABL_TYPE_FLOAT";
    case TK::Vec2: return "ABL_TYPE_FLOAT2";
    case TK::Vec3: return "ABL_TYPE_FLOAT3";
    default: throw BackendError("cuda backend: agent members must be bool, int, float, float2 or float3");
  }
}

// ---------------------------------------------------------------------------------------
// expressions
// ---------------------------------------------------------------------------------------
void CudaPrinter::args(const Expr &call) {
  for (size_t i = 0; i < call.kids.size(); i++) {
    if (i) w << ", ";
    expr(*call.kids[i]);
  }
}

void CudaPrinter::vecBinary(Op op, const Expr &l, const Expr &r) {
  const Ty &v = l.type.isVec() ? l.type : r.type;
  w << "float" << v.vecLen() << "_";
  switch (op) {
    case Op::Add: w << "add"; break;
    case Op::Sub: w << "sub"; break;
    case Op::Div: w << "div_scalar"; break;
    case Op::Mul: w << "mul_scalar"; break;
    case Op::Eq: w << "equals"; break;
    case Op::Ne: w << "not_equals"; break;
    default: throw BackendError("cuda backend: unsupported vector operator");
  }
  w << "(";
  expr(l);
  w << ", ";
  expr(r);
  w << ")";
}

void CudaPrinter::expr(const Expr &e) {
  switch (e.kind) {
    case Expr::BoolLit: w << (e.bval ? "true" : "false"); return;
    case Expr::IntLit: w << e.ival; return;
    case Expr::FloatLit:
      if (dev()) w << "ABL_R(" << floatText(e.fval) << ")";
      else w << floatText(e.fval);
      return;
    case Expr::StrLit: w << quoted(e.name); return;
    case Expr::Var: w << e.name; return;

    case Expr::Unary:
      if (e.kids[0]->type.isVec()) {
        if (e.op == Op::Pos) { expr(*e.kids[0]); return; }
        // -v is v * -1.0, exactly like the reference lowering
        w << "float" << e.kids[0]->type.vecLen() << "_mul_scalar(";
        expr(*e.kids[0]);
        w << ", " << (dev() ? "ABL_R(-1.0)" : "-1.0") << ")";
        return;
      }
      w << "(" << opSigil(e.op);
      expr(*e.kids[0]);
      w << ")";
      return;

    case Expr::Binary: {
      const Expr &l = *e.kids[0], &r = *e.kids[1];
      if (l.type.isVec() || r.type.isVec()) { vecBinary(e.op, l, r); return; }
      if (sqCompare(e)) return;
      if (e.op == Op::Mod && !(l.type.isInt() && r.type.isInt())) {
        w << (dev() ? "abl_fmod(" : "fmod(");
        expr(l); w << ", "; expr(r); w << ")";
        return;
      }
      w << "("; expr(l); w << " " << opSigil(e.op) << " "; expr(r); w << ")";
      return;
    }

    case Expr::Ternary:
      w << "("; expr(*e.kids[0]); w << " ? "; expr(*e.kids[1]); w << " : "; expr(*e.kids[2]); w << ")";
      return;

    case Expr::Member:
      expr(*e.kids[0]);
      w << ((e.kids[0]->type.isAgent() && !dev()) ? "->" : ".") << e.name;
      return;

    case Expr::Index:
      if (!dev() && e.kids[0]->type.isArray() && e.kids[0]->kind == Expr::Var &&
          e.kids[0]->sym && !e.kids[0]->sym->global) {
        // local arrays on the host are abl_array handles
        w << "(*ABL_AT("; expr(*e.kids[0]); w << ", " << storageName(e.type) << ", ";
        expr(*e.kids[1]); w << "))";
        return;
      }
      expr(*e.kids[0]); w << "["; expr(*e.kids[1]); w << "]";
      return;

    case Expr::Call: callExpr(e); return;

    case Expr::AgentCreate:
      if (dev()) {
        // only reachable as the argument of add(), handled in callExpr
        throw BackendError("cuda backend: agent creation outside add() is not supported in step functions");
      }
      w << "(" << e.name << ") {";
      w.indent();
      for (size_t i = 0; i < e.kids.size(); i++) {
        w.nl();
        w << "." << e.initNames[i] << " = ";
        expr(*e.kids[i]);
        w << ",";
      }
      w.outdent();
      w.nl();
      w << "}";
      return;

    case Expr::ArrayInit:
      w << "{ ";
      args(e);
      w << " }";
      return;

    case Expr::NewArray:
      if (dev()) throw BackendError("cuda backend: `new` arrays are not supported in step functions");
      w << "abl_array_zeroed(sizeof(" << storageName(e.elemTy) << "), ";
      expr(*e.kids[0]);
      w << ")";
      return;

    case Expr::EnvAccess:
      throw BackendError("cuda backend: unresolved environment access");
  }
}

bool CudaPrinter::isNearPair(const Expr &a, const Expr &b) const {
  if (!nearVar || nearD2.empty()) return false;
  auto is = [](const Expr &e, const Symbol *sym, const std::string &member) {
    return e.kind == Expr::Member && e.name == member && e.kids[0]->kind == Expr::Var && e.kids[0]->sym == sym;
  };
  return (is(a, nearVar, nearPos) && is(b, nearSelf, nearSelfPos)) ||
         (is(b, nearVar, nearPos) && is(a, nearSelf, nearSelfPos));
}

// dist(in.pos, nx.pos) / dist(nx.pos, in.pos) / length(in.pos - nx.pos) of the current near pair
bool CudaPrinter::isNearDistCall(const Expr &e) const {
  if (e.kind != Expr::Call || e.ckind != Expr::Builtin) return false;
  const std::string &t = e.target;
  if ((t == "dist_float2" || t == "dist_float3") && isNearPair(*e.kids[0], *e.kids[1])) return true;
  return (t == "length_float2" || t == "length_float3") && e.kids[0]->kind == Expr::Binary &&
         e.kids[0]->op == Op::Sub && isNearPair(*e.kids[0]->kids[0], *e.kids[0]->kids[1]);
}

// name of the squared-distance variable if `e` is the near pair's distance (the call itself or a
// local variable known to hold it)
const std::string *CudaPrinter::nearDistSquare(const Expr &e) const {
  if (!dev() || !nearVar || nearD2.empty()) return nullptr;
  if (isNearDistCall(e)) return &nearD2;
  if (e.kind == Expr::Var && e.sym) {
    auto it = sqAlias.find(e.sym);
    if (it != sqAlias.end()) return &it->second;
  }
  return nullptr;
}

bool CudaPrinter::assignsTo(const Stmt &s, const Symbol *sym) {
  if (s.kind == Stmt::Assign || s.kind == Stmt::AssignOp) {
    const Expr *root = s.e[0].get();
    while (root->kind == Expr::Member || root->kind == Expr::Index) root = root->kids[0].get();
    if (root->kind == Expr::Var && root->sym == sym) return true;
  }
  for (const StmtP &b : s.body) if (assignsTo(*b, sym)) return true;
  return false;
}

// `<distance> < C`, `C >= <distance>`, ... with a host-evaluable C -> comparison of the squared
// distance with kernel parameter _sql.v[k]
bool CudaPrinter::sqCompare(const Expr &e) {
  if (!dev() || !sqcmpOn() || !curStepHasLimit) return false;
  int op;
  switch (e.op) {
    case Op::Lt: op = 0; break;
    case Op::Le: op = 1; break;
    case Op::Gt: op = 2; break;
    case Op::Ge: op = 3; break;
    default: return false;
  }
  const Expr *d = e.kids[0].get(), *c = e.kids[1].get();
  const std::string *sq = nearDistSquare(*d);
  if (!sq) {
    // constant on the left: C < d  is  d > C
    std::swap(d, c);
    sq = nearDistSquare(*d);
    if (!sq) return false;
    static const int mirrored[4] = {2, 3, 0, 1};
    op = mirrored[op];
  }
  if (!hostEvaluable(*c) || !(c->type.isNum())) return false;
  std::string text = exprText(*c);
  size_t k = 0;
  for (; k < sqLimits.size(); k++) if (sqLimits[k].first == op && sqLimits[k].second == text) break;
  if (k == sqLimits.size()) {
    if (k >= 8) return false;   // ABL_SQ_LIMITS
    sqLimits.push_back({op, text});
  }
  w << "(" << *sq << (op <= 1 ? " <= " : " >= ") << "_sql.v[" << k << "])";
  return true;
}

void CudaPrinter::callExpr(const Expr &e) {
  if (e.ckind == Expr::Ctor) {
    if (e.type.isVec()) {
      w << "float" << e.type.vecLen() << "_" << (e.kids.size() == 1 ? "fill" : "create") << "(";
      args(e);
      w << ")";
    } else {
      w << "(" << typeName(e.type) << ") ";
      expr(*e.kids[0]);
    }
    return;
  }

  const std::string &t = e.target;
  if (e.ckind == Expr::Builtin) {
    if (dev() && (t == "dist_float2" || t == "dist_float3") && isNearPair(*e.kids[0], *e.kids[1])) {
      w << "abl_sqrt_narrow(" << nearD2 << ")";
      return;
    }
    if (dev() && (t == "length_float2" || t == "length_float3") && e.kids[0]->kind == Expr::Binary &&
        e.kids[0]->op == Op::Sub && isNearPair(*e.kids[0]->kids[0], *e.kids[0]->kids[1])) {
      w << "abl_sqrt_narrow(" << nearD2 << ")";
      return;
    }
    if (t == "add") {
      AgentDecl *agent = e.paramTys[0].agent;
      if (!dev()) {
        w << "*(" << agent->name << " *)abl_array_push(&agents_" << agent->name << ", sizeof("
          << agent->name << ")) = ";
        expr(*e.kids[0]);
        return;
      }
      // device: stage the new agent in the per-parent slot; the runtime appends it at commit
      const Expr &c = *e.kids[0];
      w << "{ _ctx.added = true;";
      for (size_t m = 0; m < agent->members.size(); m++) {
        const Expr *init = c.init(agent->members[m]->name);
        w << " ";
        storeMember(*agent, (int)m, "_a.add_cols", "_i", exprText(*init));
      }
      w << " }";
      return;
    }
    if (t == "save") {
      w << "abl_model_save(";
      args(e);
      w << ")";
      return;
    }
    if (t == "removeCurrent") {
      if (!dev()) throw BackendError("cuda backend: removeCurrent() outside a step function");
      w << "_ctx.dead = true";
      return;
    }
    if (t == "count") {
      w << "abl_model_count(" << agentIndex(e.paramTys[0].agent) << ")";
      return;
    }
    if (t == "count_member") {
      const Ty &mt = e.paramTys[0];
      int mi = mt.agent->memberIndex(mt.member->name);
      bool isFloat = mt.member->type.isFloat();
      w << (isFloat ? "abl_model_count_member_float(" : "abl_model_count_member_int(")
        << agentIndex(mt.agent) << ", " << mi << ", ";
      expr(*e.kids[1]);
      w << ")";
      return;
    }
    if (t == "sum") {
      const Ty &mt = e.paramTys[0];
      int mi = mt.agent->memberIndex(mt.member->name);
      const Ty &ty = mt.member->type;
      if (ty.isInt() || ty.isBool()) w << "abl_model_sum_int(" << agentIndex(mt.agent) << ", " << mi << ")";
      else if (ty.isFloat()) w << "abl_model_sum_float(" << agentIndex(mt.agent) << ", " << mi << ")";
      else if (ty.k == TK::Vec2) w << "abl_model_sum_float2(" << agentIndex(mt.agent) << ", " << mi << ")";
      else w << "abl_model_sum_float3(" << agentIndex(mt.agent) << ", " << mi << ")";
      return;
    }
    if (t == "log_csv") {
      w << "{ abl_host_log_open(\"log.csv\");";
      for (size_t i = 0; i < e.kids.size(); i++) {
        w << (e.kids[i]->type.isInt() ? " abl_host_log_int(" : " abl_host_log_float(")
          << (i == 0 ? 1 : 0) << ", ";
        expr(*e.kids[i]);
        w << ");";
      }
      w << " abl_host_log_end(); }";
      return;
    }
    if (t == "getLastExecTime") { w << "abl_model_last_exec_time()"; return; }
    if (t == "near") throw BackendError("cuda backend: near() is only supported as a for-loop range");
    if (t == "random_float" || t == "random_int") {
      w << t << "(";
      if (dev()) w << "_ctx, ";
      args(e);
      w << ")";
      return;
    }
    if (t == "min" || t == "max") {
      w << "abl_" << t << "(";
      args(e);
      w << ")";
      return;
    }
    w << t << "(";
    args(e);
    w << ")";
    return;
  }

  // user function
  w << t << "(";
  if (dev()) { w << "_ctx"; if (!e.kids.empty()) w << ", "; }
  args(e);
  w << ")";
}

// ---------------------------------------------------------------------------------------
// statements
// ---------------------------------------------------------------------------------------
void CudaPrinter::stmts(const std::vector<StmtP> &v) {
  for (const StmtP &s : v) { w.nl(); stmt(*s); }
}

void CudaPrinter::stmt(const Stmt &s) {
  switch (s.kind) {
    case Stmt::ExprS: {
      const Expr &e = *s.e[0];
      expr(e);
      bool braced = e.kind == Expr::Call && e.ckind == Expr::Builtin &&
                    (e.target == "log_csv" || (dev() && e.target == "add"));
      if (!braced) w << ";";
      return;
    }
    case Stmt::Block:
      w << "{";
      w.indent(); stmts(s.body); w.outdent();
      w.nl();
      w << "}";
      return;
    case Stmt::VarDecl: {
      const Ty &t = s.declTy;
      if (!dev() && (t.isAgent() || t.isArray())) {
        // host: agents and arrays live in a storage variable, the name is a pointer to it
        std::string store = label();
        w << (t.isArray() ? "abl_array" : t.agent->name) << " " << store;
        if (!s.e.empty()) {
          w << " = ";
          if (s.e[0]->type.isAgent() && s.e[0]->kind != Expr::AgentCreate) w << "*";
          expr(*s.e[0]);
        }
        w << ";";
        w.nl();
        w << typeName(t) << " " << s.varName << " = &" << store << ";";
        return;
      }
      w << typeName(t) << " " << s.varName;
      if (!s.e.empty()) { w << " = "; expr(*s.e[0]); }
      w << ";";
      if (dev() && sqcmpOn() && curStepHasLimit && !s.e.empty() && t.isFloat() && s.sym && nearBody &&
          innerLoopDepth == 0 && isNearDistCall(*s.e[0]) && !assignsTo(*nearBody, s.sym))
        sqAlias[s.sym] = nearD2;
      return;
    }
    case Stmt::Assign:
      if (s.e[1]->type.isAgent() && !dev()) {
        w << "*"; expr(*s.e[0]); w << " = *"; expr(*s.e[1]); w << ";";
        return;
      }
      expr(*s.e[0]); w << " = "; expr(*s.e[1]); w << ";";
      return;
    case Stmt::AssignOp: {
      const Expr &l = *s.e[0], &r = *s.e[1];
      bool special = l.type.isVec() || r.type.isVec() ||
                     (s.op == Op::Mod && !(l.type.isInt() && r.type.isInt()));
      if (special) {
        expr(l);
        w << " = ";
        if (l.type.isVec() || r.type.isVec()) vecBinary(s.op, l, r);
        else { w << (dev() ? "abl_fmod(" : "fmod("); expr(l); w << ", "; expr(r); w << ")"; }
        w << ";";
        return;
      }
      expr(l); w << " " << opSigil(s.op) << "= "; expr(r); w << ";";
      return;
    }
    case Stmt::If:
      w << "if ("; expr(*s.e[0]); w << ") ";
      stmt(*s.body[0]);
      if (s.body.size() > 1) { w << " else "; stmt(*s.body[1]); }
      return;
    case Stmt::While:
      w << "while ("; expr(*s.e[0]); w << ") ";
      innerLoopDepth++;
      stmt(*s.body[0]);
      innerLoopDepth--;
      return;
    case Stmt::For: forStmt(s); return;
    case Stmt::Return:
      if (s.e.empty()) w << "return;";
      else { w << "return "; expr(*s.e[0]); w << ";"; }
      return;
    case Stmt::Break:
      if (!nearBreakLabel.empty() && innerLoopDepth == 0) w << "goto " << nearBreakLabel << ";";
      else w << "break;";
      return;
    case Stmt::Continue:
      if (!nearContinueLabel.empty() && innerLoopDepth == 0) w << "goto " << nearContinueLabel << ";";
      else w << "continue;";
      return;
    case Stmt::Simulate:
      if (dev()) throw BackendError("cuda backend: simulate inside device code");
      w << "abl_model_simulate("; expr(*s.e[0]); w << ");";
      return;
  }
}

void CudaPrinter::forStmt(const Stmt &s) {
  if (s.forKind == Stmt::ForNear) { nearLoop(s); return; }
  struct DepthGuard { int &d; DepthGuard(int &d) : d(d) { d++; } ~DepthGuard() { d--; } } guard(innerLoopDepth);
  if (s.forKind == Stmt::ForRange) {
    std::string end = label();
    const Expr &range = *s.e[0];
    w << "for (int " << s.varName << " = "; expr(*range.kids[0]);
    w << ", " << end << " = "; expr(*range.kids[1]);
    w << "; " << s.varName << " < " << end << "; ++" << s.varName << ") ";
    stmt(*s.body[0]);
    return;
  }
  // array iteration
  const Expr &arr = *s.e[0];
  std::string idx = label();
  if (dev() || (arr.kind == Expr::Var && arr.sym && arr.sym->global)) {
    // global constant array with a compile-time length
    w << "for (int " << idx << " = 0; " << idx << " < (int)(sizeof("; expr(arr);
    w << ") / sizeof(("; expr(arr); w << ")[0])); " << idx << "++) {";
    w.indent(); w.nl();
    w << typeName(s.declTy) << " " << s.varName << " = "; expr(arr); w << "[" << idx << "];";
    w.nl(); stmt(*s.body[0]);
    w.outdent(); w.nl(); w << "}";
    return;
  }
  std::string handle = label();
  w << "abl_array* " << handle << " = "; expr(arr); w << ";";
  w.nl();
  w << "for (size_t " << idx << " = 0; " << idx << " < " << handle << "->len; " << idx << "++) {";
  w.indent(); w.nl();
  if (s.declTy.isAgent())
    w << s.declTy.agent->name << "* " << s.varName << " = ABL_AT(" << handle << ", "
      << s.declTy.agent->name << ", " << idx << ");";
  else
    w << typeName(s.declTy) << " " << s.varName << " = *ABL_AT(" << handle << ", "
      << storageName(s.declTy) << ", " << idx << ");";
  w.nl(); stmt(*s.body[0]);
  w.outdent(); w.nl(); w << "}";
}

// The DSL loop body for one accepted candidate.  d2: variable holding the squared distance of the
// near pair (dist(in.pos, nx.pos) in the body reuses it; empty: not available); breakLabel /
// continueLabel: where `break` / `continue` of the body must jump when the variant wraps the body
// in loops of its own (empty: the plain statement).
void CudaPrinter::nearLoopBody(const NearLoop &L, const std::string &d2, const std::string &breakLabel,
                           const std::string &continueLabel) {
  const std::string savedBreak = nearBreakLabel, savedContinue = nearContinueLabel;
  const int savedDepth = innerLoopDepth;
  nearBreakLabel = breakLabel;
  nearContinueLabel = continueLabel;
  innerLoopDepth = 0;
  if (!d2.empty()) setNearContext(L.s, L.agentExpr, L.pos->name, L.selfPos->name, d2);
  stmt(*L.s.body[0]);
  clearNearContext();
  nearBreakLabel = savedBreak;
  nearContinueLabel = savedContinue;
  innerLoopDepth = savedDepth;
}

// members the loop body reads are fetched only for accepted candidates
void CudaPrinter::nearLoadOthers(const NearLoop &L, const std::string &idx) {
  for (size_t m = 0; m < L.nbr->members.size(); m++) {
    if ((int)m == L.posIndex || !curFn->nearMembers.count(L.nbr->members[m]->name)) continue;
    w.nl();
    loadMember(*L.nbr, (int)m, L.s.varName + "." + L.nbr->members[m]->name, "_a.nbr.in", idx);
  }
}

// ABL_MODE 4: walk the cached list of accepted candidates (k-major, abl_cuda.h); 5 / 6: the two
// list-building passes — the cursor loop's filter with a counter instead of the body.  Leaves an
// open `else` for the ordinary variants.
void CudaPrinter::nearListLoops(const NearLoop &L) {
  const Stmt &s = L.s;
  const Expr &agentExpr = L.agentExpr, &radius = L.radius;
  AgentDecl *nbr = L.nbr;
  AgentMember *pos = L.pos, *selfPos = L.selfPos;
  const int dim = L.dim, posIndex = L.posIndex;
  const std::string &it = L.it, &sdim = L.sdim, &selfPosText = L.selfPosText;
  auto loadOthers = [&](const std::string &idx) { nearLoadOthers(L, idx); };
  (void)agentExpr; (void)radius; (void)pos; (void)selfPos; (void)dim; (void)posIndex; (void)loadOthers;
  // ABL_MODE 4: walk the cached list of accepted candidates (k-major, abl_cuda.h); 5 / 6: the
  // two list-building passes — the cursor loop's filter with a counter instead of the body
  const bool needPos = curFn->nearMembers.count(pos->name) != 0;
  std::string ptypeL = typeName(pos->type);
  w << "if (ABL_MODE == 4) {";
  w.indent(); w.nl();
  w << "const unsigned " << it << "ln = _a.nlist_cnt[_i];"; w.nl();
  w << "for (unsigned " << it << "lk = 0; " << it << "lk < " << it << "ln; " << it << "lk++) {";
  w.indent(); w.nl();
  w << "const unsigned " << it << "j = __ldg(_a.nlist_idx + (size_t)" << it << "lk * _a.nlist_stride + _i);"; w.nl();
  w << nbr->name << " " << s.varName << ";";
  if (needPos) {
    w.nl();
    loadMember(*nbr, posIndex, s.varName + "." + pos->name, "_a.nbr.in", it + "j");
    w.nl();
    w << "const abl_real " << it << "d2 = abl_sqnorm" << sdim << "(float" << sdim << "_sub(" << s.varName << "."
      << pos->name << ", " << selfPosText << "));";
  }
  loadOthers(it + "j");
  w.nl();
  nearLoopBody(L, needPos ? it + "d2" : std::string(), std::string(), std::string());
  w.outdent(); w.nl();
  w << "}";
  w.outdent(); w.nl();
  w << "} else if (ABL_MODE == 5 || ABL_MODE == 6) {";
  w.indent(); w.nl();
  w << "abl_near_iter<" << sdim << "> " << it << "b;"; w.nl();
  w << it << "b.init" << sdim << "(_a, " << selfPosText << ", true, _near_cull);"; w.nl();
  w << "unsigned " << it << "c = 0;"; w.nl();
  w << "while (" << it << "b.valid()) {";
  w.indent(); w.nl();
  w << "const unsigned " << it << "j = " << it << "b.index();"; w.nl();
  w << it << "b.next();"; w.nl();
  w << ptypeL << " " << it << "q;"; w.nl();
  loadMember(*nbr, posIndex, it + "q", "_a.nbr.in", it + "j"); w.nl();
  w << "if (abl_sqnorm" << sdim << "(float" << sdim << "_sub(" << it << "q, " << selfPosText << ")) > _near_limit) continue;"; w.nl();
  w << "if (ABL_MODE == 6) _a.nlist_idx[(size_t)" << it << "c * _a.nlist_stride + _i] = " << it << "j;"; w.nl();
  w << it << "c++;";
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << "if (ABL_MODE == 5) { _a.nlist_cnt[_i] = " << it << "c; atomicMax(_a.nlist_max, " << it << "c); }";
  w.outdent(); w.nl();
  w << "} else";
  w.nl();
}

// ABL_MODE 8 (abl_device.cuh: shadow pre-filter): the chunked loop with phase 1 in single precision
// on the pool's float4 shadow and phase 2 over per-thread lists expanded from the acceptance masks,
// in rounds the lanes of a warp enter together.  Same accepted candidates in the same order as every
// other variant.  Works for any reach (the iterator is replayed like in ABL_MODE 1).  Leaves its
// `else` open.
void CudaPrinter::nearShadowLoop(const NearLoop &L) {
  const Stmt &s = L.s;
  AgentDecl *nbr = L.nbr;
  AgentMember *pos = L.pos;
  const int dim = L.dim, posIndex = L.posIndex;
  const std::string &it = L.it, &sdim = L.sdim, &selfPosText = L.selfPosText;
  const std::string done = "_near_done" + it + "s";
  const std::string ptype = typeName(pos->type);
  w << "#ifdef ABL_HAVE_BULK_TILE"; w.nl();
  // ABL_MODE 9: masks, row table and header were written to global memory by the step's pre-filter kernel
  // (abl_prefilter_<step>, one launch earlier); an agent whose candidates did not fit them takes the cursor loop
  if (curStepSplit) w << "if (ABL_MODE == 8 || (ABL_MODE == 9 && !(_a.pf_hdr[_i] >> 31))) {";
  else w << "if (ABL_MODE == 8) {";
  w.indent(); w.nl();
  w << "// dynamic shared memory, words: [ABL_SHADOW_WORDS][blockDim.x] acceptance masks | [ABL_SHADOW_ROWS][blockDim.x] first pool"; w.nl();
  w << "// index of every row range the masks cover | [ABL_SHADOW_LIST][blockDim.x] survivors of the current round"; w.nl();
  w << "// (ABL_MODE 9: masks and row table are one column per agent in global memory, only the lists are shared memory)"; w.nl();
  w << "extern __shared__ unsigned _abl_masks[];"; w.nl();
  w << "const unsigned " << it << "str = ABL_MODE == 8 ? blockDim.x : _a.pf_stride;"; w.nl();
  w << "unsigned *const " << it << "mk = ABL_MODE == 8 ? _abl_masks + threadIdx.x : _a.pf_masks + _i;"; w.nl();
  w << "unsigned *const " << it << "rows = ABL_MODE == 8 ? " << it << "mk + ABL_SHADOW_WORDS * blockDim.x : _a.pf_rows + _i;"; w.nl();
  w << "unsigned *const " << it << "list = ABL_MODE == 8 ? " << it << "rows + ABL_SHADOW_ROWS * blockDim.x : _abl_masks + threadIdx.x;"; w.nl();
  w << "const float4 *const " << it << "sh = static_cast<const float4 *>(_a.nbr_shadow);"; w.nl();
  w << "const float " << it << "sx = (float)" << selfPosText << ".x, " << it << "sy = (float)" << selfPosText << ".y"
    << (dim == 3 ? ", " + it + "sz = (float)" + selfPosText + ".z" : std::string()) << ";"; w.nl();
  w << "const float " << it << "limf = ABL_MODE == 8 ? abl_shadow_limit(_near_limit, __ldg(_a.nbr_shadow_max), fmaxf(fabsf(" << it << "sx), "
    << (dim == 3 ? "fmaxf(fabsf(" + it + "sy), fabsf(" + it + "sz))" : "fabsf(" + it + "sy)") << ")) : 0.0f;"; w.nl();
  w << "const unsigned " << it << "wm = __activemask();   // the lanes that run this loop go through its rounds together"; w.nl();
  w << "bool " << it << "stop = false;   // `break` of the loop body"; w.nl();
  w << "bool " << it << "first = true;"; w.nl();
  w << "for (;;) {";
  w.indent(); w.nl();
  // phase 1: masks of up to ABL_SHADOW_WORDS * 32 candidates out of up to ABL_SHADOW_ROWS row ranges
  w << "unsigned " << it << "nw = 0, " << it << "nr = 0, " << it << "prev = 0xffffffffu;"; w.nl();
  w << "unsigned long long " << it << "sbits = 0ull;   // bit w: word w starts a new row range"; w.nl();
  w << "if (ABL_MODE == 9) {";
  w.indent(); w.nl();
  w << "if (" << it << "first && !" << it << "stop) { " << it << "nw = _a.pf_hdr[_i] & 0xffffu; " << it << "sbits = _a.pf_sbits[_i]; }"; w.nl();
  w << it << "first = false;";
  w.outdent(); w.nl();
  w << "} else";
  w.nl();
  w << "while (!" << it << "stop && " << it << ".valid() && " << it << "nw < ABL_SHADOW_WORDS) {";
  w.indent(); w.nl();
  w << "const unsigned " << it << "sb = " << it << ".index();"; w.nl();
  w << "if (" << it << "sb != " << it << "prev + 32u) {";
  w.indent(); w.nl();
  w << "if (" << it << "nr == ABL_SHADOW_ROWS) break;"; w.nl();
  w << it << "rows[" << it << "nr * " << it << "str] = " << it << "sb;"; w.nl();
  w << it << "nr++;"; w.nl();
  w << it << "sbits |= 1ull << " << it << "nw;";
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << it << "prev = " << it << "sb;"; w.nl();
  w << "const unsigned " << it << "sn = min(" << it << ".remaining(), 32u);"; w.nl();
  w << "unsigned " << it << "sm = 0;"; w.nl();
  w << "for (unsigned " << it << "k = 0; " << it << "k < " << it << "sn; " << it << "k++) {";
  w.indent(); w.nl();
  w << "const float4 " << it << "f = __ldg(" << it << "sh + " << it << "sb + " << it << "k);"; w.nl();
  w << "const float " << it << "dx = " << it << "f.x - " << it << "sx, " << it << "dy = " << it << "f.y - " << it << "sy"
    << (dim == 3 ? ", " + it + "dz = " + it + "f.z - " + it + "sz" : std::string()) << ";"; w.nl();
  if (dim == 3)
    w << "const float " << it << "d2f = __fmaf_rn(" << it << "dz, " << it << "dz, __fmaf_rn(" << it << "dy, " << it << "dy, " << it << "dx * " << it << "dx));";
  else
    w << "const float " << it << "d2f = __fmaf_rn(" << it << "dy, " << it << "dy, " << it << "dx * " << it << "dx);";
  w.nl();
  w << "if (!(" << it << "d2f > " << it << "limf)) " << it << "sm |= 1u << " << it << "k;";
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << it << "mk[" << it << "nw * " << it << "str] = " << it << "sm;"; w.nl();
  w << it << "nw++;"; w.nl();
  w << it << ".skip(" << it << "sn);";
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << "if (!__any_sync(" << it << "wm, " << it << "nw != 0u)) break;"; w.nl();
  // phase 2: rounds of up to ABL_SHADOW_LIST survivors, expanded from the masks
  w << "unsigned " << it << "w = 0, " << it << "m = 0, " << it << "b = 0, " << it << "ri = 0;"; w.nl();
  w << "unsigned " << it << "mn = " << it << "nw ? " << it << "mk[0] : 0u;   // mask word w, requested one fetch ahead (global memory in ABL_MODE 9)"; w.nl();
  w << "for (;;) {";
  w.indent(); w.nl();
  w << "unsigned " << it << "cnt = 0;"; w.nl();
  w << "while (" << it << "cnt < ABL_SHADOW_LIST) {";
  w.indent(); w.nl();
  w << "if (" << it << "m == 0u) {";
  w.indent(); w.nl();
  w << "if (" << it << "w == " << it << "nw) break;"; w.nl();
  w << it << "m = " << it << "mn;"; w.nl();
  w << "if ((" << it << "sbits >> " << it << "w) & 1ull) { " << it << "b = " << it << "rows[" << it << "ri * " << it << "str]; " << it << "ri++; } else " << it << "b += 32u;"; w.nl();
  w << it << "w++;"; w.nl();
  w << "if (" << it << "w < " << it << "nw) " << it << "mn = " << it << "mk[" << it << "w * " << it << "str];"; w.nl();
  w << "continue;";
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << it << "list[" << it << "cnt * blockDim.x] = " << it << "b + (__ffs(" << it << "m) - 1);"; w.nl();
  w << it << "cnt++;"; w.nl();
  w << it << "m &= " << it << "m - 1u;";
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << "__syncwarp(" << it << "wm);"; w.nl();
  // the position of the next survivor is requested before the current one is processed
  w << "unsigned " << it << "jn = " << it << "cnt ? " << it << "list[0] : 0u;"; w.nl();
  w << ptype << " " << it << "pn = " << ptype << "();"; w.nl();
  w << "if (" << it << "cnt) {";
  w.indent(); w.nl();
  loadMember(*nbr, posIndex, it + "pn", "_a.nbr.in", it + "jn");
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << "for (unsigned " << it << "q = 0; " << it << "q < " << it << "cnt; " << it << "q++) {";
  w.indent(); w.nl();
  w << "const unsigned " << it << "j = " << it << "jn;"; w.nl();
  w << nbr->name << " " << s.varName << ";"; w.nl();
  w << s.varName << "." << pos->name << " = " << it << "pn;"; w.nl();
  w << "if (" << it << "q + 1u < " << it << "cnt) {";
  w.indent(); w.nl();
  w << it << "jn = " << it << "list[(" << it << "q + 1u) * blockDim.x];"; w.nl();
  loadMember(*nbr, posIndex, it + "pn", "_a.nbr.in", it + "jn");
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << "const abl_real " << it << "d2 = abl_sqnorm" << sdim << "(float" << sdim << "_sub(" << s.varName << "."
    << pos->name << ", " << selfPosText << "));"; w.nl();
  w << "if (" << it << "d2 > _near_limit) continue;";
  nearLoadOthers(L, it + "j");
  w.nl();
  nearLoopBody(L, it + "d2", done, std::string());
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << "if (false) { " << done << ": " << it << "stop = true; " << it << "m = 0u; " << it << "w = " << it << "nw; }   // `break` of the loop body: no further candidates"; w.nl();
  w << "if (!__any_sync(" << it << "wm, " << it << "m != 0u || " << it << "w != " << it << "nw)) break;";
  w.outdent(); w.nl();
  w << "}";
  w.outdent(); w.nl();
  w << "}";
  w.outdent(); w.nl();
  w << "} else"; w.nl();
  w << "#endif"; w.nl();
}

// The pre-filter kernel of ABL_MODE 9: phase 1 of the shadow pre-filter as a kernel of its own.  It needs the
// positions only, so it runs with fewer registers than the step kernel and no shared memory, and leaves per agent
// (one column each, so that a warp reads and writes consecutive words): up to ABL_SHADOW_WORDS acceptance masks,
// the first pool index of up to ABL_SHADOW_ROWS row ranges, the 64-bit map of the words that start a new range,
// and a header (number of words; bit 31: the candidates did not fit).  Measured on circle3d 1 M (ncu): 0.86 ms for
// this kernel + 1.78 ms for the step kernel that walks the masks, against 3.05 ms for ABL_MODE 8; writing survivor
// LISTS from here instead (no expansion in the step kernel) was tried and lost (1.60 + 1.62 ms: the predicated
// appends cost more than the expansion they save).
void CudaPrinter::stepPrefilterKernel(const StepKernelCtx &C) {
  const StepInfo &si = C.si;
  FuncDecl &f = *si.fn;
  AgentDecl &self = *si.self;
  AgentMember *selfPosM = self.position();
  const int dim = C.tdim;
  const std::string sdim = std::to_string(dim);
  const std::string ptype = typeName(selfPosM->type);
  w << "#ifdef ABL_HAVE_BULK_TILE"; w.nl();
  w << "__global__ void __launch_bounds__(256, 4) abl_prefilter_" << f.emitName
    << "(const __grid_constant__ abl_step_launch _a, const abl_real _near_limit, const abl_real _near_cull) {";
  w.indent(); w.nl();
  w << "if (_a.pdl & 2) cudaTriggerProgrammaticLaunchCompletion();"; w.nl();
  w << "cudaGridDependencySynchronize();"; w.nl();
  w << "bool _boundary;"; w.nl();
  w << "unsigned _ob;"; w.nl();
  w << "const unsigned _r = abl_agent_index(_a, _boundary, _ob);"; w.nl();
  w << "if (_r == 0xffffffffu) return;"; w.nl();
  w << "const unsigned _i = _r + _ob;"; w.nl();
  w << ptype << " _p;"; w.nl();
  loadMember(self, self.memberIndex(selfPosM->name), "_p", "_a.self.in", "_i"); w.nl();
  w << "abl_near_iter<" << sdim << "> _it;"; w.nl();
  w << "_it.init" << sdim << "(_a, _p, true, _near_cull);"; w.nl();
  w << "const unsigned _str = _a.pf_stride;"; w.nl();
  w << "unsigned *const _mk = _a.pf_masks + _i, *const _rows = _a.pf_rows + _i;"; w.nl();
  w << "const float4 *const _sh = static_cast<const float4 *>(_a.nbr_shadow);"; w.nl();
  w << "const float _sx = (float)_p.x, _sy = (float)_p.y" << (dim == 3 ? ", _sz = (float)_p.z" : "") << ";"; w.nl();
  w << "const float _limf = abl_shadow_limit(_near_limit, __ldg(_a.nbr_shadow_max), fmaxf(fabsf(_sx), "
    << (dim == 3 ? "fmaxf(fabsf(_sy), fabsf(_sz))" : "fabsf(_sy)") << "));"; w.nl();
  w << "unsigned _nw = 0, _nr = 0, _prev = 0xffffffffu, _over = 0;"; w.nl();
  w << "unsigned long long _sbits = 0ull;"; w.nl();
  w << "while (_it.valid()) {";
  w.indent(); w.nl();
  w << "if (_nw == ABL_SHADOW_WORDS) { _over = 1; break; }"; w.nl();
  w << "const unsigned _sb = _it.index();"; w.nl();
  w << "if (_sb != _prev + 32u) {";
  w.indent(); w.nl();
  w << "if (_nr == ABL_SHADOW_ROWS) { _over = 1; break; }"; w.nl();
  w << "_rows[_nr * _str] = _sb;"; w.nl();
  w << "_nr++;"; w.nl();
  w << "_sbits |= 1ull << _nw;";
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << "_prev = _sb;"; w.nl();
  w << "const unsigned _sn = min(_it.remaining(), 32u);"; w.nl();
  w << "unsigned _sm = 0;"; w.nl();
  w << "for (unsigned _k = 0; _k < _sn; _k++) {";
  w.indent(); w.nl();
  w << "const float4 _f = __ldg(_sh + _sb + _k);"; w.nl();
  w << "const float _dx = _f.x - _sx, _dy = _f.y - _sy" << (dim == 3 ? ", _dz = _f.z - _sz" : "") << ";"; w.nl();
  if (dim == 3) w << "const float _d2f = __fmaf_rn(_dz, _dz, __fmaf_rn(_dy, _dy, _dx * _dx));";
  else w << "const float _d2f = __fmaf_rn(_dy, _dy, _dx * _dx);";
  w.nl();
  w << "if (!(_d2f > _limf)) _sm |= 1u << _k;";
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << "_mk[_nw * _str] = _sm;"; w.nl();
  w << "_nw++;"; w.nl();
  w << "_it.skip(_sn);";
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << "_a.pf_hdr[_i] = _nw | (_over << 31);"; w.nl();
  w << "_a.pf_sbits[_i] = _sbits;";
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << "#endif"; w.nl(); w.nl();
}

// ABL_MODE 7: the flat loop of ABL_MODE 3 over a shared-memory tile whose rows were copied by the
// TMA engine (abl_device.cuh: abl_btile_plan).  Candidate k of the thread is tile entry
// k + (k < T1 ? O0 : k < T2 ? O1 : O2); position from the tile, further members only for accepted
// candidates.  Same candidates in the same order as every other variant.  Leaves its `else` open.
void CudaPrinter::nearBulkFlatLoop(const NearLoop &L) {
  const Stmt &s = L.s;
  AgentDecl *nbr = L.nbr;
  AgentMember *pos = L.pos;
  const int posIndex = L.posIndex;
  const std::string &it = L.it, &sdim = L.sdim, &selfPosText = L.selfPosText;
  std::vector<TileCol> cols = tileColumns(*curFn, *nbr);
  w << "#ifdef ABL_HAVE_BULK_TILE"; w.nl();
  w << "if (ABL_MODE == 7 && _bt.ok) {";
  w.indent(); w.nl();
  w << "extern __shared__ __align__(16) unsigned char _abl_smem[];"; w.nl();
  w << "const unsigned char *const " << it << "tc = _abl_smem + ABL_BTILE_HDR_BYTES;"; w.nl();
  w << "for (unsigned " << it << "k = 0; " << it << "k < _bt.N; " << it << "k += 2) {";
  w.indent(); w.nl();
  w << "const bool " << it << "hB = " << it << "k + 1u < _bt.N;"; w.nl();
  w << "const unsigned " << it << "kB = " << it << "hB ? " << it << "k + 1u : " << it << "k;"; w.nl();
  w << "const unsigned " << it << "eA = " << it << "k + (" << it << "k < _bt.T1 ? _bt.O0 : (" << it << "k < _bt.T2 ? _bt.O1 : _bt.O2));"; w.nl();
  w << "const unsigned " << it << "eB = " << it << "kB + (" << it << "kB < _bt.T1 ? _bt.O0 : (" << it << "kB < _bt.T2 ? _bt.O1 : _bt.O2));"; w.nl();
  for (int h = 0; h < 2; h++) {
    const std::string H = h ? "B" : "A";
    w << "const abl_float2 " << it << "p" << H << " = reinterpret_cast<const abl_float2 *>(" << it << "tc)[" << it << "e" << H << "];"; w.nl();
  }
  for (int h = 0; h < 2; h++) {
    const std::string H = h ? "B" : "A";
    w << "const abl_real " << it << "d2" << H << " = abl_sqnorm" << sdim << "(float" << sdim << "_sub(" << it << "p" << H
      << ", " << selfPosText << "));"; w.nl();
  }
  const std::string second = "_near_bulkB" + it;
  for (int h = 0; h < 2; h++) {
    const std::string H = h ? "B" : "A";
    w << "if (" << (h ? it + "hB && " : std::string()) << "!(" << it << "d2" << H << " > _near_limit)) {";
    w.indent(); w.nl();
    w << nbr->name << " " << s.varName << ";"; w.nl();
    w << s.varName << "." << pos->name << " = " << it << "p" << H << ";";
    for (size_t q = 0; q < cols.size(); q++) {
      if (cols[q].member == posIndex) continue;
      const Ty &mt = nbr->members[cols[q].member]->type;
      const std::string dst = s.varName + "." + nbr->members[cols[q].member]->name;
      const std::string src = "reinterpret_cast<const " + cols[q].ctype + " *>(" + it + "tc + (size_t)_tile_cap * " +
                              tileOffset(cols, q) + ")[" + it + "e" + H + "]";
      w.nl();
      if (mt.k == TK::Vec3) w << dst << "." << "xyz"[cols[q].comp] << " = " << src << ";";
      else if (mt.k == TK::Bool) w << dst << " = " << src << " != 0;";
      else w << dst << " = " << src << ";";
    }
    w.nl();
    nearLoopBody(L, it + "d2" + H, std::string(), h ? std::string() : second);
    w.outdent(); w.nl();
    w << "}"; w.nl();
    if (h == 0) { w << second << ": ;"; w.nl(); }
  }
  w.outdent(); w.nl();
  w << "}";
  w.outdent(); w.nl();
  w << "} else"; w.nl();
  w << "#endif"; w.nl();
}

// ABL_MODE 2: candidates come from the shared-memory tile the kernel prologue staged
// (abl_device.cuh: abl_tile_plan); same visiting order as abl_near_iter.  Leaves its `else` open.
void CudaPrinter::nearTileLoop(const NearLoop &L) {
  const Stmt &s = L.s;
  const Expr &agentExpr = L.agentExpr, &radius = L.radius;
  AgentDecl *nbr = L.nbr;
  AgentMember *pos = L.pos, *selfPos = L.selfPos;
  const int dim = L.dim, posIndex = L.posIndex;
  const std::string &it = L.it, &sdim = L.sdim, &selfPosText = L.selfPosText;
  auto loadOthers = [&](const std::string &idx) { nearLoadOthers(L, idx); };
  (void)agentExpr; (void)radius; (void)pos; (void)selfPos; (void)dim; (void)posIndex; (void)loadOthers;
  // ABL_MODE 2: candidates come from the shared-memory tile the kernel prologue staged
  // (abl_device.cuh: abl_tile_plan); same visiting order as abl_near_iter
  std::vector<TileCol> cols = tileColumns(*curFn, *nbr);
  const std::string rows = dim == 2 ? "3" : "9";
  std::string done = "_near_done" + it + "t";
  w << "if (ABL_MODE == 2 && _tile_ok) {";
  w.indent(); w.nl();
  w << "extern __shared__ __align__(16) unsigned char _abl_smem[];"; w.nl();
  w << "const uint2 *" << it << "seg = reinterpret_cast<const uint2 *>(_abl_smem + ABL_TILE_HDR_BYTES) + threadIdx.x;"; w.nl();
  w << "const unsigned char *const " << it << "cols = _abl_smem + ABL_TILE_HDR_BYTES + " << rows << " * blockDim.x * sizeof(uint2);"; w.nl();
  // Two phases per row and chunk of up to 32 candidates: phase 1 only evaluates the radius
  // filter and collects one acceptance bit per candidate in a register; phase 2 runs the
  // loop body for the set bits, in candidate order.  A warp then pays for the body
  // max-popcount times per chunk instead of once per candidate (in the single-phase loop
  // nearly every iteration has some lane that accepts).
  w << "for (int " << it << "k = 0; " << it << "k < " << rows << "; " << it << "k++) {";
  w.indent(); w.nl();
  w << "const uint2 " << it << "sg = " << it << "seg[" << it << "k * blockDim.x];"; w.nl();
  w << "for (unsigned " << it << "b = " << it << "sg.x, " << it << "e = " << it << "sg.x + " << it << "sg.y; "
    << it << "b < " << it << "e; " << it << "b += 32) {";
  w.indent(); w.nl();
  w << "const unsigned " << it << "n = min(32u, " << it << "e - " << it << "b);"; w.nl();
  w << "unsigned " << it << "m = 0;"; w.nl();
  std::string posSrc = "reinterpret_cast<const " + cols[0].ctype + " *>(" + it + "cols)";
  auto tilePos = [&](const std::string &dst, const std::string &idx) {
    // position columns come first in the tile (offsets 0, 1, 2 scalar columns for float3)
    if (dim == 2) {
      w << dst << " = reinterpret_cast<const abl_float2 *>(" << it << "cols)[" << idx << "];";
    } else {
      for (int c = 0; c < 3; c++) {
        if (c) w.nl();
        w << dst << "." << "xyz"[c] << " = reinterpret_cast<const abl_real *>(" << it << "cols + (size_t)_tile_cap * "
          << tileOffset(cols, (size_t)c) << ")[" << idx << "];";
      }
    }
  };
  (void)posSrc;
  std::string ptypeT = typeName(pos->type);
  w << "for (unsigned " << it << "c = 0; " << it << "c < " << it << "n; " << it << "c++) {";
  w.indent(); w.nl();
  w << ptypeT << " " << it << "q;"; w.nl();
  tilePos(it + "q", it + "b + " + it + "c");
  w.nl();
  if (curStepHasLimit) {
    w << "if (!(abl_sqnorm" << sdim << "(float" << sdim << "_sub(" << it << "q, " << selfPosText
      << ")) > _near_limit)) " << it << "m |= 1u << " << it << "c;";
  } else {
    w << "if (!(dist_float" << sdim << "(" << it << "q, " << selfPosText << ") > ";
    expr(radius);
    w << ")) " << it << "m |= 1u << " << it << "c;";
  }
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << "while (" << it << "m) {";
  w.indent(); w.nl();
  w << "const unsigned " << it << "s = " << it << "b + (__ffs(" << it << "m) - 1);"; w.nl();
  w << it << "m &= " << it << "m - 1;"; w.nl();
  w << nbr->name << " " << s.varName << ";";
  auto tileLoad = [&](int member) {
    const Ty &mt = nbr->members[member]->type;
    std::string dst = s.varName + "." + nbr->members[member]->name;
    for (size_t q = 0; q < cols.size(); q++) {
      if (cols[q].member != member) continue;
      w.nl();
      std::string src = "reinterpret_cast<const " + cols[q].ctype + " *>(" + it + "cols + (size_t)_tile_cap * " +
                        tileOffset(cols, q) + ")[" + it + "s]";
      if (mt.k == TK::Vec3) w << dst << "." << "xyz"[cols[q].comp] << " = " << src << ";";
      else if (mt.k == TK::Bool) w << dst << " = " << src << " != 0;";
      else w << dst << " = " << src << ";";
    }
  };
  tileLoad(posIndex);
  if (curStepHasLimit) {
    w.nl();
    w << "const abl_real " << it << "d2 = abl_sqnorm" << sdim << "(float" << sdim << "_sub(" << s.varName << "."
      << pos->name << ", " << selfPosText << "));";
  }
  for (size_t m = 0; m < nbr->members.size(); m++) {
    if ((int)m == posIndex || !curFn->nearMembers.count(nbr->members[m]->name)) continue;
    tileLoad((int)m);
  }
  w.nl();
  nearLoopBody(L, curStepHasLimit ? it + "d2" : std::string(), done, nearContinueLabel);
  w.outdent(); w.nl();
  w << "}";
  w.outdent(); w.nl();
  w << "}";
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << done << ": ;";
  w.outdent(); w.nl();
  w << "} else {";
  w.indent(); w.nl();
}

// ABL_MODE 1: chunked two-phase loop for dense neighbourhoods.  Leaves its `else` open.
void CudaPrinter::nearChunkedLoop(const NearLoop &L) {
  const Stmt &s = L.s;
  const Expr &agentExpr = L.agentExpr, &radius = L.radius;
  AgentDecl *nbr = L.nbr;
  AgentMember *pos = L.pos, *selfPos = L.selfPos;
  const int dim = L.dim, posIndex = L.posIndex;
  const std::string &it = L.it, &sdim = L.sdim, &selfPosText = L.selfPosText;
  auto loadOthers = [&](const std::string &idx) { nearLoadOthers(L, idx); };
  (void)agentExpr; (void)radius; (void)pos; (void)selfPos; (void)dim; (void)posIndex; (void)loadOthers;
  // Dense populations (ABL_MODE 1, chosen by the launcher from the mean cell occupancy):
  // two phases per chunk of up to 32 candidates.  Phase 1 only evaluates the filter and
  // records a bit per accepted candidate; phase 2 runs the loop body for the set bits, in
  // order.  In a warp the expensive body then executes max-popcount times per chunk instead
  // of once per candidate (with the plain loop nearly every iteration has *some* lane that
  // accepts, so the whole warp pays for the body every time).
  std::string done = "_near_done" + it;
  std::string ptype = typeName(pos->type);
  w << "if (ABL_MODE == 1) {";
  w.indent(); w.nl();
  w << "// phase 1 writes one acceptance bit per candidate into shared memory (32 candidates"; w.nl();
  w << "// per word, ABL_MASK_WORDS words per thread and round); phase 2 replays the same"; w.nl();
  w << "// chunks and runs the loop body for the set bits only, in candidate order"; w.nl();
  w << "extern __shared__ unsigned _abl_masks[];"; w.nl();
  w << "unsigned " << it << "nw, " << it << "w, " << it << "m, " << it << "b;"; w.nl();
  w << "for (;;) {";
  w.indent(); w.nl();
  w << "abl_near_iter<" << sdim << "> " << it << "s = " << it << ";"; w.nl();
  w << it << "nw = 0;"; w.nl();
  w << "while (" << it << "s.valid() && " << it << "nw < ABL_MASK_WORDS) {";
  w.indent(); w.nl();
  w << "const unsigned " << it << "sb = " << it << "s.index();"; w.nl();
  w << "const unsigned " << it << "sn = min(" << it << "s.remaining(), 32u);"; w.nl();
  w << "unsigned " << it << "sm = 0;"; w.nl();
  w << "for (unsigned " << it << "k = 0; " << it << "k < " << it << "sn; " << it << "k++) {";
  w.indent(); w.nl();
  w << ptype << " " << it << "q;"; w.nl();
  loadMember(*nbr, posIndex, it + "q", "_a.nbr.in", it + "sb + " + it + "k");
  w.nl();
  w << "if (!(abl_sqnorm" << sdim << "(float" << sdim << "_sub(" << it << "q, " << selfPosText
    << ")) > _near_limit)) " << it << "sm |= 1u << " << it << "k;";
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << "_abl_masks[" << it << "nw * blockDim.x + threadIdx.x] = " << it << "sm;"; w.nl();
  w << it << "nw++;"; w.nl();
  w << it << "s.skip(" << it << "sn);";
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << "if (" << it << "nw == 0) break;"; w.nl();
  w << it << "w = 0; " << it << "m = 0; " << it << "b = 0;"; w.nl();
  w << "for (;;) {";
  w.indent(); w.nl();
  w << "while (" << it << "m == 0 && " << it << "w < " << it << "nw) {";
  w.indent(); w.nl();
  w << it << "b = " << it << ".index();"; w.nl();
  w << it << "m = _abl_masks[" << it << "w * blockDim.x + threadIdx.x];"; w.nl();
  w << it << "w++;"; w.nl();
  w << it << ".skip(min(" << it << ".remaining(), 32u));";
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << "if (" << it << "m == 0) break;"; w.nl();
  w << "const unsigned " << it << "j = " << it << "b + (__ffs(" << it << "m) - 1);"; w.nl();
  w << it << "m &= " << it << "m - 1;"; w.nl();
  w << nbr->name << " " << s.varName << ";"; w.nl();
  loadMember(*nbr, posIndex, s.varName + "." + pos->name, "_a.nbr.in", it + "j");
  w.nl();
  w << "const abl_real " << it << "d2 = abl_sqnorm" << sdim << "(float" << sdim << "_sub(" << s.varName << "."
    << pos->name << ", " << selfPosText << "));";
  loadOthers(it + "j");
  w.nl();
  nearLoopBody(L, it + "d2", done, nearContinueLabel);
  w.outdent(); w.nl();
  w << "}";
  w.outdent(); w.nl();
  w << "}"; w.nl();
  w << done << ": ;";
  w.outdent(); w.nl();
  w << "} else {";
  w.indent(); w.nl();
}

// `-C cuda.flat=true` (ABL_MODE 3, 2-D grids, reach 1): ONE flat candidate counter over the
// three row ranges instead of a cursor that switches rows.  The pool index of candidate k is
// k plus a per-row offset picked with two compares and two selects, so no lane ever leaves the
// loop body to open its next row — in the cursor loop nearly every iteration has *some* lane
// doing that, and the whole warp pays for the divergent path.  Two candidates are handled per
// iteration: both positions (and prefetched members) are requested up front and both filters
// evaluated back to back (independent dependency chains), then the bodies run in candidate
// order.  Same candidates, same order as abl_near_iter: bit-identical results.
// Leaves its `else` open.
void CudaPrinter::nearFlatLoop(const NearLoop &L) {
  const Stmt &s = L.s;
  const Expr &agentExpr = L.agentExpr, &radius = L.radius;
  AgentDecl *nbr = L.nbr;
  AgentMember *pos = L.pos, *selfPos = L.selfPos;
  const int dim = L.dim, posIndex = L.posIndex;
  const std::string &it = L.it, &sdim = L.sdim, &selfPosText = L.selfPosText;
  auto loadOthers = [&](const std::string &idx) { nearLoadOthers(L, idx); };
  (void)agentExpr; (void)radius; (void)pos; (void)selfPos; (void)dim; (void)posIndex; (void)loadOthers;
  const std::vector<int> &others = L.others;
  const bool prefetchOthers = L.prefetchOthers;
  const std::string &ptypeS = L.ptypeS;
  (void)others; (void)prefetchOthers; (void)ptypeS;
  w << "if (ABL_MODE == 3 || ABL_MODE == 7) {";   // (7: a thread or block whose rows are not staged)
  w.indent(); w.nl();
  w << "const unsigned " << it << "T1 = " << it << ".re[0] - " << it << ".rb[0];"; w.nl();
  w << "const unsigned " << it << "T2 = " << it << "T1 + (" << it << ".re[1] - " << it << ".rb[1]);"; w.nl();
  w << "const unsigned " << it << "N = " << it << "T2 + (" << it << ".re[2] - " << it << ".rb[2]);"; w.nl();
  w << "const unsigned " << it << "O0 = " << it << ".rb[0], " << it << "O1 = " << it << ".rb[1] - " << it << "T1, "
    << it << "O2 = " << it << ".rb[2] - " << it << "T2;"; w.nl();
  w << "for (unsigned " << it << "k = 0; " << it << "k < " << it << "N; " << it << "k += 2) {";
  w.indent(); w.nl();
  w << "const bool " << it << "hB = " << it << "k + 1u < " << it << "N;"; w.nl();
  w << "const unsigned " << it << "kB = " << it << "hB ? " << it << "k + 1u : " << it << "k;"; w.nl();
  w << "const unsigned " << it << "jA = " << it << "k + (" << it << "k < " << it << "T1 ? " << it << "O0 : (" << it << "k < "
    << it << "T2 ? " << it << "O1 : " << it << "O2));"; w.nl();
  w << "const unsigned " << it << "jB = " << it << "kB + (" << it << "kB < " << it << "T1 ? " << it << "O0 : (" << it << "kB < "
    << it << "T2 ? " << it << "O1 : " << it << "O2));"; w.nl();
  for (int h = 0; h < 2; h++) {
    const std::string H = h ? "B" : "A";
    w << ptypeS << " " << it << "p" << H << ";"; w.nl();
    loadMember(*nbr, posIndex, it + "p" + H, "_a.nbr.in", it + "j" + H); w.nl();
    if (prefetchOthers)
      for (int m : others) {
        w << typeName(nbr->members[m]->type) << " " << it << "m" << m << H << ";"; w.nl();
        loadMember(*nbr, m, it + "m" + std::to_string(m) + H, "_a.nbr.in", it + "j" + H); w.nl();
      }
  }
  for (int h = 0; h < 2; h++) {
    const std::string H = h ? "B" : "A";
    w << "const abl_real " << it << "d2" << H << " = abl_sqnorm" << sdim << "(float" << sdim << "_sub(" << it << "p" << H
      << ", " << selfPosText << "));"; w.nl();
  }
  const std::string second = "_near_flatB" + it;
  for (int h = 0; h < 2; h++) {
    const std::string H = h ? "B" : "A";
    w << "if (" << (h ? it + "hB && " : std::string()) << "!(" << it << "d2" << H << " > _near_limit)) {";
    w.indent(); w.nl();
    w << nbr->name << " " << s.varName << ";"; w.nl();
    w << s.varName << "." << pos->name << " = " << it << "p" << H << ";";
    if (prefetchOthers) {
      for (int m : others) { w.nl(); w << s.varName << "." << nbr->members[m]->name << " = " << it << "m" << m << H << ";"; }
    } else {
      loadOthers(it + "j" + H);
    }
    w.nl();
    nearLoopBody(L, it + "d2" + H, std::string(), h ? std::string() : second);
    w.outdent(); w.nl();
    w << "}"; w.nl();
    if (h == 0) { w << second << ": ;"; w.nl(); }
  }
  w.outdent(); w.nl();
  w << "}";
  w.outdent(); w.nl();
  w << "} else {";
  w.indent(); w.nl();
}

// ABL_MODE 0: the cursor loop (abl_near_iter), software-pipelined, optionally unrolled by two.
void CudaPrinter::nearCursorLoop(const NearLoop &L) {
  const Stmt &s = L.s;
  const Expr &agentExpr = L.agentExpr, &radius = L.radius;
  AgentDecl *nbr = L.nbr;
  AgentMember *pos = L.pos, *selfPos = L.selfPos;
  const int dim = L.dim, posIndex = L.posIndex;
  const std::string &it = L.it, &sdim = L.sdim, &selfPosText = L.selfPosText;
  auto loadOthers = [&](const std::string &idx) { nearLoadOthers(L, idx); };
  (void)agentExpr; (void)radius; (void)pos; (void)selfPos; (void)dim; (void)posIndex; (void)loadOthers;
  const std::vector<int> &others = L.others;
  const bool prefetchOthers = L.prefetchOthers;
  const std::string &ptypeS = L.ptypeS;
  (void)others; (void)prefetchOthers; (void)ptypeS;
  // `-C cuda.unroll=true`: the loop is unrolled by two with alternating prefetch registers, so
  // that handing the prefetched candidate to the body costs no register copies (8 moves per
  // iteration for a double-precision float2 position + float2 member)
  const bool unroll = config.getBool("cuda.unroll", false);
  auto setName = [&](int set, const std::string &base) { return unroll ? base + (set ? "B" : "A") : base; };
  auto prefetch = [&](int set) {
    w << "if (" << it << ".valid()) {";
    w.indent(); w.nl();
    loadMember(*nbr, posIndex, setName(set, it + "p"), "_a.nbr.in", it + ".index()");
    if (prefetchOthers) {
      for (int m : others) {
        w.nl();
        loadMember(*nbr, m, setName(set, it + "m" + std::to_string(m)), "_a.nbr.in", it + ".index()");
      }
    }
    w.outdent(); w.nl();
    w << "}";
  };
  for (int set = 0; set < (unroll ? 2 : 1); set++) {
    w << ptypeS << " " << setName(set, it + "p") << ";"; w.nl();
    if (prefetchOthers)
      for (int m : others) { w << typeName(nbr->members[m]->type) << " " << setName(set, it + "m" + std::to_string(m)) << ";"; w.nl(); }
  }
  prefetch(0);
  w.nl();
  // one candidate: take it from register set `set`, request the next one into the other set
  auto half = [&](int set, const std::string &skip) {
    w << "const unsigned " << it << "j = " << it << ".index();";
    w.nl();
    w << nbr->name << " " << s.varName << ";";
    w.nl();
    w << s.varName << "." << pos->name << " = " << setName(set, it + "p") << ";"; w.nl();
    if (prefetchOthers)
      for (int m : others) { w << s.varName << "." << nbr->members[m]->name << " = " << setName(set, it + "m" + std::to_string(m)) << ";"; w.nl(); }
    w << it << ".next();"; w.nl();
    prefetch(unroll ? 1 - set : 0);
    w.nl();
    if (curStepHasLimit) {
      w << "const abl_real " << it << "d2 = abl_sqnorm" << sdim << "(float" << sdim << "_sub(" << s.varName << "."
        << pos->name << ", " << selfPosText << "));";
      w.nl();
      w << "if (" << it << "d2 > _near_limit) " << skip;
    } else {
      w << "if (dist_float" << sdim << "(" << s.varName << "." << pos->name << ", " << selfPosText << ") > ";
      expr(radius);
      w << ") " << skip;
    }
    if (!prefetchOthers) loadOthers(it + "j");
    w.nl();
    nearLoopBody(L, curStepHasLimit ? it + "d2" : std::string(), std::string(), nearContinueLabel);
  };
  if (!unroll) {
    w << "while (" << it << ".valid()) {";
    w.indent(); w.nl();
    half(0, "continue;");
    w.outdent(); w.nl();
    w << "}";
  } else {
    const std::string second = "_near_second" + it;
    w << "while (" << it << ".valid()) {";
    w.indent(); w.nl();
    w << "{";
    w.indent(); w.nl();
    std::string savedContinue = nearContinueLabel;
    nearContinueLabel = second;
    half(0, "goto " + second + ";");
    nearContinueLabel = savedContinue;
    w.outdent(); w.nl();
    w << "}"; w.nl();
    w << second << ": ;"; w.nl();
    w << "if (!" << it << ".valid()) break;"; w.nl();
    half(1, "continue;");
    w.outdent(); w.nl();
    w << "}";
  }
}

// for (T nx : near(agent, radius)) body   — device only
void CudaPrinter::nearLoop(const Stmt &s) {
  if (!dev() || !curStep)
    throw BackendError("cuda backend: for-near loops are only supported directly inside step functions");
  const Expr &call = *s.e[0];
  const Expr &agentExpr = *call.kids[0];
  const Expr &radius = *call.kids[1];
  AgentDecl *nbr = s.declTy.agent;
  AgentMember *pos = nbr->position();
  AgentDecl *selfTy = agentExpr.type.agent;
  AgentMember *selfPos = selfTy ? selfTy->position() : nullptr;
  if (!selfPos) throw BackendError("cuda backend: near() needs an agent with a position");
  int dim = pos->type.vecLen();
  std::string it = label();
  std::string sdim = std::to_string(dim);

  int posIndex = nbr->memberIndex(pos->name);
  std::string selfPosText = exprText(agentExpr) + "." + selfPos->name;

  // Variants are template instances of one kernel (`if (ABL_MODE == k)`), printed in this order:
  // list walk / list building (4, 5, 6), shared-memory tile (2), chunked (1), flat (3), cursor (0);
  // every emitter but the last leaves an `else` open, closed at the end of this function.
  w << "{";
  w.indent(); w.nl();
  NearLoop L{s, agentExpr, radius, nbr, pos, selfPos, dim, posIndex, it, sdim, selfPosText, {}, false, std::string()};
  if (curStepList) nearListLoops(L);
  w << "{";
  w.indent(); w.nl();
  const bool tile = curStepTile;
  if (curStepBulk) nearBulkFlatLoop(L);
  if (tile) nearTileLoop(L);
  w << "abl_near_iter<" << sdim << "> " << it << ";";
  w.nl();
  if (curStepFlat) w << "if (ABL_MODE == 3 || ABL_MODE == 7) " << it << ".rows" << sdim << "(_a, " << selfPosText << ", true, _near_cull); else ";
  w << it << ".init" << sdim << "(_a, " << selfPosText << ", true, _near_cull);";
  w.nl();
  // Radius filter: inclusive radius, self included, same operand order as the reference's
  // filter (CPrinter.cpp:166-169): dist(nx.pos, in.pos) > radius -> skip.  When the radius
  // is a host-evaluable constant the launcher precomputes the equivalent bound on the
  // squared distance (abl_near_sq_limit) and the kernel skips the square root.
  if (curStepDense) nearShadowLoop(L);
  if (curStepHasLimit) nearChunkedLoop(L);
  // Software-pipelined candidate loop: the position of the *next* candidate is requested
  // before the current one is tested and processed, so the load latency (L1 miss -> L2) is
  // overlapped with the body instead of stalling every iteration.  The iterator is advanced
  // at the top, which also makes `continue` in the body do the right thing.
  // Besides the position, up to two further neighbour members are fetched one candidate
  // ahead as well (speculatively: a rejected candidate wastes the load, an accepted one no
  // longer waits for a dependent L2 round trip).
  std::vector<int> &others = L.others;
  int otherCols = 0;
  for (size_t m = 0; m < nbr->members.size(); m++) {
    if ((int)m == posIndex || !curFn->nearMembers.count(nbr->members[m]->name)) continue;
    others.push_back((int)m);
    otherCols += nbr->members[m]->type.k == TK::Vec3 ? 3 : 1;
  }
  // (only floating-point members: integer/bool flags usually guard cheap bodies, where the
  // extra load per rejected candidate costs more than the stall it hides — measured on
  // game_of_life)
  bool allFloat = true;
  for (int m : others) {
    const Ty &mt = nbr->members[m]->type;
    allFloat = allFloat && (mt.isFloat() || mt.isVec());
  }
  const bool prefetchOthers = !others.empty() && otherCols <= 2 && allFloat;
  std::string ptypeS = typeName(pos->type);
  L.prefetchOthers = prefetchOthers;
  L.ptypeS = ptypeS;
  if (curStepFlat) nearFlatLoop(L);
  nearCursorLoop(L);
  if (curStepFlat) { w.outdent(); w.nl(); w << "}"; }
  if (curStepHasLimit) { w.outdent(); w.nl(); w << "}"; }
  if (tile) { w.outdent(); w.nl(); w << "}"; }
  w.outdent(); w.nl();
  w << "}";
  w.outdent(); w.nl();
  w << "}";
}

std::vector<CudaPrinter::TileCol> CudaPrinter::tileColumns(const FuncDecl &f, const AgentDecl &nbr) const {
  std::vector<TileCol> cols;
  AgentMember *pos = nbr.position();
  auto add = [&](int m) {
    const Ty &t = nbr.members[m]->type;
    int c = columnOf(nbr, m);
    switch (t.k) {
      case TK::Vec2: cols.push_back({m, 0, c, "abl_float2", "sizeof(abl_float2)"}); break;
      case TK::Vec3:
        for (int k = 0; k < 3; k++) cols.push_back({m, k, c + k, "abl_real", "sizeof(abl_real)"});
        break;
      case TK::Float: cols.push_back({m, 0, c, "abl_real", "sizeof(abl_real)"}); break;
      case TK::Int: cols.push_back({m, 0, c, "int", "sizeof(int)"}); break;
      case TK::Bool: cols.push_back({m, 0, c, "unsigned char", "1"}); break;
      default: throw BackendError("cuda backend: member type " + t.str() + " cannot be staged");
    }
  };
  int posIndex = nbr.memberIndex(pos->name);
  add(posIndex);
  for (size_t m = 0; m < nbr.members.size(); m++) {
    if ((int)m == posIndex || !f.nearMembers.count(nbr.members[m]->name)) continue;
    add((int)m);
  }
  return cols;
}

const Stmt *CudaPrinter::findNearStmt(const std::vector<StmtP> &body) const {
  for (const StmtP &s : body) {
    if (s->kind == Stmt::For && s->forKind == Stmt::ForNear) return s.get();
    if (const Stmt *r = findNearStmt(s->body)) return r;
  }
  return nullptr;
}

// ---------------------------------------------------------------------------------------
// declarations
// ---------------------------------------------------------------------------------------
// Element type of a global constant array.  The reference prints the DSL's base type name for
// arrays (CPrinter::printType, CPrinter.cpp:20-24: `float FORCE[] = {...}`), so a table of
// `float` is stored in SINGLE precision even when abl_float is double; elements widen when they
// are read.  Kept, because it defines the values the reference `c` backend computes with.
std::string CudaPrinter::arrayElemName(const Ty &base) const {
  if (base.isFloat()) return "float";
  return typeName(base);
}

void CudaPrinter::constDecl(const ConstDecl &c) {
  const Ty &t = c.type;
  if (!dev()) {
    // same text as the reference: `double W = 141.421;`, vectors as brace initialisers
    Ty base = c.isArray ? t.elem() : t;
    w << (c.isArray ? arrayElemName(base) : typeName(base)) << " " << c.name << (c.isArray ? "[]" : "") << " = ";
    if (base.isVec() && !c.isArray) {
      w << "{";
      args(*c.init);
      w << "};";
    } else {
      expr(*c.init);
      w << ";";
    }
    return;
  }
  if (c.isArray) {
    Ty base = t.elem();
    w << "__device__ const " << arrayElemName(base) << " " << c.name << "[] = ";
    if (base.isVec()) throw BackendError("cuda backend: vector constant arrays are not supported");
    expr(*c.init);
    w << ";";
    return;
  }
  if (t.isVec()) {
    w << "__device__ const " << typeName(t) << " " << c.name << " = {";
    args(*c.init);
    w << "};";
    return;
  }
  if (t.isString()) return;
  w << "static constexpr " << typeName(t) << " " << c.name << " = ";
  expr(*c.init);
  w << ";";
}

void CudaPrinter::function(const FuncDecl &f) {
  curFn = &f;
  if (dev()) {
    w << "__device__ " << typeName(f.retTy) << " " << f.emitName << "(abl_ctx& _ctx";
    for (const Param &p : f.params) {
      w << ", ";
      if (p.type.isAgent()) w << "const " << p.type.agent->name << "& " << p.name;
      else w << typeName(p.type) << " " << p.name;
    }
    w << ") {";
  } else {
    w << typeName(f.retTy) << " " << f.emitName << "(";
    for (size_t i = 0; i < f.params.size(); i++) {
      if (i) w << ", ";
      w << typeName(f.params[i].type) << " " << f.params[i].name;
    }
    if (f.params.empty()) w << "void";
    w << ") {";
  }
  w.indent(); stmts(f.body); w.outdent();
  w.nl();
  w << "}";
  curFn = nullptr;
}

void CudaPrinter::agentStruct(const AgentDecl &a) {
  w << "typedef struct {";
  w.indent();
  for (auto &m : a.members) { w.nl(); w << typeName(m->type) << " " << m->name << ";"; }
  w.outdent();
  w.nl();
  w << "} " << a.name << ";";
}

// ---------------------------------------------------------------------------------------
// step analysis: which members of `in` are read / of `out` are written
// ---------------------------------------------------------------------------------------
void CudaPrinter::scanStepExpr(const Expr &e, StepInfo &si, const FuncDecl &f) {
  const Param &p = f.params[0];
  if (e.kind == Expr::Member && e.kids[0]->kind == Expr::Var) {
    const Expr &base = *e.kids[0];
    if (base.sym && base.sym == p.sym) { si.reads.insert(e.name); return; }
    if (base.sym && base.sym == p.outSym) { si.reads.insert(e.name); return; }  // reads copy-through value
  }
  if (e.kind == Expr::Var && e.sym && (e.sym == p.sym)) si.wholeIn = true;
  if (e.kind == Expr::Var && e.sym && (e.sym == p.outSym)) { si.wholeOut = true; si.wholeIn = true; }
  if (e.kind == Expr::Call && e.ckind == Expr::Builtin && e.target == "near") {
    // near(in, r): only the position is needed, not the whole record
    const Expr &a = *e.kids[0];
    if (a.kind == Expr::Var && a.sym == p.sym) {
      if (AgentMember *pm = p.type.agent->position()) si.reads.insert(pm->name);
      scanStepExpr(*e.kids[1], si, f);
      return;
    }
  }
  for (const ExprP &k : e.kids) scanStepExpr(*k, si, f);
}

static const Expr *outMemberRoot(const Expr &lhs, const Param &p) {
  // out.m, out.m.x -> the Member node directly on `out`
  const Expr *e = &lhs;
  while (e->kind == Expr::Member || e->kind == Expr::Index) {
    const Expr &base = *e->kids[0];
    if (e->kind == Expr::Member && base.kind == Expr::Var && base.sym && base.sym == p.outSym) return e;
    e = &base;
  }
  return nullptr;
}

void CudaPrinter::scanStepStmt(const Stmt &s, StepInfo &si, const FuncDecl &f) {
  const Param &p = f.params[0];
  if (s.kind == Stmt::Assign || s.kind == Stmt::AssignOp) {
    const Expr *root = outMemberRoot(*s.e[0], p);
    if (root) {
      bool identity = false;
      if (s.kind == Stmt::Assign && root == s.e[0].get()) {
        const Expr &r = *s.e[1];
        identity = r.kind == Expr::Member && r.name == root->name && r.kids[0]->kind == Expr::Var &&
                   r.kids[0]->sym == p.sym;
      }
      if (!identity) si.writes.insert(root->name);
      if (s.kind == Stmt::AssignOp || root != s.e[0].get()) si.reads.insert(root->name);
      scanStepExpr(*s.e[1], si, f);
      return;
    }
    if (s.e[0]->kind == Expr::Var && s.e[0]->sym == p.outSym) {
      si.wholeOut = true;
      scanStepExpr(*s.e[1], si, f);
      return;
    }
  }
  for (const ExprP &e : s.e) scanStepExpr(*e, si, f);
  for (const StmtP &b : s.body) scanStepStmt(*b, si, f);
}

void CudaPrinter::analyseSteps() {
  steps.clear();
  if (!script.simulate) return;
  for (FuncDecl *f : script.simulate->steps) {
    StepInfo si;
    si.fn = f;
    si.self = f->stepAgent();
    for (const StmtP &s : f->body) scanStepStmt(*s, si, *f);
    if (si.wholeIn) for (auto &m : si.self->members) si.reads.insert(m->name);
    if (si.wholeOut) for (auto &m : si.self->members) si.writes.insert(m->name);
    f->writtenMembers = si.writes;
    steps.push_back(si);
  }
}

void CudaPrinter::reachableExpr(const Expr &e, std::set<const FuncDecl *> &seen) {
  if (e.kind == Expr::Call && e.ckind == Expr::User && e.callee) reachable(e.callee, seen);
  for (const ExprP &k : e.kids) reachableExpr(*k, seen);
}
void CudaPrinter::reachableStmt(const Stmt &s, std::set<const FuncDecl *> &seen) {
  for (const ExprP &e : s.e) reachableExpr(*e, seen);
  for (const StmtP &b : s.body) reachableStmt(*b, seen);
}
void CudaPrinter::reachable(const FuncDecl *f, std::set<const FuncDecl *> &seen) {
  if (!seen.insert(f).second) return;
  for (const StmtP &s : f->body) reachableStmt(*s, seen);
}

// ---------------------------------------------------------------------------------------
// host program
// ---------------------------------------------------------------------------------------
std::string CudaPrinter::hostSource() {
  target = Target::Host;
  w = Writer();
  anon = 0;
  analyseSteps();
  w << "/* Generated by OpenABL (cuda backend, sm_100a). Host program. */";
  w.nl();
  w << "#include <stdio.h>"; w.nl();
  w << "#include \"abl_host.h\""; w.nl(); w.nl();

  for (AgentDecl *a : script.agents) {
    agentStruct(*a);
    w.nl();
    w << "static const abl_member_desc " << a->name << "_members[] = {";
    w.indent();
    for (auto &m : a->members) {
      w.nl();
      w << "{ " << typeTag(m->type) << ", offsetof(" << a->name << ", " << m->name << "), \""
        << m->name << "\", " << (m->isPosition ? 1 : 0) << " },";
    }
    w.nl();
    w << "{ ABL_TYPE_END, 0, NULL, 0 }";
    w.outdent(); w.nl();
    w << "};"; w.nl();
    w << "abl_array agents_" << a->name << ";"; w.nl(); w.nl();
  }
  w << "abl_host_type abl_model_types[] = {";
  w.indent();
  for (AgentDecl *a : script.agents) {
    w.nl();
    w << "{ { \"" << a->name << "\", " << a->name << "_members, " << a->members.size()
      << ", sizeof(" << a->name << ") }, &agents_" << a->name << ", -1 },";
  }
  w.nl();
  w << "{ { NULL, NULL, 0, 0 }, NULL, -1 }";
  w.outdent(); w.nl();
  w << "};"; w.nl();
  w << "const int abl_model_n_types = " << script.agents.size() << ";"; w.nl();
  w << "const int abl_model_use_float = " << (useFloat ? 1 : 0) << ";"; w.nl(); w.nl();

  hostSeqSupport();

  for (ConstDecl *c : script.consts) { constDecl(*c); w.nl(); }
  w.nl();

  FuncDecl *mainFn = script.mainFunc;
  for (FuncDecl *f : script.funcs) {
    if (f->isStep() || f == mainFn) continue;
    function(*f);
    w.nl();
  }
  w.nl();
  hostVisualize();
  hostSimulate();

  // main(): whole user main as abl_model_main, plus the pre-simulate part on its own
  curFn = mainFn;
  w << "int abl_model_main(void) {";
  w.indent(); stmts(mainFn->body); w.nl();
  w << "return 0;";
  w.outdent(); w.nl();
  w << "}"; w.nl(); w.nl();

  w << "/* statements of main() that precede `simulate` (population set-up only) */"; w.nl();
  w << "int abl_model_populate(void) {";
  w.indent();
  w.nl();
  w << "/* start from a clean slate so that repeated calls rebuild the same population */";
  w.nl();
  w << "for (int t = 0; t < abl_model_n_types; t++) abl_model_types[t].agents->len = 0;";
  w.nl();
  w << "abl_host_rng_reset();";
  for (const StmtP &s : mainFn->body) {
    if (s->kind == Stmt::Simulate) break;
    w.nl(); stmt(*s);
  }
  w.nl();
  w << "return 0;";
  w.outdent(); w.nl();
  w << "}"; w.nl(); w.nl();
  curFn = nullptr;

  w << "#ifndef ABL_MODEL_NO_MAIN"; w.nl();
  w << "int main(void) { return abl_model_main(); }"; w.nl();
  w << "#endif"; w.nl();
  return w.str();
}

// `-C visualize=true` (the reference's option name, FlameGPUBackend.cpp:277; its JVM back ends draw the model in a
// window, MasonPrinter.cpp:704-827): the model's getColor / getSize hooks — analysed by every back end, dropped by all
// but mason / dmason (AnalysisVisitor.cpp:215-222, src/Sema.cpp) — are printed as host functions, and the `simulate`
// statement writes one picture per frame interval: frames/frame_<timestep>.ppm (abl_host_frame_*, asset/cuda/abl_host.c).
void CudaPrinter::hostVisualize() {
  if (!visualize()) return;
  std::map<const AgentDecl *, std::string> colorOf, sizeOf;
  for (Decl &d : script.decls) {
    if (d.kind != Decl::Func || !d.func || !d.func->skipped) continue;
    FuncDecl &f = *d.func;
    if ((f.name != "getColor" && f.name != "getSize") || f.params.size() != 1 || !f.params[0].type.isAgent() ||
        !f.params[0].type.agent)
      continue;
    const AgentDecl *a = f.params[0].type.agent;
    const std::string name = "abl_vis_" + f.name + "_" + a->name;
    (f.name == "getColor" ? colorOf : sizeOf)[a] = name;
    curFn = &f;
    // (agent-typed variables are pointers to records in host code)
    w << "static " << typeName(f.retTy) << " " << name << "(const " << a->name << " *" << f.params[0].name << ") {";
    w.indent(); stmts(f.body); w.outdent();
    w.nl();
    w << "}"; w.nl();
    curFn = nullptr;
  }
  EnvDecl *env = script.env;
  char buf[256];
  w << "/* one picture of the host arrays as they are (after abl_model_populate or abl_model_download) */"; w.nl();
  w << "int abl_model_write_frame(const char *path) {"; w.nl();
  w << "    abl_frame fr;"; w.nl();
  snprintf(buf, sizeof buf, "    if (abl_host_frame_begin(&fr, 500, %.17g, %.17g, %.17g, %.17g)) return 1;",
           env ? env->envMin.v[0] : 0.0, env ? env->envMin.v[1] : 0.0, env ? env->envMax.v[0] : 1.0, env ? env->envMax.v[1] : 1.0);
  w << buf; w.nl();
  for (AgentDecl *a : script.agents) {
    AgentMember *pos = a->position();
    if (!pos) continue;
    w << "    for (size_t i = 0; i < agents_" << a->name << ".len; i++) {"; w.nl();
    w << "        const " << a->name << " *a = (const " << a->name << " *)agents_" << a->name << ".data + i;"; w.nl();
    w << "        abl_host_frame_dot(&fr, (double)a->" << pos->name << ".x, (double)a->" << pos->name << ".y, ";
    if (colorOf.count(a)) w << "(int)" << colorOf[a] << "(a)"; else w << "0";
    w << ", ";
    if (sizeOf.count(a)) w << "(double)" << sizeOf[a] << "(a)"; else w << "1.0";
    w << ");"; w.nl();
    w << "    }"; w.nl();
  }
  w << "    return abl_host_frame_end(&fr, path);"; w.nl();
  w << "}"; w.nl(); w.nl();
}

void CudaPrinter::hostSeqSupport() {
  w << "/* runtime handle used by `simulate` and by reductions in the sequential step */"; w.nl();
  w << "static abl_runtime *abl_rt = NULL;"; w.nl();
  w << "void abl_model_set_runtime(abl_runtime *rt) { abl_rt = rt; }"; w.nl();
  w << "abl_runtime *abl_model_runtime(void) { return abl_rt; }"; w.nl();
  w << "extern int abl_model_setup(abl_runtime *rt);"; w.nl();
  w << "extern int abl_model_parallel_steps(abl_runtime *rt);"; w.nl();
  w << "#define ABL_UNUSED __attribute__((unused))"; w.nl();
  w << "static ABL_UNUSED int abl_model_count(int t) { int r = 0; abl_host_check(abl_cuda_count(abl_rt, abl_model_types[t].pool, &r), \"count\"); return r; }"; w.nl();
  w << "static ABL_UNUSED int abl_model_count_member_int(int t, int m, int v) { int r = 0; abl_host_check(abl_cuda_count_member_int(abl_rt, abl_model_types[t].pool, m, v, &r), \"count\"); return r; }"; w.nl();
  w << "static ABL_UNUSED int abl_model_count_member_float(int t, int m, double v) { int r = 0; abl_host_check(abl_cuda_count_member_float(abl_rt, abl_model_types[t].pool, m, v, &r), \"count\"); return r; }"; w.nl();
  w << "static ABL_UNUSED int abl_model_sum_int(int t, int m) { int r = 0; abl_host_check(abl_cuda_sum_int(abl_rt, abl_model_types[t].pool, m, &r), \"sum\"); return r; }"; w.nl();
  w << "static ABL_UNUSED abl_real abl_model_sum_float(int t, int m) { double r = 0; abl_host_check(abl_cuda_sum_float(abl_rt, abl_model_types[t].pool, m, 0, &r), \"sum\"); return (abl_real)r; }"; w.nl();
  w << "static ABL_UNUSED abl_float2 abl_model_sum_float2(int t, int m) { double x = 0, y = 0; abl_host_check(abl_cuda_sum_float(abl_rt, abl_model_types[t].pool, m, 0, &x), \"sum\"); abl_host_check(abl_cuda_sum_float(abl_rt, abl_model_types[t].pool, m, 1, &y), \"sum\"); return float2_create((abl_real)x, (abl_real)y); }"; w.nl();
  w << "static ABL_UNUSED abl_float3 abl_model_sum_float3(int t, int m) { double v[3] = {0, 0, 0}; for (int k = 0; k < 3; k++) abl_host_check(abl_cuda_sum_float(abl_rt, abl_model_types[t].pool, m, k, &v[k]), \"sum\"); return float3_create((abl_real)v[0], (abl_real)v[1], (abl_real)v[2]); }"; w.nl();
  w << "static ABL_UNUSED abl_real abl_model_last_exec_time(void) { double s = 0; abl_host_check(abl_cuda_last_exec_time(abl_rt, &s), \"getLastExecTime\"); return (abl_real)s; }"; w.nl();
  w << "static ABL_UNUSED void abl_model_save(const char *path) {"; w.nl();
  {
    // save() format: the reference runtime knows three (asset/c/libabl.h:211-215); its `c` backend
    // always passes SAVE_JSON (CPrinter.cpp:74-77), the FLAME backends write their initial states
    // as XML.  Selected at compile time with -C cuda.save_format=json|flame_xml|flamegpu_xml.
    std::string fmt = config.getString("cuda.save_format", "json");
    if (fmt == "json") { w << "    abl_host_save_json(abl_model_types, abl_model_n_types, path);"; }
    else if (fmt == "flame_xml") { w << "    abl_host_save_flame_xml(abl_model_types, abl_model_n_types, path, 0);"; }
    else if (fmt == "flamegpu_xml") { w << "    abl_host_save_flame_xml(abl_model_types, abl_model_n_types, path, 1);"; }
    else throw BackendError("cuda backend: unknown cuda.save_format \"" + fmt + "\" (json, flame_xml, flamegpu_xml)");
    w.nl();
  }
  w << "    if (getenv(\"ABL_DUMP_STATE\")) { char raw[4096]; snprintf(raw, sizeof raw, \"%s.bin\", path); abl_host_save_raw(abl_model_types, abl_model_n_types, raw); }"; w.nl();
  w << "}"; w.nl();
  w.nl();
}

void CudaPrinter::hostSimulate() {
  FuncDecl *seq = script.simulate ? script.simulate->seqStep : nullptr;
  w << "/* one timestep: all parallel step functions in order, then the sequential step */"; w.nl();
  w << "int abl_model_timestep(abl_runtime *rt) {"; w.nl();
  w << "    abl_rt = rt;"; w.nl();
  w << "    abl_host_check(abl_cuda_begin_timestep(rt), \"begin_timestep\");"; w.nl();
  w << "    abl_host_check(abl_model_parallel_steps(rt), \"step\");"; w.nl();
  if (seq) { w << "    " << seq->emitName << "();"; w.nl(); }
  w << "    abl_host_check(abl_cuda_end_timestep(rt), \"end_timestep\");"; w.nl();
  w << "    return 0;"; w.nl();
  w << "}"; w.nl(); w.nl();

  w << "/* the sequential step alone (harnesses that drive the parallel step functions one by one) */"; w.nl();
  w << "int abl_model_sequential_step(abl_runtime *rt) {"; w.nl();
  w << "    abl_rt = rt;"; w.nl();
  if (seq) { w << "    " << seq->emitName << "();"; w.nl(); }
  w << "    return 0;"; w.nl();
  w << "}"; w.nl(); w.nl();

  w << "/* host record arrays are page-locked while the runtime uses them (direct DMA) */"; w.nl();
  w << "int abl_model_upload(abl_runtime *rt) {"; w.nl();
  w << "    for (int t = 0; t < abl_model_n_types; t++) {"; w.nl();
  w << "        abl_host_type *ty = &abl_model_types[t];"; w.nl();
  w << "        abl_host_check(abl_cuda_pin_host(rt, ty->agents->data, ty->agents->cap * ty->desc.stride), \"pin\");"; w.nl();
  w << "        abl_host_check(abl_cuda_upload(rt, ty->pool, ty->agents->data, ty->agents->len), \"upload\");"; w.nl();
  w << "    }"; w.nl();
  w << "    return 0;"; w.nl();
  w << "}"; w.nl();
  w << "int abl_model_download(abl_runtime *rt) {"; w.nl();
  w << "    for (int t = 0; t < abl_model_n_types; t++) {"; w.nl();
  w << "        abl_host_type *ty = &abl_model_types[t];"; w.nl();
  w << "        size_t n = 0;"; w.nl();
  w << "        abl_host_check(abl_cuda_pool_size(rt, ty->pool, &n), \"pool_size\");"; w.nl();
  w << "        if (n > ty->agents->cap) {"; w.nl();
  w << "            abl_host_check(abl_cuda_unpin_host(rt, ty->agents->data), \"unpin\");"; w.nl();
  w << "            abl_array_resize(ty->agents, ty->desc.stride, n);"; w.nl();
  w << "            abl_host_check(abl_cuda_pin_host(rt, ty->agents->data, ty->agents->cap * ty->desc.stride), \"pin\");"; w.nl();
  w << "        }"; w.nl();
  w << "        ty->agents->len = n;"; w.nl();
  w << "        abl_host_check(abl_cuda_download(rt, ty->pool, ty->agents->data, n, &n), \"download\");"; w.nl();
  w << "    }"; w.nl();
  w << "    return 0;"; w.nl();
  w << "}"; w.nl();
  w << "void abl_model_unpin(abl_runtime *rt) {"; w.nl();
  w << "    for (int t = 0; t < abl_model_n_types; t++) abl_cuda_unpin_host(rt, abl_model_types[t].agents->data);"; w.nl();
  w << "}"; w.nl(); w.nl();

  // `-C cuda.gpus=N` / ABL_CUDA_GPUS=N: the same simulate statement on N devices of this machine.  The decomposition is part
  // of the generated program, as in the reference's distributed backends (DMasonPrinter.cpp:30-32 dmason.grid_rows/cols,
  // FlameMainPrinter.cpp:40-41 `mpirun -np 4`); the runtime does the work (abl_cuda_group_simulate).
  const int gpus = config.getInt("cuda.gpus", 1);
  if (gpus < 1) throw BackendError("cuda backend: cuda.gpus must be at least 1");
  if (gpus > 1 && visualize()) throw BackendError("cuda backend: visualize=true is not supported with cuda.gpus > 1");
  bool addsAtRunTime = false;
  for (const StepInfo &si : steps) addsAtRunTime = addsAtRunTime || si.fn->addedAgent != nullptr;
  if (gpus > 1 && (seq || addsAtRunTime))
    throw BackendError("cuda backend: cuda.gpus > 1 is not supported for models with a sequential step or run-time add(): "
                       "their timestep needs the host between step functions");
  w << "/* one timestep of the parallel step functions; callable from several host threads, one runtime each */"; w.nl();
  w << "static int abl_model_timestep_mt(abl_runtime *rt) {"; w.nl();
  w << "    int rc;"; w.nl();
  w << "    if ((rc = abl_cuda_begin_timestep(rt))) return rc;"; w.nl();
  w << "    if ((rc = abl_model_parallel_steps(rt))) return rc;"; w.nl();
  w << "    return abl_cuda_end_timestep(rt);"; w.nl();
  w << "}"; w.nl(); w.nl();
  w << "/* the `simulate` statement: upload, run, download (replaces the reference's inline"; w.nl();
  w << "   double-buffered OpenMP loop) */"; w.nl();
  w << "static void abl_model_simulate(int timesteps) {"; w.nl();
  w << "    abl_config cfg;"; w.nl();
  w << "    abl_cuda_default_config(&cfg);"; w.nl();
  w << "    cfg.use_float = abl_model_use_float;"; w.nl();
  w << "    cfg.block_size = " << config.getInt("cuda.block_size", 0) << ";"; w.nl();
  w << "    cfg.tile_neighbours = " << (config.getBool("cuda.tile", false) ? 1 : 0) << ";"; w.nl();
  w << "    if (getenv(\"ABL_CUDA_DEVICE\")) cfg.device = atoi(getenv(\"ABL_CUDA_DEVICE\"));"; w.nl();
  w << "    int gpus = " << gpus << ";   /* -C cuda.gpus */"; w.nl();
  w << "    if (getenv(\"ABL_CUDA_GPUS\")) gpus = atoi(getenv(\"ABL_CUDA_GPUS\"));"; w.nl();
  w << "    if (gpus > 1) {"; w.nl();
  if (visualize()) { w << "        fprintf(stderr, \"ABL_CUDA_GPUS > 1: no frames are written (visualize=true draws single-device runs)\\n\");"; w.nl(); }
  if (seq || addsAtRunTime) {
    w << "        fprintf(stderr, \"ABL_CUDA_GPUS > 1 is not supported for this model (sequential step or run-time add())\\n\");"; w.nl();
    w << "        exit(1);"; w.nl();
  } else {
    w << "        int pools[" << script.agents.size() << "], has_pos[" << script.agents.size() << "];"; w.nl();
    w << "        void *data[" << script.agents.size() << "]; size_t len[" << script.agents.size() << "], stride[" << script.agents.size() << "];"; w.nl();
    w << "        /* pool indices follow the order of registration in abl_model_setup: agent declaration order */"; w.nl();
    w << "        for (int t = 0; t < abl_model_n_types; t++) {"; w.nl();
    w << "            abl_host_type *ty = &abl_model_types[t];"; w.nl();
    w << "            pools[t] = t; data[t] = ty->agents->data; len[t] = ty->agents->len; stride[t] = ty->desc.stride;"; w.nl();
    w << "            has_pos[t] = 0;"; w.nl();
    w << "            for (int m = 0; m < ty->desc.n_members; m++) if (ty->desc.members[m].is_pos) has_pos[t] = 1;"; w.nl();
    w << "        }"; w.nl();
    w << "        abl_group_population pop = { abl_model_n_types, pools, data, len, stride, has_pos };"; w.nl();
    w << "        abl_host_check(abl_cuda_group_simulate(&cfg, gpus, abl_model_setup, abl_model_timestep_mt, timesteps, &pop), \"simulate on several GPUs\");"; w.nl();
    w << "        return;"; w.nl();
  }
  w << "    }"; w.nl();
  w << "    abl_runtime *rt = NULL;"; w.nl();
  w << "    abl_host_check(abl_cuda_create(&rt, &cfg), \"create\");"; w.nl();
  w << "    abl_host_check(abl_model_setup(rt), \"setup\");"; w.nl();
  w << "    abl_model_upload(rt);"; w.nl();
  if (visualize()) {
    const int every = config.getInt("cuda.frame_interval", 1);
    if (every < 1) throw BackendError("cuda backend: cuda.frame_interval must be at least 1");
    w << "    /* -C visualize=true: frames/frame_<timestep>.ppm of the initial state and after every " << every << " timestep(s) */"; w.nl();
    w << "    abl_host_make_dir(\"frames\");"; w.nl();
    w << "    abl_model_write_frame(\"frames/frame_00000.ppm\");"; w.nl();
    w << "    for (int t = 0; t < timesteps; t++) {"; w.nl();
    w << "        abl_model_timestep(rt);"; w.nl();
    w << "        if ((t + 1) % " << every << " == 0 || t + 1 == timesteps) {"; w.nl();
    w << "            char frame_path[64];"; w.nl();
    w << "            abl_model_download(rt);"; w.nl();
    w << "            snprintf(frame_path, sizeof frame_path, \"frames/frame_%05d.ppm\", t + 1);"; w.nl();
    w << "            abl_model_write_frame(frame_path);"; w.nl();
    w << "        }"; w.nl();
    w << "    }"; w.nl();
  } else {
    w << "    for (int t = 0; t < timesteps; t++) abl_model_timestep(rt);"; w.nl();
  }
  w << "    abl_model_download(rt);"; w.nl();
  w << "    abl_model_unpin(rt);"; w.nl();
  w << "    abl_rt = NULL;"; w.nl();
  w << "    abl_host_check(abl_cuda_destroy(rt), \"destroy\");"; w.nl();
  w << "}"; w.nl(); w.nl();
}

// ---------------------------------------------------------------------------------------
// device program
// ---------------------------------------------------------------------------------------
void CudaPrinter::loadMember(const AgentDecl &a, int m, const std::string &dst,
                             const std::string &view, const std::string &idx) {
  const Ty &t = a.members[m]->type;
  int c = columnOf(a, m);
  if (t.k == TK::Vec3) w << dst << " = abl_ld3(" << view << ", " << c << ", " << idx << ");";
  else w << dst << " = abl_ld<" << typeName(t) << ">(" << view << "[" << c << "], " << idx << ");";
}

void CudaPrinter::storeMember(const AgentDecl &a, int m, const std::string &cols,
                              const std::string &idx, const std::string &value) {
  const Ty &t = a.members[m]->type;
  int c = columnOf(a, m);
  if (t.k == TK::Vec3) w << "abl_st3(" << cols << ", " << c << ", " << idx << ", " << value << ");";
  else w << "abl_st<" << typeName(t) << ">(" << cols << "[" << c << "], " << idx << ", " << value << ");";
}

const Expr *CudaPrinter::findNearRadius(const std::vector<StmtP> &body) {
  for (const StmtP &s : body) {
    if (s->kind == Stmt::For && s->forKind == Stmt::ForNear) return s->e[0]->kids[1].get();
    if (const Expr *r = findNearRadius(s->body)) return r;
  }
  return nullptr;
}

// true if `e` only involves literals, scalar global constants and arithmetic: the launcher
// (host code in the same .cu) can then evaluate it with the arithmetic the kernel would use
bool CudaPrinter::hostEvaluable(const Expr &e) {
  switch (e.kind) {
    case Expr::BoolLit: case Expr::IntLit: case Expr::FloatLit: return true;
    case Expr::Var: return e.sym && e.sym->global && !e.type.isVec() && !e.type.isArray() && (e.type.isNum() || e.type.isBool());
    case Expr::Unary: case Expr::Binary: case Expr::Ternary:
      if (e.type.isVec()) return false;
      for (const ExprP &k : e.kids) if (!hostEvaluable(*k)) return false;
      return true;
    case Expr::Call:
      if (e.ckind != Expr::Ctor || e.type.isVec()) return false;
      return hostEvaluable(*e.kids[0]);
    default: return false;
  }
}

// Cached neighbour lists (`-C cuda.nlist=true`): the accepted candidates of an agent are the same
// in every timestep when no step function of the model writes the position of, removes or adds
// agents of either type of the for-near loop — and the step itself neither removes nor adds
// (the list-building launches run the function body without its side effects on the pool only).
bool CudaPrinter::stepListEligible(const StepInfo &si) const {
  const FuncDecl &f = *si.fn;
  if (!config.getBool("cuda.nlist", false) || !f.nearAgent || !curStepHasLimit) return false;
  if (f.usesRemoval || f.addedAgent) return false;
  AgentDecl *types[2] = { si.self, f.nearAgent };
  for (AgentDecl *t : types) {
    AgentMember *pos = t ? t->position() : nullptr;
    if (!pos) return false;
    for (const StepInfo &o : steps) {
      if (o.self == t && (o.writes.count(pos->name) || o.fn->usesRemoval)) return false;
      if (o.fn->addedAgent == t) return false;
    }
  }
  return true;
}

// The __global__ wrapper of a step function: thread -> agent, loads of the members the step reads,
// the tile staging of ABL_MODE 2, the call, stores of the written members, the fused histogram of
// the next binning and the slab epilogue.
void CudaPrinter::stepKernelWrapper(const StepKernelCtx &C) {
  const StepInfo &si = C.si;
  FuncDecl &f = *si.fn;
  AgentDecl &self = *si.self;
  const Param &p = f.params[0];
  const Expr *radius = C.radius;
  AgentMember *selfPosM = self.position();
  const std::vector<TileCol> &tcols = C.tcols;
  const int tdim = C.tdim;
  const std::string &trows = C.trows;
  const bool sql = C.sql;
  (void)p; (void)radius; (void)selfPosM; (void)tcols; (void)tdim; (void)trows; (void)sql;
  w << "template <int ABL_MODE>"; w.nl();
  // (two resident CTAs of 256 threads are all the shadow pre-filter's shared memory allows: let it have the registers)
  w << "__global__ void __launch_bounds__(256, ABL_MODE == 8 ? 2 : 0) abl_kernel_" << f.emitName
    << "(const __grid_constant__ abl_step_launch _a, const abl_real _near_limit, const abl_real _near_cull, "
    << (sql ? "const abl_sq_limits _sql, " : "") << "const unsigned _tile_cap) {";
  w.indent(); w.nl();
  if (curStepBulk) {
    // the mbarrier of the bulk-staged tile: no global memory involved, so this overlaps the tail of the preceding kernel
    w << "#ifdef ABL_HAVE_BULK_TILE"; w.nl();
    w << "if (ABL_MODE == 7) { extern __shared__ __align__(16) unsigned char _abl_smem0[]; abl_btile_begin(_abl_smem0); }"; w.nl();
    w << "#endif"; w.nl();
  }
  // (pdl bit 1: the blocks of the NEXT kernel of the stream may become resident while this kernel's last wave
  // is still running — they wait in their own cudaGridDependencySynchronize until this grid has completed)
  w << "if (_a.pdl & 2) cudaTriggerProgrammaticLaunchCompletion();"; w.nl();
  w << "cudaGridDependencySynchronize();   // programmatic dependent launch: wait for the preceding kernel"; w.nl();
  // _r: index inside the launched (owned) range, _i: index in the pool's columns
  w << "bool _boundary;"; w.nl();
  w << "unsigned _ob;"; w.nl();
  w << "const unsigned _r = abl_agent_index(_a, _boundary, _ob);"; w.nl();
  w << "const unsigned _i = _r + _ob;"; w.nl();
  std::set<std::string> loads = si.reads;
  for (const std::string &m : si.writes) loads.insert(m);
  // the slab epilogue routes by the agent's position: always have it in registers
  if (selfPosM) loads.insert(selfPosM->name);
  if (curStepTile) {
    // tiled kernels keep surplus threads of the last block alive for the cooperative staging
    w << "const bool _active = _r != 0xffffffffu;"; w.nl();
    w << "if (ABL_MODE != 2 && ABL_MODE != 7 && !_active) return;"; w.nl();
    w << self.name << " " << p.name << " = {};"; w.nl();
    w << "if (_active) {";
    w.indent();
  } else {
    w << "if (_r == 0xffffffffu) return;"; w.nl();
    w << self.name << " " << p.name << ";";
  }
  for (size_t m = 0; m < self.members.size(); m++) {
    if (!loads.count(self.members[m]->name)) continue;
    w.nl();
    loadMember(self, (int)m, p.name + "." + self.members[m]->name, "_a.self.in", "_i");
  }
  w.nl();
  if (curStepTile) {
    w.outdent(); w << "}"; w.nl();
    w << "bool _tile_ok = false;"; w.nl();
    w << "if (ABL_MODE == 2) {";
    w.indent(); w.nl();
    w << "extern __shared__ __align__(16) unsigned char _abl_smem[];"; w.nl();
    w << "abl_near_iter<" << tdim << "> _rows;"; w.nl();
    w << "_rows.rows" << tdim << "(_a, " << p.name << "." << selfPosM->name << ", _active, _near_cull);"; w.nl();
    w << "_tile_ok = abl_tile_plan<" << tdim << ">(_rows, _tile_cap, _abl_smem);"; w.nl();
    w << "if (_tile_ok) {";
    w.indent(); w.nl();
    w << "unsigned char *const _cols = _abl_smem + ABL_TILE_HDR_BYTES + " << trows << " * blockDim.x * sizeof(uint2);"; w.nl();
    w << "const unsigned _total = reinterpret_cast<const abl_tile_hdr *>(_abl_smem)->total;"; w.nl();
    w << "for (unsigned _e = threadIdx.x; _e < _total; _e += blockDim.x) {";
    w.indent(); w.nl();
    w << "const unsigned _g = abl_tile_src<" << tdim << ">(_abl_smem, _e);";
    for (size_t q = 0; q < tcols.size(); q++) {
      w.nl();
      w << "reinterpret_cast<" << tcols[q].ctype << " *>(_cols + (size_t)_tile_cap * " << tileOffset(tcols, q) << ")[_e] = "
        << "abl_ld<" << tcols[q].ctype << ">(_a.nbr.in[" << tcols[q].column << "], _g);";
    }
    w.outdent(); w.nl();
    w << "}"; w.nl();
    w << "__syncthreads();";
    w.outdent(); w.nl();
    w << "}"; w.nl();
    w << "if (!_active) return;";
    w.outdent(); w.nl();
    w << "}"; w.nl();
  } else {
    w << "const bool _tile_ok = false;"; w.nl();
  }
  if (curStepBulk) {
    // ABL_MODE 7: warp 0 plans the row ranges of the block from its first and last agent and starts
    // the bulk copies; every thread fetches its own row ranges meanwhile and then waits on the mbarrier
    const int pc = columnOf(self, self.memberIndex(selfPosM->name));
    w << "#ifdef ABL_HAVE_BULK_TILE"; w.nl();
    w << "abl_btile_flat2 _bt;"; w.nl();
    w << "_bt.ok = false; _bt.T1 = _bt.T2 = _bt.N = _bt.O0 = _bt.O1 = _bt.O2 = 0;"; w.nl();
    w << "if (ABL_MODE == 7) {";
    w.indent(); w.nl();
    w << "extern __shared__ __align__(16) unsigned char _abl_smem[];"; w.nl();
    w << "const abl_btile_cols _TC = { " << tcols.size() << ", {";
    for (size_t q = 0; q < tcols.size(); q++) w << (q ? ", " : "") << tcols[q].column;
    w << "}, {";
    for (size_t q = 0; q < tcols.size(); q++) w << (q ? ", " : "") << "(int)" << tcols[q].bytes;
    w << "}, {";
    for (size_t q = 0; q < tcols.size(); q++) w << (q ? ", " : "") << "(unsigned)" << tileOffset(tcols, q);
    w << "}, (unsigned)" << tileOffset(tcols, tcols.size()) << " };"; w.nl();
    w << "if (threadIdx.x < 32) {";
    w.indent(); w.nl();
    w << "unsigned _bf = 0, _bl = 0;"; w.nl();
    w << "const bool _any = abl_block_span(_a, _bf, _bl);"; w.nl();
    w << "const abl_real *const _pc = static_cast<const abl_real *>(_a.self.in[" << pc << "]);"; w.nl();
    w << "abl_real _pf[2] = {0, 0}, _pl[2] = {0, 0};"; w.nl();
    w << "if (_any) { _pf[0] = __ldg(_pc + 2 * (size_t)_bf); _pf[1] = __ldg(_pc + 2 * (size_t)_bf + 1); "
      << "_pl[0] = __ldg(_pc + 2 * (size_t)_bl); _pl[1] = __ldg(_pc + 2 * (size_t)_bl + 1); }"; w.nl();
    w << "abl_btile_plan<2>(_a, _pf, _pl, _any, _tile_cap, _abl_smem, _TC);";
    w.outdent(); w.nl();
    w << "}"; w.nl();
    w << "if (!_active) return;"; w.nl();
    w << "abl_near_iter<2> _rows;"; w.nl();
    w << "_rows.rows2(_a, " << p.name << "." << selfPosM->name << ", true, _near_cull);"; w.nl();
    w << "_tile_ok = abl_btile_wait(_abl_smem, (_a.pdl & 4) != 0);"; w.nl();
    w << "_bt = abl_btile_thread2(_rows, _abl_smem, _tile_ok);";
    w.outdent(); w.nl();
    w << "}"; w.nl();
    w << "#endif"; w.nl();
  }
  w << self.name << " " << p.outName << " = " << p.name << ";"; w.nl();
  w << "abl_ctx _ctx;"; w.nl();
  if (f.usesRng) { w << "abl_ctx_init(_ctx, _a.seed, _a.timestep, _a.step_index, _a.self.id[_i]);"; w.nl(); }
  else { w << "_ctx.rng = 0; _ctx.dead = false; _ctx.added = false;"; w.nl(); }
  w << f.emitName << "<ABL_MODE>(_ctx, _a, _i, _near_limit, _near_cull, " << (sql ? "_sql, " : "") << "_tile_cap, _tile_ok" << (curStepBulk ? " ABL_BT_ARG" : "") << ", " << p.name << ", " << p.outName << ");";
  if (curStepList) { w.nl(); w << "if (ABL_MODE == 5 || ABL_MODE == 6) return;   // list-building launch: nothing of the step is stored"; }
  AgentMember *selfPos = self.position();
  for (size_t m = 0; m < self.members.size(); m++) {
    if (!si.writes.count(self.members[m]->name)) continue;
    w.nl();
    storeMember(self, (int)m, "_a.self.out", "_i", p.outName + "." + self.members[m]->name);
  }
  if (selfPos && si.writes.count(selfPos->name)) {
    // fused histogram of the next binning (no-op unless the runtime asks for it)
    w.nl();
    w << "abl_bin_epilogue" << selfPos->type.vecLen() << "(_a, _r, " << p.outName << "." << selfPos->name << ");";
  }
  // slab decomposition: route this agent's record to the neighbouring slabs (no-op otherwise)
  w.nl();
  if (selfPos) w << "abl_slab_epilogue" << selfPos->type.vecLen() << "(_a, _i, " << p.outName << "." << selfPos->name << ", _boundary);";
  if (f.usesRemoval) { w.nl(); w << "_a.dead[_i] = _ctx.dead ? 1 : 0;"; }
  if (f.addedAgent) { w.nl(); w << "_a.add_flag[_i] = _ctx.added ? 1 : 0;"; }
  if (selfPos) { w.nl(); w << "abl_slab_block_done(_a, _boundary);"; }
  w.outdent(); w.nl();
  w << "}"; w.nl(); w.nl();

}

// The host-side launcher registered with the runtime (abl_step_desc.launch): bounds on the squared
// distance, culling range, choice of the kernel variant (density rule or run-time tuner) and of
// the block size, launch.
void CudaPrinter::stepLauncher(const StepKernelCtx &C) {
  const StepInfo &si = C.si;
  FuncDecl &f = *si.fn;
  AgentDecl &self = *si.self;
  const Param &p = f.params[0];
  const Expr *radius = C.radius;
  AgentMember *selfPosM = self.position();
  const std::vector<TileCol> &tcols = C.tcols;
  const int tdim = C.tdim;
  const std::string &trows = C.trows;
  const bool sql = C.sql;
  (void)p; (void)radius; (void)selfPosM; (void)tcols; (void)tdim; (void)trows; (void)sql;
  w << "static int abl_last_mode_" << f.emitName << " = -1;   // ABL_MODE of the most recent launch (abl_model_step_variant)"; w.nl();
  w << "static int abl_launch_" << f.emitName << "(const abl_step_launch *_args) {"; w.nl();
  w << "    abl_step_launch _copy = *_args;   // block counts of the boundary parts are filled in below"; w.nl();
  if (!curStepDense) { w << "    if (_args->probe) return 0;   // no variant of this step uses the single-precision shadow"; w.nl(); }
  w << "    abl_step_launch *a = &_copy;"; w.nl();
  w << "    int bs = a->block_size > 0 && a->block_size <= 256 ? a->block_size : 0;"; w.nl();
  if (curStepHasLimit) {
    // host-side evaluation of the radius with the kernel's own arithmetic and constants
    Target saved = target;
    // (function-local statics with an initialiser: initialised once, thread-safe — the group driver calls the
    // launcher from one host thread per device)
    w << "    static const abl_real limit = abl_near_sq_limit((abl_real)(";
    expr(*radius);
    w << "));"; w.nl();
    target = saved;
  } else {
    w << "    const abl_real limit = 0;"; w.nl();
  }
  if (sql) {
    w << "    static const abl_sq_limits sql = [] {"; w.nl();
    w << "        abl_sq_limits s;"; w.nl();
    w << "        memset(&s, 0, sizeof s);"; w.nl();
    for (size_t k = 0; k < sqLimits.size(); k++) {
      w << "        s.v[" << k << "] = abl_sq_cmp_limit(" << sqLimits[k].first << ", (abl_real)(" << sqLimits[k].second << "));"; w.nl();
    }
    w << "        return s;"; w.nl();
    w << "    }();"; w.nl();
  }
  // cell-range culling for radii well below the cell size (`-C cuda.cull=false` turns it off)
  if (curStepHasLimit && config.getBool("cuda.cull", true)) {
    w << "    const abl_real cull = abl_near_cull_range(limit, a->grid.cell_size);"; w.nl();
  } else {
    w << "    const abl_real cull = ABL_R(-1.0);"; w.nl();
  }
  const std::string lim = sql ? "limit, cull, sql" : "limit, cull";
  const std::string K = "abl_kernel_" + f.emitName;
  // Which loop: the density rule for pinned runs, a timed choice otherwise.  row_occ = mean number
  // of candidates in one visited row (cells of a row after culling x agents per cell).
  w << "    const double occ = a->grid.n_cells ? (double)a->nbr.n / (double)a->grid.n_cells : 0.0;"; w.nl();
  w << "    const double row_cells = cull > ABL_R(0.0) ? fmin(3.0, 1.0 + 2.0 * (double)cull / a->grid.cell_size) : 2.0 * a->reach + 1.0;"; w.nl();
  w << "    const double row_occ = occ * row_cells;"; w.nl();
  if (curStepHasLimit && config.getBool("cuda.rowcull", false)) {
    // `-C cuda.rowcull=true` (off by default): dense rows (8 candidates and more), one cell of reach, radius not clearly below
    // the cell size: every row is narrowed along x to the cells within reach of the agent, 24 % fewer candidates per AGENT in
    // 3-D.  Measured on B200 (circle3d 1 M): 3.37 -> 4.02 ms per step — the lanes of a warp then walk different ranges, the
    // warp still covers their union, and the chunked loops lose their lock step (23 instead of 30 active lanes in phase 1).
    w << "    a->row_cull = (a->reach == 1 && row_occ >= 8.0 && !(cull > ABL_R(0.0)) && limit > ABL_R(0.0) && limit < (abl_real)INFINITY) ? (double)limit : 0.0;"; w.nl();
  }
  w << "    int dev = 0;"; w.nl();
  w << "    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= ABL_TUNE_DEVICES) dev = 0;"; w.nl();
  w << "    const bool can_flat = " << (curStepFlat ? "a->reach == 1" : "false") << ";"; w.nl();
  if (curStepHasLimit) {
    // Dense neighbourhoods: chunked two-phase loop.  Against the cursor loop it pays from a mean row
    // of 8 candidates (round 1); against the flat loop only for crowded rows — measured on B200 in
    // round 2 (profiles/r2/): game_of_life (7 per row) flat 0.43 ms, chunked 0.94 ms; every loop of
    // predator_prey 4 M (rows of 2 .. 43) flat, 1.30 ms per timestep against 1.72 ms by the old rule.
    w << "    const bool chunked = row_occ >= (can_flat && a->flat_loop != 0 ? 64.0 : 8.0);"; w.nl();
  } else {
    w << "    const bool chunked = false;"; w.nl();
  }
  w << "    int mode = chunked ? 1 : (a->flat_loop > 0 && can_flat ? 3 : 0);"; w.nl();
  w << "    bool timed = false;"; w.nl();
  w << "    static abl_tuner tune[ABL_TUNE_DEVICES];"; w.nl();
  if (curStepList) {
    // cached neighbour lists: count / fill launches of the runtime, then the list walk
    w << "    const bool listed = a->nlist_phase != 0 || a->nlist_cnt != nullptr;"; w.nl();
    w << "    if (listed) mode = a->nlist_phase == 1 ? 5 : a->nlist_phase == 2 ? 6 : 4;"; w.nl();
  } else {
    w << "    const bool listed = false;"; w.nl();
  }
  if (curStepDense) {
    // the runtime's question before the launch proper (nothing is launched): would this launch use the single-precision shadow?
    // (2: it would also use the pre-filter kernel's scratch columns, ABL_MODE 9 — the candidates of an agent fit ABL_SHADOW_WORDS masks)
    w << "    const bool split = " << (curStepSplit ? "a->reach == 1 && row_occ * " + std::string(tdim == 3 ? "9.0" : "3.0") + " <= 0.75 * 32.0 * ABL_SHADOW_WORDS" : "false") << ";"; w.nl();
    w << "    if (a->probe) return (a->flat_loop != 0 && chunked && !listed) ? (split ? 2 : 1) : 0;"; w.nl();
  }
  if (curStepTile) {
    // opt-in: stage the block's candidate rows in shared memory (ABL_MODE 2), sparse neighbourhoods only
    w << "    const unsigned tile_entry = " << tileOffset(tcols, tcols.size()) << ";"; w.nl();
    w << "    const int tile_bs = bs ? bs : 128;"; w.nl();
    w << "    const unsigned tile_cap = (a->tile_neighbours && !chunked && !listed && tile_bs % 32 == 0) ? abl_tile_capacity(a, " << trows << ", tile_bs, tile_entry, 64u * 1024u) : 0u;"; w.nl();
    w << "    if (tile_cap) {"; w.nl();
    w << "        const unsigned tile_grid = abl_grid_blocks(a, tile_bs);"; w.nl();
    w << "        const size_t smem = ABL_TILE_HDR_BYTES + (size_t)" << trows << " * tile_bs * sizeof(uint2) + (size_t)tile_cap * tile_entry;"; w.nl();
    w << "        static bool tile_set[ABL_TUNE_DEVICES];   // the attribute is per device"; w.nl();
    w << "        if (!tile_set[dev]) { cudaFuncSetAttribute(" << K << "<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); tile_set[dev] = true; }"; w.nl();
    w << "        return (int)abl_launch_kernel(a, " << K << "<2>, tile_grid, tile_bs, smem, *a, " << lim << ", tile_cap);"; w.nl();
    w << "    }"; w.nl();
  }
  if (curStepDense) {
    // dense rows: the chunked loop with its filter on the single-precision shadow (ABL_MODE 8) when the runtime keeps one.
    // The runtime first asks (probe != 0, nothing is launched) whether this launch would use the shadow, and only then builds it.
    w << "#ifdef ABL_HAVE_BULK_TILE"; w.nl();
    w << "    if (a->nbr_shadow != nullptr && a->flat_loop != 0 && chunked && !listed) {"; w.nl();
    w << "        const int dbs = bs ? bs : 256;"; w.nl();
    if (curStepSplit) {
      w << "        if (split && a->pf_masks != nullptr) {"; w.nl();
      w << "            // phase 1 as a kernel of its own, then the step kernel walks the masks it left in global memory"; w.nl();
      w << "            const unsigned pgrid = abl_grid_blocks(a, 256);"; w.nl();
      w << "            int prc = (int)abl_launch_kernel(a, abl_prefilter_" << f.emitName << ", pgrid, 256, 0, *a, limit, cull);"; w.nl();
      w << "            const unsigned sgrid = abl_grid_blocks(a, dbs);"; w.nl();
      w << "            if (prc == 0) prc = (int)abl_launch_kernel(a, " << K << "<9>, sgrid, dbs, (size_t)ABL_SHADOW_LIST * dbs * sizeof(unsigned), *a, " << lim << ", 0u);"; w.nl();
      w << "            if (prc == 0) { abl_last_mode_" << f.emitName << " = 9; return 0; }"; w.nl();
      w << "            (void)cudaGetLastError();"; w.nl();
      w << "        }"; w.nl();
    }
    w << "        const size_t dsmem = (size_t)(ABL_SHADOW_WORDS + ABL_SHADOW_ROWS + ABL_SHADOW_LIST) * dbs * sizeof(unsigned);"; w.nl();
    w << "        static size_t dense_set[ABL_TUNE_DEVICES];   // opt-in shared memory size, per device"; w.nl();
    w << "        if (dsmem > 48u * 1024u && dense_set[dev] < dsmem) { cudaFuncSetAttribute(" << K << "<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem); dense_set[dev] = dsmem; }"; w.nl();
    w << "        abl_last_mode_" << f.emitName << " = 8;"; w.nl();
    w << "        const unsigned dgrid = abl_grid_blocks(a, dbs);"; w.nl();
    w << "        const int drc = (int)abl_launch_kernel(a, " << K << "<8>, dgrid, dbs, dsmem, *a, " << lim << ", 0u);"; w.nl();
    w << "        if (drc == 0) return 0;"; w.nl();
    w << "        (void)cudaGetLastError();   // could not be launched: the chunked loop below"; w.nl();
    w << "    }"; w.nl();
    w << "#endif"; w.nl();
  }
  if (curStepBulk) {
    // the launcher's rule for sparse 2-D loops: flat loop, from a TMA-staged tile when the block's rows fit
    w << "#ifdef ABL_HAVE_BULK_TILE"; w.nl();
    // (only for entries of 32 bytes and more — boids2d in double: +9 % at 1 M agents, +13 % at 16 M;
    // lighter loops do not wait for their loads enough to repay the planning: boids2d in float
    // neutral, game_of_life (17 bytes) 13 % slower — profiles/r2/r2b_*.json)
    w << "    const unsigned bentry = " << tileOffset(tcols, tcols.size()) << ";"; w.nl();
    w << "    if (a->bulk_tile && (a->bulk_tile > 1 || bentry >= 32u) && a->flat_loop > 0 && can_flat && !chunked && !listed) {"; w.nl();
    w << "        const int bbs = bs ? bs : 128;"; w.nl();
    w << "        const unsigned bcap = bbs % 32 == 0 ? abl_tile_capacity(a, 3, bbs, bentry, 96u * 1024u) : 0u;"; w.nl();
    w << "        if (bcap) {"; w.nl();
    w << "            const size_t bsmem = ABL_BTILE_HDR_BYTES + (size_t)bcap * bentry;"; w.nl();
    w << "            static size_t bulk_set[ABL_TUNE_DEVICES];   // opt-in shared memory size, per device"; w.nl();
    w << "            if (bsmem > 48u * 1024u && bulk_set[dev] < bsmem) { cudaFuncSetAttribute(" << K << "<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem); bulk_set[dev] = bsmem; }"; w.nl();
    w << "            abl_last_mode_" << f.emitName << " = 7;"; w.nl();
    w << "            const unsigned bgrid = abl_grid_blocks(a, bbs);"; w.nl();
    w << "            const int brc = (int)abl_launch_kernel(a, " << K << "<7>, bgrid, bbs, bsmem, *a, " << lim << ", bcap);"; w.nl();
    w << "            if (brc == 0) return 0;"; w.nl();
    w << "            (void)cudaGetLastError();   // could not be launched: the flat loop over global memory below"; w.nl();
    w << "        }"; w.nl();
    w << "    }"; w.nl();
    w << "#endif"; w.nl();
  }
  if (curStepHasLimit) {
    // a->flat_loop < 0 (the runtime's default): the plausible variants are timed over the first
    // launches (abl_device.cuh: abl_tuner) — cursor loop and flat loop unless the rows are
    // crowded, the chunked loop unless they are nearly empty
    w << "    if (a->flat_loop < 0 && !listed) {"; w.nl();
    w << "        int var[ABL_TUNE_VARIANTS], nvar = 0;"; w.nl();
    w << "        unsigned vmask = 0;"; w.nl();
    w << "        if (row_occ <= 48.0) { var[nvar++] = 0; vmask |= 1u; if (can_flat) { var[nvar++] = 3; vmask |= 2u; } }"; w.nl();
    w << "        if (row_occ >= 2.0) { var[nvar++] = 1; vmask |= 4u; }"; w.nl();
    w << "        timed = nvar > 1;"; w.nl();
    w << "        mode = var[abl_tune_begin(tune, vmask, nvar, \"" << f.emitName << ": candidate loop (0 cursor, 3 flat, 1 chunked; in that order)\", a->stream)];"; w.nl();
    w << "    }"; w.nl();
  }
  // block size 0 = automatic: 128 threads for the per-candidate loops, 256 for the chunked one
  // (measured on circle3d 1 M: 3.9 ms against 4.5 ms per step)
  w << "    if (bs == 0) bs = mode == 1 ? 256 : 128;"; w.nl();
  w << "    const unsigned grid = abl_grid_blocks(a, bs);"; w.nl();
  w << "    int rc;"; w.nl();
  w << "    abl_last_mode_" << f.emitName << " = mode;"; w.nl();
  w << "    switch (mode) {"; w.nl();
  if (curStepHasLimit) {
    w << "    case 1: {"; w.nl();
    w << "        static bool smem_set[ABL_TUNE_DEVICES];   // the attribute is per device"; w.nl();
    w << "        if (!smem_set[dev]) { cudaFuncSetAttribute(" << K << "<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ABL_MASK_WORDS * 256 * (int)sizeof(unsigned)); smem_set[dev] = true; }"; w.nl();
    w << "        rc = (int)abl_launch_kernel(a, " << K << "<1>, grid, bs, (size_t)ABL_MASK_WORDS * bs * sizeof(unsigned), *a, " << lim << ", 0u);"; w.nl();
    w << "        break;"; w.nl();
    w << "    }"; w.nl();
  }
  if (curStepFlat) { w << "    case 3: rc = (int)abl_launch_kernel(a, " << K << "<3>, grid, bs, 0, *a, " << lim << ", 0u); break;"; w.nl(); }
  if (curStepList) {
    for (int mlist = 4; mlist <= 6; mlist++) {
      w << "    case " << mlist << ": rc = (int)abl_launch_kernel(a, " << K << "<" << mlist << ">, grid, bs, 0, *a, " << lim << ", 0u); break;"; w.nl();
    }
  }
  w << "    default: rc = (int)abl_launch_kernel(a, " << K << "<0>, grid, bs, 0, *a, " << lim << ", 0u); break;"; w.nl();
  w << "    }"; w.nl();
  // a trial variant that cannot even be launched must not stop the simulation: back to the cursor loop
  w << "    if (rc != 0 && timed && mode != 0) {"; w.nl();
  w << "        abl_tune_abort(tune);"; w.nl();
  w << "        timed = false;"; w.nl();
  w << "        abl_last_mode_" << f.emitName << " = 0;"; w.nl();
  w << "        const unsigned g0 = abl_grid_blocks(a, 128);   // fills in the block counts of *a: its own statement, before *a is copied"; w.nl();
  w << "        rc = (int)abl_launch_kernel(a, " << K << "<0>, g0, 128, 0, *a, " << lim << ", 0u);"; w.nl();
  w << "    }"; w.nl();
  w << "    if (timed) abl_tune_end(tune, a->stream);"; w.nl();
  w << "    return rc;"; w.nl();
  w << "}"; w.nl(); w.nl();
}

void CudaPrinter::stepKernel(const StepInfo &si, int index) {
  FuncDecl &f = *si.fn;
  AgentDecl &self = *si.self;
  const Param &p = f.params[0];
  curFn = &f;
  curStep = &si;
  const Expr *radius = f.nearAgent ? findNearRadius(f.body) : nullptr;
  curStepHasLimit = radius && hostEvaluable(*radius);
  // Shared-memory tiling needs the loop to range over the neighbourhood of the stepped agent
  // itself (`near(in, r)`): the kernel prologue plans the tile from in.pos of every thread.
  const Stmt *nearStmt = f.nearAgent ? findNearStmt(f.body) : nullptr;
  AgentMember *selfPosM = self.position();
  curStepTile = false;
  std::vector<TileCol> tcols;
  int tdim = 0;
  if (nearStmt && selfPosM && nearStmt->declTy.agent && nearStmt->declTy.agent->position()) {
    const Expr &ae = *nearStmt->e[0]->kids[0];
    curStepTile = ae.kind == Expr::Var && ae.sym && ae.sym == p.sym;
    if (curStepTile) {
      tcols = tileColumns(f, *nearStmt->declTy.agent);
      tdim = nearStmt->declTy.agent->position()->type.vecLen();
    }
  }
  const std::string trows = tdim == 2 ? "3" : "9";
  curStepList = stepListEligible(si) && nearStmt && nearStmt->e[0]->kids[0]->kind == Expr::Var &&
                nearStmt->e[0]->kids[0]->sym == p.sym;
  if (curStepList) listSteps.insert(&f);
  curStepFlat = config.getBool("cuda.flat", true) && curStepHasLimit && nearStmt && nearStmt->declTy.agent &&
                nearStmt->declTy.agent->position() && nearStmt->declTy.agent->position()->type.vecLen() == 2;

  curStepBulk = curStepTile && curStepFlat && tdim == 2 && config.getBool("cuda.bulk", true) &&
                tcols.size() <= 8;   // ABL_BTILE_MAX_COLS
  curStepDense = curStepHasLimit && nearStmt && !useFloat && config.getBool("cuda.dense", true);
  if (curStepDense) denseSteps.insert(&f);
  curStepSplit = curStepDense && curStepTile && (tdim == 2 || tdim == 3) && config.getBool("cuda.split", true);

  // the user's step function
  w << "template <int ABL_MODE>"; w.nl();
  sqLimits.clear();
  const bool sql = sqcmpOn() && curStepHasLimit;   // comparison bounds on the squared distance travel as a kernel parameter
  w << "__device__ __forceinline__ void " << f.emitName << "(abl_ctx& _ctx, const abl_step_launch& _a, unsigned _i, const abl_real _near_limit, const abl_real _near_cull, "
    << (sql ? "const abl_sq_limits& _sql, " : "") << "const unsigned _tile_cap, const bool _tile_ok" << (curStepBulk ? " ABL_BT_PARAM" : "") << ", const "
    << self.name << "& " << p.name << ", " << self.name << "& " << p.outName << ") {";
  w.indent(); stmts(f.body); w.outdent();
  w.nl();
  w << "}"; w.nl(); w.nl();

  StepKernelCtx ctx{si, radius, tcols, tdim, trows, sql};
  stepKernelWrapper(ctx);
  if (curStepSplit) stepPrefilterKernel(ctx);
  stepLauncher(ctx);
  (void)index;
  curFn = nullptr;
  curStep = nullptr;
  curStepHasLimit = false;
  curStepTile = false;
  curStepFlat = false;
  curStepList = false;
  curStepBulk = false;
  curStepDense = false;
  curStepSplit = false;
}

std::string CudaPrinter::kernelSource() {
  target = Target::Device;
  w = Writer();
  anon = 0;
  analyseSteps();
  w << "// Generated by OpenABL (cuda backend, sm_100a). Device program."; w.nl();
  if (useFloat) { w << "#ifndef ABL_USE_FLOAT"; w.nl(); w << "#define ABL_USE_FLOAT 1"; w.nl(); w << "#endif"; w.nl(); }
  w << "#include <cuda_runtime.h>"; w.nl();
  w << "#include \"abl_device.cuh\""; w.nl(); w.nl();

  for (AgentDecl *a : script.agents) {
    w << "struct " << a->name << " {";
    w.indent();
    for (auto &m : a->members) { w.nl(); w << typeName(m->type) << " " << m->name << ";"; }
    w.outdent(); w.nl();
    w << "};"; w.nl();
  }
  w.nl();
  for (ConstDecl *c : script.consts) {
    if (c->type.isString()) continue;
    constDecl(*c);
    w.nl();
  }
  w.nl();

  // device functions reachable from step functions, in declaration order
  std::set<const FuncDecl *> used;
  for (const StepInfo &si : steps) reachable(si.fn, used);
  for (FuncDecl *f : script.funcs) {
    if (f->isStep() || f->isSeqStep() || f->isMain() || !used.count(f)) continue;
    function(*f);
    w.nl();
  }
  w.nl();
  for (size_t i = 0; i < steps.size(); i++) stepKernel(steps[i], (int)i);

  // registration with the runtime
  EnvDecl *env = script.env;
  w << "struct abl_host_type_ { abl_agent_desc desc; void *agents; int pool; };"; w.nl();
  w << "extern \"C\" abl_host_type_ abl_model_types[];"; w.nl();
  w << "static int abl_model_step_ids[" << (steps.empty() ? 1 : steps.size()) << "];"; w.nl();
  w << "extern \"C\" const int abl_model_n_steps = " << steps.size() << ";"; w.nl();
  w << "extern \"C\" const char *abl_model_step_name(int s) {"; w.nl();
  w << "    static const char *names[] = {";
  for (const StepInfo &si : steps) w << " \"" << si.fn->name << "\",";
  w << " \"\" };"; w.nl();
  w << "    return names[s];"; w.nl();
  w << "}"; w.nl(); w.nl();

  w << "extern \"C\" int abl_model_step_pool(int s) {"; w.nl();
  w << "    static const int types[] = {";
  for (const StepInfo &si : steps) w << " " << agentIndex(si.self) << ",";
  w << " -1 };"; w.nl();
  w << "    return abl_model_types[types[s]].pool;"; w.nl();
  w << "}"; w.nl(); w.nl();

  w << "/* bit 0: the step function calls removeCurrent(), bit 1: it adds agents at run time */"; w.nl();
  w << "extern \"C\" int abl_model_step_flags(int s) {"; w.nl();
  w << "    static const int flags[] = {";
  for (const StepInfo &si : steps) w << " " << ((si.fn->usesRemoval ? 1 : 0) | (si.fn->addedAgent ? 2 : 0)) << ",";
  w << " 0 };"; w.nl();
  w << "    return flags[s];"; w.nl();
  w << "}"; w.nl(); w.nl();

  w << "/* ABL_MODE of the kernel the most recent launch of step function s used (-1: not launched yet): 0 cursor"; w.nl();
  w << " * loop, 1 chunked, 2 shared-memory tile, 3 flat loop, 4 neighbour-list walk — after the tuning phase"; w.nl();
  w << " * this is the variant the launcher's run-time tuner kept */"; w.nl();
  w << "extern \"C\" int abl_model_step_variant(int s) {"; w.nl();
  w << "    switch (s) {"; w.nl();
  for (size_t i = 0; i < steps.size(); i++) {
    w << "    case " << i << ": return abl_last_mode_" << steps[i].fn->emitName << ";"; w.nl();
  }
  w << "    default: return -1;"; w.nl();
  w << "    }"; w.nl();
  w << "}"; w.nl(); w.nl();

  w << "extern \"C\" int abl_model_setup(abl_runtime *rt) {"; w.nl();
  w << "    int rc;"; w.nl();
  if (env && env->dim > 0) {
    char buf[512];
    snprintf(buf, sizeof buf, "    const double env_min[3] = { %.17g, %.17g, %.17g };", env->envMin.v[0],
             env->envMin.v[1], env->dim == 3 ? env->envMin.v[2] : 0.0);
    w << buf; w.nl();
    snprintf(buf, sizeof buf, "    const double env_max[3] = { %.17g, %.17g, %.17g };", env->envMax.v[0],
             env->envMax.v[1], env->dim == 3 ? env->envMax.v[2] : 0.0);
    w << buf; w.nl();
    snprintf(buf, sizeof buf, "    if ((rc = abl_cuda_set_environment(rt, %d, env_min, env_max, %.17g))) return rc;",
             env->dim, env->granularity.valid() ? env->granularity.num() : 1.0);
    w << buf; w.nl();
  }
  w << "    for (int t = 0; abl_model_types[t].desc.name; t++)"; w.nl();
  w << "        if ((rc = abl_cuda_add_pool(rt, &abl_model_types[t].desc, &abl_model_types[t].pool))) return rc;"; w.nl();
  for (size_t i = 0; i < steps.size(); i++) {
    const StepInfo &si = steps[i];
    FuncDecl &f = *si.fn;
    unsigned mask = 0;
    for (size_t m = 0; m < si.self->members.size(); m++)
      if (si.writes.count(si.self->members[m]->name)) mask |= 1u << m;
    double radius = f.nearRadius.valid() && f.nearRadius.isNum() ? f.nearRadius.num() : 0.0;
    if (f.nearAgent && !(f.nearRadius.valid() && f.nearRadius.isNum())) {
      // dynamic radius: cover it with the cell size (it must not exceed the granularity)
      radius = env && env->granularity.valid() ? env->granularity.num() : 0.0;
    }
    char buf[256];
    w << "    {"; w.nl();
    w << "        abl_step_desc d;"; w.nl();
    w << "        d.name = \"" << f.name << "\";"; w.nl();
    w << "        d.self_pool = abl_model_types[" << agentIndex(si.self) << "].pool;"; w.nl();
    if (f.nearAgent) w << "        d.nbr_pool = abl_model_types[" << agentIndex(f.nearAgent) << "].pool;";
    else w << "        d.nbr_pool = -1;";
    w.nl();
    snprintf(buf, sizeof buf, "        d.radius = %.17g;", radius);
    w << buf; w.nl();
    w << "        d.written_members = " << mask << "u;"; w.nl();
    w << "        d.uses_removal = " << (f.usesRemoval ? 1 : 0) << ";"; w.nl();
    if (f.addedAgent) w << "        d.added_pool = abl_model_types[" << agentIndex(f.addedAgent) << "].pool;";
    else w << "        d.added_pool = -1;";
    w.nl();
    w << "        d.launch = abl_launch_" << f.emitName << ";"; w.nl();
    w << "        d.nlist = " << (listSteps.count(&f) ? 1 : 0) << ";"; w.nl();
    w << "        d.shadow = " << (denseSteps.count(&f) ? 1 : 0) << ";"; w.nl();
    w << "        if ((rc = abl_cuda_register_step(rt, &d, &abl_model_step_ids[" << i << "]))) return rc;"; w.nl();
    w << "    }"; w.nl();
  }
  w << "    return 0;"; w.nl();
  w << "}"; w.nl(); w.nl();

  w << "extern \"C\" int abl_model_run_step(abl_runtime *rt, int s) { return abl_cuda_step(rt, abl_model_step_ids[s]); }"; w.nl();
  w << "extern \"C\" int abl_model_parallel_steps(abl_runtime *rt) {"; w.nl();
  w << "    int rc;"; w.nl();
  w << "    for (int s = 0; s < abl_model_n_steps; s++)"; w.nl();
  w << "        if ((rc = abl_cuda_step(rt, abl_model_step_ids[s]))) return rc;"; w.nl();
  w << "    return 0;"; w.nl();
  w << "}"; w.nl();
  return w.str();
}

std::string buildScript(const BackendContext &ctx, bool useFloat) {
  std::string asset = ctx.assetDir + "/cuda";
  std::string def = useFloat ? " -DABL_USE_FLOAT=1" : "";
  std::ostringstream s;
  s << "#!/bin/sh\n"
       "# Generated by OpenABL (cuda backend). Builds libmodel.so (+ ./main) for NVIDIA B200.\n"
       "set -e\n"
       "cd \"$(dirname \"$0\")\"\n"
       "NVCC=${NVCC:-nvcc}\n"
       "CC=${CC:-gcc}\n"
       "ASSET=\"" << asset << "\"\n"
       "CUDA_LIB=${CUDA_LIB:-/usr/local/cuda/lib64}\n"
       "ARCH=\"-gencode arch=compute_100a,code=sm_100a\"\n"
       "# -fmad=false: the reference C path has no FMA contraction; results must match it\n"
       "$NVCC $ARCH -O3 -lineinfo -fmad=false -std=c++17 -diag-suppress 177" << def << " -Xcompiler -fPIC -I. -c model_kernels.cu -o model_kernels.o\n"
       "$CC -O2 -std=c99" << def << " -fPIC -I. -c model_host.c -o model_host.o\n"
       "$CC -O2 -std=c99" << def << " -fPIC -I. -DABL_MODEL_NO_MAIN -c model_host.c -o model_host_lib.o\n"
       "$CC -O2 -std=c99" << def << " -fPIC -I. -c abl_host.c -o abl_host.o\n"
       "if [ -f \"$ASSET/libabl_cuda.so\" ]; then\n"
       "  RT_DIR=\"$ASSET\"\n"
       "else\n"
       "  $NVCC $ARCH -O3 -lineinfo -std=c++17 --cudart shared -Xcompiler -fPIC -I. -shared -o libabl_cuda.so \"$ASSET/abl_runtime.cu\" -lnccl\n"
       "  RT_DIR=\"$(pwd)\"\n"
       "fi\n"
       "$NVCC $ARCH --cudart shared -shared -o libmodel.so model_kernels.o model_host_lib.o abl_host.o -L\"$RT_DIR\" -labl_cuda -Xlinker -rpath -Xlinker \"$RT_DIR\" -Xlinker -rpath -Xlinker \"$CUDA_LIB\"\n"
       "$NVCC $ARCH --cudart shared -o main model_kernels.o model_host.o abl_host.o -L\"$RT_DIR\" -labl_cuda -Xlinker -rpath -Xlinker \"$RT_DIR\" -Xlinker -rpath -Xlinker \"$CUDA_LIB\" -lm\n";
  return s.str();
}

}  // namespace

void CudaBackend::generate(Script &script, const BackendContext &ctx) {
  if (!script.mainFunc) throw BackendError("cuda backend: script has no main function");
  bool useFloat = ctx.config.getBool("use_float", false);

  CudaPrinter printer(script, useFloat, ctx.config);
  std::string host = printer.hostSource();
  std::string kernels = printer.kernelSource();

  const std::string &out = ctx.outputDir;
  writeToFile(out + "/model_host.c", host);
  writeToFile(out + "/model_kernels.cu", kernels);

  std::string asset = ctx.assetDir + "/cuda";
  std::string abi = asset + "/abl_cuda.h";
  if (!fileExists(abi)) abi = ctx.assetDir + "/../include/abl_cuda.h";
  if (!fileExists(abi)) throw std::runtime_error("abl_cuda.h not found (looked in " + asset + " and " + ctx.assetDir + "/../include)");
  copyFile(abi, out + "/abl_cuda.h");
  copyFile(asset + "/abl_device.cuh", out + "/abl_device.cuh");
  copyFile(asset + "/abl_slab.cuh", out + "/abl_slab.cuh");
  copyFile(asset + "/abl_host.h", out + "/abl_host.h");
  copyFile(asset + "/abl_host.c", out + "/abl_host.c");
  writeToFile(out + "/build.sh", buildScript(ctx, useFloat));
  writeToFile(out + "/run.sh", "#!/bin/sh\ncd \"$(dirname \"$0\")\"\n./main\n");
  makeFileExecutable(out + "/build.sh");
  makeFileExecutable(out + "/run.sh");
}

}  // namespace abl
