#include "Parser.hpp"

#include <cstdlib>
#include <cstring>

namespace abl {

namespace {

enum class T : uint8_t {
  End, Ident, Int, Float, Str, Bool,
  // keywords
  KwAgent, KwBreak, KwContinue, KwElse, KwEnvironment, KwIf, KwFor, KwNew, KwParam,
  KwPosition, KwReturn, KwSequential, KwSimulate, KwStep, KwWhile,
  // punctuation / operators
  Plus, Minus, Star, Slash, Percent, Amp, Caret, Pipe, Assign, Bang, Tilde, Question,
  Dot, Comma, Colon, Semi, Lt, Gt, LParen, RParen, LBracket, RBracket, LBrace, RBrace,
  DotDot, Arrow, EqEq, NotEq, Le, Ge, Shl, Shr, AndAnd, OrOr,
  PlusEq, MinusEq, StarEq, SlashEq, PercentEq, AmpEq, CaretEq, PipeEq, ShlEq, ShrEq,
  Bad,
};

struct Token {
  T kind = T::End;
  std::string text;
  long ival = 0;
  double fval = 0;
  bool bval = false;
  int line = 1;     // reported begin line (see Parser.hpp about the column-1 quirk)
  int endLine = 1;  // physical line the token ends on
};

struct Keyword { const char *word; T kind; };
const Keyword kKeywords[] = {
  {"agent", T::KwAgent}, {"break", T::KwBreak}, {"continue", T::KwContinue},
  {"else", T::KwElse}, {"environment", T::KwEnvironment}, {"if", T::KwIf},
  {"for", T::KwFor}, {"new", T::KwNew}, {"param", T::KwParam},
  {"position", T::KwPosition}, {"return", T::KwReturn}, {"sequential", T::KwSequential},
  {"simulate", T::KwSimulate}, {"step", T::KwStep}, {"while", T::KwWhile},
};

struct Punct { const char *text; T kind; };
// Longest first so that maximal munch works with a linear scan.
const Punct kPuncts[] = {
  {"<<=", T::ShlEq}, {">>=", T::ShrEq},
  {"..", T::DotDot}, {"->", T::Arrow}, {"==", T::EqEq}, {"!=", T::NotEq}, {"<=", T::Le},
  {">=", T::Ge}, {"<<", T::Shl}, {">>", T::Shr}, {"&&", T::AndAnd}, {"||", T::OrOr},
  {"+=", T::PlusEq}, {"-=", T::MinusEq}, {"*=", T::StarEq}, {"/=", T::SlashEq},
  {"%=", T::PercentEq}, {"&=", T::AmpEq}, {"^=", T::CaretEq}, {"|=", T::PipeEq},
  {"+", T::Plus}, {"-", T::Minus}, {"*", T::Star}, {"/", T::Slash}, {"%", T::Percent},
  {"&", T::Amp}, {"^", T::Caret}, {"|", T::Pipe}, {"=", T::Assign}, {"!", T::Bang},
  {"~", T::Tilde}, {"?", T::Question}, {".", T::Dot}, {",", T::Comma}, {":", T::Colon},
  {";", T::Semi}, {"<", T::Lt}, {">", T::Gt}, {"(", T::LParen}, {")", T::RParen},
  {"[", T::LBracket}, {"]", T::RBracket}, {"{", T::LBrace}, {"}", T::RBrace},
};

bool isIdStart(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_'; }
bool isDigit(char c) { return c >= '0' && c <= '9'; }
bool isHex(char c) { return isDigit(c) || (c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F'); }

class Lexer {
public:
  explicit Lexer(const std::string &src) : s(src) {}

  Token next() {
    Token t;
    int begin = line;  // location start := end of previous token
    for (;;) {
      char c = peek(0);
      if (c == ' ' || c == '\t' || c == '\r') {
        while (peek(0) == ' ' || peek(0) == '\t' || peek(0) == '\r') p++;
        begin = line;
      } else if (c == '\n') {
        line++; p++;           // newlines move the end, not the start
      } else if (c == '/' && peek(1) == '/') {
        while (p < s.size() && s[p] != '\n') p++;
      } else if (c == '/' && peek(1) == '*') {
        p += 2;
        while (p < s.size() && !(s[p] == '*' && peek(1) == '/')) {
          if (s[p] == '\n') line++;
          p++;
        }
        if (p < s.size()) p += 2;
        begin = line;
      } else {
        break;
      }
    }
    t.line = begin;
    t.endLine = line;
    if (p >= s.size()) { t.kind = T::End; return t; }

    char c = s[p];
    if (isIdStart(c)) {
      size_t b = p;
      while (p < s.size() && (isIdStart(s[p]) || isDigit(s[p]))) p++;
      t.text = s.substr(b, p - b);
      t.kind = T::Ident;
      if (t.text == "true" || t.text == "false") {
        t.kind = T::Bool; t.bval = t.text == "true";
        return t;
      }
      for (const Keyword &k : kKeywords)
        if (t.text == k.word) { t.kind = k.kind; break; }
      return t;
    }
    if (isDigit(c) || (c == '.' && isDigit(peek(1)))) return number(t);
    if (c == '"') return string(t);
    for (const Punct &pu : kPuncts) {
      size_t n = strlen(pu.text);
      if (s.compare(p, n, pu.text) == 0) {
        p += n; t.kind = pu.kind; t.text = pu.text;
        return t;
      }
    }
    t.kind = T::Bad; t.text = std::string(1, c); p++;
    return t;
  }

private:
  char peek(size_t o) const { return p + o < s.size() ? s[p + o] : '\0'; }

  Token number(Token t) {
    size_t b = p;
    if (s[p] == '0' && (peek(1) == 'x' || peek(1) == 'X') && isHex(peek(2))) {
      p += 2;
      while (isHex(peek(0))) p++;
      t.kind = T::Int; t.text = s.substr(b, p - b);
      t.ival = strtol(t.text.c_str(), nullptr, 16);
      return t;
    }
    while (isDigit(peek(0))) p++;
    bool isFloat = false;
    if (peek(0) == '.' && peek(1) == '.' && p > b) {
      // "0..n" is the integer 0 followed by the range operator
    } else if (peek(0) == '.') {
      isFloat = true; p++;
      while (isDigit(peek(0))) p++;
    }
    if (isFloat && (peek(0) == 'e' || peek(0) == 'E')) {
      size_t q = p + 1;
      if (q < s.size() && (s[q] == '+' || s[q] == '-')) q++;
      if (q < s.size() && isDigit(s[q])) {
        while (q < s.size() && isDigit(s[q])) q++;
        p = q;
      }
    }
    t.text = s.substr(b, p - b);
    if (isFloat) { t.kind = T::Float; t.fval = strtod(t.text.c_str(), nullptr); }
    else { t.kind = T::Int; t.ival = strtol(t.text.c_str(), nullptr, 10); }
    return t;
  }

  Token string(Token t) {
    p++;  // opening quote
    std::string out;
    while (p < s.size() && s[p] != '"') {
      if (s[p] == '\\' && p + 1 < s.size()) { out.push_back(s[p + 1]); p += 2; }
      else out.push_back(s[p++]);
    }
    if (p < s.size()) p++;
    t.kind = T::Str; t.text = out;
    return t;
  }

  const std::string &s;
  size_t p = 0;
  int line = 1;
};

struct Bail {};  // thrown on the first syntax error

// Binary operator table: token -> (op, precedence level); higher binds tighter.
struct BinInfo { T tok; Op op; int prec; };
const BinInfo kBinary[] = {
  {T::OrOr, Op::Or, 1}, {T::AndAnd, Op::And, 2}, {T::Pipe, Op::BitOr, 3},
  {T::Caret, Op::BitXor, 4}, {T::Amp, Op::BitAnd, 5}, {T::EqEq, Op::Eq, 6},
  {T::NotEq, Op::Ne, 6}, {T::Lt, Op::Lt, 7}, {T::Le, Op::Le, 7}, {T::Gt, Op::Gt, 7},
  {T::Ge, Op::Ge, 7}, {T::DotDot, Op::Range, 8}, {T::Shl, Op::Shl, 9}, {T::Shr, Op::Shr, 9},
  {T::Plus, Op::Add, 10}, {T::Minus, Op::Sub, 10}, {T::Star, Op::Mul, 11},
  {T::Slash, Op::Div, 11}, {T::Percent, Op::Mod, 11},
};

struct AssignInfo { T tok; Op op; };
const AssignInfo kAssignOps[] = {
  {T::PlusEq, Op::Add}, {T::MinusEq, Op::Sub}, {T::StarEq, Op::Mul}, {T::SlashEq, Op::Div},
  {T::PercentEq, Op::Mod}, {T::AmpEq, Op::BitAnd}, {T::CaretEq, Op::BitXor},
  {T::PipeEq, Op::BitOr}, {T::ShlEq, Op::Shl}, {T::ShrEq, Op::Shr},
};

class Parser {
public:
  Parser(const std::string &src, ParseError &err) : lex(src), err(err) {
    cur = lex.next();
    ahead = lex.next();
  }

  std::unique_ptr<Script> run() {
    auto script = std::unique_ptr<Script>(new Script);
    script->line = 1;
    try {
      while (cur.kind != T::End) script->decls.push_back(declaration());
    } catch (const Bail &) {
      return nullptr;
    }
    return script;
  }

private:
  Lexer lex;
  ParseError &err;
  Token cur, ahead;
  int prevEndLine = 1;

  void advance() {
    prevEndLine = cur.endLine;
    cur = ahead;
    ahead = lex.next();
  }
  [[noreturn]] void fail(const std::string &what) {
    std::string got = cur.kind == T::End ? "end of file" : "\"" + cur.text + "\"";
    err.msg = "syntax error, unexpected " + got + (what.empty() ? "" : ", expecting " + what);
    err.line = cur.line;
    throw Bail{};
  }
  bool at(T k) const { return cur.kind == k; }
  bool accept(T k) { if (at(k)) { advance(); return true; } return false; }
  void expect(T k, const char *what) { if (!accept(k)) fail(what); }
  std::string ident(int *line = nullptr) {
    if (!at(T::Ident)) fail("identifier");
    std::string s = cur.text;
    if (line) *line = cur.line;
    advance();
    return s;
  }

  // ---- declarations -------------------------------------------------------
  Decl declaration() {
    Decl d;
    if (at(T::KwAgent)) { d.kind = Decl::Agent; d.agent = agentDecl(); }
    else if (at(T::KwEnvironment)) { d.kind = Decl::Env; d.env = envDecl(); }
    else if (at(T::KwStep) || at(T::KwSequential)) { d.kind = Decl::Func; d.func = stepDecl(); }
    else if (at(T::KwParam)) { d.kind = Decl::ConstD; d.cnst = constDecl(); }
    else if (at(T::Ident) && ahead.kind == T::Ident) {
      // `type name (` is a function, anything else a global constant
      Token save = cur;
      (void)save;
      d = typedDecl();
    } else {
      fail("declaration");
    }
    return d;
  }

  std::unique_ptr<AgentDecl> agentDecl() {
    auto a = std::unique_ptr<AgentDecl>(new AgentDecl);
    a->line = cur.line;
    advance();
    a->name = ident();
    expect(T::LBrace, "\"{\"");
    while (!at(T::RBrace)) {
      auto m = std::unique_ptr<AgentMember>(new AgentMember);
      // An absent `position` keyword is an empty production: its location is the
      // end of whatever precedes it.
      m->line = at(T::KwPosition) ? cur.line : prevEndLine;
      m->isPosition = accept(T::KwPosition);
      m->typeName = ident(&m->typeLine);
      m->name = ident();
      expect(T::Semi, "\";\"");
      a->members.push_back(std::move(m));
    }
    advance();
    return a;
  }

  std::unique_ptr<EnvDecl> envDecl() {
    auto e = std::unique_ptr<EnvDecl>(new EnvDecl);
    e->line = cur.line;
    advance();
    expect(T::LBrace, "\"{\"");
    while (!at(T::RBrace)) {
      int l;
      e->names.push_back(ident(&l));
      e->lines.push_back(l);
      expect(T::Colon, "\":\"");
      e->values.push_back(expression());
      if (!accept(T::Comma)) break;
    }
    expect(T::RBrace, "\"}\"");
    return e;
  }

  void paramList(FuncDecl &f) {
    expect(T::LParen, "\"(\"");
    if (!at(T::RParen)) {
      do {
        Param p;
        p.line = cur.line;
        p.typeName = ident(&p.typeLine);
        p.name = ident(&p.nameLine);
        if (accept(T::Arrow)) p.outName = ident(&p.outLine);
        f.params.push_back(std::move(p));
      } while (accept(T::Comma));
    }
    expect(T::RParen, "\")\"");
  }

  void funcBody(FuncDecl &f) {
    expect(T::LBrace, "\"{\"");
    while (!at(T::RBrace)) f.body.push_back(statement());
    advance();
  }

  std::unique_ptr<FuncDecl> stepDecl() {
    auto f = std::unique_ptr<FuncDecl>(new FuncDecl);
    f->line = cur.line;
    if (accept(T::KwSequential)) {
      expect(T::KwStep, "\"step\"");
      f->kind = FuncDecl::SeqStep;
    } else {
      advance();
      f->kind = FuncDecl::Step;
    }
    f->retTypeName = "void";
    f->retLine = 1;
    f->name = ident();
    paramList(*f);
    funcBody(*f);
    return f;
  }

  Decl typedDecl() {
    Decl d;
    int line = cur.line, typeLine = cur.line;
    std::string typeName = ident();
    if (ahead.kind == T::LParen) {
      auto f = std::unique_ptr<FuncDecl>(new FuncDecl);
      f->line = line; f->retLine = typeLine; f->retTypeName = typeName;
      f->name = ident();
      paramList(*f);
      funcBody(*f);
      d.kind = Decl::Func; d.func = std::move(f);
    } else {
      d.kind = Decl::ConstD;
      d.cnst = constTail(line, typeLine, typeName, false);
    }
    return d;
  }

  std::unique_ptr<ConstDecl> constDecl() {
    int line = cur.line;
    advance();  // param
    int typeLine;
    std::string typeName = ident(&typeLine);
    return constTail(line, typeLine, typeName, true);
  }

  std::unique_ptr<ConstDecl> constTail(int line, int typeLine, const std::string &typeName,
                                       bool isParam) {
    auto c = std::unique_ptr<ConstDecl>(new ConstDecl);
    c->line = line; c->typeLine = typeLine; c->typeName = typeName; c->isParam = isParam;
    c->name = ident(&c->nameLine);
    if (accept(T::LBracket)) { expect(T::RBracket, "\"]\""); c->isArray = true; }
    expect(T::Assign, "\"=\"");
    if (at(T::LBrace)) {
      ExprP arr(new Expr(Expr::ArrayInit, cur.line));
      advance();
      do {
        if (at(T::RBrace)) break;  // trailing comma
        arr->kids.push_back(expression());
      } while (accept(T::Comma));
      if (arr->kids.empty()) fail("expression");
      expect(T::RBrace, "\"}\"");
      c->init = std::move(arr);
    } else {
      c->init = expression();
    }
    expect(T::Semi, "\";\"");
    return c;
  }

  // ---- statements ---------------------------------------------------------
  StmtP statement() {
    int line = cur.line;
    switch (cur.kind) {
      case T::LBrace: {
        StmtP s(new Stmt(Stmt::Block, line));
        advance();
        while (!at(T::RBrace)) s->body.push_back(statement());
        advance();
        return s;
      }
      case T::KwIf: {
        StmtP s(new Stmt(Stmt::If, line));
        advance();
        expect(T::LParen, "\"(\"");
        s->e.push_back(expression());
        expect(T::RParen, "\")\"");
        s->body.push_back(statement());
        if (accept(T::KwElse)) s->body.push_back(statement());
        return s;
      }
      case T::KwWhile: {
        StmtP s(new Stmt(Stmt::While, line));
        advance();
        expect(T::LParen, "\"(\"");
        s->e.push_back(expression());
        expect(T::RParen, "\")\"");
        s->body.push_back(statement());
        return s;
      }
      case T::KwFor: {
        StmtP s(new Stmt(Stmt::For, line));
        advance();
        expect(T::LParen, "\"(\"");
        s->typeName = ident(&s->typeLine);
        s->varName = ident(&s->varLine);
        expect(T::Colon, "\":\"");
        s->e.push_back(expression());
        expect(T::RParen, "\")\"");
        s->body.push_back(statement());
        return s;
      }
      case T::KwReturn: {
        StmtP s(new Stmt(Stmt::Return, line));
        advance();
        if (!at(T::Semi)) s->e.push_back(expression());
        expect(T::Semi, "\";\"");
        return s;
      }
      case T::KwBreak: {
        advance(); expect(T::Semi, "\";\"");
        return StmtP(new Stmt(Stmt::Break, line));
      }
      case T::KwContinue: {
        advance(); expect(T::Semi, "\";\"");
        return StmtP(new Stmt(Stmt::Continue, line));
      }
      case T::KwSimulate: {
        StmtP s(new Stmt(Stmt::Simulate, line));
        advance();
        expect(T::LParen, "\"(\"");
        s->e.push_back(expression());
        expect(T::RParen, "\")\"");
        expect(T::LBrace, "\"{\"");
        s->stepNames.push_back(ident());
        while (accept(T::Comma)) {
          if (at(T::RBrace)) break;
          s->stepNames.push_back(ident());
        }
        expect(T::RBrace, "\"}\"");
        return s;
      }
      default: break;
    }

    if (at(T::Ident) && ahead.kind == T::Ident) {
      StmtP s(new Stmt(Stmt::VarDecl, line));
      s->typeName = ident(&s->typeLine);
      s->varName = ident(&s->varLine);
      if (accept(T::Assign)) s->e.push_back(expression());
      expect(T::Semi, "\";\"");
      return s;
    }

    ExprP lhs = expression();
    int startLine = lastStart;
    if (accept(T::Assign)) {
      StmtP s(new Stmt(Stmt::Assign, startLine));
      s->e.push_back(std::move(lhs));
      s->e.push_back(expression());
      expect(T::Semi, "\";\"");
      return s;
    }
    for (const AssignInfo &ai : kAssignOps) {
      if (at(ai.tok)) {
        advance();
        StmtP s(new Stmt(Stmt::AssignOp, startLine));
        s->op = ai.op;
        s->e.push_back(std::move(lhs));
        s->e.push_back(expression());
        expect(T::Semi, "\";\"");
        return s;
      }
    }
    StmtP s(new Stmt(Stmt::ExprS, startLine));
    s->e.push_back(std::move(lhs));
    expect(T::Semi, "\";\"");
    return s;
  }

  // ---- expressions --------------------------------------------------------
  // `lastStart` is the line on which the most recently parsed expression
  // syntactically begins; it differs from node->line only for "(expr)".
  int lastStart = 1;

  ExprP expression() { return binary(0); }

  ExprP binary(int minPrec) {
    ExprP lhs = unary();
    int start = lastStart;
    for (;;) {
      const BinInfo *bi = nullptr;
      for (const BinInfo &b : kBinary) if (cur.kind == b.tok) { bi = &b; break; }
      if (bi && bi->prec >= minPrec && bi->prec >= 1) {
        advance();
        ExprP rhs = binary(bi->prec + 1);
        if (bi->op == Op::Range && at(T::DotDot)) fail("");  // `..` is non-associative
        ExprP n(new Expr(Expr::Binary, start));
        n->op = bi->op;
        n->kids.push_back(std::move(lhs));
        n->kids.push_back(std::move(rhs));
        lhs = std::move(n);
      } else if (at(T::Question) && minPrec <= 0) {
        advance();
        ExprP n(new Expr(Expr::Ternary, start));
        n->kids.push_back(std::move(lhs));
        n->kids.push_back(binary(0));
        expect(T::Colon, "\":\"");
        n->kids.push_back(binary(0));
        lhs = std::move(n);
      } else {
        break;
      }
    }
    lastStart = start;
    return lhs;
  }

  ExprP unary() {
    Op op;
    bool isUnary = true;
    switch (cur.kind) {
      case T::Bang: op = Op::Not; break;
      case T::Tilde: op = Op::BitNot; break;
      case T::Plus: op = Op::Pos; break;
      case T::Minus: op = Op::Neg; break;
      default: isUnary = false; op = Op::Pos; break;
    }
    if (!isUnary) return postfix();
    int line = cur.line;
    advance();
    ExprP n(new Expr(Expr::Unary, line));
    n->op = op;
    n->kids.push_back(unary());
    lastStart = line;
    return n;
  }

  ExprP postfix() {
    ExprP e = primary();
    int start = lastStart;
    for (;;) {
      if (accept(T::Dot)) {
        ExprP n(new Expr(Expr::Member, start));
        n->name = ident();
        n->kids.push_back(std::move(e));
        e = std::move(n);
      } else if (accept(T::LBracket)) {
        ExprP n(new Expr(Expr::Index, start));
        n->kids.push_back(std::move(e));
        n->kids.push_back(expression());
        expect(T::RBracket, "\"]\"");
        e = std::move(n);
      } else {
        break;
      }
    }
    lastStart = start;
    return e;
  }

  void memberInits(Expr &n) {
    while (!at(T::RBrace)) {
      int l;
      n.initNames.push_back(ident(&l));
      n.initLines.push_back(l);
      expect(T::Colon, "\":\"");
      n.kids.push_back(expression());
      if (!accept(T::Comma)) break;
    }
    expect(T::RBrace, "\"}\"");
  }

  ExprP primary() {
    int line = cur.line;
    ExprP n;
    switch (cur.kind) {
      case T::Bool: n.reset(new Expr(Expr::BoolLit, line)); n->bval = cur.bval; advance(); break;
      case T::Int: n.reset(new Expr(Expr::IntLit, line)); n->ival = cur.ival; advance(); break;
      case T::Float: n.reset(new Expr(Expr::FloatLit, line)); n->fval = cur.fval; advance(); break;
      case T::Str: n.reset(new Expr(Expr::StrLit, line)); n->name = cur.text; advance(); break;
      case T::LParen: {
        advance();
        n = expression();
        expect(T::RParen, "\")\"");
        break;
      }
      case T::KwEnvironment: {
        advance();
        expect(T::Dot, "\".\"");
        n.reset(new Expr(Expr::EnvAccess, line));
        n->name = ident();
        break;
      }
      case T::KwNew: {
        advance();
        n.reset(new Expr(Expr::NewArray, line));
        n->name = ident();
        expect(T::LBracket, "\"[\"");
        n->kids.push_back(expression());
        expect(T::RBracket, "\"]\"");
        break;
      }
      case T::Ident: {
        std::string name = cur.text;
        advance();
        if (accept(T::LParen)) {
          n.reset(new Expr(Expr::Call, line));
          n->name = name;
          if (!at(T::RParen)) {
            do {
              if (at(T::RParen)) break;  // trailing comma
              n->kids.push_back(expression());
            } while (accept(T::Comma));
          }
          expect(T::RParen, "\")\"");
        } else if (accept(T::LBrace)) {
          n.reset(new Expr(Expr::AgentCreate, line));
          n->name = name;
          memberInits(*n);
        } else {
          n.reset(new Expr(Expr::Var, line));
          n->name = name;
        }
        break;
      }
      default:
        fail("expression");
    }
    lastStart = line;
    return n;
  }
};

}  // namespace

std::unique_ptr<Script> parseScript(const std::string &text, ParseError &err) {
  Parser p(text, err);
  return p.run();
}

}  // namespace abl
