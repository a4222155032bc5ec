// Parser shim for building the UNMODIFIED reference compiler without flex/bison.
//
// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  The reference's 24 hand-written
// .cpp files are compiled where they lie under /root/reference/src; the three
// generated artefacts they need (location.hh, Parser.hpp, and the parser TU) are
// replaced by this directory.  This TU parses the model with this repository's own
// hand-written parser (src/Parser.cpp) and rebuilds the result as a reference AST
// through the reference's public node constructors, i.e. exactly what the actions of
// reference src/Parser.y:162-364 would have constructed.  Everything downstream —
// analysis, constant folding, CPrinter, libabl — is then genuine reference code.
#include <cstdio>
#include <iostream>
#include <string>

#include "ParserContext.hpp"   // reference
#include "../../src/Parser.hpp"  // ours

using namespace OpenABL;
namespace R = OpenABL::AST;

namespace {

R::Location L(int line) { return R::Location((unsigned)line); }

R::UnaryOp unop(abl::Op op) {
  switch (op) {
    case abl::Op::Neg: return R::UnaryOp::MINUS;
    case abl::Op::Pos: return R::UnaryOp::PLUS;
    case abl::Op::Not: return R::UnaryOp::LOGICAL_NOT;
    default: return R::UnaryOp::BITWISE_NOT;
  }
}

R::BinaryOp binop(abl::Op op) {
  switch (op) {
    case abl::Op::Add: return R::BinaryOp::ADD;
    case abl::Op::Sub: return R::BinaryOp::SUB;
    case abl::Op::Mul: return R::BinaryOp::MUL;
    case abl::Op::Div: return R::BinaryOp::DIV;
    case abl::Op::Mod: return R::BinaryOp::MOD;
    case abl::Op::BitAnd: return R::BinaryOp::BITWISE_AND;
    case abl::Op::BitXor: return R::BinaryOp::BITWISE_XOR;
    case abl::Op::BitOr: return R::BinaryOp::BITWISE_OR;
    case abl::Op::Shl: return R::BinaryOp::SHIFT_LEFT;
    case abl::Op::Shr: return R::BinaryOp::SHIFT_RIGHT;
    case abl::Op::Eq: return R::BinaryOp::EQUALS;
    case abl::Op::Ne: return R::BinaryOp::NOT_EQUALS;
    case abl::Op::Lt: return R::BinaryOp::SMALLER;
    case abl::Op::Le: return R::BinaryOp::SMALLER_EQUALS;
    case abl::Op::Gt: return R::BinaryOp::GREATER;
    case abl::Op::Ge: return R::BinaryOp::GREATER_EQUALS;
    case abl::Op::And: return R::BinaryOp::LOGICAL_AND;
    case abl::Op::Or: return R::BinaryOp::LOGICAL_OR;
    default: return R::BinaryOp::RANGE;
  }
}

R::Expression *conv(const abl::Expr &e);

R::MemberInitList *convInits(const abl::Expr &e) {
  auto *list = new R::MemberInitList();
  for (size_t i = 0; i < e.kids.size(); i++)
    list->emplace_back(new R::MemberInitEntry(e.initNames[i], conv(*e.kids[i]), L(e.initLines[i])));
  return list;
}

R::ExpressionList *convList(const std::vector<abl::ExprP> &v) {
  auto *list = new R::ExpressionList();
  for (const auto &k : v) list->emplace_back(conv(*k));
  return list;
}

R::Expression *conv(const abl::Expr &e) {
  using E = abl::Expr;
  switch (e.kind) {
    case E::BoolLit: return new R::BoolLiteral(e.bval, L(e.line));
    case E::IntLit: return new R::IntLiteral(e.ival, L(e.line));
    case E::FloatLit: return new R::FloatLiteral(e.fval, L(e.line));
    case E::StrLit: return new R::StringLiteral(e.name, L(e.line));
    case E::Var: return new R::VarExpression(new R::Var(e.name, L(e.line)), L(e.line));
    case E::Unary: return new R::UnaryOpExpression(unop(e.op), conv(*e.kids[0]), L(e.line));
    case E::Binary:
      return new R::BinaryOpExpression(binop(e.op), conv(*e.kids[0]), conv(*e.kids[1]), L(e.line));
    case E::Call: return new R::CallExpression(e.name, convList(e.kids), L(e.line));
    case E::Member: return new R::MemberAccessExpression(conv(*e.kids[0]), e.name, L(e.line));
    case E::EnvAccess: return new R::EnvironmentAccessExpression(e.name, L(e.line));
    case E::Index:
      return new R::ArrayAccessExpression(conv(*e.kids[0]), conv(*e.kids[1]), L(e.line));
    case E::Ternary:
      return new R::TernaryExpression(conv(*e.kids[0]), conv(*e.kids[1]), conv(*e.kids[2]), L(e.line));
    case E::AgentCreate: return new R::AgentCreationExpression(e.name, convInits(e), L(e.line));
    case E::ArrayInit: return new R::ArrayInitExpression(convList(e.kids), L(e.line));
    case E::NewArray:
      return new R::NewArrayExpression(new R::SimpleType(e.name, L(e.line)), conv(*e.kids[0]), L(e.line));
  }
  return nullptr;
}

R::Statement *conv(const abl::Stmt &s);

R::StatementList *convStmts(const std::vector<abl::StmtP> &v) {
  auto *list = new R::StatementList();
  for (const auto &k : v) list->emplace_back(conv(*k));
  return list;
}

R::Statement *conv(const abl::Stmt &s) {
  using S = abl::Stmt;
  switch (s.kind) {
    case S::ExprS: return new R::ExpressionStatement(conv(*s.e[0]), L(s.line));
    case S::Block: return new R::BlockStatement(convStmts(s.body), L(s.line));
    case S::VarDecl:
      return new R::VarDeclarationStatement(new R::SimpleType(s.typeName, L(s.typeLine)),
                                            new R::Var(s.varName, L(s.varLine)),
                                            s.e.empty() ? nullptr : conv(*s.e[0]), L(s.line));
    case S::If:
      return new R::IfStatement(conv(*s.e[0]), conv(*s.body[0]),
                                s.body.size() > 1 ? conv(*s.body[1]) : nullptr, L(s.line));
    case S::While: return new R::WhileStatement(conv(*s.e[0]), conv(*s.body[0]), L(s.line));
    case S::For:
      return new R::ForStatement(new R::SimpleType(s.typeName, L(s.typeLine)),
                                 new R::Var(s.varName, L(s.varLine)), conv(*s.e[0]),
                                 conv(*s.body[0]), L(s.line));
    case S::Return: return new R::ReturnStatement(s.e.empty() ? nullptr : conv(*s.e[0]), L(s.line));
    case S::Break: return new R::BreakStatement(L(s.line));
    case S::Continue: return new R::ContinueStatement(L(s.line));
    case S::Simulate:
      return new R::SimulateStatement(conv(*s.e[0]), new R::IdentList(s.stepNames), L(s.line));
    case S::Assign: return new R::AssignStatement(conv(*s.e[0]), conv(*s.e[1]), L(s.line));
    case S::AssignOp:
      return new R::AssignOpStatement(binop(s.op), conv(*s.e[0]), conv(*s.e[1]), L(s.line));
  }
  return nullptr;
}

R::Declaration *conv(const abl::Decl &d) {
  switch (d.kind) {
    case abl::Decl::Agent: {
      auto *members = new R::AgentMemberList();
      for (const auto &m : d.agent->members)
        members->emplace_back(new R::AgentMember(
            m->isPosition, new R::SimpleType(m->typeName, L(m->typeLine)), m->name, L(m->line)));
      return new R::AgentDeclaration(d.agent->name, members, L(d.agent->line));
    }
    case abl::Decl::Func: {
      const abl::FuncDecl &f = *d.func;
      auto *params = new R::ParamList();
      for (const abl::Param &p : f.params)
        params->emplace_back(new R::Param(
            new R::SimpleType(p.typeName, L(p.typeLine)), new R::Var(p.name, L(p.nameLine)),
            p.outName.empty() ? nullptr : new R::Var(p.outName, L(p.outLine)), L(p.line)));
      auto kind = f.kind == abl::FuncDecl::Step ? R::FunctionDeclaration::STEP
                : f.kind == abl::FuncDecl::SeqStep ? R::FunctionDeclaration::SEQ_STEP
                                                   : R::FunctionDeclaration::NORMAL;
      return new R::FunctionDeclaration(new R::SimpleType(f.retTypeName, L(f.retLine)), f.name,
                                        params, convStmts(f.body), kind, L(f.line));
    }
    case abl::Decl::ConstD: {
      const abl::ConstDecl &c = *d.cnst;
      return new R::ConstDeclaration(new R::SimpleType(c.typeName, L(c.typeLine)),
                                     new R::Var(c.name, L(c.nameLine)), conv(*c.init), c.isArray,
                                     c.isParam, L(c.line));
    }
    case abl::Decl::Env: {
      const abl::EnvDecl &e = *d.env;
      auto *list = new R::MemberInitList();
      for (size_t i = 0; i < e.names.size(); i++)
        list->emplace_back(new R::MemberInitEntry(e.names[i], conv(*e.values[i]), L(e.lines[i])));
      return new R::EnvironmentDeclaration(list, L(e.line));
    }
  }
  return nullptr;
}

}  // namespace

namespace OpenABL {

int Parser::parse() {
  std::string text;
  char buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, ctx.file)) > 0) text.append(buf, n);
  abl::ParseError perr;
  auto script = abl::parseScript(text, perr);
  if (!script) {
    std::cerr << "Parse error: " << perr.msg << " on line " << perr.line << std::endl;
    return 1;
  }
  auto *decls = new R::DeclarationList();
  for (const abl::Decl &d : script->decls) decls->emplace_back(conv(d));
  ctx.script = new R::Script(decls, L(1));
  return 0;
}

void ParserContext::initLexer() {}

}  // namespace OpenABL
