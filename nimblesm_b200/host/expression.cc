// nimblesm_b200/host/expression.cc — see expression.h.
#include "expression.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

namespace nimble_b200 {

struct Expression::Node
{
  enum Kind { CONST, VAR, ADD, SUB, MUL, DIV, MOD, POW, IPOW, NEG, FUNC, COND, LT, LE, GT, GE, EQ, AND, OR, XOR, NOT, BCONST };
  Kind                  kind = CONST;
  double                value = 0.0;
  int                   index = 0;  // VAR: 0..3 = x y z t; FUNC: function id; IPOW: exponent
  std::shared_ptr<Node> a, b, c;
};

namespace {

using NodeP = std::shared_ptr<Expression::Node>;
using Kind  = Expression::Node::Kind;

const char* const kFuncs[] = {"sin",  "cos",  "tan",  "erf",  "exp",   "log",   "abs",   "asin", "acos",
                              "atan", "sqrt", "cbrt", "erfc", "ceil", "round", "floor", "log10"};

double
call(int id, double v)
{
  switch (id) {
    case 0: return std::sin(v);
    case 1: return std::cos(v);
    case 2: return std::tan(v);
    case 3: return std::erf(v);
    case 4: return std::exp(v);
    case 5: return std::log(v);
    case 6: return std::fabs(v);
    case 7: return std::asin(v);
    case 8: return std::acos(v);
    case 9: return std::atan(v);
    case 10: return std::sqrt(v);
    case 11: return std::cbrt(v);
    case 12: return std::erfc(v);
    case 13: return std::ceil(v);
    case 14: return std::round(v);
    case 15: return std::floor(v);
    default: return std::log10(v);
  }
}

// a window into the (lower-cased) expression text
struct Span
{
  const char* p;
  int         n;
  std::string
  str() const
  {
    return std::string(p, (size_t)n);
  }
  Span
  sub(int from, int to) const
  {
    return Span{p + from, to - from};
  }
  bool
  is(const char* lit) const
  {
    return str() == lit;
  }
};

void
trim(Span& s)
{
  while (s.n > 1 && s.p[0] == ' ') ++s.p, --s.n;
  while (s.n > 1 && s.p[s.n - 1] == ' ') --s.n;
  if (s.n == 1 && s.p[0] == ' ') s.n = 0;
}

// index of the ')' matching the '(' at i, or -1
int
closing(const Span& s, int i)
{
  int depth = 0;
  for (int k = i; k < s.n; ++k) {
    if (s.p[k] == '(') ++depth;
    if (s.p[k] == ')' && --depth == 0) return k;
  }
  return -1;
}

void
format(Span& s)
{
  if (s.n == 0) return;
  trim(s);
  while (s.n >= 2 && s.p[0] == '(' && closing(s, 0) == s.n - 1) {  // redundant outer parentheses
    ++s.p;
    s.n -= 2;
  }
}

// last top-level occurrence of c scanning from the right (the reference's `r << c`), or -1
int
find_last(const Span& s, char c)
{
  int pos = s.n - 1;
  while (pos > 0 && s.p[pos] != c) {
    if (s.p[pos] == ')') {
      int depth = 1;
      while (depth != 0 && pos > 0) {
        char ch = s.p[--pos];
        if (ch == ')') ++depth;
        if (ch == '(') --depth;
      }
    } else {
      --pos;
    }
  }
  return (pos >= 0 && s.n > 0 && s.p[pos] == c) ? pos : -1;
}

// first top-level occurrence scanning from the left (`r >> c`), or -1
int
find_first(const Span& s, char c)
{
  int pos = 0;
  while (pos < s.n - 1 && s.p[pos] != c) {
    if (s.p[pos] == '(') {
      int k = closing(s, pos);
      pos   = k < 0 ? s.n - 1 : k;
      if (pos < s.n - 1 && s.p[pos] != c) ++pos;
    } else {
      ++pos;
    }
  }
  return (s.n > 0 && s.p[pos] == c) ? pos : -1;
}

bool
is_digit(char c)
{
  return c >= '0' && c <= '9';
}

bool
is_integer(const Span& s)
{
  if (s.n == 0) return false;
  if (s.n == 1 && (s.p[0] == '-' || s.p[0] == '+')) return false;
  if (s.p[0] != '-' && s.p[0] != '+' && !is_digit(s.p[0])) return false;
  for (int i = 1; i < s.n; ++i)
    if (!is_digit(s.p[i])) return false;
  return true;
}

bool
is_number(const Span& s)
{
  if (s.n == 0) return false;
  int  pos     = 0;
  bool decimal = false, digits = false;
  if (s.p[0] == '-' || s.p[0] == '+') ++pos;
  while (pos < s.n) {
    char c = s.p[pos++];
    if (is_digit(c)) {
      digits = true;
      continue;
    }
    if (c == '.' && !decimal) {
      decimal = true;
      continue;
    }
    if (c == 'e' && digits) return is_integer(s.sub(pos, s.n));
    return false;
  }
  return true;
}

NodeP
make(Kind k, NodeP a = nullptr, NodeP b = nullptr, NodeP c = nullptr)
{
  NodeP n = std::make_shared<Expression::Node>();
  n->kind = k, n->a = a, n->b = b, n->c = c;
  return n;
}

NodeP
parse(Span s);

NodeP
parse_bool(Span s)
{
  format(s);
  if (s.is("true") || s.is("false")) {
    NodeP n  = make(Kind::BCONST);
    n->value = s.is("true") ? 1.0 : 0.0;
    return n;
  }
  int k;
  if ((k = find_last(s, '|')) >= 0) return make(Kind::OR, parse_bool(s.sub(0, k)), parse_bool(s.sub(k + 1, s.n)));
  if ((k = find_last(s, '^')) >= 0) return make(Kind::XOR, parse_bool(s.sub(0, k)), parse_bool(s.sub(k + 1, s.n)));
  if ((k = find_last(s, '&')) >= 0) return make(Kind::AND, parse_bool(s.sub(0, k)), parse_bool(s.sub(k + 1, s.n)));
  if ((k = find_last(s, '!')) >= 0) return make(Kind::NOT, parse_bool(s.sub(1, s.n)));
  if ((k = find_last(s, '>')) >= 0) {
    if (k + 1 < s.n && s.p[k + 1] == '=') return make(Kind::GE, parse(s.sub(0, k)), parse(s.sub(k + 2, s.n)));
    return make(Kind::GT, parse(s.sub(0, k)), parse(s.sub(k + 1, s.n)));
  }
  if ((k = find_last(s, '<')) >= 0) {
    if (k + 1 < s.n && s.p[k + 1] == '=') return make(Kind::LE, parse(s.sub(0, k)), parse(s.sub(k + 2, s.n)));
    return make(Kind::LT, parse(s.sub(0, k)), parse(s.sub(k + 1, s.n)));
  }
  if ((k = find_last(s, '=')) >= 1 && s.p[k - 1] == '=') return make(Kind::EQ, parse(s.sub(k + 1, s.n)), parse(s.sub(0, k - 1)));
  throw std::invalid_argument("Unable to parse \"" + s.str() + "\" as boolean");
}

NodeP
parse_operation(const Span& s)
{
  int k;
  if ((k = find_first(s, '?')) >= 0) {
    // matching ':' of this '?' (nested ternaries count)
    int depth = 1, pos = k;
    while (depth != 0 && pos < s.n) {
      char c = s.p[++pos];
      if (c == '?') ++depth;
      if (c == ':') --depth;
    }
    if (depth > 0) throw std::invalid_argument("Couldn't find matching : for ternary operator in " + s.str());
    if (pos != s.n) return make(Kind::COND, parse_bool(s.sub(0, k)), parse(s.sub(k + 1, pos)), parse(s.sub(pos + 1, s.n)));
  }
  if ((k = find_last(s, '%')) >= 0) return make(Kind::MOD, parse(s.sub(0, k)), parse(s.sub(k + 1, s.n)));
  if ((k = find_last(s, '+')) >= 0) return make(Kind::ADD, parse(s.sub(0, k)), parse(s.sub(k + 1, s.n)));
  if ((k = find_last(s, '-')) >= 0) {
    // a subtraction, not a negation: look at the character before the run of '-' / blanks
    int pos = k;
    while (pos > 0 && (s.p[pos] == '-' || s.p[pos] == ' ')) --pos;
    if (pos > 0 || s.p[pos] != '-') {
      char c = s.p[pos];
      if (c != '*' && c != '/' && c != '^' && c != 'e' && c != '=' && c != '<' && c != '~' && c != '>') {
        while (s.p[pos] != '-') ++pos;
        return make(Kind::SUB, parse(s.sub(0, pos)), parse(s.sub(pos + 1, s.n)));
      }
    }
  }
  if ((k = find_last(s, '*')) >= 0) return make(Kind::MUL, parse(s.sub(0, k)), parse(s.sub(k + 1, s.n)));
  if ((k = find_last(s, '/')) >= 0) return make(Kind::DIV, parse(s.sub(0, k)), parse(s.sub(k + 1, s.n)));
  if ((k = find_first(s, '^')) >= 0) {
    NodeP base  = parse(s.sub(0, k));
    Span  power = s.sub(k + 1, s.n);
    format(power);
    if (is_integer(power)) {
      NodeP n  = make(Kind::IPOW, base);
      n->index = (int)std::strtol(power.str().c_str(), nullptr, 10);
      return n;
    }
    return make(Kind::POW, base, parse(power));
  }
  if (s.n > 0 && s.p[0] == '-') return make(Kind::NEG, parse(s.sub(1, s.n)));
  return nullptr;
}

NodeP
parse_function(const Span& s)
{
  char c = s.p[0];
  if (is_digit(c) || c == '(' || c == '-' || s.n < 3) return nullptr;
  int pos = 0;
  while (pos < s.n && s.p[pos] != '(' && s.p[pos] != ' ') {
    c = s.p[pos];
    if (c == '*' || c == '/' || c == '^' || c == '+' || c == '=' || c == '<' || c == '~' || c == '>') return nullptr;
    ++pos;
  }
  int open = find_first(s, '(');
  if (open < 0) return nullptr;
  int close = closing(s, open);
  if (close < 0) throw std::invalid_argument("Mismatched parethesis in " + s.str());
  if (close < s.n - 1) return nullptr;
  const std::string name = s.sub(0, open).str();
  for (size_t id = 0; id < sizeof(kFuncs) / sizeof(kFuncs[0]); ++id)
    if (name == kFuncs[id]) {
      NodeP n  = make(Kind::FUNC, parse(s.sub(open, s.n)));
      n->index = (int)id;
      return n;
    }
  return nullptr;
}

NodeP
parse(Span s)
{
  format(s);
  if (s.n == 0) throw std::invalid_argument("Can't parse empty string");
  // constants
  if (s.is("e") || s.is("pi") || s.is("tau") || is_number(s)) {
    NodeP n  = make(Kind::CONST);
    n->value = s.is("e") ? M_E : s.is("pi") ? M_PI : s.is("tau") ? M_PI * 2 : std::stod(s.str());
    return n;
  }
  if (NodeP n = parse_operation(s)) return n;
  if (NodeP n = parse_function(s)) return n;
  static const char* const vars[] = {"x", "y", "z", "t"};
  for (int i = 0; i < 4; ++i)
    if (s.is(vars[i])) {
      NodeP n  = make(Kind::VAR);
      n->index = i;
      return n;
    }
  throw std::invalid_argument("Unable to parse \"" + s.str() + "\"");
}

double
ev(const Expression::Node* n, const double* v)
{
  switch (n->kind) {
    case Kind::CONST:
    case Kind::BCONST: return n->value;
    case Kind::VAR: return v[n->index];
    case Kind::ADD: return ev(n->a.get(), v) + ev(n->b.get(), v);
    case Kind::SUB: return ev(n->a.get(), v) - ev(n->b.get(), v);
    case Kind::MUL: return ev(n->a.get(), v) * ev(n->b.get(), v);
    case Kind::DIV: return ev(n->a.get(), v) / ev(n->b.get(), v);
    case Kind::MOD: return std::fmod(ev(n->a.get(), v), ev(n->b.get(), v));
    case Kind::POW: return std::pow(ev(n->a.get(), v), ev(n->b.get(), v));
    case Kind::IPOW: return std::pow(ev(n->a.get(), v), n->index);
    case Kind::NEG: return -ev(n->a.get(), v);
    case Kind::FUNC: return call(n->index, ev(n->a.get(), v));
    case Kind::COND: return ev(n->a.get(), v) != 0.0 ? ev(n->b.get(), v) : ev(n->c.get(), v);
    case Kind::LT: return ev(n->a.get(), v) < ev(n->b.get(), v) ? 1.0 : 0.0;
    case Kind::LE: return ev(n->a.get(), v) <= ev(n->b.get(), v) ? 1.0 : 0.0;
    case Kind::GT: return ev(n->a.get(), v) > ev(n->b.get(), v) ? 1.0 : 0.0;
    case Kind::GE: return ev(n->a.get(), v) >= ev(n->b.get(), v) ? 1.0 : 0.0;
    case Kind::EQ: return ev(n->a.get(), v) == ev(n->b.get(), v) ? 1.0 : 0.0;
    case Kind::AND: return (ev(n->a.get(), v) != 0.0) & (ev(n->b.get(), v) != 0.0) ? 1.0 : 0.0;
    case Kind::OR: return (ev(n->a.get(), v) != 0.0) | (ev(n->b.get(), v) != 0.0) ? 1.0 : 0.0;
    case Kind::XOR: return (ev(n->a.get(), v) != 0.0) ^ (ev(n->b.get(), v) != 0.0) ? 1.0 : 0.0;
    case Kind::NOT: return ev(n->a.get(), v) != 0.0 ? 0.0 : 1.0;
  }
  return 0.0;
}

bool
uses(const Expression::Node* n, int lo, int hi)
{
  if (!n) return false;
  if (n->kind == Kind::VAR && n->index >= lo && n->index <= hi) return true;
  return uses(n->a.get(), lo, hi) || uses(n->b.get(), lo, hi) || uses(n->c.get(), lo, hi);
}

// canonical text of a sub-tree (slot de-duplication only)
std::string
show(const Expression::Node* n)
{
  if (!n) return "";
  char buf[64];
  switch (n->kind) {
    case Kind::CONST:
    case Kind::BCONST: std::snprintf(buf, sizeof buf, "%a", n->value); return buf;
    case Kind::VAR: return std::string(1, "xyzt"[n->index]);
    default: break;
  }
  std::snprintf(buf, sizeof buf, "(%d:%d ", (int)n->kind, n->index);
  return std::string(buf) + show(n->a.get()) + "," + show(n->b.get()) + "," + show(n->c.get()) + ")";
}

struct Emitter
{
  std::vector<int32_t>                 code;
  std::vector<double>                  consts;
  std::vector<NodeP>                   slot_nodes;
  std::vector<std::string>             slot_keys;
  std::vector<NodeP>                   entry_nodes;  // sub-trees of (x, y, z) alone that need libm: one value per BC entry
  std::vector<std::string>             entry_keys;
  int                                  depth = 0, max_seen = 0;
  void
  push(int op, int arg)
  {
    code.push_back(op | (arg << 8));
    if (++depth > max_seen) max_seen = depth;
  }
  void
  apply(int op, int pops)
  {
    code.push_back(op);
    depth += 1 - pops;
  }
};

// op codes of include/nsm_b200.h (nsm_bc_op); kept numeric here: the host layer does not include the CUDA ABI
enum { OP_CONST = 0, OP_X = 1, OP_SLOT = 4, OP_ADD = 5, OP_SUB = 6, OP_MUL = 7, OP_DIV = 8, OP_FMOD = 9, OP_NEG = 10, OP_SQRT = 11,
       OP_ABS = 12, OP_FLOOR = 13, OP_CEIL = 14, OP_ROUND = 15, OP_LT = 16, OP_LE = 17, OP_GT = 18, OP_GE = 19, OP_EQ = 20,
       OP_AND = 21, OP_OR = 22, OP_XOR = 23, OP_NOT = 24, OP_SELECT = 25, OP_ENTRYCONST = 26 };

// every operation of the sub-tree has a correctly rounded / exact device counterpart
bool
exact_on_device(const Expression::Node* n)
{
  if (!n) return true;
  switch (n->kind) {
    case Kind::CONST:
    case Kind::BCONST:
    case Kind::VAR: return true;
    case Kind::ADD:
    case Kind::SUB:
    case Kind::MUL:
    case Kind::DIV:
    case Kind::MOD:
    case Kind::NEG:
    case Kind::LT:
    case Kind::LE:
    case Kind::GT:
    case Kind::GE:
    case Kind::EQ:
    case Kind::AND:
    case Kind::OR:
    case Kind::XOR:
    case Kind::NOT:
    case Kind::COND: break;
    case Kind::FUNC:
      if (n->index != 6 && n->index != 10 && n->index != 13 && n->index != 14 && n->index != 15) return false;
      break;
    default: return false;  // POW / IPOW
  }
  return exact_on_device(n->a.get()) && exact_on_device(n->b.get()) && exact_on_device(n->c.get());
}

bool
emit(const NodeP& n, Emitter& e)
{
  if (n->kind == Kind::CONST || n->kind == Kind::BCONST) {
    e.consts.push_back(n->value);
    e.push(OP_CONST, (int)e.consts.size() - 1);
    return true;
  }
  if (!uses(n.get(), 0, 2)) {  // a function of t alone (or of constants): evaluated by the host, once per step
    const std::string key = show(n.get());
    size_t            k   = 0;
    while (k < e.slot_keys.size() && e.slot_keys[k] != key) ++k;
    if (k == e.slot_keys.size()) {
      e.slot_keys.push_back(key);
      e.slot_nodes.push_back(n);
    }
    e.push(OP_SLOT, (int)k);
    return true;
  }
  if (!uses(n.get(), 3, 3) && !exact_on_device(n.get())) {
    // a function of the position alone that goes through libm or pow (sin(3*x), x^2, exp(-y)): it does not change in
    // time, so the host evaluates it ONCE per boundary-condition entry at set-up -- with glibc's bits, exactly where the
    // reference evaluates it every step (src/nimble_boundary_condition_manager.h:166-201) -- and the device program
    // reads it as a per-entry constant
    const std::string key = show(n.get());
    size_t            k   = 0;
    while (k < e.entry_keys.size() && e.entry_keys[k] != key) ++k;
    if (k == e.entry_keys.size()) {
      e.entry_keys.push_back(key);
      e.entry_nodes.push_back(n);
    }
    e.push(OP_ENTRYCONST, (int)k);
    return true;
  }
  auto binary = [&](int op) {
    if (!emit(n->a, e) || !emit(n->b, e)) return false;
    e.apply(op, 2);
    return true;
  };
  auto unary = [&](int op) {
    if (!emit(n->a, e)) return false;
    e.apply(op, 1);
    return true;
  };
  switch (n->kind) {
    case Kind::VAR: e.push(OP_X + n->index, 0); return true;  // x, y, z (t never reaches here)
    case Kind::ADD: return binary(OP_ADD);
    case Kind::SUB: return binary(OP_SUB);
    case Kind::MUL: return binary(OP_MUL);
    case Kind::DIV: return binary(OP_DIV);
    case Kind::MOD: return binary(OP_FMOD);
    case Kind::NEG: return unary(OP_NEG);
    case Kind::LT: return binary(OP_LT);
    case Kind::LE: return binary(OP_LE);
    case Kind::GT: return binary(OP_GT);
    case Kind::GE: return binary(OP_GE);
    case Kind::EQ: return binary(OP_EQ);
    case Kind::AND: return binary(OP_AND);
    case Kind::OR: return binary(OP_OR);
    case Kind::XOR: return binary(OP_XOR);
    case Kind::NOT: return unary(OP_NOT);
    case Kind::COND:
      if (!emit(n->a, e) || !emit(n->b, e) || !emit(n->c, e)) return false;
      e.apply(OP_SELECT, 3);
      return true;
    case Kind::FUNC:
      switch (n->index) {  // kFuncs: only the correctly rounded / exact ones have a device counterpart
        case 6: return unary(OP_ABS);
        case 10: return unary(OP_SQRT);
        case 13: return unary(OP_CEIL);
        case 14: return unary(OP_ROUND);
        case 15: return unary(OP_FLOOR);
        default: return false;
      }
    default: return false;  // POW / IPOW mixing position and time: glibc's pow is not reproducible on the device
  }
}

}  // namespace

bool
Expression::compile(std::vector<int32_t>& code, std::vector<double>& consts, std::vector<Expression>& slots,
                    std::vector<Expression>& entry_constants, int max_depth) const
{
  if (!root_) return false;
  Emitter e;
  for (const Expression& s : slots) {
    e.slot_keys.push_back(show(s.root_.get()));
    e.slot_nodes.push_back(s.root_);
  }
  for (const Expression& s : entry_constants) {
    e.entry_keys.push_back(show(s.root_.get()));
    e.entry_nodes.push_back(s.root_);
  }
  const size_t const_base = consts.size();
  if (!emit(root_, e) || e.max_seen > max_depth) return false;
  for (int32_t w : e.code) {
    const int op = w & 0xff, arg = w >> 8;
    code.push_back(op == OP_CONST ? (op | ((arg + (int)const_base) << 8)) : w);
  }
  consts.insert(consts.end(), e.consts.begin(), e.consts.end());
  for (size_t k = slots.size(); k < e.slot_nodes.size(); ++k) slots.push_back(Expression(e.slot_nodes[k], e.slot_keys[k]));
  for (size_t k = entry_constants.size(); k < e.entry_nodes.size(); ++k) entry_constants.push_back(Expression(e.entry_nodes[k], e.entry_keys[k]));
  return true;
}

Expression::Expression(const std::string& text) : text_(text)
{
  if (text.empty()) return;
  std::string lower(text);
  for (char& c : lower)
    if (c >= 'A' && c <= 'Z') c = (char)(c + 32);
  // the tree keeps no pointers into `lower`
  root_ = parse(Span{lower.c_str(), (int)lower.size()});
}

double
Expression::eval(double x, double y, double z, double t) const
{
  if (!root_) return 0.0;
  const double v[4] = {x, y, z, t};
  return ev(root_.get(), v);
}

bool
Expression::depends_on_time() const
{
  return uses(root_.get(), 3, 3);
}

bool
Expression::depends_on_position() const
{
  return uses(root_.get(), 0, 2);
}

}  // namespace nimble_b200
