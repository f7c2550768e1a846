// nimblesm_b200/host/expression.cc — see expression.h.
#include "expression.h"

#include <cmath>
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

}  // namespace

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
