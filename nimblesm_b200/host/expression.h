// nimblesm_b200/host/expression.h — arithmetic expressions in x, y, z, t for boundary-condition magnitudes
// (the role of ExpressionParsing::BoundaryConditionFunctor, src/nimble_expression_parser.h:694-760).
//
// eval() runs on the host with libm, exactly where the reference evaluates magnitudes
// (src/nimble_boundary_condition_manager.h:104-113,166-201): the bits of cos()/exp() must be glibc's for parity.
// compile() turns the tree into a device program (include/nsm_b200.h, nsm_bc_op) when its position-dependent part
// is IEEE-exact arithmetic; every sub-tree of t alone becomes a "slot" the host evaluates once per step, so
// cos(t*pi/T)*x needs one scalar per step from the host instead of one value per boundary node.  The PARSE TREE follows the reference's splitting rules
// (src/nimble_expression_parser.h:583-640): the text is split at the LAST top-level '+', else the last binary
// '-', else the last '*', else the last '/', else the FIRST '^' -- so "a*b/c" is a*(b/c) and "a+b-c" is
// a+(b-c), unlike C -- since the association order decides the rounding of the result.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace nimble_b200 {

class Expression
{
 public:
  Expression() = default;
  explicit Expression(const std::string& text);  // throws std::invalid_argument on a parse error
  bool
  empty() const
  {
    return !root_;
  }
  double
  eval(double x, double y, double z, double t) const;
  bool
  depends_on_time() const;
  bool
  depends_on_position() const;
  const std::string&
  text() const
  {
    return text_;
  }
  struct Node;

  // Appends this expression's device program to `code` (words op | arg << 8), its constants to `consts` and the
  // host-evaluated sub-expressions of t to `slots` (shared between programs: an identical text reuses its slot).
  // Returns false, leaving the outputs untouched, when some position-dependent node has no bit-exact device
  // counterpart (transcendental functions and pow of x, y, z) or the value stack would exceed `max_depth`.
  // entry_constants: sub-trees of (x, y, z) alone that need libm or pow; the caller evaluates each once per
  // boundary-condition entry (NSM_BCOP_ENTRYCONST).  Only sub-trees that MIX position and time through such a
  // function (sin(x*t)) leave the expression without a device form.
  bool
  compile(std::vector<int32_t>& code, std::vector<double>& consts, std::vector<Expression>& slots,
          std::vector<Expression>& entry_constants, int max_depth) const;

 private:
  Expression(std::shared_ptr<Node> root, std::string text) : text_(std::move(text)), root_(std::move(root)) {}
  std::string           text_;
  std::shared_ptr<Node> root_;
};

}  // namespace nimble_b200
