// nimblesm_b200/host/expression.h — arithmetic expressions in x, y, z, t for boundary-condition magnitudes
// (the role of ExpressionParsing::BoundaryConditionFunctor, src/nimble_expression_parser.h:694-760).
//
// Host-only: magnitudes are evaluated on the host with libm, exactly where the reference evaluates them
// (src/nimble_boundary_condition_manager.h:104-113,166-201), and uploaded as per-step tables, because the bits
// of cos()/exp() must be glibc's for parity.  The PARSE TREE follows the reference's splitting rules
// (src/nimble_expression_parser.h:583-640): the text is split at the LAST top-level '+', else the last binary
// '-', else the last '*', else the last '/', else the FIRST '^' -- so "a*b/c" is a*(b/c) and "a+b-c" is
// a+(b-c), unlike C -- since the association order decides the rounding of the result.
#pragma once
#include <memory>
#include <string>

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

 private:
  std::string           text_;
  std::shared_ptr<Node> root_;
};

}  // namespace nimble_b200
