// nimblesm_b200/host/boundary_condition.cc — see boundary_condition.h.
#include "boundary_condition.h"

#include <cstdlib>

#include <algorithm>
#include <sstream>
#include <stdexcept>

namespace nimble_b200 {

bool
BoundaryCondition::Initialize(int dim, std::string bc_string, std::map<int, std::string> const& node_set_names,
                              std::map<int, std::string> const& side_set_names)
{
  dim_ = dim;
  std::istringstream ss(bc_string);
  std::string        type("undefined"), coordinate("undefined");
  ss >> type;
  if (type == "initial_velocity")
    bc_type_ = INITIAL_VELOCITY;
  else if (type == "prescribed_velocity")
    bc_type_ = PRESCRIBED_VELOCITY;
  else if (type == "prescribed_displacement")
    bc_type_ = PRESCRIBED_DISPLACEMENT;
  else if (type == "prescribed_traction")
    bc_type_ = PRESCRIBED_TRACTION;
  else
    throw std::invalid_argument("Error processing boundary condition, unknown boundary condition type: " + type);
  const bool neumann = bc_type_ == PRESCRIBED_TRACTION;
  if (neumann)
    ss >> side_set_name_;
  else
    ss >> node_set_name_;
  ss >> coordinate;
  const long quotes = std::count(bc_string.begin(), bc_string.end(), '"');
  if (quotes == 2) {
    has_expression_    = true;
    const size_t first = bc_string.find('"'), last = bc_string.rfind('"');
    expression_string_ = bc_string.substr(first + 1, last - first - 1);
    expression_        = Expression(expression_string_);
  } else if (quotes == 0) {
    has_expression_ = false;
    ss >> magnitude_;
  } else {
    throw std::invalid_argument("Error processing boundary condition, illegal number of quotes: " + bc_string);
  }
  std::transform(coordinate.begin(), coordinate.end(), coordinate.begin(), ::tolower);
  auto find_id = [](std::map<int, std::string> const& names, std::string const& name) {
    for (auto const& kv : names)
      if (kv.second == name) return kv.first;
    return -1;
  };
  bool valid = true;
  if (neumann) {
    side_set_id_ = find_id(side_set_names, side_set_name_);
    valid        = side_set_id_ != -1;
  } else {
    node_set_id_ = find_id(node_set_names, node_set_name_);
    valid        = node_set_id_ != -1;
  }
  if (coordinate == "x")
    coordinate_ = 0;
  else if (coordinate == "y")
    coordinate_ = 1;
  else if (coordinate == "z")
    coordinate_ = 2;
  else
    throw std::invalid_argument("Error processing boundary condition, unknown coordinate: " + coordinate);
  return valid;
}

void
BoundaryConditionManager::Initialize(std::map<int, std::string> const& node_set_names,
                                     std::map<int, std::vector<int>> const& node_sets,
                                     std::map<int, std::string> const& side_set_names,
                                     std::map<int, std::vector<int>> const& side_sets,
                                     std::vector<std::string> const& bc_strings, int dim,
                                     std::string const& time_integration_scheme)
{
  node_set_names_ = node_set_names, node_sets_ = node_sets;
  side_set_names_ = side_set_names, side_sets_ = side_sets;
  dim_            = dim;
  if (time_integration_scheme == "explicit")
    scheme_ = EXPLICIT;
  else if (time_integration_scheme == "quasistatic")
    scheme_ = QUASISTATIC;
  else
    throw std::invalid_argument("BoundaryConditionManager: unknown time integration scheme " + time_integration_scheme);
  boundary_conditions_.clear();
  for (auto const& s : bc_strings) {
    BoundaryCondition bc;
    if (bc.Initialize(dim_, s, node_set_names_, side_set_names_)) boundary_conditions_.push_back(bc);
  }
  table_ = DeviceTable();
  time_dependent_ = false;
  for (size_t b = 0; b < boundary_conditions_.size(); ++b) {
    const BoundaryCondition& bc = boundary_conditions_[b];
    if (bc.bc_type_ == BoundaryCondition::PRESCRIBED_TRACTION)
      throw std::invalid_argument("prescribed_traction is outside the hex8 explicit path of this build (side sets)");
    if (bc.bc_type_ != BoundaryCondition::PRESCRIBED_VELOCITY && bc.bc_type_ != BoundaryCondition::PRESCRIBED_DISPLACEMENT)
      continue;
    auto it = node_sets_.find(bc.node_set_id_);
    if (it == node_sets_.end()) continue;
    for (int n : it->second) {
      table_.node.push_back(n);
      table_.comp.push_back(bc.coordinate_);
      table_.kind.push_back(bc.bc_type_ == BoundaryCondition::PRESCRIBED_VELOCITY ? 0 : 1);
      table_.bc_index.push_back((int)b);
    }
    if (bc.has_expression_ && bc.expression_.depends_on_time()) time_dependent_ = true;
  }
  // device programs for the time-dependent magnitudes: all of them or none (one code path per run)
  programs_ = DevicePrograms();
  const char* force_host = std::getenv("NSM_B200_HOST_BC");
  if (time_dependent_ && dim_ == 3 && !(force_host && force_host[0] == '1')) {
    DevicePrograms   p;
    std::vector<int> program_of_bc(boundary_conditions_.size(), -1);
    bool             ok = true;
    for (size_t b = 0; b < boundary_conditions_.size() && ok; ++b) {
      const BoundaryCondition& bc = boundary_conditions_[b];
      if (bc.bc_type_ != BoundaryCondition::PRESCRIBED_VELOCITY && bc.bc_type_ != BoundaryCondition::PRESCRIBED_DISPLACEMENT)
        continue;
      if (!bc.has_expression_ || !bc.expression_.depends_on_time()) continue;
      ok = bc.expression_.compile(p.code, p.consts, p.slots, p.entry_constants, 16);
      if (ok) {
        program_of_bc[b] = (int)p.offsets.size() - 1;
        p.offsets.push_back((int)p.code.size());
      }
    }
    if (ok) {
      for (int b : table_.bc_index) p.program_of_entry.push_back(program_of_bc[b]);
      p.active  = true;
      programs_ = std::move(p);
    }
  }
}

void
BoundaryConditionManager::ApplyInitialConditions(const Viewify<2>& X, Viewify<2> velocity) const
{
  for (auto const& bc : boundary_conditions_) {
    if (bc.bc_type_ != BoundaryCondition::INITIAL_VELOCITY) continue;
    auto it = node_sets_.find(bc.node_set_id_);
    if (it == node_sets_.end()) continue;
    for (int n : it->second)
      velocity(n, bc.coordinate_) =
          bc.has_expression_ ? bc.expression_.eval(X(n, 0), X(n, 1), dim_ == 3 ? X(n, 2) : 0.0, 0.0) : bc.magnitude_;
  }
}

void
BoundaryConditionManager::ApplyKinematicBC(double time_current, double time_previous, const Viewify<2>& X,
                                           Viewify<2> displacement, Viewify<2> velocity) const
{
  const double delta_t = time_current - time_previous;
  for (auto const& bc : boundary_conditions_) {
    auto it = node_sets_.find(bc.node_set_id_);
    if (it == node_sets_.end()) continue;
    const int c = bc.coordinate_;
    if (bc.bc_type_ == BoundaryCondition::PRESCRIBED_VELOCITY) {
      for (int n : it->second) {
        const double v =
            bc.has_expression_ ? bc.expression_.eval(X(n, 0), X(n, 1), dim_ == 3 ? X(n, 2) : 0.0, time_current) : bc.magnitude_;
        velocity(n, c) = v;
        if (scheme_ == QUASISTATIC) displacement(n, c) += v * delta_t;
      }
    } else if (bc.bc_type_ == BoundaryCondition::PRESCRIBED_DISPLACEMENT && delta_t > 0.0) {
      for (int n : it->second) {
        const double d =
            bc.has_expression_ ? bc.expression_.eval(X(n, 0), X(n, 1), dim_ == 3 ? X(n, 2) : 0.0, time_current) : bc.magnitude_;
        velocity(n, c) = (d - displacement(n, c)) / delta_t;
        if (scheme_ == QUASISTATIC) displacement(n, c) = d;
      }
    }
  }
}

void
BoundaryConditionManager::EvaluateEntryConstants(const Viewify<2>& X, double* values) const
{
  // values[j][k]: entry constant j (a function of the position alone) at the node of table entry k
  const size_t n = table_.node.size();
  for (size_t j = 0; j < programs_.entry_constants.size(); ++j)
    for (size_t k = 0; k < n; ++k) {
      const int nd      = table_.node[k];
      values[j * n + k] = programs_.entry_constants[j].eval(X(nd, 0), X(nd, 1), dim_ == 3 ? X(nd, 2) : 0.0, 0.0);
    }
}

void
BoundaryConditionManager::EvaluateMagnitudes(double t, const Viewify<2>& X, double* values, bool skip_program_entries) const
{
  const size_t n = table_.node.size();
  for (size_t k = 0; k < n; ++k) {
    const BoundaryCondition& bc = boundary_conditions_[table_.bc_index[k]];
    const int                nd = table_.node[k];
    if (skip_program_entries && programs_.active && programs_.program_of_entry[k] >= 0) {
      values[k] = 0.0;  // rewritten on the device every step
      continue;
    }
    values[k] = bc.has_expression_ ? bc.expression_.eval(X(nd, 0), X(nd, 1), dim_ == 3 ? X(nd, 2) : 0.0, t) : bc.magnitude_;
  }
}

}  // namespace nimble_b200
