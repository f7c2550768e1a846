// nimblesm_b200/host/boundary_condition.h — nimble::BoundaryCondition / BoundaryConditionManager
// (src/nimble_boundary_condition.{h,cc}, src/nimble_boundary_condition_manager.{h,cc}) for the explicit scheme.
//
// The reference applies every BC on host views, node by node, twice per step.  Here the manager (1) keeps the
// same parse / validity rules and host application (used at t = 0 and by callers that drive the reference
// sequence through ModelData), and (2) flattens the kinematic BCs into the device table of
// nsm_b200_set_bc_table (deck order, later entries win) and evaluates their magnitudes per step on the host
// -- constants once, expression(x,y,z,t) BCs into a [steps][entries] table -- so that whole runs of steps
// execute on the GPU without a host round trip.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "expression.h"
#include "view.h"

namespace nimble_b200 {

class GenesisMesh;

class BoundaryCondition
{
 public:
  enum Boundary_Condition_Type
  {
    UNDEFINED               = 0,
    INITIAL_VELOCITY        = 1,
    PRESCRIBED_VELOCITY     = 2,
    PRESCRIBED_DISPLACEMENT = 3,
    PRESCRIBED_TRACTION     = 4
  };
  // false: the node set is not known on this rank (the BC is then dropped, :72-76 of the manager .cc);
  // throws std::invalid_argument on a malformed string (src/nimble_boundary_condition.cc:65-160)
  bool
  Initialize(int dim, std::string bc_string, std::map<int, std::string> const& node_set_names,
             std::map<int, std::string> const& side_set_names);

  int                     dim_{0};
  std::string             node_set_name_{"unknown"};
  int                     node_set_id_{-1};
  std::string             side_set_name_{"unknown"};
  int                     side_set_id_{-1};
  int                     coordinate_{-1};
  double                  magnitude_{0.0};
  Boundary_Condition_Type bc_type_{UNDEFINED};
  bool                    has_expression_{false};
  std::string             expression_string_{""};
  Expression              expression_;
};

class BoundaryConditionManager
{
 public:
  enum Time_Integration_Scheme
  {
    EXPLICIT    = 0,
    QUASISTATIC = 1
  };
  void
  Initialize(std::map<int, std::string> const& node_set_names, std::map<int, std::vector<int>> const& node_sets,
             std::map<int, std::string> const& side_set_names, std::map<int, std::vector<int>> const& side_sets,
             std::vector<std::string> const& bc_strings, int dim, std::string const& time_integration_scheme);
  bool
  IsPeriodicRVEProblem() const
  {
    return false;
  }
  const std::vector<BoundaryCondition>&
  GetBoundaryConditions() const
  {
    return boundary_conditions_;
  }

  // host application on [n][3] views, the reference's loops (src/nimble_boundary_condition_manager.h:93-204)
  void
  ApplyInitialConditions(const Viewify<2>& reference_coordinates, Viewify<2> velocity) const;
  void
  ApplyKinematicBC(double time_current, double time_previous, const Viewify<2>& reference_coordinates,
                   Viewify<2> displacement, Viewify<2> velocity) const;

  // ---- device tables ---------------------------------------------------------------------------------
  // One entry per (kinematic BC, node of its set), in deck order.
  struct DeviceTable
  {
    std::vector<int> node, comp, kind;  // kind: 0 = prescribed velocity, 1 = prescribed displacement
    std::vector<int> bc_index;          // entry -> boundary_conditions_ index
  };
  const DeviceTable&
  GetDeviceTable() const
  {
    return table_;
  }
  bool
  HasTimeDependentMagnitudes() const
  {
    return time_dependent_;
  }
  // magnitudes of all table entries at time t (x, y, z = reference coordinates of the entry's node); with
  // skip_program_entries the entries served by a device program (below) are left at 0
  void
  EvaluateMagnitudes(double t, const Viewify<2>& reference_coordinates, double* values, bool skip_program_entries = false) const;

  // ---- device programs (include/nsm_b200.h, nsm_b200_set_bc_programs) ----------------------------------
  // When every time-dependent expression compiles (Expression::compile), the step loop evaluates the magnitudes on
  // the device: per step the host supplies only the values of `slots` (sub-expressions of t alone).  Otherwise
  // (or with NSM_B200_HOST_BC=1 in the environment) active == false and the host evaluates one row of magnitudes
  // per step, as the reference does.
  struct DevicePrograms
  {
    bool                 active = false;
    std::vector<int>     offsets{0};        // [n_programs + 1]
    std::vector<int32_t> code;
    std::vector<double>  consts;
    std::vector<Expression> slots;            // functions of t alone: one host value per step
    std::vector<Expression> entry_constants;  // functions of (x, y, z) alone through libm / pow: one host value per entry
    std::vector<int>     program_of_entry;  // [table entries], -1 = host magnitude
  };
  const DevicePrograms&
  GetDevicePrograms() const
  {
    return programs_;
  }
  // values [entry_constants.size()][table entries] (NSM_BCOP_ENTRYCONST)
  void
  EvaluateEntryConstants(const Viewify<2>& X, double* values) const;
  void
  EvaluateSlots(double t, double* values) const
  {
    for (size_t k = 0; k < programs_.slots.size(); ++k) values[k] = programs_.slots[k].eval(0.0, 0.0, 0.0, t);
  }

 private:
  std::map<int, std::string>      node_set_names_, side_set_names_;
  std::map<int, std::vector<int>> node_sets_, side_sets_;
  std::vector<BoundaryCondition>  boundary_conditions_;
  int                             dim_{3};
  Time_Integration_Scheme         scheme_{EXPLICIT};
  DeviceTable                     table_;
  bool                            time_dependent_{false};
  DevicePrograms                  programs_;
};

}  // namespace nimble_b200
