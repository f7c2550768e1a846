// nimblesm_b200/host/model_data.h — the model-data layer: FieldIds, ModelDataBase and the B200 ModelData.
//
// ModelDataBase repeats the virtual interface of nimble::ModelDataBase (src/nimble_model_data_base.h:86-352);
// nimble_b200::ModelData is the subclass the reference would select in DataManager::Initialize
// (src/nimble_data_manager.cc:105-114) instead of nimble::ModelData / nimble_kokkos::ModelData.  Nodal fields
// live on the device as SoA fp64; the host keeps pinned [n][3] mirrors so that Viewify views handed to callers
// stay valid (src/nimble_model_data.cc:540-546).  The mirrors are synchronised explicitly:
//   UpdateWithNewVelocity / UpdateWithNewDisplacement   host -> device   (the reference's deep_copies,
//                                                       src/nimble_kokkos_model_data.cc:1730-1740)
//   ComputeInternalForce(displacement, force)           host -> device -> host
//   AdvanceOnDevice(n, ...)                             whole steps on the device, then PullNodalFields()
#pragma once
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "block.h"
#include "device.h"
#include "view.h"

namespace nimble_b200 {

class DataManager;
class GenesisMesh;

struct FieldIds
{
  int deformation_gradient = -1;
  int stress               = -1;
  int unrotated_stress     = -1;
  int reference_coordinates = -1;
  int displacement          = -1;
  int velocity              = -1;
  int acceleration          = -1;
  int lumped_mass    = -1;
  int internal_force = -1;
  int contact_force  = -1;
  int external_force = -1;
};

class ModelDataBase
{
 public:
  ModelDataBase()          = default;
  virtual ~ModelDataBase() = default;
  virtual int
  AllocateNodeData(Length length, std::string label, int num_objects) = 0;
  int
  GetFieldIdChecked(const std::string& field_label) const;  // throws std::runtime_error: Field "x" not allocated
  virtual int
  GetFieldId(const std::string& field_label) const = 0;
  virtual void
  InitializeBlocks(DataManager& data_manager, const std::shared_ptr<MaterialFactoryBase>& material_factory_base) = 0;
  virtual void
  UpdateStates(const DataManager& data_manager) = 0;
  Viewify<1>
  GetScalarNodeData(const std::string& label)
  {
    return GetScalarNodeData(GetFieldIdChecked(label));
  }
  virtual Viewify<1>
  GetScalarNodeData(int field_id) = 0;
  Viewify<2>
  GetVectorNodeData(const std::string& label)
  {
    return GetVectorNodeData(GetFieldIdChecked(label));
  }
  virtual Viewify<2>
  GetVectorNodeData(int field_id) = 0;
  virtual void
  ComputeLumpedMass(DataManager& data_manager) = 0;
  virtual void
  InitializeExodusOutput(DataManager& data_manager) = 0;
  virtual void
  WriteExodusOutput(DataManager& data_manager, double time_current) = 0;
  virtual void
  ComputeExternalForce(DataManager&, double, double, bool)
  {
  }
  virtual void
  ComputeInternalForce(DataManager& data_manager, double time_previous, double time_current, bool is_output_step,
                       const Viewify<2>& displacement, Viewify<2>& force) = 0;
  virtual void
  ApplyInitialConditions(DataManager& data_manager);
  virtual void
  ApplyKinematicConditions(DataManager& data_manager, double time_current, double time_previous);
  virtual void
  UpdateWithNewVelocity(DataManager&, double)
  {
  }
  virtual void
  UpdateWithNewDisplacement(DataManager&, double)
  {
  }
  int
  GetDimension() const
  {
    return dim_;
  }
  void
  SetDimension(int dim);
  void
  SetCriticalTimeStep(double time_step)
  {
    critical_time_step_ = time_step;
  }
  double
  GetCriticalTimeStep() const
  {
    return critical_time_step_;
  }
  void
  SetReferenceCoordinates(const GenesisMesh& mesh);
  const std::vector<std::string>&
  GetNodeDataLabelsForOutput() const
  {
    return output_node_component_labels_;
  }
  const std::map<int, std::vector<std::string>>&
  GetElementDataLabels() const
  {
    return element_component_labels_;
  }
  const std::map<int, std::vector<std::string>>&
  GetElementDataLabelsForOutput() const
  {
    return output_element_component_labels_;
  }
  const std::map<int, std::vector<std::string>>&
  GetDerivedElementDataLabelsForOutput() const
  {
    return derived_output_element_data_labels_;
  }

 protected:
  int                                     dim_ = 3;
  double                                  critical_time_step_ = 0.0;
  std::vector<std::string>                output_node_component_labels_;
  std::map<int, std::vector<std::string>> element_component_labels_;
  std::map<int, std::vector<std::string>> output_element_component_labels_;
  std::map<int, std::vector<std::string>> derived_output_element_data_labels_;
};

class ModelData : public ModelDataBase
{
 public:
  // assembly: NSM_ASSEMBLY_ATOMIC (fastest) or NSM_ASSEMBLY_ORDERED (nodal sums in the serial reference's order,
  // bit-reproducible); flags: NSM_FLAG_* of include/nsm_b200.h
  explicit ModelData(int device = 0, int assembly = NSM_ASSEMBLY_ORDERED, unsigned flags = NSM_FLAG_CACHE_REF_JACOBIAN);
  ~ModelData() override;

  int
  AllocateNodeData(Length length, std::string label, int num_objects) override;
  int
  GetFieldId(const std::string& field_label) const override;
  void
  InitializeBlocks(DataManager& data_manager, const std::shared_ptr<MaterialFactoryBase>& material_factory_base) override;
  // N <- N+1 (src/nimble_model_data.h:104-107): a record swap on the device for blocks whose material has state
  // variables (nsm_b200_update_states); nothing to roll for the reference's two stateless models, whose F / sigma are
  // written once per requesting call
  void
  UpdateStates(const DataManager&) override;
  using ModelDataBase::GetScalarNodeData;
  using ModelDataBase::GetVectorNodeData;
  Viewify<1>
  GetScalarNodeData(int field_id) override;
  Viewify<2>
  GetVectorNodeData(int field_id) override;
  void
  ComputeLumpedMass(DataManager& data_manager) override;
  void
  InitializeExodusOutput(DataManager& data_manager) override;
  void
  WriteExodusOutput(DataManager& data_manager, double time_current) override;
  void
  ComputeExternalForce(DataManager& data_manager, double time_previous, double time_current, bool is_output_step) override;
  void
  ComputeInternalForce(DataManager& data_manager, double time_previous, double time_current, bool is_output_step,
                       const Viewify<2>& displacement, Viewify<2>& force) override;
  void
  ApplyKinematicConditions(DataManager& data_manager, double time_current, double time_previous) override;
  void
  UpdateWithNewVelocity(DataManager& data_manager, double dt) override;
  void
  UpdateWithNewDisplacement(DataManager& data_manager, double dt) override;

  // ---- B200 extensions -------------------------------------------------------------------------------
  DeviceContext&
  Device()
  {
    return *device_;
  }
  int
  DeviceIndex() const
  {
    return device_index_;
  }
  int
  Assembly() const
  {
    return assembly_;
  }
  std::map<int, std::shared_ptr<Block>>&
  GetBlocks()
  {
    return blocks_;
  }
  // n whole steps of the explicit loop body on the device (src/integrators/explicit_time_integrator.cc:177-278)
  // starting from the DEVICE state; boundary-condition magnitudes are evaluated on the host for every step of
  // the run.  store_ipt_last: the last step is an output step.
  void
  AdvanceOnDevice(DataManager& data_manager, int n_steps, double& time_current, double user_time_step, bool store_ipt_last);
  // accumulated device time of AdvanceOnDevice: element kernels / node passes incl. the shared-node exchange
  double
  DeviceForceSeconds() const
  {
    return device_force_seconds_;
  }
  double
  DeviceUpdateSeconds() const
  {
    return device_update_seconds_;
  }
  double
  DeviceContactSeconds() const
  {
    return device_contact_seconds_;
  }
  // contact entities were sent to the device (ContactManager::CreateContactEntities): the fused steps carry the contact
  // term and PullNodalFields also brings the contact force home
  void
  SetContactOnDevice(bool on)
  {
    contact_on_device_ = on;
  }
  void
  PushNodalFields();  // host mirrors (u, v, a) -> device
  void
  PullNodalFields();  // device (u, v, a, f_int) -> host mirrors
  std::vector<double>&
  GetElementDataNew(int block_id);  // [elem][8][15 + n_state], refreshed from the device
  void
  SpecifyOutputFields(const std::string& output_field_string);

 private:
  struct Field
  {
    std::string label;
    Length      length;
    int         num_objects;
    double*     data = nullptr;  // pinned host memory
  };
  int
  device_field(const std::string& label) const;
  int                                   device_index_;
  int                                   assembly_;
  unsigned                              flags_;
  std::unique_ptr<DeviceContext>        device_;
  std::vector<Field>                    fields_;
  std::map<std::string, int>            field_ids_;
  std::map<int, std::shared_ptr<Block>> blocks_;
  std::vector<int>                      block_ids_;
  std::map<int, std::vector<double>>    element_data_np1_;
  double                                device_force_seconds_ = 0.0, device_update_seconds_ = 0.0, device_contact_seconds_ = 0.0;
  std::vector<double>                   bc_values_;
  std::vector<double>                   bc_slots_;
  bool                                  bc_table_sent_ = false, bc_programs_sent_ = false;
  int                                   num_nodes_     = 0;
  bool                                  contact_on_device_ = false;
};

}  // namespace nimble_b200
