// nimblesm_b200/host/model_data.cc — see model_data.h.
#include "model_data.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "data_manager.h"

namespace nimble_b200 {

namespace {

// GetComponentLabels (src/nimble_data_utils.cc:159-205), 3-D
std::vector<std::string>
component_labels(const std::string& label, Length length)
{
  static const char* const vec[3] = {"_x", "_y", "_z"};
  static const char* const sym[6] = {"_xx", "_yy", "_zz", "_xy", "_yz", "_zx"};
  static const char* const ful[9] = {"_xx", "_yy", "_zz", "_xy", "_yz", "_zx", "_yx", "_zy", "_xz"};
  std::vector<std::string> out;
  if (length == SCALAR) out.push_back(label);
  if (length == VECTOR)
    for (auto s : vec) out.push_back(label + s);
  if (length == SYMMETRIC_TENSOR)
    for (auto s : sym) out.push_back(label + s);
  if (length == FULL_TENSOR)
    for (auto s : ful) out.push_back(label + s);
  return out;
}

template <class C, class T>
bool
contains(const C& c, const T& v)
{
  return std::find(c.begin(), c.end(), v) != c.end();
}

const char* const kDeviceFieldLabels[NSM_FIELD_COUNT] = {"lumped_mass",  "reference_coordinate", "displacement", "velocity",
                                                         "acceleration", "internal_force",       "external_force", "contact_force"};

}  // namespace

// ---- ModelDataBase ---------------------------------------------------------------------------------------
int
ModelDataBase::GetFieldIdChecked(const std::string& field_label) const
{
  const int id = GetFieldId(field_label);
  if (id < 0) throw std::runtime_error("Field \"" + field_label + "\" not allocated");
  return id;
}

void
ModelDataBase::SetDimension(int dim)
{
  if (dim != 3) throw std::invalid_argument("\nError: the B200 hex8 path is three-dimensional, ModelData::SetDimension(" + std::to_string(dim) + ")\n");
  dim_ = dim;
}

void
ModelDataBase::SetReferenceCoordinates(const GenesisMesh& mesh)
{
  // src/nimble_model_data_base.cc:60-81
  const double* x = mesh.GetCoordinatesX();
  const double* y = mesh.GetCoordinatesY();
  const double* z = mesh.GetCoordinatesZ();
  Viewify<2>    X = GetVectorNodeData("reference_coordinate");
  const int     n = (int)mesh.GetNumNodes();
  for (int i = 0; i < n; ++i) X(i, 0) = x[i], X(i, 1) = y[i], X(i, 2) = z[i];
}

void
ModelDataBase::ApplyInitialConditions(DataManager& data_manager)
{
  auto bc = data_manager.GetBoundaryConditionManager();
  bc->ApplyInitialConditions(GetVectorNodeData("reference_coordinate"), GetVectorNodeData("velocity"));
}

void
ModelDataBase::ApplyKinematicConditions(DataManager& data_manager, double time_current, double time_previous)
{
  auto bc = data_manager.GetBoundaryConditionManager();
  bc->ApplyKinematicBC(time_current, time_previous, GetVectorNodeData("reference_coordinate"), GetVectorNodeData("displacement"),
                       GetVectorNodeData("velocity"));
}

// ---- ModelData (B200) ------------------------------------------------------------------------------------
ModelData::ModelData(int device, int assembly, unsigned flags) : device_index_(device), assembly_(assembly), flags_(flags)
{
  device_.reset(new DeviceContext(device));
}

ModelData::~ModelData()
{
  for (Field& f : fields_) nsm_b200_host_free(f.data);
}

int
ModelData::AllocateNodeData(Length length, std::string label, int num_objects)
{
  auto it = field_ids_.find(label);
  if (it != field_ids_.end()) return it->second;
  Field f;
  f.label = label, f.length = length, f.num_objects = num_objects;
  const int64_t count = (int64_t)num_objects * (int)length;
  f.data              = (double*)nsm_b200_host_alloc(std::max<int64_t>(count, 1) * (int64_t)sizeof(double));
  if (!f.data) throw std::runtime_error("ModelData::AllocateNodeData: pinned host allocation failed for " + label);
  std::fill(f.data, f.data + count, 0.0);
  const int id      = (int)fields_.size();
  field_ids_[label] = id;
  fields_.push_back(f);
  num_nodes_ = num_objects;
  return id;
}

int
ModelData::GetFieldId(const std::string& field_label) const
{
  auto it = field_ids_.find(field_label);
  return it == field_ids_.end() ? -1 : it->second;
}

Viewify<1>
ModelData::GetScalarNodeData(int field_id)
{
  Field& f = fields_.at(field_id);
  return Viewify<1>(f.data, f.num_objects);
}

Viewify<2>
ModelData::GetVectorNodeData(int field_id)
{
  Field& f = fields_.at(field_id);
  return Viewify<2>(f.data, {f.num_objects, 3}, {3, 1});
}

int
ModelData::device_field(const std::string& label) const
{
  for (int i = 0; i < NSM_FIELD_COUNT; ++i)
    if (label == kDeviceFieldLabels[i]) return i;
  return -1;
}

void
ModelData::InitializeBlocks(DataManager& data_manager, const std::shared_ptr<MaterialFactoryBase>& material_factory_base)
{
  const GenesisMesh& mesh   = data_manager.GetMesh();
  const Parser&      parser = data_manager.GetParser();
  // EmplaceBlocks (src/nimble_model_data.cc:416-437)
  for (int block_id : mesh.GetBlockIds()) {
    const std::string params = parser.GetModelMaterialParameters(block_id);
    if (params == "none")
      throw std::invalid_argument("\nError: no \"element block\" / \"material parameters\" entry in the input deck for block " +
                                  std::to_string(block_id) + "\n");
    if (mesh.GetElementType(block_id) != "HEX")
      throw std::invalid_argument("\nError: the B200 path handles hex8 blocks only (block " + std::to_string(block_id) + ")\n");
    auto block = std::make_shared<Block>();
    block->SetDeviceIndex(device_index_);
    block->Initialize(params, *material_factory_base);
    std::vector<std::pair<std::string, Length>> labels;
    block->GetDataLabelsAndLengths(labels);
    std::vector<std::string> comps;
    for (auto const& ll : labels) {
      auto c = component_labels(ll.first, ll.second);
      comps.insert(comps.end(), c.begin(), c.end());
    }
    blocks_[block_id] = block;
    block_ids_.push_back(block_id);
    element_component_labels_[block_id]           = comps;
    output_element_component_labels_[block_id]    = {};
    derived_output_element_data_labels_[block_id] = {};
  }
  // device model: nodes, blocks in ascending id, assembly tables
  DeviceContext& d = *device_;
  d.check(nsm_b200_set_nodes(d.get(), mesh.GetNumNodes(), mesh.GetCoordinatesX(), mesh.GetCoordinatesY(), mesh.GetCoordinatesZ()),
          "ModelData::InitializeBlocks (nodes)");
  for (auto const& kv : blocks_) {
    const Material&           m  = *kv.second->GetMaterialPointer();
    const std::vector<double> mp = m.DeviceParameters();
    d.check(nsm_b200_add_block_params(d.get(), kv.first, mesh.GetNumElementsInBlock(kv.first), mesh.GetConnectivity(kv.first), m.Kind(),
                                      (int)mp.size(), mp.data()),
            "ModelData::InitializeBlocks (block)");
  }
  d.check(nsm_b200_finalize(d.get(), assembly_, flags_), "ModelData::InitializeBlocks (finalize)");
  // shared-node exchange between the ranks of this run
  auto comm = data_manager.GetVectorCommunicator();
  if (comm && comm->NumRanks() > 1) comm->ConnectDevices(d);
  // AllocateInitializeElementData (src/nimble_model_data.cc:451-485)
  SpecifyOutputFields(parser.GetOutputFieldString());
}

void
ModelData::SpecifyOutputFields(const std::string& output_field_string)
{
  // src/nimble_model_data.cc:215-368, same precedence of interpretations
  std::vector<std::string> node_fields, node_components, elem_fields, elem_components, ipt_fields, ipt_components;
  std::map<std::string, Length> length_of;
  for (const Field& f : fields_) {
    node_fields.push_back(f.label);
    length_of[f.label] = f.length;
    for (auto const& c : component_labels(f.label, f.length)) node_components.push_back(c);
  }
  // per-point fields: F, sigma and the state variables of the blocks' materials (ELEMENT fields with an integration
  // point prefix in the reference's data_fields_, src/nimble_model_data.cc:233-260)
  std::vector<std::pair<std::string, Length>> point_fields = {{"deformation_gradient", FULL_TENSOR}, {"stress", SYMMETRIC_TENSOR}};
  for (auto const& kv : blocks_) {
    const Material& m = *kv.second->GetMaterialPointer();
    for (int i = 0; i < m.NumStateVariables(); ++i) {
      char label[MaterialParameters::MAX_MAT_MODEL_STR_LEN];
      m.GetStateVariableLabel(i, label);
      if (std::find_if(point_fields.begin(), point_fields.end(), [&](auto const& pf) { return pf.first == label; }) == point_fields.end())
        point_fields.emplace_back(label, SCALAR);
    }
  }
  for (int q = 1; q <= 8; ++q)
    for (auto const& base : point_fields) {
      char prefix[16];
      snprintf(prefix, sizeof prefix, "ipt%02d_", q);
      const std::string label = prefix + base.first;
      ipt_fields.push_back(label);
      length_of[label] = base.second;
      for (auto const& c : component_labels(label, base.second)) ipt_components.push_back(c);
      if (!contains(elem_fields, base.first)) {
        elem_fields.push_back(base.first);
        length_of[base.first] = base.second;
        for (auto const& c : component_labels(base.first, base.second)) elem_components.push_back(c);
      }
    }
  // a label is output for a block only if the block carries it (src/nimble_model_data.cc:284-306): `c` without the
  // integration-point prefix (volume averages) or with it
  auto on_block = [&](int id, const std::string& c, bool averaged) {
    for (auto const& have : element_component_labels_.at(id))
      if ((averaged ? have.substr(6) : have) == c) return true;  // "iptNN_" is six characters
    return false;
  };
  std::istringstream ss(output_field_string);
  std::string        req;
  while (ss >> req) {
    if (contains(node_components, req)) {
      output_node_component_labels_.push_back(req);
    } else if (contains(node_fields, req)) {
      for (auto const& c : component_labels(req, length_of[req])) output_node_component_labels_.push_back(c);
    } else if (contains(elem_components, req)) {
      for (int id : block_ids_)
        if (on_block(id, req, true)) derived_output_element_data_labels_[id].push_back(req);
    } else if (contains(ipt_components, req)) {
      for (int id : block_ids_)
        if (on_block(id, req, false)) output_element_component_labels_[id].push_back(req);
    } else if (contains(elem_fields, req)) {
      for (auto const& c : component_labels(req, length_of[req]))
        for (int id : block_ids_)
          if (on_block(id, c, true) && !contains(derived_output_element_data_labels_[id], c)) derived_output_element_data_labels_[id].push_back(c);
    } else if (contains(ipt_fields, req)) {
      for (auto const& c : component_labels(req, length_of[req]))
        for (int id : block_ids_)
          if (on_block(id, c, false)) output_element_component_labels_[id].push_back(c);
    } else if (req == "volume") {
      for (int id : block_ids_) derived_output_element_data_labels_[id].push_back(req);
    } else {
      throw std::invalid_argument("\nError:  ModelData::SpecifyOutputFields(), unable to process requested output \"" + req + "\".\n");
    }
  }
}

namespace {
// see RankGroup::SetLockstep
void
enter_exchange_call(DataManager& data_manager)
{
  auto vc    = data_manager.GetVectorCommunicator();
  auto group = vc ? vc->Group() : nullptr;
  if (group && group->Lockstep()) group->Barrier();
}
}  // namespace

void
ModelData::ComputeLumpedMass(DataManager& data_manager)
{
  enter_exchange_call(data_manager);
  // ModelData::ComputeLumpedMass (src/nimble_model_data.cc:495-530): mass from the reference configuration, critical
  // time step from the current one (displacement as it stands on the host), shared-node sum on the device
  DeviceContext& d = *device_;
  d.check(nsm_b200_upload_field(d.get(), NSM_FIELD_DISPLACEMENT, fields_.at(GetFieldIdChecked("displacement")).data),
          "ModelData::ComputeLumpedMass (displacement)");
  double dt = 0.0;
  d.check(nsm_b200_compute_lumped_mass(d.get(), &dt), "ModelData::ComputeLumpedMass");
  d.check(nsm_b200_download_field(d.get(), NSM_FIELD_LUMPED_MASS, fields_.at(GetFieldIdChecked("lumped_mass")).data),
          "ModelData::ComputeLumpedMass (download)");
  enter_exchange_call(data_manager);
  SetCriticalTimeStep(dt);
}

void
ModelData::ComputeExternalForce(DataManager&, double, double, bool)
{
  // the serial reference zeroes the external force (src/nimble_model_data.cc:609-618); the device treats an
  // external force that was never uploaded as identically zero
  const int id = GetFieldId("external_force");
  if (id >= 0) GetVectorNodeData(id).zero();
}

void
ModelData::ComputeInternalForce(DataManager& data_manager, double, double, bool is_output_step, const Viewify<2>& displacement,
                                Viewify<2>& force)
{
  DeviceContext& d = *device_;
  enter_exchange_call(data_manager);
  d.check(nsm_b200_internal_force_host(d.get(), displacement.data(), force.data(), is_output_step ? 1 : 0),
          "ModelData::ComputeInternalForce");
  enter_exchange_call(data_manager);
}

void
ModelData::ApplyKinematicConditions(DataManager& data_manager, double time_current, double time_previous)
{
  ModelDataBase::ApplyKinematicConditions(data_manager, time_current, time_previous);
}

void
ModelData::UpdateWithNewVelocity(DataManager&, double)
{
  DeviceContext& d = *device_;
  d.check(nsm_b200_upload_field(d.get(), NSM_FIELD_VELOCITY, fields_.at(GetFieldIdChecked("velocity")).data),
          "ModelData::UpdateWithNewVelocity");
}

void
ModelData::UpdateWithNewDisplacement(DataManager&, double)
{
  DeviceContext& d = *device_;
  d.check(nsm_b200_upload_field(d.get(), NSM_FIELD_DISPLACEMENT, fields_.at(GetFieldIdChecked("displacement")).data),
          "ModelData::UpdateWithNewDisplacement");
}

void
ModelData::PushNodalFields()
{
  DeviceContext& d = *device_;
  for (const char* label : {"displacement", "velocity", "acceleration"}) {
    const int id = GetFieldId(label);
    if (id < 0) continue;
    d.check(nsm_b200_upload_field_async(d.get(), device_field(label), fields_[id].data), "ModelData::PushNodalFields");
  }
  d.check(nsm_b200_sync(d.get()), "ModelData::PushNodalFields (sync)");
}

void
ModelData::PullNodalFields()
{
  DeviceContext& d = *device_;
  for (const char* label : {"displacement", "velocity", "acceleration", "internal_force", "contact_force"}) {
    const int id = GetFieldId(label);
    if (id < 0 || (std::string(label) == "contact_force" && !contact_on_device_)) continue;
    d.check(nsm_b200_download_field_async(d.get(), device_field(label), fields_[id].data), "ModelData::PullNodalFields");
  }
  d.check(nsm_b200_sync(d.get()), "ModelData::PullNodalFields (sync)");
}

void
ModelData::AdvanceOnDevice(DataManager& data_manager, int n_steps, double& time_current, double user_time_step, bool store_ipt_last)
{
  if (n_steps <= 0) return;
  DeviceContext& d  = *device_;
  auto           bc = data_manager.GetBoundaryConditionManager();
  const auto&    table = bc->GetDeviceTable();
  const int64_t  n_bc  = (int64_t)table.node.size();
  if (!bc_table_sent_) {
    d.check(nsm_b200_set_bc_table(d.get(), n_bc, table.node.data(), table.comp.data(), table.kind.data()),
            "ModelData::AdvanceOnDevice (boundary-condition table)");
    bc_table_sent_ = true;
  }
  if (n_bc > 0) {
    Viewify<2> X = GetVectorNodeData("reference_coordinate");
    const auto& programs = bc->GetDevicePrograms();
    if (programs.active) {
      // time-dependent magnitudes are evaluated on the device: one row of host magnitudes for the other entries
      // and, per step, the sub-expressions of t (accumulated like the loop does, :192-195)
      if (!bc_programs_sent_) {
        d.check(nsm_b200_set_bc_programs(d.get(), (int)programs.offsets.size() - 1, programs.offsets.data(), programs.code.data(),
                                         (int)programs.consts.size(), programs.consts.data(), (int)programs.slots.size(), n_bc,
                                         programs.program_of_entry.data()),
                "ModelData::AdvanceOnDevice (boundary-condition programs)");
        if (!programs.entry_constants.empty()) {
          std::vector<double> ec(programs.entry_constants.size() * (size_t)n_bc);
          bc->EvaluateEntryConstants(X, ec.data());
          d.check(nsm_b200_set_bc_entry_constants(d.get(), (int)programs.entry_constants.size(), n_bc, ec.data()),
                  "ModelData::AdvanceOnDevice (boundary-condition entry constants)");
        }
        bc_programs_sent_ = true;
      }
      bc_values_.resize((size_t)n_bc);
      bc->EvaluateMagnitudes(time_current, X, bc_values_.data(), true);
      d.check(nsm_b200_set_bc_values(d.get(), n_bc, bc_values_.data()), "ModelData::AdvanceOnDevice (magnitudes)");
      const size_t n_slots = programs.slots.size();
      bc_slots_.resize((size_t)n_steps * std::max<size_t>(n_slots, 1));
      double t = time_current;
      for (int s = 0; s < n_steps; ++s) {
        t += user_time_step;
        bc->EvaluateSlots(t, bc_slots_.data() + (size_t)s * n_slots);
      }
      d.check(nsm_b200_set_bc_slots_steps(d.get(), n_steps, (int)n_slots, bc_slots_.data()), "ModelData::AdvanceOnDevice (slots)");
    } else if (bc->HasTimeDependentMagnitudes()) {
      // one row per step, evaluated at the time_current of that step, accumulated like the loop does (:192-195)
      bc_values_.resize((size_t)n_steps * n_bc);
      double t = time_current;
      for (int s = 0; s < n_steps; ++s) {
        t += user_time_step;
        bc->EvaluateMagnitudes(t, X, bc_values_.data() + (size_t)s * n_bc);
      }
      d.check(nsm_b200_set_bc_values_steps(d.get(), n_steps, n_bc, bc_values_.data()), "ModelData::AdvanceOnDevice (magnitudes)");
    } else {
      bc_values_.resize((size_t)n_bc);
      bc->EvaluateMagnitudes(time_current, X, bc_values_.data());
      d.check(nsm_b200_set_bc_values(d.get(), n_bc, bc_values_.data()), "ModelData::AdvanceOnDevice (magnitudes)");
    }
  }
  enter_exchange_call(data_manager);
  // device time of the element kernels ("Force calculation") and of the node passes + shared-node exchange ("Time
  // Integration Scheme" / vector reduction), CUDA events on the context's stream: the figures of the timing summary
  d.check(nsm_b200_profile(d.get(), 1), "ModelData::AdvanceOnDevice (profile)");
  d.check(nsm_b200_step(d.get(), n_steps, &time_current, user_time_step, store_ipt_last ? 1 : 0), "ModelData::AdvanceOnDevice");
  double  elem_ms = 0.0, node_ms = 0.0;
  int64_t n_prof  = 0;
  d.check(nsm_b200_profile_read(d.get(), &elem_ms, &node_ms, &n_prof), "ModelData::AdvanceOnDevice (profile)");
  device_force_seconds_ += 1e-3 * elem_ms * (double)n_prof;
  device_update_seconds_ += 1e-3 * node_ms * (double)n_prof;
  if (contact_on_device_) {
    double contact_ms = 0.0;
    d.check(nsm_b200_profile_read_contact(d.get(), &contact_ms), "ModelData::AdvanceOnDevice (profile)");
    device_contact_seconds_ += 1e-3 * contact_ms * (double)n_prof;
  }
  enter_exchange_call(data_manager);
}

void
ModelData::UpdateStates(const DataManager&)
{
  DeviceContext& d = *device_;
  d.check(nsm_b200_update_states(d.get()), "ModelData::UpdateStates");
}

std::vector<double>&
ModelData::GetElementDataNew(int block_id)
{
  DeviceContext&       d = *device_;
  std::vector<double>& v = element_data_np1_[block_id];
  v.resize((size_t)nsm_b200_num_elements(d.get(), block_id) * 8 * nsm_b200_element_data_stride(d.get(), block_id));
  d.check(nsm_b200_get_element_data(d.get(), block_id, v.data()), "ModelData::GetElementDataNew");
  return v;
}

void
ModelData::InitializeExodusOutput(DataManager&)
{
}

void
ModelData::WriteExodusOutput(DataManager& data_manager, double time_current)
{
  // ModelData::WriteExodusOutput (src/nimble_model_data.cc:557-596): nodal components, per-point components and
  // the volume-averaged ("derived") components of every block, taken from the HOST mirrors (callers pull the
  // device state first) and from the device's integration-point data.
  DeviceContext&                   d = *device_;
  std::vector<double>              global_data;
  std::vector<std::vector<double>> node_out(output_node_component_labels_.size());
  for (size_t k = 0; k < output_node_component_labels_.size(); ++k) {
    const std::string& want = output_node_component_labels_[k];
    for (const Field& f : fields_) {
      auto comps = component_labels(f.label, f.length);
      for (size_t c = 0; c < comps.size(); ++c)
        if (comps[c] == want) {
          node_out[k].resize(f.num_objects);
          for (int i = 0; i < f.num_objects; ++i) node_out[k][i] = f.data[(size_t)i * (int)f.length + c];
        }
    }
  }
  std::map<int, std::vector<std::vector<double>>> elem_out, derived_out;
  // the derived kernel reads the device displacement: make it the one the caller sees on the host
  d.check(nsm_b200_upload_field(d.get(), NSM_FIELD_DISPLACEMENT, fields_.at(GetFieldIdChecked("displacement")).data),
          "ModelData::WriteExodusOutput (displacement)");
  for (int id : block_ids_) {
    const int64_t ne = nsm_b200_num_elements(d.get(), id);
    auto&         eo = elem_out[id];
    auto&         dv = derived_out[id];
    const auto&   want_e = output_element_component_labels_.at(id);
    const auto&   want_d = derived_output_element_data_labels_.at(id);
    if (!want_e.empty()) {
      // the requested columns are split out on the device; only they cross the bus
      const auto&          comps = element_component_labels_.at(id);
      std::vector<int32_t> offs(want_e.size());
      for (size_t k = 0; k < want_e.size(); ++k) offs[k] = (int32_t)(std::find(comps.begin(), comps.end(), want_e[k]) - comps.begin());
      std::vector<double> flat(want_e.size() * (size_t)ne);
      d.check(nsm_b200_get_element_components(d.get(), id, (int)offs.size(), offs.data(), flat.data()),
              "ModelData::WriteExodusOutput (integration-point components)");
      eo.resize(want_e.size());
      for (size_t k = 0; k < want_e.size(); ++k) eo[k].assign(flat.begin() + k * (size_t)ne, flat.begin() + (k + 1) * (size_t)ne);
    }
    if (!want_d.empty()) {
      const int           record = nsm_b200_element_data_stride(d.get(), id);
      std::vector<double> flat((size_t)(1 + record) * ne);
      d.check(nsm_b200_derived_element_data(d.get(), id, flat.data()), "ModelData::WriteExodusOutput (derived)");
      // rows: volume, then the block's per-point labels of point 1 without their prefix (F 9, sigma 6, state scalars)
      std::vector<std::string> order{"volume"};
      for (int k = 0; k < record; ++k) order.push_back(element_component_labels_.at(id)[k].substr(6));
      dv.resize(want_d.size());
      for (size_t k = 0; k < want_d.size(); ++k) {
        const size_t row = std::find(order.begin(), order.end(), want_d[k]) - order.begin();
        dv[k].assign(flat.begin() + row * ne, flat.begin() + (row + 1) * ne);
      }
    }
  }
  // the file I/O of this plane runs on the writer thread while the integrator goes on stepping
  data_manager.GetExodusOutput()->WriteStepAsync(time_current, std::move(global_data), std::move(node_out),
                                                 output_element_component_labels_, std::move(elem_out),
                                                 derived_output_element_data_labels_, std::move(derived_out));
}

}  // namespace nimble_b200
