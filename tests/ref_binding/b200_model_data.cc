// tests/ref_binding/b200_model_data.cc — see b200_model_data.h.
#include "b200_model_data.h"

#include <sstream>
#include <stdexcept>

#include "nimble_data_manager.h"
#include "nimble_genesis_mesh.h"
#include "nimble_material.h"
#include "nimble_material_factory_base.h"
#include "nimble_parser.h"
#include "nsm_b200.h"

namespace nsm_binding {

B200ModelData::B200ModelData(int device, int assembly, unsigned flags) : assembly_(assembly), flags_(flags)
{
  if (nsm_b200_create(device, &ctx_)) throw std::runtime_error(std::string("nsm_b200_create: ") + nsm_b200_last_error(nullptr));
}

B200ModelData::~B200ModelData()
{
  nsm_b200_destroy(ctx_);
}

void
B200ModelData::check(int status, const char* what) const
{
  // NIMBLE_ABORT would call abort(); the test harness wants the message
  if (status) throw std::runtime_error(std::string(what) + ": " + nsm_b200_last_error(ctx_));
}

void
B200ModelData::InitializeBlocks(nimble::DataManager& data_manager, const std::shared_ptr<nimble::MaterialFactoryBase>& factory)
{
  // blocks, materials, element-data labels and containers, output selection: the reference's own code
  nimble::ModelData::InitializeBlocks(data_manager, factory);
  // ... and the same model on the device
  const nimble::GenesisMesh& mesh   = data_manager.GetMesh();
  const nimble::Parser&      parser = data_manager.GetParser();
  check(nsm_b200_set_nodes(ctx_, (int64_t)mesh.GetNumNodes(), mesh.GetCoordinatesX(), mesh.GetCoordinatesY(), mesh.GetCoordinatesZ()),
        "nsm_b200_set_nodes");
  for (int id : mesh.GetBlockIds()) {
    const std::string& text = parser.GetModelMaterialParameters(id);
    std::istringstream in(text);
    std::string        name;
    in >> name;
    auto                values = factory->parse_material_params_string(text);  // the reference's own parser / key validation
    std::vector<double> params = {values.at("bulk_modulus"), values.at("shear_modulus"), values.at("density")};
    int                 kind;
    if (name == "elastic")
      kind = NSM_MAT_ELASTIC;
    else if (name == "neohookean")
      kind = NSM_MAT_NEOHOOKEAN;
    else if (name == "j2_plasticity") {
      kind = NSM_MAT_J2_PLASTICITY;
      params.push_back(values.at("yield_stress"));
      params.push_back(values.at("hardening_modulus"));
    } else
      throw std::invalid_argument("B200ModelData: no device kernel for material model " + name);
    // the label order and record size must be the reference's (src/nimble_block.cc:84-108)
    if ((int)GetElementDataLabels().at(id).size() != 8 * (15 + nsm_b200_material_num_state(kind)))
      throw std::logic_error("B200ModelData: element-data record size differs from the reference's for block " + std::to_string(id));
    check(nsm_b200_add_block_params(ctx_, id, mesh.GetNumElementsInBlock(id), mesh.GetConnectivity(id), kind, (int)params.size(),
                                    params.data()),
          "nsm_b200_add_block_params");
  }
  check(nsm_b200_finalize(ctx_, assembly_, flags_), "nsm_b200_finalize");
}

void
B200ModelData::ComputeLumpedMass(nimble::DataManager& data_manager)
{
  // src/nimble_model_data.cc:495-530: mass from the reference configuration, critical time step from the current one
  auto displacement = GetVectorNodeData("displacement");
  check(nsm_b200_upload_field(ctx_, NSM_FIELD_DISPLACEMENT, displacement.data()), "upload displacement");
  double dt = 0.0;
  check(nsm_b200_compute_lumped_mass(ctx_, &dt), "nsm_b200_compute_lumped_mass");
  check(nsm_b200_download_field(ctx_, NSM_FIELD_LUMPED_MASS, GetNodeData(GetFieldId("lumped_mass"))), "download lumped_mass");
  SetCriticalTimeStep(dt);
}

void
B200ModelData::ComputeInternalForce(nimble::DataManager&, double, double, bool is_output_step, const nimble::Viewify<2>& displacement,
                                    nimble::Viewify<2>& force)
{
  // src/nimble_model_data.cc:620-667.  The Viewify<2> arguments are the integrator's [n][3] host views.
  check(nsm_b200_internal_force_host(ctx_, displacement.data(), force.data(), is_output_step ? 1 : 0), "nsm_b200_internal_force_host");
  if (is_output_step) PullElementData();
}

void
B200ModelData::PullElementData()
{
  for (auto& kv : GetBlocks()) {
    std::vector<double>& np1 = GetElementDataNew(kv.first);
    check(nsm_b200_get_element_data(ctx_, kv.first, np1.data()), "nsm_b200_get_element_data");
  }
}

void
B200ModelData::UpdateStates(const nimble::DataManager& data_manager)
{
  nimble::ModelData::UpdateStates(data_manager);  // the host containers' swap (src/nimble_model_data.h:104-107)
  check(nsm_b200_update_states(ctx_), "nsm_b200_update_states");
}

long
B200ModelData::DeviceLaunches() const
{
  return (long)nsm_b200_launch_count(ctx_);
}

}  // namespace nsm_binding
