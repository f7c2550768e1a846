// nimblesm_b200/host/block.cc — see block.h.
#include "block.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace nimble_b200 {

// One block alone on the device: all nodes the caller's coordinate array can be indexed with.
struct BlockBase::Device
{
  DeviceContext       ctx;
  const double*       key_coords = nullptr;
  const int*          key_conn   = nullptr;
  int                 key_elems  = 0;
  int                 num_nodes  = 0;
  std::vector<double> scratch;
  explicit Device(int dev) : ctx(dev) {}
};

BlockBase::Device&
BlockBase::device_for(const double* X, int num_elem, const int* conn) const
{
  if (device_ && device_->key_coords == X && device_->key_conn == conn && device_->key_elems == num_elem) return *device_;
  device_ = std::make_shared<Device>(0);
  Device& d = *device_;
  d.key_coords = X, d.key_conn = conn, d.key_elems = num_elem;
  int max_node = -1;
  for (long i = 0; i < (long)num_elem * 8; ++i) max_node = std::max(max_node, conn[i]);
  d.num_nodes = max_node + 1;
  std::vector<double> x(d.num_nodes), y(d.num_nodes), z(d.num_nodes);
  for (int n = 0; n < d.num_nodes; ++n) x[n] = X[3 * n], y[n] = X[3 * n + 1], z[n] = X[3 * n + 2];
  d.ctx.check(nsm_b200_set_nodes(d.ctx.get(), d.num_nodes, x.data(), y.data(), z.data()), "Block: set_nodes");
  d.ctx.check(nsm_b200_add_block(d.ctx.get(), 1, num_elem, conn, material_->Kind(), material_->GetBulkModulus(),
                                 material_->GetShearModulus(), material_->GetDensity()),
              "Block: add_block");
  // ORDERED assembly: nodal sums in ascending element order == the serial loop of the reference (src/nimble_block.cc:434)
  d.ctx.check(nsm_b200_finalize(d.ctx.get(), NSM_ASSEMBLY_ORDERED, 0), "Block: finalize");
  d.scratch.assign((size_t)d.num_nodes * 3, 0.0);
  return d;
}

double
BlockBase::ComputeCriticalTimeStep(const Viewify<2>& X, const Viewify<2>& u, int num_elem, const int* elem_conn) const
{
  Device& d = device_for(X.data(), num_elem, elem_conn);
  d.ctx.check(nsm_b200_upload_field(d.ctx.get(), NSM_FIELD_DISPLACEMENT, u.data()), "Block: upload displacement");
  double dt = 0.0;
  d.ctx.check(nsm_b200_compute_lumped_mass(d.ctx.get(), &dt), "Block: critical time step");
  return dt;
}

void
Block::Initialize(std::string const& model_material_parameters, MaterialFactoryBase& factory)
{
  model_material_parameters_ = model_material_parameters;
  InstantiateMaterialModel(factory);
  InstantiateElement();
}

void
Block::InstantiateMaterialModel(MaterialFactoryBase& factory)
{
  factory.parse_and_create(model_material_parameters_, NumIntegrationPointsPerElement());
  material_ = factory.get_material();
}

void
Block::GetDataLabelsAndLengths(std::vector<std::pair<std::string, Length>>& out) const
{
  for (int ipt = 1; ipt <= NumIntegrationPointsPerElement(); ++ipt) {
    char prefix[16];
    snprintf(prefix, sizeof prefix, "ipt%02d_", ipt);  // AddIntegrationPointPrefix (src/nimble_data_utils.cc)
    out.emplace_back(std::string(prefix) + "deformation_gradient", FULL_TENSOR);
    out.emplace_back(std::string(prefix) + "stress", SYMMETRIC_TENSOR);
  }
}

void
Block::ComputeLumpedMassMatrix(const double* X, int num_elem, const int* elem_conn, double* lumped_mass) const
{
  Device& d = device_for(X, num_elem, elem_conn);
  std::fill(d.scratch.begin(), d.scratch.end(), 0.0);
  d.ctx.check(nsm_b200_upload_field(d.ctx.get(), NSM_FIELD_DISPLACEMENT, d.scratch.data()), "Block: upload displacement");
  d.ctx.check(nsm_b200_compute_lumped_mass(d.ctx.get(), nullptr), "Block::ComputeLumpedMassMatrix");
  std::vector<double> m(d.num_nodes);
  d.ctx.check(nsm_b200_download_field(d.ctx.get(), NSM_FIELD_LUMPED_MASS, m.data()), "Block: download lumped mass");
  for (int n = 0; n < d.num_nodes; ++n) lumped_mass[n] += m[n];
}

void
Block::InitializeElementData(int num_elem_in_block, std::vector<double>& elem_data_n, std::vector<double>& elem_data_np1) const
{
  const size_t per_elem = 8 * 15;
  elem_data_n.assign(num_elem_in_block * per_elem, 0.0);
  for (int e = 0; e < num_elem_in_block; ++e)
    for (int q = 0; q < 8; ++q)
      for (int k = 0; k < 3; ++k) elem_data_n[e * per_elem + q * 15 + k] = 1.0;  // F = identity (xx, yy, zz first)
  elem_data_np1 = elem_data_n;
}

void
Block::ComputeInternalForce(const double* X, const double* displacement, const double*, double* internal_force, double, double,
                            int num_elem, const int* elem_conn, const int*, std::vector<std::string> const&,
                            std::vector<double> const&, std::vector<double>& elem_data_np1, DataManager*, bool,
                            bool compute_stress_only) const
{
  Device& d = device_for(X, num_elem, elem_conn);
  // the reference stores F and sigma of every point on every call (src/nimble_block.cc:355-368)
  d.ctx.check(nsm_b200_internal_force_host(d.ctx.get(), displacement, d.scratch.data(), 1), "Block::ComputeInternalForce");
  if (!compute_stress_only)
    for (size_t i = 0; i < d.scratch.size(); ++i) internal_force[i] += d.scratch[i];
  elem_data_np1.resize((size_t)num_elem * 120);
  d.ctx.check(nsm_b200_get_element_data(d.ctx.get(), 1, elem_data_np1.data()), "Block: element data");
}

void
Block::ComputeDerivedElementData(const double* X, const double* displacement, int num_elem, const int* elem_conn,
                                 std::vector<double> const&, std::vector<std::vector<double>>& derived) const
{
  // uses the integration-point data of the last ComputeInternalForce on this block (the reference passes the
  // same values back in through elem_data_np1)
  Device& d = device_for(X, num_elem, elem_conn);
  d.ctx.check(nsm_b200_upload_field(d.ctx.get(), NSM_FIELD_DISPLACEMENT, displacement), "Block: upload displacement");
  std::vector<double> flat((size_t)16 * num_elem);
  d.ctx.check(nsm_b200_derived_element_data(d.ctx.get(), 1, flat.data()), "Block::ComputeDerivedElementData");
  derived.assign(16, std::vector<double>());
  for (int k = 0; k < 16; ++k) derived[k].assign(flat.begin() + (size_t)k * num_elem, flat.begin() + (size_t)(k + 1) * num_elem);
}

void
BlockMaterialInterface::ComputeStress() const
{
  for (const BlockData& b : blocks) {
    auto it = arrays.find(b.id);
    if (it == arrays.end() || b.num_elems == 0) continue;
    const Material* m = b.material_device;
    device.check(nsm_b200_compute_stress(device.get(), m->Kind(), m->GetBulkModulus(), m->GetShearModulus(),
                                         (int64_t)b.num_elems * b.num_points_per_elem, it->second.deformation_gradient_np1,
                                         it->second.stress_np1),
                 "BlockMaterialInterface::ComputeStress");
  }
}

}  // namespace nimble_b200
