// nimblesm_b200/host/block.cc — see block.h.
#include "block.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace nimble_b200 {

// One block alone on the device: all nodes the caller's coordinate array can be indexed with.  The device model is
// rebuilt whenever the CONTENT of the coordinates or of the connectivity changes (a caller may mutate its arrays in
// place, or reuse a freed buffer for another mesh of the same size): the key is a hash of both arrays.
struct BlockBase::Device
{
  DeviceContext       ctx;
  uint64_t            key       = 0;
  int                 key_elems = -1;
  int                 num_nodes = 0;
  std::vector<double> scratch;
  explicit Device(int dev) : ctx(dev) {}
};

namespace {
uint64_t
fnv1a(const void* data, size_t bytes, uint64_t h)
{
  const unsigned char* p = (const unsigned char*)data;
  for (size_t i = 0; i < bytes; ++i) h = (h ^ p[i]) * 1099511628211ULL;
  return h;
}
}  // namespace

BlockBase::Device&
BlockBase::device_for(const double* X, int num_elem, const int* conn) const
{
  int max_node = -1;
  for (long i = 0; i < (long)num_elem * 8; ++i) max_node = std::max(max_node, conn[i]);
  const int num_nodes = max_node + 1;
  uint64_t  key       = fnv1a(conn, (size_t)num_elem * 8 * sizeof(int), 1469598103934665603ULL);
  key                 = fnv1a(X, (size_t)num_nodes * 3 * sizeof(double), key);
  if (device_ && device_->key == key && device_->key_elems == num_elem && device_->num_nodes == num_nodes) return *device_;
  device_ = std::make_shared<Device>(device_index_);
  Device& d = *device_;
  d.key = key, d.key_elems = num_elem;
  d.num_nodes = num_nodes;
  std::vector<double> x(d.num_nodes), y(d.num_nodes), z(d.num_nodes);
  for (int n = 0; n < d.num_nodes; ++n) x[n] = X[3 * n], y[n] = X[3 * n + 1], z[n] = X[3 * n + 2];
  d.ctx.check(nsm_b200_set_nodes(d.ctx.get(), d.num_nodes, x.data(), y.data(), z.data()), "Block: set_nodes");
  const std::vector<double> params = material_->DeviceParameters();
  d.ctx.check(nsm_b200_add_block_params(d.ctx.get(), 1, num_elem, conn, material_->Kind(), (int)params.size(), params.data()),
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
    for (int i = 0; i < material_->NumStateVariables(); ++i) {
      char label[MaterialParameters::MAX_MAT_MODEL_STR_LEN];
      material_->GetStateVariableLabel(i, label);
      out.emplace_back(std::string(prefix) + label, SCALAR);
    }
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
  const int    record   = NumDataPerIntegrationPoint();
  const size_t per_elem = 8 * (size_t)record;
  elem_data_n.assign(num_elem_in_block * per_elem, 0.0);
  for (int e = 0; e < num_elem_in_block; ++e)
    for (int q = 0; q < 8; ++q) {
      for (int k = 0; k < 3; ++k) elem_data_n[e * per_elem + q * record + k] = 1.0;  // F = identity (xx, yy, zz first)
      for (int k = 15; k < record; ++k) elem_data_n[e * per_elem + q * record + k] = material_->GetStateVariableInitialValue(k - 15);
    }
  elem_data_np1 = elem_data_n;
}

void
Block::ComputeInternalForce(const double* X, const double* displacement, const double*, double* internal_force, double, double,
                            int num_elem, const int* elem_conn, const int*, std::vector<std::string> const&,
                            std::vector<double> const& elem_data_n, std::vector<double>& elem_data_np1, DataManager*, bool,
                            bool compute_stress_only) const
{
  Device& d = device_for(X, num_elem, elem_conn);
  const size_t per_elem = 8 * (size_t)NumDataPerIntegrationPoint();
  if (material_->NumStateVariables() > 0) {  // F_n, sigma_n, state_n come from the caller's N container, as in the reference
    if (elem_data_n.size() != (size_t)num_elem * per_elem) throw std::invalid_argument("Block::ComputeInternalForce: elem_data_n has the wrong size");
    d.ctx.check(nsm_b200_set_element_data(d.ctx.get(), 1, 1, elem_data_n.data()), "Block: previous element data");
  }
  // the reference stores F and sigma of every point on every call (src/nimble_block.cc:355-368)
  d.ctx.check(nsm_b200_internal_force_host(d.ctx.get(), displacement, d.scratch.data(), 1), "Block::ComputeInternalForce");
  if (!compute_stress_only)
    for (size_t i = 0; i < d.scratch.size(); ++i) internal_force[i] += d.scratch[i];
  elem_data_np1.resize((size_t)num_elem * per_elem);
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
  const int           rows = 1 + NumDataPerIntegrationPoint();
  std::vector<double> flat((size_t)rows * num_elem);
  d.ctx.check(nsm_b200_derived_element_data(d.ctx.get(), 1, flat.data()), "Block::ComputeDerivedElementData");
  derived.assign(rows, std::vector<double>());
  for (int k = 0; k < rows; ++k) derived[k].assign(flat.begin() + (size_t)k * num_elem, flat.begin() + (size_t)(k + 1) * num_elem);
}

void
BlockMaterialInterface::ComputeStress() const
{
  for (const BlockData& b : blocks) {
    auto it = arrays.find(b.id);
    if (it == arrays.end() || b.num_elems == 0) continue;
    const Material*           m      = b.material_device;
    const std::vector<double> params = m->DeviceParameters();
    const Arrays&             a      = it->second;
    device.check(nsm_b200_compute_stress_state(device.get(), m->Kind(), (int)params.size(), params.data(),
                                               (int64_t)b.num_elems * b.num_points_per_elem, a.deformation_gradient_n,
                                               a.deformation_gradient_np1, a.stress_n, a.state_n, a.stress_np1, a.state_np1),
                 "BlockMaterialInterface::ComputeStress");
  }
}

}  // namespace nimble_b200
