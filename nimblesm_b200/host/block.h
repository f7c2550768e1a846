// nimblesm_b200/host/block.h — element-block layer with the surface of nimble::BlockBase / nimble::Block
// (src/nimble_block_base.h:57-119, src/nimble_block.h:60-171) and the stress seam BlockData /
// BlockMaterialInterface(Factory) (src/nimble_block_material_interface_base.h:52-84,
// src/nimble_block_material_interface_factory_base.h:59-72, src/nimble_kokkos_block_material_interface.{h,cc}).
//
// In the B200 build the model-level path (ModelData::ComputeInternalForce / the fused step) keeps ALL blocks
// resident in one device context.  The per-block entry points below keep the reference signatures (host arrays,
// one block per call) for callers and tests written against nimble::Block; they run on a block-private device
// context that is built on first use and reused while the caller passes the same mesh arrays.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "device.h"
#include "material.h"
#include "view.h"

namespace nimble_b200 {

enum Length { LENGTH_0 = 0, SCALAR = 1, VECTOR = 3, SYMMETRIC_TENSOR = 6, FULL_TENSOR = 9 };  // src/nimble_data_utils.h

class DataManager;

class BlockBase
{
 public:
  BlockBase()          = default;
  virtual ~BlockBase() = default;
  virtual void
  InstantiateElement() = 0;
  double
  GetDensity() const
  {
    return material_->GetDensity();
  }
  double
  GetBulkModulus() const
  {
    return material_->GetBulkModulus();
  }
  double
  GetShearModulus() const
  {
    return material_->GetShearModulus();
  }
  std::shared_ptr<Material>
  GetMaterialPointer() const
  {
    return material_;
  }
  // BlockBase::ComputeCriticalTimeStep (src/nimble_block_base.cc:51-84) on [n][3] host views
  virtual double
  ComputeCriticalTimeStep(const Viewify<2>& node_reference_coordinates, const Viewify<2>& node_displacements, int num_elem,
                          const int* elem_conn) const;
  // CUDA device of the per-block entry points below (the owning ModelData's; default 0)
  void
  SetDeviceIndex(int device)
  {
    device_index_ = device;
  }

 protected:
  int                       device_index_ = 0;
  std::string               model_material_parameters_ = "none";
  std::shared_ptr<Material> material_;
  // block-private device context for the per-block entry points
  struct Device;
  Device&
  device_for(const double* reference_coordinates, int num_elem, const int* elem_conn) const;
  mutable std::shared_ptr<Device> device_;
};

class Block : public BlockBase
{
 public:
  Block()           = default;
  ~Block() override = default;
  void
  Initialize(std::string const& model_material_parameters, MaterialFactoryBase& factory);
  void
  InstantiateMaterialModel(MaterialFactoryBase& factory);
  void
  InstantiateElement() override
  {
  }  // hex8 with 2x2x2 Gauss points is compiled into the kernels (src/nimble_block.cc:78-82)
  int
  NumIntegrationPointsPerElement() const
  {
    return 8;
  }
  int
  NumNodesPerElement() const
  {
    return 8;
  }
  // iptNN_deformation_gradient (FULL_TENSOR), iptNN_stress (SYMMETRIC_TENSOR), then the material's state variables
  // (SCALAR) for NN = 01..08 (src/nimble_block.cc:84-108)
  void
  GetDataLabelsAndLengths(std::vector<std::pair<std::string, Length>>& data_labels_and_lengths) const;
  // lumped_mass[node] += element contribution (src/nimble_block.cc:110-146)
  void
  ComputeLumpedMassMatrix(const double* reference_coordinates, int num_elem, const int* elem_conn, double* lumped_mass) const;
  // doubles per integration point: 15 + the material's state variables
  int
  NumDataPerIntegrationPoint() const
  {
    return 15 + material_->NumStateVariables();
  }
  // F = identity, sigma = 0, state variables at their initial values, in both states (src/nimble_block.cc:148-207)
  void
  InitializeElementData(int num_elem_in_block, std::vector<double>& elem_data_n, std::vector<double>& elem_data_np1) const;
  // internal_force[3*node+i] += f; elem_data_np1 receives F / sigma / state of every integration point; a material
  // with state variables reads F_n / sigma_n / state_n from elem_data_n (src/nimble_block.cc:297-368, 388-436).
  // reference_coordinates / displacement / internal_force are [n][3] AoS.
  void
  ComputeInternalForce(const double* reference_coordinates, const double* displacement, const double* velocity,
                       double* internal_force, double time_previous, double time_current, int num_elem, const int* elem_conn,
                       const int* elem_global_ids, std::vector<std::string> const& elem_data_labels,
                       std::vector<double> const& elem_data_n, std::vector<double>& elem_data_np1, DataManager* data_manager,
                       bool is_output_step, bool compute_stress_only = false) const;
  // volume and volume averages of the integration-point fields (src/nimble_block.cc:438-497);
  // derived_elem_data[k][elem], k: 0 = volume, 1..9 = F components, 10..15 = sigma components, 16.. = state variables
  void
  ComputeDerivedElementData(const double* reference_coordinates, const double* displacement, int num_elem, const int* elem_conn,
                            std::vector<double> const& elem_data_np1, std::vector<std::vector<double>>& derived_elem_data) const;
};

// ---- stress seam ---------------------------------------------------------------------------------------
struct BlockData
{
  BlockData(BlockBase* block_, Material* material_d_, const int block_id_, const int num_block_elems_,
            const int num_points_per_block_elem_)
      : block(block_), material_device(material_d_), id(block_id_), num_elems(num_block_elems_),
        num_points_per_elem(num_points_per_block_elem_)
  {
  }
  BlockBase* block               = nullptr;
  Material*  material_device     = nullptr;  // descriptor; the arithmetic is the device kernel selected by its kind
  int        id                  = 0;
  int        num_elems           = 0;
  int        num_points_per_elem = 0;
};

class BlockMaterialInterfaceBase
{
 public:
  virtual ~BlockMaterialInterfaceBase() = default;
  virtual void
  ComputeStress() const = 0;
};

class ModelDataBase;
struct FieldIds;

// MDRange(elem, ipt) -> Material::GetStress of the reference (src/nimble_kokkos_block_material_interface.cc:65-119,
// 138-194): here one stress_kernel launch per block over its [num_elems * num_points][9] deformation gradients.
class BlockMaterialInterface : public BlockMaterialInterfaceBase
{
 public:
  struct Arrays
  {
    const double* deformation_gradient_np1;  // [num_elems][num_points][9]
    double*       stress_np1;                // [num_elems][num_points][6]
    // read by materials with state variables only (the other views of compute_block_stress,
    // src/nimble_kokkos_block_material_interface.cc:86-118)
    const double* deformation_gradient_n = nullptr;  // [num_elems][num_points][9]
    const double* stress_n               = nullptr;  // [num_elems][num_points][6]
    const double* state_n                = nullptr;  // [num_elems][num_points][n_state]
    double*       state_np1              = nullptr;
  };
  BlockMaterialInterface(double time_n_, double time_np1_, const std::vector<BlockData>& blocks_,
                         const std::map<int, Arrays>& arrays_, DeviceContext& device_)
      : time_n(time_n_), time_np1(time_np1_), blocks(blocks_), arrays(arrays_), device(device_)
  {
  }
  void
  ComputeStress() const override;

 protected:
  double                 time_n, time_np1;
  std::vector<BlockData> blocks;
  std::map<int, Arrays>  arrays;
  DeviceContext&         device;
};

class BlockMaterialInterfaceFactoryBase
{
 public:
  virtual ~BlockMaterialInterfaceFactoryBase() = default;
  virtual std::shared_ptr<BlockMaterialInterfaceBase>
  create(double time_n, double time_np1, const FieldIds& field_ids, const std::vector<BlockData>& blocks,
         ModelDataBase* model_data_ptr) const = 0;
};

}  // namespace nimble_b200
