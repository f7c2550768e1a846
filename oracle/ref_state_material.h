// oracle/ref_state_material.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A history-dependent nimble::Material for the checker build (oracle/_ref/libnimble_ref.so).  The reference ships
// no material that carries state variables (src/nimble_material.cc:60,218: both report 0), yet its block functor,
// element-data containers and ModelData::UpdateStates carry per-integration-point state N / N+1 for any material
// that does (src/nimble_block.cc:84-108, 168-183, 297-307, 324-337, 355-368; src/nimble_model_data.h:104-107).
// To make that plumbing the ORACLE of the B200 state-variable slot, this subclass plugs into the UNMODIFIED
// nimble_block.cc / nimble_model_data.cc through the reference's own extension points (Material virtuals,
// MaterialFactoryBase::create / add_valid_double_parameter_name).
//
// Model "j2_plasticity": small-strain J2 plasticity with linear isotropic hardening in incremental form; it reads
// all four inputs of Material::GetStress (F_n, F_np1, sigma_n, state_n).  Parameters: density, bulk_modulus,
// shear_modulus, yield_stress, hardening_modulus.  State: equivalent_plastic_strain, von_mises_stress.  The
// operation sequence below is the contract the plain-C oracle (oracle/hex8_oracle.c: h8o_stress_j2) and the device
// code (nimblesm_b200/csrc/hex8_math.cuh: stress_j2) reproduce bit for bit.
#pragma once
#include <memory>

#include "nimble_material.h"
#include "nimble_material_factory.h"

namespace nsm_oracle {

class J2PlasticityMaterial : public nimble::Material
{
 public:
  explicit J2PlasticityMaterial(nimble::MaterialParameters const& p);
  int
  NumStateVariables() const override
  {
    return 2;
  }
  void
  GetStateVariableLabel(int index, char label[nimble::MaterialParameters::MAX_MAT_MODEL_STR_LEN]) const override;
  double
  GetStateVariableInitialValue(int) const override
  {
    return 0.0;
  }
  double
  GetDensity() const override
  {
    return density_;
  }
  double
  GetBulkModulus() const override
  {
    return bulk_modulus_;
  }
  double
  GetShearModulus() const override
  {
    return shear_modulus_;
  }
  void
  GetStress(int elem_id, int num_pts, double time_previous, double time_current, const double* deformation_gradient_n,
            const double* deformation_gradient_np1, const double* stress_n, double* stress_np1, const double* state_data_n,
            double* state_data_np1, nimble::DataManager& data_manager, bool is_output_step) override;
  void
  GetTangent(int num_pts, double* material_tangent) const override;

 protected:
  void
  GetStress(double, double, nimble::Viewify<1, const double>&, nimble::Viewify<1, const double>&, nimble::Viewify<1, const double>&,
            nimble::Viewify<1>) const override;

 private:
  double density_, bulk_modulus_, shear_modulus_, yield_stress_, hardening_modulus_;
};

// nimble::MaterialFactory + "j2_plasticity"
class StateMaterialFactory : public nimble::MaterialFactory
{
 public:
  StateMaterialFactory();

 protected:
  void
  create() override;
};

}  // namespace nsm_oracle
