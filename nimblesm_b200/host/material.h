// nimblesm_b200/host/material.h — host-side material description and factories with the surface of
// nimble::MaterialParameters / Material / MaterialFactoryBase / MaterialFactory
// (src/nimble_material.h:61-314, src/nimble_material_factory_base.{h,cc}, src/nimble_material_factory.{h,cc}).
//
// A Material here is a DESCRIPTOR (model kind + moduli + density): the constitutive arithmetic lives in the
// CUDA kernels (csrc/hex8_math.cuh: stress_elastic / stress_neohookean).  GetStress keeps the reference's
// host-array signature and runs the points through the device stress seam (nsm_b200_compute_stress).
#pragma once
#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/nsm_b200.h"

namespace nimble_b200 {

class DeviceContext;

class MaterialParameters
{
 public:
  static const int MAX_NUM_MAT_PARAM     = 64;  // src/nimble_material.h:64-65
  static const int MAX_MAT_MODEL_STR_LEN = 64;
  MaterialParameters() = default;
  MaterialParameters(const std::string& material_name, const std::map<std::string, std::string>& string_params,
                     const std::map<std::string, double>& double_params, int num_material_points = 0)
      : material_name_(material_name), string_params_(string_params), double_params_(double_params),
        num_material_points_(num_material_points)
  {
  }
  void
  AddParameter(const char* name, double value)
  {
    double_params_[name] = value;
  }
  void
  AddStringParameter(const char* name, const char* value)
  {
    string_params_[name] = value;
  }
  bool
  IsParameter(const char* name) const
  {
    return double_params_.count(name) != 0;
  }
  bool
  IsStringParameter(const char* name) const
  {
    return string_params_.count(name) != 0;
  }
  std::string
  GetMaterialName(bool upper_case = false) const;
  int
  GetNumParameters() const
  {
    return (int)double_params_.size();
  }
  double
  GetParameterValue(const char* name) const;  // throws std::invalid_argument when absent
  int
  GetNumStringParameters() const
  {
    return (int)string_params_.size();
  }
  const std::string&
  GetStringParameterValue(const char* name) const;  // throws std::invalid_argument when absent
  const std::map<std::string, double>&
  GetParameters() const
  {
    return double_params_;
  }
  const std::map<std::string, std::string>&
  GetStringParameters() const
  {
    return string_params_;
  }
  int
  GetNumMaterialPoints() const
  {
    return num_material_points_;
  }

 private:
  std::string                        material_name_;
  std::map<std::string, std::string> string_params_;
  std::map<std::string, double>      double_params_;
  int                                num_material_points_ = 0;
};

class Material
{
 public:
  Material(const MaterialParameters& params, nsm_material_kind kind);
  virtual ~Material() = default;
  virtual bool
  IsNGPLAMEModel() const
  {
    return false;
  }
  // Material::NumStateVariables / GetStateVariableLabel / GetStateVariableInitialValue (src/nimble_material.h:214-225):
  // answered by the device library for the material kind (0 for the reference's two models, src/nimble_material.cc:60,218)
  virtual int
  NumStateVariables() const
  {
    return nsm_b200_material_num_state(kind_);
  }
  virtual void
  GetStateVariableLabel(int index, char label[MaterialParameters::MAX_MAT_MODEL_STR_LEN]) const;
  virtual double
  GetStateVariableInitialValue(int index) const
  {
    return nsm_b200_material_state_initial_value(kind_, index);
  }
  // {bulk_modulus, shear_modulus, density, model-specific...} in the order nsm_b200_add_block_params takes them
  virtual std::vector<double>
  DeviceParameters() const
  {
    return {bulk_modulus_, shear_modulus_, density_};
  }
  double
  GetDensity() const
  {
    return density_;
  }
  double
  GetBulkModulus() const
  {
    return bulk_modulus_;
  }
  double
  GetShearModulus() const
  {
    return shear_modulus_;
  }
  nsm_material_kind
  Kind() const
  {
    return kind_;
  }
  const MaterialParameters&
  Parameters() const
  {
    return params_;
  }
  // Material::GetStress (src/nimble_material.h:241-256): (F_n, F_np1, sigma_n, state_n) [num_pts][9 / 9 / 6 / n_state]
  // -> (stress_np1 [num_pts][6], state_np1); the reference's two models ignore the N arguments, a material with
  // state variables reads them all.
  void
  GetStress(int elem_id, int num_pts, double time_previous, double time_current, const double* deformation_gradient_n,
            const double* deformation_gradient_np1, const double* stress_n, double* stress_np1, const double* state_data_n,
            double* state_data_np1, DeviceContext& device, bool is_output_step = false) const;

 protected:
  MaterialParameters params_;
  nsm_material_kind  kind_;
  double             density_ = 0.0, bulk_modulus_ = 0.0, shear_modulus_ = 0.0;
};

class ElasticMaterial : public Material
{
 public:
  explicit ElasticMaterial(const MaterialParameters& p) : Material(p, NSM_MAT_ELASTIC) {}
};

class NeohookeanMaterial : public Material
{
 public:
  explicit NeohookeanMaterial(const MaterialParameters& p) : Material(p, NSM_MAT_NEOHOOKEAN) {}
};

// The history-dependent model of the state-variable slot (include/nsm_b200.h, NSM_MAT_J2_PLASTICITY): small-strain
// J2 plasticity with linear isotropic hardening; parameters density, bulk_modulus, shear_modulus, yield_stress,
// hardening_modulus; state equivalent_plastic_strain, von_mises_stress.
class J2PlasticityMaterial : public Material
{
 public:
  explicit J2PlasticityMaterial(const MaterialParameters& p)
      : Material(p, NSM_MAT_J2_PLASTICITY), yield_stress_(p.GetParameterValue("yield_stress")), hardening_modulus_(p.GetParameterValue("hardening_modulus"))
  {
  }
  std::vector<double>
  DeviceParameters() const override
  {
    return {bulk_modulus_, shear_modulus_, density_, yield_stress_, hardening_modulus_};
  }

 private:
  double yield_stress_, hardening_modulus_;
};

class MaterialFactoryBase
{
 public:
  MaterialFactoryBase();
  virtual ~MaterialFactoryBase() = default;
  void
  add_valid_double_parameter_name(const char* name)
  {
    if (std::find(valid_double_parameter_names_.begin(), valid_double_parameter_names_.end(), name) ==
        valid_double_parameter_names_.end())
      valid_double_parameter_names_.push_back(name);
  }
  void
  add_valid_string_parameter_name(const char* name)
  {
    if (std::find(valid_string_parameter_names_.begin(), valid_string_parameter_names_.end(), name) ==
        valid_string_parameter_names_.end())
      valid_string_parameter_names_.push_back(name);
  }
  virtual std::shared_ptr<Material>
  get_material() const
  {
    return material;
  }
  virtual void
  parse_and_create(const std::string& mat_params, int num_points)
  {
    material_params = ParseMaterialParametersString(mat_params, num_points);
    create();
  }
  virtual void
  parse_and_create(const std::string& mat_params)
  {
    parse_and_create(mat_params, 0);
  }
  virtual std::map<std::string, double>
  parse_material_params_string(const std::string& mat_params)
  {
    return ParseMaterialParametersString(mat_params, 0)->GetParameters();
  }

 protected:
  virtual void
  create() = 0;
  std::shared_ptr<MaterialParameters>
  ParseMaterialParametersString(const std::string& material_parameters, int num_material_points = 0) const;
  std::shared_ptr<Material>                 material;
  std::shared_ptr<const MaterialParameters> material_params;

 private:
  std::vector<std::string> valid_double_parameter_names_;
  std::vector<std::string> valid_string_parameter_names_;
};

class MaterialFactory : public MaterialFactoryBase
{
 public:
  MaterialFactory() = default;

 protected:
  void
  create() override;
};

}  // namespace nimble_b200
