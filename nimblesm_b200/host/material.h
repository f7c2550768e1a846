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
  bool
  IsParameter(const char* name) const
  {
    return double_params_.count(name) != 0;
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
  const std::map<std::string, double>&
  GetParameters() const
  {
    return double_params_;
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
  virtual int
  NumStateVariables() const
  {
    return 0;  // src/nimble_material.cc:60,218
  }
  virtual void
  GetStateVariableLabel(int, char*) const
  {
  }
  virtual double
  GetStateVariableInitialValue(int) const
  {
    return 0.0;
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
  // Material::GetStress (src/nimble_material.h:196-214): deformation_gradient_np1 [num_pts][9] ->
  // stress_np1 [num_pts][6]; the N-state arguments are accepted and unused, as in both reference models.
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
