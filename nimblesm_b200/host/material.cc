// nimblesm_b200/host/material.cc — see material.h.
#include "material.h"

#include <sstream>
#include <stdexcept>

#include "device.h"

namespace nimble_b200 {

std::string
MaterialParameters::GetMaterialName(bool upper_case) const
{
  std::string name = material_name_;
  if (upper_case) std::transform(name.begin(), name.end(), name.begin(), [](unsigned char c) { return (char)std::toupper(c); });
  return name;
}

double
MaterialParameters::GetParameterValue(const char* name) const
{
  auto it = double_params_.find(name);
  if (it == double_params_.end()) {
    std::string msg = "Error in GetParameterValue() for material " + material_name_ + "; parameter '" + name + "' not found";
    throw std::invalid_argument(msg);
  }
  return it->second;
}

const std::string&
MaterialParameters::GetStringParameterValue(const char* name) const
{
  auto it = string_params_.find(name);
  if (it == string_params_.end()) {
    std::string msg = "Error in GetStringParameterValue() for material " + material_name_ + "; parameter '" + name + "' not found";
    throw std::invalid_argument(msg);
  }
  return it->second;
}

Material::Material(const MaterialParameters& params, nsm_material_kind kind) : params_(params), kind_(kind)
{
  // src/nimble_material.cc:52-60, 210-218
  density_       = params_.GetParameterValue("density");
  bulk_modulus_  = params_.GetParameterValue("bulk_modulus");
  shear_modulus_ = params_.GetParameterValue("shear_modulus");
}

void
Material::GetStateVariableLabel(int index, char label[MaterialParameters::MAX_MAT_MODEL_STR_LEN]) const
{
  const char* l = nsm_b200_material_state_label(kind_, index);
  if (!l) throw std::invalid_argument("Material::GetStateVariableLabel: bad index " + std::to_string(index));
  std::snprintf(label, MaterialParameters::MAX_MAT_MODEL_STR_LEN, "%s", l);
}

void
Material::GetStress(int, int num_pts, double, double, const double* deformation_gradient_n, const double* deformation_gradient_np1,
                    const double* stress_n, double* stress_np1, const double* state_data_n, double* state_data_np1,
                    DeviceContext& device, bool) const
{
  const std::vector<double> params = DeviceParameters();
  device.check(nsm_b200_compute_stress_state(device.get(), kind_, (int)params.size(), params.data(), num_pts, deformation_gradient_n,
                                             deformation_gradient_np1, stress_n, state_data_n, stress_np1, state_data_np1),
               "Material::GetStress");
}

MaterialFactoryBase::MaterialFactoryBase()
{
  // NeohookeanMaterial / ElasticMaterial::register_supported_material_parameters (src/nimble_material.cc:128-134, 312-318)
  add_valid_double_parameter_name("bulk_modulus");
  add_valid_double_parameter_name("shear_modulus");
  add_valid_double_parameter_name("density");
  // the history-dependent model of the state-variable slot
  add_valid_double_parameter_name("yield_stress");
  add_valid_double_parameter_name("hardening_modulus");
}

std::shared_ptr<MaterialParameters>
MaterialFactoryBase::ParseMaterialParametersString(const std::string& material_parameters, int num_material_points) const
{
  std::istringstream       in(material_parameters);
  std::vector<std::string> tokens;
  for (std::string t; in >> t;) tokens.push_back(t);
  if (tokens.size() < 2) throw std::invalid_argument("material string needs a model name and parameters: '" + material_parameters + "'");
  // key-value pairs after the model name; a key is a double parameter, a string parameter, or an error
  // (src/nimble_material_factory_base.cc:62-98; the first occurrence of a key wins, as std::map::insert there)
  std::map<std::string, double>      doubles;
  std::map<std::string, std::string> strings;
  for (size_t i = 1; i < tokens.size(); i += 2) {
    const std::string& key       = tokens[i];
    const bool         is_double = std::find(valid_double_parameter_names_.begin(), valid_double_parameter_names_.end(), key) !=
                           valid_double_parameter_names_.end();
    const bool is_string = !is_double && std::find(valid_string_parameter_names_.begin(), valid_string_parameter_names_.end(), key) !=
                                             valid_string_parameter_names_.end();
    if (!is_double && !is_string) throw std::invalid_argument("Invalid material parameter encountered: '" + key + "'");
    if (i + 1 >= tokens.size()) throw std::invalid_argument("material parameter '" + key + "' has no value");
    if (is_double)
      doubles.insert(std::make_pair(key, std::stod(tokens[i + 1])));
    else
      strings.insert(std::make_pair(key, tokens[i + 1]));
  }
  return std::make_shared<MaterialParameters>(tokens.front(), strings, doubles, num_material_points);
}

void
MaterialFactory::create()
{
  const std::string name = material_params->GetMaterialName(false);
  if (name == "neohookean")
    material = std::make_shared<NeohookeanMaterial>(*material_params);
  else if (name == "elastic")
    material = std::make_shared<ElasticMaterial>(*material_params);
  else if (name == "j2_plasticity")
    material = std::make_shared<J2PlasticityMaterial>(*material_params);
  else
    throw std::invalid_argument("\nError in Block::InstantiateMaterialModel(), invalid material model name.\n");
}

}  // namespace nimble_b200
