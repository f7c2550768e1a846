// oracle/ref_state_material.cc — TEST INFRASTRUCTURE; see ref_state_material.h.
#include "ref_state_material.h"

#include <cmath>
#include <cstring>
#include <stdexcept>

#include "nimble_utils.h"

namespace nsm_oracle {

J2PlasticityMaterial::J2PlasticityMaterial(nimble::MaterialParameters const& p)
    : density_(p.GetParameterValue("density")),
      bulk_modulus_(p.GetParameterValue("bulk_modulus")),
      shear_modulus_(p.GetParameterValue("shear_modulus")),
      yield_stress_(p.GetParameterValue("yield_stress")),
      hardening_modulus_(p.GetParameterValue("hardening_modulus"))
{
}

void
J2PlasticityMaterial::GetStateVariableLabel(int index, char label[nimble::MaterialParameters::MAX_MAT_MODEL_STR_LEN]) const
{
  std::strcpy(label, index == 0 ? "equivalent_plastic_strain" : "von_mises_stress");
}

void
J2PlasticityMaterial::GetStress(int, int num_pts, double, double, const double* Fn, const double* Fnp1, const double* sn,
                                double* snp1, const double* state_n, double* state_np1, nimble::DataManager&, bool)
{
  using namespace nimble;
  const double two_mu = 2.0 * shear_modulus_;
  const double lambda = bulk_modulus_ - 2.0 * shear_modulus_ / 3.0;
  for (int pt = 0; pt < num_pts; ++pt, Fn += 9, Fnp1 += 9, sn += 6, snp1 += 6, state_n += 2, state_np1 += 2) {
    double de[6], t[6];
    de[K_S_XX] = Fnp1[K_F_XX] - Fn[K_F_XX];
    de[K_S_YY] = Fnp1[K_F_YY] - Fn[K_F_YY];
    de[K_S_ZZ] = Fnp1[K_F_ZZ] - Fn[K_F_ZZ];
    de[K_S_XY] = 0.5 * ((Fnp1[K_F_XY] + Fnp1[K_F_YX]) - (Fn[K_F_XY] + Fn[K_F_YX]));
    de[K_S_YZ] = 0.5 * ((Fnp1[K_F_YZ] + Fnp1[K_F_ZY]) - (Fn[K_F_YZ] + Fn[K_F_ZY]));
    de[K_S_ZX] = 0.5 * ((Fnp1[K_F_ZX] + Fnp1[K_F_XZ]) - (Fn[K_F_ZX] + Fn[K_F_XZ]));
    const double tr = de[K_S_XX] + de[K_S_YY] + de[K_S_ZZ];
    t[K_S_XX] = sn[K_S_XX] + (two_mu * de[K_S_XX] + lambda * tr);
    t[K_S_YY] = sn[K_S_YY] + (two_mu * de[K_S_YY] + lambda * tr);
    t[K_S_ZZ] = sn[K_S_ZZ] + (two_mu * de[K_S_ZZ] + lambda * tr);
    t[K_S_XY] = sn[K_S_XY] + two_mu * de[K_S_XY];
    t[K_S_YZ] = sn[K_S_YZ] + two_mu * de[K_S_YZ];
    t[K_S_ZX] = sn[K_S_ZX] + two_mu * de[K_S_ZX];
    const double p  = (t[K_S_XX] + t[K_S_YY] + t[K_S_ZZ]) / 3.0;
    const double s0 = t[K_S_XX] - p, s1 = t[K_S_YY] - p, s2 = t[K_S_ZZ] - p;
    const double s3 = t[K_S_XY], s4 = t[K_S_YZ], s5 = t[K_S_ZX];
    const double j2 = 0.5 * (s0 * s0 + s1 * s1 + s2 * s2) + (s3 * s3 + s4 * s4 + s5 * s5);
    const double q  = std::sqrt(3.0 * j2);
    const double eqps_n = state_n[0];
    const double f      = q - (yield_stress_ + hardening_modulus_ * eqps_n);
    if (f > 0.0) {  // radial return
      const double dgamma = f / (3.0 * shear_modulus_ + hardening_modulus_);
      const double scale  = 1.0 - (3.0 * shear_modulus_ * dgamma) / q;
      snp1[K_S_XX] = p + scale * s0;
      snp1[K_S_YY] = p + scale * s1;
      snp1[K_S_ZZ] = p + scale * s2;
      snp1[K_S_XY] = scale * s3;
      snp1[K_S_YZ] = scale * s4;
      snp1[K_S_ZX] = scale * s5;
      state_np1[0] = eqps_n + dgamma;
      state_np1[1] = scale * q;
    } else {
      for (int i = 0; i < 6; ++i) snp1[i] = t[i];
      state_np1[0] = eqps_n;
      state_np1[1] = q;
    }
  }
}

void
J2PlasticityMaterial::GetStress(double, double, nimble::Viewify<1, const double>&, nimble::Viewify<1, const double>&,
                                nimble::Viewify<1, const double>&, nimble::Viewify<1>) const
{
  throw std::logic_error("J2PlasticityMaterial: the stateless GetStress overload cannot carry state");
}

void
J2PlasticityMaterial::GetTangent(int, double*) const
{
  throw std::logic_error("J2PlasticityMaterial::GetTangent: explicit dynamics only");
}

StateMaterialFactory::StateMaterialFactory() : nimble::MaterialFactory()
{
  add_valid_double_parameter_name("yield_stress");
  add_valid_double_parameter_name("hardening_modulus");
}

void
StateMaterialFactory::create()
{
  if (material_params->GetMaterialName(false) == "j2_plasticity")
    material = std::make_shared<J2PlasticityMaterial>(*material_params);
  else
    nimble::MaterialFactory::create();
}

}  // namespace nsm_oracle
