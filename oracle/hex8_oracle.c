/* oracle/hex8_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see hex8_oracle.h).
 *
 * CPU restatement of the hex8 explicit-dynamics hot path of NimbleSM's serial build.  Operation order
 * follows the reference source so that results are bit-identical to it under IEEE fp64 without FMA
 * contraction (checked in tests/test_oracle.py against the reference's own compiled objects).
 * Tensor storage orders (src/nimble_utils.h:86-110): full = xx,yy,zz,xy,yz,zx,yx,zy,xz;
 * symmetric = xx,yy,zz,xy,yz,zx.
 */
#include "hex8_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { FXX = 0, FYY = 1, FZZ = 2, FXY = 3, FYZ = 4, FZX = 5, FYX = 6, FZY = 7, FXZ = 8 };
enum { SXX = 0, SYY = 1, SZZ = 2, SXY = 3, SYZ = 4, SZX = 5 };

/* Parent-domain signs of the 8 nodes; the 2x2x2 Gauss points use the same sign pattern
 * (src/nimble_element.cc:58-90 and :113-120). */
static const double SGN[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                 {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};

static double g_N[64];
static double g_dN[192];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

/* HexElement::HexElement / ShapeFunctionValues / ShapeFunctionDerivatives
 * (src/nimble_element.cc:55-171).  The Gauss abscissa is the 15-digit literal of :58, and each entry
 * is c*(1±r)*(1±s)*(1±t) evaluated left to right with c = 1/8. */
void
h8o_shape_tables(double shape_vals[64], double shape_derivs[192])
{
  const double g = 0.577350269189626;
  const double c = 1.0 / 8.0;
  for (int q = 0; q < 8; ++q) {
    const double r = SGN[q][0] * g, s = SGN[q][1] * g, t = SGN[q][2] * g;
    for (int j = 0; j < 8; ++j) {
      const double fr = 1.0 + SGN[j][0] * r, fs = 1.0 + SGN[j][1] * s, ft = 1.0 + SGN[j][2] * t;
      shape_vals[8 * q + j]            = c * fr * fs * ft;
      shape_derivs[24 * q + 3 * j + 0] = (SGN[j][0] * c) * fs * ft;
      shape_derivs[24 * q + 3 * j + 1] = (SGN[j][1] * c) * fr * ft;
      shape_derivs[24 * q + 3 * j + 2] = (SGN[j][2] * c) * fr * fs;
    }
  }
}

static void
init_tables(void)
{
  h8o_shape_tables(g_N, g_dN);
}

/* Invert3x3 (src/nimble_utils.h:1229-1268): cofactor inverse, nine true divisions by det. */
double
h8o_invert3x3(const double m[3][3], double inv[3][3])
{
  const double c0 = m[1][1] * m[2][2] - m[1][2] * m[2][1];
  const double c1 = m[1][0] * m[2][2] - m[1][2] * m[2][0];
  const double c2 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
  const double c3 = m[0][1] * m[2][2] - m[0][2] * m[2][1];
  const double c4 = m[0][0] * m[2][2] - m[2][0] * m[0][2];
  const double c5 = m[0][0] * m[2][1] - m[0][1] * m[2][0];
  const double c6 = m[0][1] * m[1][2] - m[0][2] * m[1][1];
  const double c7 = m[0][0] * m[1][2] - m[0][2] * m[1][0];
  const double c8 = m[0][0] * m[1][1] - m[0][1] * m[1][0];
  const double det = m[0][0] * c0 - m[0][1] * c1 + m[0][2] * c2;
  inv[0][0] = c0 / det;
  inv[0][1] = -1.0 * c3 / det;
  inv[0][2] = c6 / det;
  inv[1][0] = -1.0 * c1 / det;
  inv[1][1] = c4 / det;
  inv[1][2] = -1.0 * c7 / det;
  inv[2][0] = c2 / det;
  inv[2][1] = -1.0 * c5 / det;
  inv[2][2] = c8 / det;
  return det;
}

/* Jacobian-like sum  J[i][k] = sum_j x_j[i] * dN_j/dxi_k  at Gauss point q, nodes ascending
 * (src/nimble_element.h:463-471). */
static void
param_gradient(const double x[24], int q, double J[3][3])
{
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) J[i][k] = 0.0;
  for (int j = 0; j < 8; ++j) {
    const double* d = &g_dN[24 * q + 3 * j];
    for (int i = 0; i < 3; ++i) {
      J[i][0] += x[3 * j + i] * d[0];
      J[i][1] += x[3 * j + i] * d[1];
      J[i][2] += x[3 * j + i] * d[2];
    }
  }
}

/* HexElement::ComputeDeformationGradients serial wrapper + _impl
 * (src/nimble_element.cc:333-351, src/nimble_element.h:430-502): the wrapper forms disp = cur - ref and
 * the kernel re-adds it (cc = ref + disp), which is reproduced literally. */
void
h8o_def_grad(const double ref[24], const double cur[24], double F[72])
{
  pthread_once(&g_once, init_tables);
  double cc[24];
  for (int i = 0; i < 24; ++i) {
    const double d = cur[i] - ref[i];
    cc[i]          = ref[i] + d;
  }
  for (int q = 0; q < 8; ++q) {
    double a[3][3], b[3][3], binv[3][3], fg[3][3];
    param_gradient(cc, q, a);
    param_gradient(ref, q, b);
    h8o_invert3x3(b, binv);
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) fg[j][k] = a[j][0] * binv[0][k] + a[j][1] * binv[1][k] + a[j][2] * binv[2][k];
    double* Fq = &F[9 * q];
    Fq[FXX] = fg[0][0];
    Fq[FXY] = fg[0][1];
    Fq[FXZ] = fg[0][2];
    Fq[FYX] = fg[1][0];
    Fq[FYY] = fg[1][1];
    Fq[FYZ] = fg[1][2];
    Fq[FZX] = fg[2][0];
    Fq[FZY] = fg[2][1];
    Fq[FZZ] = fg[2][2];
  }
}

/* ElasticMaterial::GetStress (src/nimble_material.cc:95-126): small-strain Hooke law on sym(F) - I. */
void
h8o_stress_elastic(double bulk, double shear, const double F[9], double sig[6])
{
  const double two_mu = 2.0 * shear;
  const double lambda = bulk - 2.0 * shear / 3.0;
  double       e[6];
  e[SXX] = F[FXX] - 1.0;
  e[SYY] = F[FYY] - 1.0;
  e[SZZ] = F[FZZ] - 1.0;
  e[SXY] = 0.5 * (F[FXY] + F[FYX]);
  e[SYZ] = 0.5 * (F[FYZ] + F[FZY]);
  e[SZX] = 0.5 * (F[FZX] + F[FXZ]);
  const double tr = e[SXX] + e[SYY] + e[SZZ];
  sig[SXX] = two_mu * e[SXX] + lambda * tr;
  sig[SYY] = two_mu * e[SYY] + lambda * tr;
  sig[SZZ] = two_mu * e[SZZ] + lambda * tr;
  sig[SXY] = two_mu * e[SXY];
  sig[SYZ] = two_mu * e[SYZ];
  sig[SZX] = two_mu * e[SZX];
}

/* The history-dependent material of the state-variable slot ("j2_plasticity": small-strain J2, linear isotropic
 * hardening, incremental form).  The reference ships no material with state, so the operation sequence is DEFINED by
 * the test-only nimble::Material subclass that plugs into the reference's own block / element-data plumbing
 * (oracle/ref_state_material.cc, J2PlasticityMaterial::GetStress); this is its restatement, bit for bit
 * (tests/test_oracle.py).  params = {bulk, shear, yield_stress, hardening_modulus}; state = {equivalent plastic
 * strain, von Mises stress}. */
void
h8o_stress_j2(const double params[4], const double Fn[9], const double Fnp1[9], const double sn[6], const double state_n[2],
              double snp1[6], double state_np1[2])
{
  const double bulk = params[0], shear = params[1], yield = params[2], hard = params[3];
  const double two_mu = 2.0 * shear;
  const double lambda = bulk - 2.0 * shear / 3.0;
  double       de[6], t[6];
  de[SXX] = Fnp1[FXX] - Fn[FXX];
  de[SYY] = Fnp1[FYY] - Fn[FYY];
  de[SZZ] = Fnp1[FZZ] - Fn[FZZ];
  de[SXY] = 0.5 * ((Fnp1[FXY] + Fnp1[FYX]) - (Fn[FXY] + Fn[FYX]));
  de[SYZ] = 0.5 * ((Fnp1[FYZ] + Fnp1[FZY]) - (Fn[FYZ] + Fn[FZY]));
  de[SZX] = 0.5 * ((Fnp1[FZX] + Fnp1[FXZ]) - (Fn[FZX] + Fn[FXZ]));
  const double tr = de[SXX] + de[SYY] + de[SZZ];
  t[SXX] = sn[SXX] + (two_mu * de[SXX] + lambda * tr);
  t[SYY] = sn[SYY] + (two_mu * de[SYY] + lambda * tr);
  t[SZZ] = sn[SZZ] + (two_mu * de[SZZ] + lambda * tr);
  t[SXY] = sn[SXY] + two_mu * de[SXY];
  t[SYZ] = sn[SYZ] + two_mu * de[SYZ];
  t[SZX] = sn[SZX] + two_mu * de[SZX];
  const double p  = (t[SXX] + t[SYY] + t[SZZ]) / 3.0;
  const double s0 = t[SXX] - p, s1 = t[SYY] - p, s2 = t[SZZ] - p;
  const double s3 = t[SXY], s4 = t[SYZ], s5 = t[SZX];
  const double j2 = 0.5 * (s0 * s0 + s1 * s1 + s2 * s2) + (s3 * s3 + s4 * s4 + s5 * s5);
  const double q  = sqrt(3.0 * j2);
  const double eqps_n = state_n[0];
  const double f      = q - (yield + hard * eqps_n);
  if (f > 0.0) { /* radial return */
    const double dgamma = f / (3.0 * shear + hard);
    const double scale  = 1.0 - (3.0 * shear * dgamma) / q;
    snp1[SXX]    = p + scale * s0;
    snp1[SYY]    = p + scale * s1;
    snp1[SZZ]    = p + scale * s2;
    snp1[SXY]    = scale * s3;
    snp1[SYZ]    = scale * s4;
    snp1[SZX]    = scale * s5;
    state_np1[0] = eqps_n + dgamma;
    state_np1[1] = scale * q;
  } else {
    for (int i = 0; i < 6; ++i) snp1[i] = t[i];
    state_np1[0] = eqps_n;
    state_np1[1] = q;
  }
}

int
h8o_num_state(int material)
{
  return material == H8O_J2_PLASTICITY ? 2 : 0;
}

/* Cos_Of_Acos_Divided_By_3 (src/nimble_utils.h:650-665): rational (6,5) fit on [0,1]. */
static double
cos_third_acos(double x)
{
  const double x2 = x * x;
  const double x4 = x2 * x2;
  return (0.866025403784438713 + 2.12714890259493060 * x +
          ((1.89202064815951569 + 0.739603278343401613 * x) * x2 +
           (0.121973926953064794 + x * (0.00655637626263929360 + 0.0000390884982780803443 * x)) * x4)) /
         (1.0 + 2.26376989330935617 * x +
          ((1.80461009751278976 + 0.603976798217196003 * x) * x2 +
           (0.0783255761115461708 + 0.00268525944538021629 * x) * x4));
}

static double
times_sign_of(double x, double y) /* MultiplySign, src/nimble_utils.h:131-137 */
{
  return x * (1 - 2 * (y < 0));
}

static double
sel0(int cond, double v) /* if_then_else_zero, src/nimble_utils.h:159-165 */
{
  return cond ? v : 0.0;
}

/* Eigen_Sym33_NonUnit (src/nimble_utils.h:667-857): closed-form symmetric 3x3 eigen-solver;
 * eigenvectors are NOT normalised; degenerate input (c2 >= -1e-30 c1^2) returns (c1, identity). */
void
h8o_eigen_sym33(const double A[6], double eval[3], double v0[3], double v1[3], double v2[3])
{
  double       cxx = A[SXX], cyy = A[SYY], czz = A[SZZ];
  const double cxy = A[SXY], cyz = A[SYZ], czx = A[SZX];

  const double c1 = (cxx + cyy + czz) / 3.0;
  cxx -= c1;
  cyy -= c1;
  czz -= c1;

  const double cxy2 = cxy * cxy, cyz2 = cyz * cyz, czx2 = czx * czx, cxxcyy = cxx * cyy;
  const double c2   = cxxcyy + cyy * czz + czz * cxx - cxy2 - cyz2 - czx2;

  const double three_over_a = -3.0 / c2;
  const double root_toa     = sqrt(three_over_a);
  const double c3           = cxx * cyz2 + cyy * czx2 - 2.0 * cxy * cyz * czx + czz * (cxy2 - cxxcyy);
  const double rr           = -0.5 * c3 * three_over_a * root_toa;
  const double absrr        = fabs(rr);
  const double arg          = absrr < 1.0 ? absrr : 1.0;
  const double two_cos      = 2.0 * times_sign_of(cos_third_acos(arg), rr);
  double       e2           = two_cos / root_toa;

  const double r0[3] = {cxx - e2, cxy, czx};
  const double r1[3] = {cxy, cyy - e2, cyz};
  const double r2[3] = {czx, cyz, czz - e2};

  /* QR with column pivoting through branch-free selects (:716-771) */
  const double k0 = r0[0] * r0[0] + cxy2 + czx2;
  const double k1 = cxy2 + r1[1] * r1[1] + cyz2;
  const double k2 = czx2 + cyz2 + r2[2] * r2[2];
  const int    k0gk1 = k1 <= k0, k0gk2 = k2 <= k0, k1gk2 = k2 <= k1;
  const int    big0 = k0gk1 && k0gk2;
  const int    big1 = k1gk2 && !k0gk1;
  const int    big2 = !(big0 || big1);

  double p[3], s[3], t[3];
  for (int i = 0; i < 3; ++i) {
    p[i] = sel0(big0, r0[i]) + sel0(big1, r1[i]) + sel0(big2, r2[i]);
    s[i] = big0 ? r1[i] : r0[i];
    t[i] = big2 ? r1[i] : r2[i];
  }
  const double ipp = 1.0 / (sel0(big0, k0) + sel0(big1, k1) + sel0(big2, k2));
  const double ps  = ipp * (p[0] * s[0] + p[1] * s[1] + p[2] * s[2]);
  const double pt  = ipp * (p[0] * t[0] + p[1] * t[1] + p[2] * t[2]);
  for (int i = 0; i < 3; ++i) s[i] -= ps * p[i];
  for (int i = 0; i < 3; ++i) t[i] -= pt * p[i];

  const double a0     = s[0] * s[0] + s[1] * s[1] + s[2] * s[2];
  const double a1     = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
  const int    a0lea1 = a0 <= a1;
  double       w[3];
  for (int i = 0; i < 3; ++i) w[i] = a0lea1 ? t[i] : s[i];
  const double iww = 1.0 / (a0lea1 ? a1 : a0);

  v2[0] = p[1] * w[2] - p[2] * w[1];
  v2[1] = p[2] * w[0] - p[0] * w[2];
  v2[2] = p[0] * w[1] - p[1] * w[0];

  const double Ap0 = cxx * p[0] + cxy * p[1] + czx * p[2];
  const double Ap1 = cxy * p[0] + cyy * p[1] + cyz * p[2];
  const double Ap2 = czx * p[0] + cyz * p[1] + czz * p[2];
  const double Aw0 = cxx * w[0] + cxy * w[1] + czx * w[2];
  const double Aw1 = cxy * w[0] + cyy * w[1] + cyz * w[2];
  const double Aw2 = czx * w[0] + cyz * w[1] + czz * w[2];

  double       mxx   = (p[0] * Ap0 + p[1] * Ap1 + p[2] * Ap2) * ipp;
  const double pAw   = (p[0] * Aw0 + p[1] * Aw1 + p[2] * Aw2);
  double       myy   = (w[0] * Aw0 + w[1] * Aw1 + w[2] * Aw2) * iww;
  const double mxy2  = pAw * pAw * iww * ipp;

  /* Wilkinson shift on the 2x2 remainder (:796-803) */
  const double hb  = 0.5 * (mxx - myy);
  const double sq  = times_sign_of(sqrt(hb * hb + mxy2), hb);
  double       e0  = myy + hb - sq;
  double       e1  = mxx + myy - e0;
  mxx -= e0;
  myy -= e0;
  const double mxx2 = mxx * mxx, myy2 = myy * myy;
  const double f1   = (mxx2 < myy2) ? pAw * iww : mxx;
  const double f2   = (mxx2 < myy2) ? myy : ipp * pAw;
  for (int i = 0; i < 3; ++i) v0[i] = f1 * w[i] - f2 * p[i];
  const int both_zero = (mxx2 == 0.0) && (mxy2 == 0.0);
  for (int i = 0; i < 3; ++i) v0[i] = both_zero ? w[i] : v0[i];

  v1[0] = v2[1] * v0[2] - v2[2] * v0[1];
  v1[1] = v2[2] * v0[0] - v2[0] * v0[2];
  v1[2] = v2[0] * v0[1] - v2[1] * v0[0];

  e0 += c1;
  e1 += c1;
  e2 += c1;

  const double tol = (c1 * c1) * (-1.0e-30);
  const int    ok  = c2 < tol;
  eval[0] = ok ? e0 : c1;
  eval[1] = ok ? e1 : c1;
  eval[2] = ok ? e2 : c1;
  static const double I3[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; ++i) {
    v0[i] = ok ? v0[i] : I3[0][i];
    v1[i] = ok ? v1[i] : I3[1][i];
    v2[i] = ok ? v2[i] : I3[2][i];
  }
}

/* Invert_Full33 (src/nimble_utils.h:523-551): cofactors over the determinant, the signs applied as "-1.0 * minor / det";
 * returns the determinant.  Known answers: unit_tests/test_nimble_utils.cc:47-93 (tests/test_oracle.py). */
double
h8o_invert_full33(const double F[9], double G[9])
{
  const double m0  = F[FYY] * F[FZZ] - F[FYZ] * F[FZY];
  const double m1  = F[FYX] * F[FZZ] - F[FYZ] * F[FZX];
  const double m2  = F[FYX] * F[FZY] - F[FYY] * F[FZX];
  const double m3  = F[FXY] * F[FZZ] - F[FXZ] * F[FZY];
  const double m4  = F[FXX] * F[FZZ] - F[FZX] * F[FXZ];
  const double m5  = F[FXX] * F[FZY] - F[FXY] * F[FZX];
  const double m6  = F[FXY] * F[FYZ] - F[FXZ] * F[FYY];
  const double m7  = F[FXX] * F[FYZ] - F[FXZ] * F[FYX];
  const double m8  = F[FXX] * F[FYY] - F[FXY] * F[FYX];
  const double det = F[FXX] * m0 - F[FXY] * m1 + F[FXZ] * m2;
  G[FXX] = m0 / det;
  G[FXY] = -1.0 * m3 / det;
  G[FXZ] = m6 / det;
  G[FYX] = -1.0 * m1 / det;
  G[FYY] = m4 / det;
  G[FYZ] = -1.0 * m7 / det;
  G[FZX] = m2 / det;
  G[FZY] = -1.0 * m5 / det;
  G[FZZ] = m8 / det;
  return det;
}

/* Left stretch V of F = V R as computed by Polar_Decomp (src/nimble_utils.h:859-908) through
 * Invert_Full33 (:523-551) and Square_Full33T_Full33 (:194-205): eigen-decompose (F^-1)^T (F^-1) = V^-2
 * and rebuild V = sum_i v_i v_i^T / (sqrt(lambda_i) |v_i|^2).  The rotation product of :907 only feeds a
 * debug check in the caller and is not part of any result. */
void
h8o_polar_left_stretch(const double F[9], double V[6])
{
  double G[9];
  h8o_invert_full33(F, G);

  double C[6]; /* G^T G */
  C[SXX] = G[FXX] * G[FXX] + G[FYX] * G[FYX] + G[FZX] * G[FZX];
  C[SYY] = G[FXY] * G[FXY] + G[FYY] * G[FYY] + G[FZY] * G[FZY];
  C[SZZ] = G[FXZ] * G[FXZ] + G[FYZ] * G[FYZ] + G[FZZ] * G[FZZ];
  C[SXY] = G[FXX] * G[FXY] + G[FYX] * G[FYY] + G[FZX] * G[FZY];
  C[SYZ] = G[FXY] * G[FXZ] + G[FYY] * G[FYZ] + G[FZY] * G[FZZ];
  C[SZX] = G[FXX] * G[FXZ] + G[FYX] * G[FYZ] + G[FZX] * G[FZZ];

  double lam[3], a[3], b[3], c[3];
  h8o_eigen_sym33(C, lam, a, b, c);
  for (int i = 0; i < 3; ++i) lam[i] = lam[i] < 0.0 ? 0.0 : lam[i];

  const double la = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
  const double lb = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
  const double lc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
  const double wa = 1.0 / (sqrt(lam[0]) * la);
  const double wb = 1.0 / (sqrt(lam[1]) * lb);
  const double wc = 1.0 / (sqrt(lam[2]) * lc);

  V[SXX] = wa * a[0] * a[0] + wb * b[0] * b[0] + wc * c[0] * c[0];
  V[SYY] = wa * a[1] * a[1] + wb * b[1] * b[1] + wc * c[1] * c[1];
  V[SZZ] = wa * a[2] * a[2] + wb * b[2] * b[2] + wc * c[2] * c[2];
  V[SXY] = wa * a[0] * a[1] + wb * b[0] * b[1] + wc * c[0] * c[1];
  V[SYZ] = wa * a[1] * a[2] + wb * b[1] * b[2] + wc * c[1] * c[2];
  V[SZX] = wa * a[2] * a[0] + wb * b[2] * b[0] + wc * c[2] * c[0];
}

/* NeohookeanMaterial::GetStress (src/nimble_material.cc:252-310). */
void
h8o_stress_neohookean(double bulk, double shear, const double F[9], double sig[6])
{
  double v[6];
  h8o_polar_left_stretch(F, v);
  const double J = v[SXX] * v[SYY] * v[SZZ] + 2.0 * v[SXY] * v[SYZ] * v[SZX] - v[SXX] * v[SYZ] * v[SYZ] -
                   v[SYY] * v[SZX] * v[SZX] - v[SZZ] * v[SXY] * v[SXY];
  const double cj  = cbrt(J);
  const double fac = 1.0 / (cj * cj);
  const double p   = 0.5 * bulk * (J - 1.0 / J);

  double bxx = v[SXX] * v[SXX] + v[SXY] * v[SXY] + v[SZX] * v[SZX];
  double byy = v[SXY] * v[SXY] + v[SYY] * v[SYY] + v[SYZ] * v[SYZ];
  double bzz = v[SZX] * v[SZX] + v[SYZ] * v[SYZ] + v[SZZ] * v[SZZ];
  double bxy = v[SXX] * v[SXY] + v[SXY] * v[SYY] + v[SZX] * v[SYZ];
  double byz = v[SXY] * v[SZX] + v[SYY] * v[SYZ] + v[SYZ] * v[SZZ];
  double bzx = v[SZX] * v[SXX] + v[SYZ] * v[SXY] + v[SZZ] * v[SZX];
  bxx = fac * bxx;
  byy = fac * byy;
  bzz = fac * bzz;
  bxy = fac * bxy;
  byz = fac * byz;
  bzx = fac * bzx;
  const double tr = bxx + byy + bzz;
  bxx = bxx - tr / 3.0;
  byy = byy - tr / 3.0;
  bzz = bzz - tr / 3.0;
  sig[SXX] = p + shear * bxx / J;
  sig[SYY] = p + shear * byy / J;
  sig[SZZ] = p + shear * bzz / J;
  sig[SXY] = shear * bxy / J;
  sig[SYZ] = shear * byz / J;
  sig[SZX] = shear * bzx / J;
}

/* HexElement::ComputeNodalForces + _impl (src/nimble_element.cc:457-472, src/nimble_element.h:540-625):
 * f_node -= (dN/dx . sigma) * (detJ * w), Gauss points outermost, w = 1. */
void
h8o_nodal_forces(const double cur[24], const double sig[48], double f[24])
{
  pthread_once(&g_once, init_tables);
  double acc[24];
  for (int i = 0; i < 24; ++i) acc[i] = 0.0;
  for (int q = 0; q < 8; ++q) {
    double a[3][3], ai[3][3];
    double cc[24];
    for (int i = 0; i < 24; ++i) cc[i] = cur[i] + 0.0; /* wrapper passes a zero displacement (:462-463) */
    param_gradient(cc, q, a);
    const double  det = h8o_invert3x3(a, ai);
    const double* s   = &sig[6 * q];
    for (int n = 0; n < 8; ++n) {
      const double* d  = &g_dN[24 * q + 3 * n];
      const double  g1 = d[0] * ai[0][0] + d[1] * ai[1][0] + d[2] * ai[2][0];
      const double  g2 = d[0] * ai[0][1] + d[1] * ai[1][1] + d[2] * ai[2][1];
      const double  g3 = d[0] * ai[0][2] + d[1] * ai[1][2] + d[2] * ai[2][2];
      double        f1 = g1 * s[SXX] + g2 * s[SXY] + g3 * s[SZX];
      double        f2 = g1 * s[SXY] + g2 * s[SYY] + g3 * s[SYZ];
      double        f3 = g1 * s[SZX] + g2 * s[SYZ] + g3 * s[SZZ];
      f1 *= det * 1.0;
      f2 *= det * 1.0;
      f3 *= det * 1.0;
      acc[3 * n + 0] -= f1;
      acc[3 * n + 1] -= f2;
      acc[3 * n + 2] -= f3;
    }
  }
  for (int i = 0; i < 24; ++i) f[i] = acc[i];
}

/* HexElement::ComputeLumpedMass / ComputeConsistentMass_impl (src/nimble_element.cc:173-193,
 * src/nimble_element.h:266-309): row sums of the consistent mass, loops nested (i, j, gauss point). */
void
h8o_lumped_mass(double density, const double ref[24], double m[8])
{
  pthread_once(&g_once, init_tables);
  double det[8];
  for (int q = 0; q < 8; ++q) {
    double a[3][3], ai[3][3];
    param_gradient(ref, q, a);
    det[q] = h8o_invert3x3(a, ai);
  }
  for (int i = 0; i < 8; ++i) {
    m[i] = 0.0;
    for (int j = 0; j < 8; ++j) {
      double mij = 0.0;
      for (int q = 0; q < 8; ++q) mij += 1.0 * density * g_N[8 * q + i] * g_N[8 * q + j] * det[q];
      m[i] += mij;
    }
  }
}

/* HexElement::ComputeCharacteristicLength (src/nimble_element.cc:221-260), including the quirk that
 * the box maxima start at 0.0 (:230). */
double
h8o_char_length(const double x[24])
{
  double xmax = 0.0, ymax = 0.0, zmax = 0.0;
  double xmin = DBL_MAX, ymin = DBL_MAX, zmin = DBL_MAX, dmin2 = DBL_MAX;
  for (int n = 0; n < 8; ++n) {
    const double nx = x[3 * n], ny = x[3 * n + 1], nz = x[3 * n + 2];
    if (nx < xmin) xmin = nx;
    if (nx > xmax) xmax = nx;
    if (ny < ymin) ymin = ny;
    if (ny > ymax) ymax = ny;
    if (nz < zmin) zmin = nz;
    if (nz > zmax) zmax = nz;
    for (int m = n + 1; m < 8; ++m) {
      const double mx = x[3 * m], my = x[3 * m + 1], mz = x[3 * m + 2];
      const double d2 = (nx - mx) * (nx - mx) + (ny - my) * (ny - my) + (nz - mz) * (nz - mz);
      if (d2 < dmin2) dmin2 = d2;
    }
  }
  double len = sqrt(dmin2);
  double box = xmax - xmin;
  if (ymax - ymin < box) box = ymax - ymin;
  if (zmax - zmin < box) box = zmax - zmin;
  if (box < len) len = box;
  return len;
}

/* HexElement::ComputeVolumeAverage + ComputeVolumeAverageQuantities_impl (src/nimble_element.cc:262-282,
 * src/nimble_element.h:343-392): volume = sum detJ (unit weights), averages over the current
 * configuration.  q is [8][nq]. */
void
h8o_volume_average(const double cur[24], int nq, const double* q, double* volume, double* avg)
{
  pthread_once(&g_once, init_tables);
  double vol = 0.0;
  for (int i = 0; i < nq; ++i) avg[i] = 0.0;
  double cc[24];
  for (int i = 0; i < 24; ++i) cc[i] = cur[i] + 0.0;
  for (int g = 0; g < 8; ++g) {
    double a[3][3], ai[3][3];
    param_gradient(cc, g, a);
    const double det = h8o_invert3x3(a, ai);
    vol += det;
    for (int i = 0; i < nq; ++i) avg[i] += q[g * nq + i] * 1.0 * det;
  }
  for (int i = 0; i < nq; ++i) avg[i] /= vol;
  *volume = vol;
}

/* ---------------------------------------------------------------------------------------------
 * block level
 * ------------------------------------------------------------------------------------------- */
static void
gather(const double* ref, const double* disp, const int* en, double X[24], double x[24])
{
  /* ComputeInternalForceFunctor gather (src/nimble_block.cc:309-316): cur = ref + disp */
  for (int j = 0; j < 8; ++j) {
    const long n = en[j];
    for (int i = 0; i < 3; ++i) {
      X[3 * j + i] = ref[3 * n + i];
      x[3 * j + i] = ref[3 * n + i] + disp[3 * n + i];
    }
  }
}

/* Block::ComputeInternalForce / functor (src/nimble_block.cc:278-436): elements in ascending order,
 * gather -> F -> stress -> store -> nodal force -> scatter-add. */
void
h8o_block_internal_force(int material, double bulk, double shear, const double* ref, const double* disp, long n_elem,
                         const int* conn, double* f, double* elem_data)
{
  for (long e = 0; e < n_elem; ++e) {
    const int* en = &conn[8 * e];
    double     X[24], x[24], F[72], sig[48], fe[24];
    gather(ref, disp, en, X, x);
    h8o_def_grad(X, x, F);
    for (int q = 0; q < 8; ++q) {
      if (material == H8O_ELASTIC)
        h8o_stress_elastic(bulk, shear, &F[9 * q], &sig[6 * q]);
      else
        h8o_stress_neohookean(bulk, shear, &F[9 * q], &sig[6 * q]);
    }
    if (elem_data) {
      double* d = &elem_data[120 * e];
      for (int q = 0; q < 8; ++q) {
        memcpy(&d[15 * q], &F[9 * q], 9 * sizeof(double));
        memcpy(&d[15 * q + 9], &sig[6 * q], 6 * sizeof(double));
      }
    }
    h8o_nodal_forces(x, sig, fe);
    for (int j = 0; j < 8; ++j)
      for (int i = 0; i < 3; ++i) f[3L * en[j] + i] += fe[3 * j + i];
  }
}

/* The same functor for any material, with the state plumbing of src/nimble_block.cc:297-307, 324-337, 355-368:
 * records are [8][15 + n_state] per element (F 9, sigma 6, state scalars; src/nimble_block.cc:84-108); F_n, sigma_n
 * and state_n are read from elem_data_n, the new record goes to elem_data_np1 (both required when n_state > 0).
 * params = {bulk, shear, material-specific...}. */
void
h8o_block_internal_force_state(int material, const double* params, const double* ref, const double* disp, long n_elem,
                               const int* conn, double* f, const double* elem_data_n, double* elem_data_np1)
{
  const int ns = h8o_num_state(material), stride = 15 + ns;
  for (long e = 0; e < n_elem; ++e) {
    const int* en = &conn[8 * e];
    double     X[24], x[24], F[72], sig[48], fe[24];
    gather(ref, disp, en, X, x);
    h8o_def_grad(X, x, F);
    for (int q = 0; q < 8; ++q) {
      if (material == H8O_ELASTIC)
        h8o_stress_elastic(params[0], params[1], &F[9 * q], &sig[6 * q]);
      else if (material == H8O_NEOHOOKEAN)
        h8o_stress_neohookean(params[0], params[1], &F[9 * q], &sig[6 * q]);
      else {
        const double* rn = &elem_data_n[(8 * e + q) * stride];
        double*       rp = &elem_data_np1[(8 * e + q) * stride];
        h8o_stress_j2(params, &rn[0], &F[9 * q], &rn[9], &rn[15], &sig[6 * q], &rp[15]);
      }
    }
    if (elem_data_np1) {
      for (int q = 0; q < 8; ++q) {
        double* d = &elem_data_np1[(8 * e + q) * stride];
        memcpy(&d[0], &F[9 * q], 9 * sizeof(double));
        memcpy(&d[9], &sig[6 * q], 6 * sizeof(double));
      }
    }
    h8o_nodal_forces(x, sig, fe);
    for (int j = 0; j < 8; ++j)
      for (int i = 0; i < 3; ++i) f[3L * en[j] + i] += fe[3 * j + i];
  }
}

/* Block::ComputeDerivedElementData for records of `stride` doubles per point: out [(1 + stride)][n_elem]. */
void
h8o_block_derived_stride(const double* ref, const double* disp, long n_elem, const int* conn, const double* elem_data,
                         int stride, double* out)
{
  for (long e = 0; e < n_elem; ++e) {
    double X[24], x[24], vol, avg[64];
    gather(ref, disp, &conn[8 * e], X, x);
    h8o_volume_average(x, stride, &elem_data[8L * stride * e], &vol, avg);
    out[e] = vol;
    for (int k = 0; k < stride; ++k) out[(long)(k + 1) * n_elem + e] = avg[k];
  }
}

/* Block::ComputeLumpedMassMatrix (src/nimble_block.cc:110-146). */
void
h8o_block_lumped_mass(double density, const double* ref, long n_elem, const int* conn, double* mass)
{
  for (long e = 0; e < n_elem; ++e) {
    const int* en = &conn[8 * e];
    double     X[24], m[8];
    for (int j = 0; j < 8; ++j)
      for (int i = 0; i < 3; ++i) X[3 * j + i] = ref[3L * en[j] + i];
    h8o_lumped_mass(density, X, m);
    for (int j = 0; j < 8; ++j) mass[en[j]] += m[j];
  }
}

/* BlockBase::ComputeCriticalTimeStep (src/nimble_block_base.cc:51-84). */
double
h8o_block_critical_dt(double bulk, double density, const double* ref, const double* disp, long n_elem, const int* conn)
{
  const double c  = sqrt(bulk / density);
  double       dt = DBL_MAX;
  for (long e = 0; e < n_elem; ++e) {
    double X[24], x[24];
    gather(ref, disp, &conn[8 * e], X, x);
    const double t = h8o_char_length(x) / c;
    if (t < dt) dt = t;
  }
  return dt;
}

/* Block::ComputeDerivedElementData (src/nimble_block.cc:438-497) for the full set of 15 ipt fields. */
void
h8o_block_derived(const double* ref, const double* disp, long n_elem, const int* conn, const double* elem_data,
                  double* out)
{
  for (long e = 0; e < n_elem; ++e) {
    double X[24], x[24], vol, avg[15];
    gather(ref, disp, &conn[8 * e], X, x);
    h8o_volume_average(x, 15, &elem_data[120 * e], &vol, avg);
    out[e] = vol;
    for (int k = 0; k < 15; ++k) out[(long)(k + 1) * n_elem + e] = avg[k];
  }
}

/* ---------------------------------------------------------------------------------------------
 * node level
 * ------------------------------------------------------------------------------------------- */
/* Viewify::operator+= of an AXPYResult (src/nimble_view.h:181-214): prod = alpha*1.0; y += prod*x. */
void
h8o_axpy(long n, double alpha, const double* x, double* y)
{
  const double prod = alpha * 1.0;
  for (long i = 0; i < n; ++i) y[i] += prod * x[i];
}

/* acceleration loop (src/integrators/explicit_time_integrator.cc:250-257). */
void
h8o_accel(long n_nodes, const double* mass, const double* f_int, const double* f_ext, double* a)
{
  for (long k = 0; k < n_nodes; ++k) {
    const double rm = 1.0 / mass[k];
    for (int i = 0; i < 3; ++i) a[3 * k + i] = rm * (f_int[3 * k + i] + (f_ext ? f_ext[3 * k + i] : 0.0));
  }
}

/* ---------------------------------------------------------------------------------------------
 * whole steps, threaded over contiguous element chunks with private force buffers
 * ------------------------------------------------------------------------------------------- */
typedef struct
{
  int           material;
  double        bulk, shear;
  const double *ref, *disp;
  const int*    conn;
  long          e0, e1, n0, n1;
  double*       fpriv; /* covers nodes [n0, n1) */
} chunk_t;

static void*
chunk_force(void* arg)
{
  chunk_t* c = (chunk_t*)arg;
  memset(c->fpriv, 0, sizeof(double) * 3 * (size_t)(c->n1 - c->n0));
  h8o_block_internal_force(c->material, c->bulk, c->shear, c->ref, c->disp, c->e1 - c->e0, c->conn + 8 * c->e0,
                           c->fpriv - 3 * c->n0, NULL);
  return NULL;
}

double
h8o_bench_steps(int material, double bulk, double shear, long n_nodes, const double* ref, long n_elem, const int* conn,
                const double* mass, double* u, double* v, double* a, double* f, double dt, int steps, int threads)
{
  if (threads < 1) threads = 1;
  chunk_t*   ch  = (chunk_t*)calloc((size_t)threads, sizeof(chunk_t));
  pthread_t* tid = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
  for (int t = 0; t < threads; ++t) {
    chunk_t* c = &ch[t];
    c->material = material, c->bulk = bulk, c->shear = shear, c->ref = ref, c->disp = u, c->conn = conn;
    c->e0 = n_elem * t / threads, c->e1 = n_elem * (t + 1) / threads;
    c->n0 = n_nodes, c->n1 = 0;
    for (long i = 8 * c->e0; i < 8 * c->e1; ++i) {
      if (conn[i] < c->n0) c->n0 = conn[i];
      if (conn[i] + 1 > c->n1) c->n1 = conn[i] + 1;
    }
    if (c->e1 == c->e0) c->n0 = c->n1 = 0;
    c->fpriv = (double*)malloc(sizeof(double) * 3 * (size_t)(c->n1 - c->n0) + 8);
  }
  const long   ndof = 3 * n_nodes;
  const double hdt  = 0.5 * dt;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int s = 0; s < steps; ++s) {
    h8o_axpy(ndof, hdt, a, v);
    h8o_axpy(ndof, dt, v, u);
    for (int t = 1; t < threads; ++t) pthread_create(&tid[t], NULL, chunk_force, &ch[t]);
    chunk_force(&ch[0]);
    for (int t = 1; t < threads; ++t) pthread_join(tid[t], NULL);
    memset(f, 0, sizeof(double) * (size_t)ndof);
    for (int t = 0; t < threads; ++t) {
      const double* src = ch[t].fpriv - 3 * ch[t].n0;
      for (long i = 3 * ch[t].n0; i < 3 * ch[t].n1; ++i) f[i] += src[i];
    }
    h8o_accel(n_nodes, mass, f, NULL, a);
    h8o_axpy(ndof, hdt, a, v);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  for (int t = 0; t < threads; ++t) free(ch[t].fpriv);
  free(ch);
  free(tid);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
