/* oracle/contact_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of NimbleSM's penalty contact (node-to-triangle, the `contact_force` term of the explicit
 * loop, src/integrators/explicit_time_integrator.cc:232-249).  Only tests/ may load it.  Each function cites the
 * reference file:line it follows (paths relative to /root/reference).  IEEE fp64 in the reference's source order;
 * build with -ffp-contract=off (oracle/Makefile).
 *
 * What the reference does per step (ArborXSerialContactManager::ComputeSerialContactForce,
 * src/contact/serial/arborx_serial_contact_manager.cc:147-196): coordinates of the contact sub-model = model
 * coordinates + displacement; every contact entity refreshes its vertices and its inflated bounding box; a box-box
 * search (ArborX BVH, boxes held in FLOAT by ArborX::Point/Box) pairs secondary NODES with primary TRIANGLES; each pair
 * is projected (ContactManager::Projection) and, when the node lies inside the facet and has penetrated by less than
 * the facet's characteristic length, the penalty force is spread over the facet's nodes and the node
 * (PenaltyContactEnforcement::EnforceContact).  The result is a SUM over all such pairs, so the search algorithm does
 * not matter, only which pairs pass the box test and the projection: this file walks all pairs.
 *
 * Parity status: the reference's contact manager itself cannot be compiled here (Kokkos + ArborX are absent and
 * src/nimble_contact_manager.h declares Kokkos views unconditionally).  PINNED in two ways (tests/test_oracle.py):
 * (1) bit for bit against oracle/ref_contact.cc, which drives the reference's own unmodified nimble::ContactEntity
 * objects (src/nimble_contact_entity.{h,cc} do compile: vertex refresh, bounding box, force spreading and scatter are
 * the reference's code) around the restated Projection; (2) against the reference's gold files
 * test/contact/cubes_contact and sphere_plate_contact (displacement and contact_force at the exodiff tolerances).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "hex8_oracle.h"

/* CrossProduct (src/nimble_utils.h:393-400) */
static void
cross3(const double* u, const double* v, double* r)
{
  r[0] = u[1] * v[2] - u[2] * v[1];
  r[1] = u[2] * v[0] - u[0] * v[2];
  r[2] = u[0] * v[1] - u[1] * v[0];
}

/* Characteristic lengths (ContactManager::CreateContactEntities, src/nimble_contact_manager.cc:288-330 for the
 * secondary nodes: sqrt of the largest squared edge of any secondary face holding the node; CreateContactNodesAndFaces,
 * :1062-1077 for the primary faces: largest edge length).  quads hold node ids of `coord` ([n][3], MODEL coordinates). */
void
h8o_contact_char_lengths(const double* coord, long n_prim, const int* prim_quads, double* prim_len, long n_sec_faces,
                         const int* sec_quads, long n_nodes, double* node_len /* [n_nodes], 0 where no secondary face */)
{
  for (long f = 0; f < n_prim; ++f) {
    const int* face = prim_quads + 4 * f;
    double     mx   = -DBL_MAX;
    for (int i = 0; i < 4; ++i) {
      const int    a = face[i], b = (i + 1 < 4) ? face[i + 1] : face[0];
      const double e = sqrt((coord[3 * b] - coord[3 * a]) * (coord[3 * b] - coord[3 * a]) +
                            (coord[3 * b + 1] - coord[3 * a + 1]) * (coord[3 * b + 1] - coord[3 * a + 1]) +
                            (coord[3 * b + 2] - coord[3 * a + 2]) * (coord[3 * b + 2] - coord[3 * a + 2]));
      if (e > mx) mx = e;
    }
    prim_len[f] = mx;
  }
  for (long i = 0; i < n_nodes; ++i) node_len[i] = 0.0;
  for (long f = 0; f < n_sec_faces; ++f) {
    const int* face = sec_quads + 4 * f;
    double     mx2  = -DBL_MAX;
    for (int i = 0; i < 4; ++i) {
      const int    a = face[i], b = (i + 1 < 4) ? face[i + 1] : face[0];
      const double e2 = (coord[3 * b] - coord[3 * a]) * (coord[3 * b] - coord[3 * a]) +
                        (coord[3 * b + 1] - coord[3 * a + 1]) * (coord[3 * b + 1] - coord[3 * a + 1]) +
                        (coord[3 * b + 2] - coord[3 * a + 2]) * (coord[3 * b + 2] - coord[3 * a + 2]);
      if (e2 > mx2) mx2 = e2;
    }
    const double len = sqrt(mx2);
    for (int i = 0; i < 4; ++i)
      if (node_len[face[i]] < len) node_len[face[i]] = len; /* "always use the maximum characteristic length" */
  }
}

typedef struct
{
  double p1[3], p2[3], p3[3];
  double char_len;
  float  lo[3], hi[3]; /* ArborX::Box: float corners */
  int    n1, n2, nf[4];
} Tri;

/* ContactEntity::SetBoundingBox (src/nimble_contact_entity.cc:76-125): min/max of the vertices, inflated by
 * inflation_factor (0.15) * char_len; then narrowed to float as ArborX::Point does (src/contact/arborx_utils.h:85-90) */
static void
inflate_box(const double* lo, const double* hi, double char_len, float* flo, float* fhi)
{
  const double inflation_length = 0.15 * char_len;
  for (int d = 0; d < 3; ++d) {
    flo[d] = (float)(lo[d] - inflation_length);
    fhi[d] = (float)(hi[d] + inflation_length);
  }
}

/* ContactManager::Projection (src/nimble_contact_manager.cc:1549-1620), tolerance 1.e-8 (nimble_contact_manager.h:297) */
static int
projection(const double* p, const Tri* t, double* gap, double* normal, double* bary)
{
  const double tol = 1.e-8;
  double       u[3], v[3], w[3], n[3], cr[3];
  for (int i = 0; i < 3; ++i) {
    u[i] = t->p2[i] - t->p1[i];
    v[i] = t->p3[i] - t->p1[i];
    w[i] = p[i] - t->p1[i];
  }
  cross3(u, v, n);
  const double n_squared = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
  cross3(u, w, cr);
  const double alpha3 = (cr[0] * n[0] + cr[1] * n[1] + cr[2] * n[2]) / n_squared;
  cross3(w, v, cr);
  const double alpha2 = (cr[0] * n[0] + cr[1] * n[1] + cr[2] * n[2]) / n_squared;
  const double alpha1 = 1.0 - alpha2 - alpha3;
  const double tol2   = 1.0 + tol;
  const int    a1 = (alpha1 > -tol && alpha1 < tol2), a2 = (alpha2 > -tol && alpha2 < tol2), a3 = (alpha3 > -tol && alpha3 < tol2);
  if (!(a1 && a2 && a3)) return 0;
  const double xp = alpha1 * t->p1[0] + alpha2 * t->p2[0] + alpha3 * t->p3[0];
  const double yp = alpha1 * t->p1[1] + alpha2 * t->p2[1] + alpha3 * t->p3[1];
  const double zp = alpha1 * t->p1[2] + alpha2 * t->p2[2] + alpha3 * t->p3[2];
  const double dx = p[0] - xp, dy = p[1] - yp, dz = p[2] - zp;
  const double s  = 1.0 / sqrt(n_squared);
  normal[0] = n[0] * s, normal[1] = n[1] * s, normal[2] = n[2] * s;
  *gap      = dx * normal[0] + dy * normal[1] + dz * normal[2];
  bary[0] = alpha1, bary[1] = alpha2, bary[2] = alpha3;
  return (*gap < 0.0) && (*gap > -t->char_len); /* inside but not through */
}

/* The projection alone, for the known answers of the reference's unit test (unit_tests/projection_node_to_face.cc:146-279):
 * node[3], tri[9] = the facet's three vertices; gap / normal[3] / bary[3] are written only when the projection falls
 * inside the facet's barycentric window, as in the reference; returns the reference's `in` flag. */
int
h8o_contact_projection(const double* node, const double* tri, double char_len, double* gap, double* normal, double* bary)
{
  Tri t;
  memset(&t, 0, sizeof t);
  for (int d = 0; d < 3; ++d) t.p1[d] = tri[d], t.p2[d] = tri[3 + d], t.p3[d] = tri[6 + d];
  t.char_len = char_len;
  return projection(node, &t, gap, normal, bary);
}

/* One evaluation of the contact force.
 *   ref, disp, contact_force: [n_nodes][3]; contact_force is overwritten (zero away from the contact surfaces,
 *     ContactManager::GetForces, src/nimble_contact_manager.cc:732-748)
 *   prim_quads [n_prim][4]: skin faces of the primary blocks, each split into 4 triangles (node i, node i+1, face
 *     centre) as CreateContactNodesAndFaces does (:1043-1190); prim_len their characteristic lengths
 *   sec_nodes [n_sec], sec_len [n_sec]: contact nodes of the secondary blocks and their characteristic lengths
 *   status (optional): [4*n_prim + n_sec] contact_status flags, triangles first
 * Pairs are visited node-major, triangles ascending; returns the number of enforced pairs. */
long
h8o_contact_force(double penalty, long n_nodes, const double* ref, const double* disp, long n_prim, const int* prim_quads,
                  const double* prim_len, long n_sec, const int* sec_nodes, const double* sec_len, double* contact_force,
                  unsigned char* status)
{
  const long n_tri = 4 * n_prim;
  Tri*       tri   = (Tri*)malloc((size_t)(n_tri > 0 ? n_tri : 1) * sizeof(Tri));
  long       pairs = 0;
  memset(contact_force, 0, (size_t)n_nodes * 3 * sizeof(double));
  if (status) memset(status, 0, (size_t)(n_tri + n_sec));
  /* ApplyDisplacements (:750-786): coord = model_coord + displacement; ContactEntity::SetCoordinates
   * (src/nimble_contact_entity.h:236-258): the third vertex is the mean of the face's four nodes */
  for (long f = 0; f < n_prim; ++f) {
    const int* q = prim_quads + 4 * f;
    double     c[4][3], ctr[3];
    for (int i = 0; i < 4; ++i)
      for (int d = 0; d < 3; ++d) c[i][d] = ref[3 * q[i] + d] + disp[3 * q[i] + d];
    for (int d = 0; d < 3; ++d) ctr[d] = (c[0][d] + c[1][d] + c[2][d] + c[3][d]) / 4.0;
    for (int k = 0; k < 4; ++k) {
      Tri*      t = &tri[4 * f + k];
      const int a = k, b = (k + 1) % 4;
      double    lo[3], hi[3];
      for (int d = 0; d < 3; ++d) {
        t->p1[d] = c[a][d], t->p2[d] = c[b][d], t->p3[d] = ctr[d];
        lo[d] = hi[d] = t->p1[d];
        if (t->p2[d] < lo[d]) lo[d] = t->p2[d];
        if (t->p2[d] > hi[d]) hi[d] = t->p2[d];
        if (t->p3[d] < lo[d]) lo[d] = t->p3[d];
        if (t->p3[d] > hi[d]) hi[d] = t->p3[d];
      }
      t->char_len = prim_len[f];
      t->n1 = q[a], t->n2 = q[b];
      for (int i = 0; i < 4; ++i) t->nf[i] = q[i];
      inflate_box(lo, hi, t->char_len, t->lo, t->hi);
    }
  }
  for (long s = 0; s < n_sec; ++s) {
    const int node = sec_nodes[s];
    double    p[3];
    float     lo[3], hi[3];
    for (int d = 0; d < 3; ++d) p[d] = ref[3 * node + d] + disp[3 * node + d];
    inflate_box(p, p, sec_len[s], lo, hi);
    for (long j = 0; j < n_tri; ++j) {
      const Tri* t = &tri[j];
      /* ArborX::intersects(Box) on float boxes: closed intervals overlap in every direction */
      if (hi[0] < t->lo[0] || lo[0] > t->hi[0] || hi[1] < t->lo[1] || lo[1] > t->hi[1] || hi[2] < t->lo[2] || lo[2] > t->hi[2]) continue;
      double gap = 0.0, normal[3] = {0., 0., 0.}, bary[3] = {0., 0., 0.};
      if (!projection(p, t, &gap, normal, bary)) continue;
      if (status) status[j] = 1, status[n_tri + s] = 1;
      ++pairs;
      /* PenaltyContactEnforcement::EnforceContact (src/nimble_contact_manager.h:94-128), details::getContactForce
       * (:80-85), ContactEntity::SetNodalContactForces (src/nimble_contact_entity.h:420-447) and
       * ScatterForceToContactManagerForceVector (:260-299): face first, then the node */
      const double scale = penalty * gap;
      double       cf[3];
      for (int i = 0; i < 3; ++i) cf[i] = scale * normal[i];
      for (int d = 0; d < 3; ++d) contact_force[3 * t->n1 + d] += bary[0] * cf[d];
      for (int d = 0; d < 3; ++d) contact_force[3 * t->n2 + d] += bary[1] * cf[d];
      for (int i = 0; i < 4; ++i)
        for (int d = 0; d < 3; ++d) contact_force[3 * t->nf[i] + d] += (bary[2] * cf[d]) / 4.0;
      for (int d = 0; d < 3; ++d) contact_force[3 * node + d] += -cf[d];
    }
  }
  free(tri);
  return pairs;
}

/* the acceleration line with contact (src/integrators/explicit_time_integrator.cc:243-249) */
void
h8o_accel_contact(long n_nodes, const double* mass, const double* f_int, const double* f_ext, const double* f_contact, double* a)
{
  for (long i = 0; i < n_nodes; ++i) {
    const double oneOverM = 1.0 / mass[i];
    for (int c = 0; c < 3; ++c) a[3 * i + c] = oneOverM * (f_int[3 * i + c] + (f_ext ? f_ext[3 * i + c] : 0.0) + f_contact[3 * i + c]);
  }
}
