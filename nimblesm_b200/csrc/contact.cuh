// nimblesm_b200/csrc/contact.cuh — penalty contact on the device (sm_100a, fp64, no FMA contraction): the
// `contact_force` term of the explicit loop (src/integrators/explicit_time_integrator.cc:232-249).
//
// What the reference computes per step (ArborXSerialContactManager::ComputeSerialContactForce,
// src/contact/serial/arborx_serial_contact_manager.cc:147-196): every secondary NODE against every primary TRIANGLE
// (a skin quad split in four around its centre, ContactManager::CreateContactNodesAndFaces,
// src/nimble_contact_manager.cc:1043-1190) whose inflated bounding boxes intersect (ArborX BVH on float boxes); pairs
// that ContactManager::Projection (:1549-1620) places inside the facet, penetrating by less than the facet's
// characteristic length, receive penalty * gap along the facet normal, spread over the facet's nodes by the barycentric
// coordinates (the centre's share goes to the quad's four nodes in quarters) and, negated, to the node
// (PenaltyContactEnforcement::EnforceContact, src/nimble_contact_manager.h:94-128).  The result is a sum over all
// accepted pairs: the search structure decides only how fast the pairs are found.
//
// Execution plan (B200): the contact surfaces are 2-D, so the work is tiny next to the element kernel and latency, not
// bandwidth, is what counts -- three launches per evaluation, no host synchronisation, no sort:
//   contact_update_kernel   thread per quad / per contact node: current coordinates (X + u), the quad's centre, the four
//                           triangles' inflated boxes (narrowed to float exactly as ArborX::Point does), the grid-wide
//                           bounding box of the triangles and the largest box extent (warp -> CTA -> one atomic per
//                           quantity), contact force of the touched nodes cleared, hash heads cleared
//   contact_bin_kernel      thread per QUAD (its four triangles share three quarters of their boxes): cell of the
//                           minimum corner of the union box on a uniform grid of pitch h >= every box extent, pushed
//                           on the chain of its hash bucket (atomicExch, no scan, no sort); thread per contact node:
//                           the nodes whose box meets the triangles' bounding box are compacted into a list
//   contact_pair_kernel     one WARP per LISTED contact node (a device-filling grid strides over the list), in rounds:
//                           lane l < 27 advances along the chain of neighbour cell l -- a box of extent <= h anchored
//                           in cell n can only meet boxes anchored in n + {-1,0,1}^3 -- to its next quad with a float-box
//                           hit, tested triangle by triangle exactly as ArborX tests them; the finds are compacted and
//                           all 32 lanes share out the (quad, triangle) items: projection and enforcement in the
//                           reference's operation order (so each pair's force has the oracle's bits),
//                           red.global.add.f64 into the nodal contact force (the order of the sum over pairs is not
//                           fixed: noise ~1e-16, as in the reference's own Kokkos::atomic_add scatter).
// What bounds it and what each step of the above bought (544 -> 106 us at 640 k triangles / 160 k nodes): DESIGN.md §3.8.
#pragma once
#include <float.h>
#include <stdint.h>

namespace nsm {

struct alignas(32) QuadBin
{
  long long cell[3];
  int       next;
  int       pad;
};

struct ContactArgs
{
  int64_t       n_quads, n_sec;
  const int*    quad;      // [n_quads][4] node ids (internal numbering), Exodus face order
  const double* quad_len;  // [n_quads] characteristic length (largest edge in the model configuration)
  const int*    sec_node;  // [n_sec]
  const double* sec_len;   // [n_sec]
  int64_t       n_surf;    // nodes of the contact sub-model (of the quads and the contact nodes), each once
  const int*    surf_node; // [n_surf]
  const double* X[3];
  const double* u[3];
  double*       fc[3];     // nodal contact force, SoA
  double        penalty;
  // per-evaluation scratch
  double*    quad_xyz;  // [n_quads][15]: the four corners and the centre, current configuration
  float*     tri_box;   // [4 n_quads][6]: lo xyz, hi xyz
  QuadBin*   bin;       // [n_quads] cell of the quad's union box + next quad of the hash chain
  int*       head;      // [table_mask + 1]
  unsigned   table_mask;
  unsigned*  red;       // [8] this evaluation: ordered-float min corner x, y, z of the triangles' boxes; max extent of any
                        //     box (float bits, >= 0); ordered-float max corner x, y, z of the triangles' boxes; pad
  unsigned*  red_next;  // [8] the next evaluation's, reset here
  int*       near_list; // [n_sec] contact nodes whose box meets the bounding box of all triangles (this evaluation)
  unsigned long long* counters;  // [0] enforced pairs, [1] pairs that passed the box test, [4] length of near_list
  unsigned char*      status;    // [4 n_quads + n_sec] contact_status flags of this evaluation
  // ORDERED assembly: the seven nodal contributions of every accepted pair are FILED with their place in the serial
  // order (contact node, triangle, slot) instead of being added atomically; contact_ordered_sum_kernel adds them per
  // target node in that order, so the contact force has the serial walk's bits (counters[5] = contributions filed,
  // counters[6] = pairs that did not fit and were added atomically)
  int                 ordered;
  long long           contrib_cap;
  unsigned long long* contrib_key;     // [cap] contact node << 33 | triangle << 3 | slot
  int*                contrib_target;  // [cap]
  double*             contrib_val;     // [cap][3]
};

__device__ __forceinline__ unsigned
ordered_float(float f)
{  // monotone map float -> unsigned (so that min/max work through integer atomics)
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ float
float_of_ordered(unsigned o)
{
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// ContactEntity::SetBoundingBox (src/nimble_contact_entity.cc:76-125): vertex min/max inflated by 0.15 * char_len;
// narrowed to float as ArborX::Point holds it (src/contact/arborx_utils.h:85-90)
__device__ __forceinline__ void
inflate_to_float(const double lo[3], const double hi[3], double char_len, float* out)
{
  const double inflation_length = 0.15 * char_len;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    out[d]     = (float)(lo[d] - inflation_length);
    out[3 + d] = (float)(hi[d] + inflation_length);
  }
}

__global__ void __launch_bounds__(256)
contact_update_kernel(const ContactArgs p)
{
  // The per-quad records leave through shared memory: a thread's 15 doubles / 24 floats are 120 / 96 bytes apart from
  // its neighbour's, and stored lane by lane every instruction touched 32 sectors -- the load/store unit's queue was
  // where this kernel waited (ncu: lg_throttle 40 of 90 stall cycles per issue, profiles/r02S_*).  Staged, the CTA
  // writes its 256 records as contiguous runs.
  __shared__ double   stage[256 * 15];
  __shared__ unsigned part[8][7];
  const int64_t t   = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  const int     l   = threadIdx.x;
  for (int64_t i = t; i <= (int64_t)p.table_mask; i += nth) p.head[i] = -1;
  // the contact force of every node of the contact sub-model, once (the quads share their nodes four ways)
  for (int64_t i = t; i < p.n_surf; i += nth) {
    const int nd = p.surf_node[i];
    p.fc[0][nd] = 0.0, p.fc[1][nd] = 0.0, p.fc[2][nd] = 0.0;
  }
  for (int64_t i = t; i < 4 * p.n_quads + p.n_sec; i += nth) p.status[i] = 0;
  if (t < 8) p.red_next[t] = t < 3 ? 0xffffffffu : 0u;
  if (t < 2) p.counters[t] = 0ull;
  if (t == 2) p.counters[4] = 0ull;
  if (t == 3) p.counters[5] = p.counters[6] = 0ull;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, ext = 0.0f;
  float box[4][6];
  const int64_t block_base = (int64_t)blockIdx.x * blockDim.x;
  const int     n_here     = (int)max((int64_t)0, min((int64_t)blockDim.x, p.n_quads - block_base));  // quads of this CTA
  if (t < p.n_quads) {
    // ContactManager::ApplyDisplacements (src/nimble_contact_manager.cc:750-786) + ContactEntity::SetCoordinates
    // (src/nimble_contact_entity.h:236-258): the third vertex of every triangle is the mean of the quad's nodes
    double c[4][3], ctr[3];
    const int4 qn = *reinterpret_cast<const int4*>(p.quad + 4 * t);
    const int  node[4] = {qn.x, qn.y, qn.z, qn.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int d = 0; d < 3; ++d) c[k][d] = p.X[d][node[k]] + p.u[d][node[k]];
#pragma unroll
    for (int d = 0; d < 3; ++d) ctr[d] = (c[0][d] + c[1][d] + c[2][d] + c[3][d]) / 4.0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int d = 0; d < 3; ++d) stage[15 * l + 3 * k + d] = c[k][d];
#pragma unroll
    for (int d = 0; d < 3; ++d) stage[15 * l + 12 + d] = ctr[d];
    const double len = p.quad_len[t];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = (k + 1) & 3;
      double    blo[3], bhi[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        blo[d] = bhi[d] = c[k][d];
        if (c[b][d] < blo[d]) blo[d] = c[b][d];
        if (c[b][d] > bhi[d]) bhi[d] = c[b][d];
        if (ctr[d] < blo[d]) blo[d] = ctr[d];
        if (ctr[d] > bhi[d]) bhi[d] = ctr[d];
      }
      inflate_to_float(blo, bhi, len, box[k]);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        lo[d] = fminf(lo[d], box[k][d]);
        hi[d] = fmaxf(hi[d], box[k][3 + d]);
      }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) ext = fmaxf(ext, hi[d] - lo[d]);  // the quad is binned by the union of its triangles' boxes
  }
  __syncthreads();
  for (int i = l; i < 15 * n_here; i += blockDim.x) p.quad_xyz[15 * block_base + i] = stage[i];
  __syncthreads();
  float* fstage = reinterpret_cast<float*>(stage);
  if (t < p.n_quads) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int d = 0; d < 6; ++d) fstage[24 * l + 6 * k + d] = box[k][d];
  }
  __syncthreads();
  for (int i = l; i < 24 * n_here; i += blockDim.x) p.tri_box[24 * block_base + i] = fstage[i];
  if (t < p.n_sec) {
    const int nd = p.sec_node[t];
    double    x[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) x[d] = p.X[d][nd] + p.u[d][nd];
    float nbox[6];
    inflate_to_float(x, x, p.sec_len[t], nbox);
#pragma unroll
    for (int d = 0; d < 3; ++d) ext = fmaxf(ext, nbox[3 + d] - nbox[d]);  // (the grid is anchored at the triangles' boxes only)
  }
  // grid-wide corners of the triangles' boxes and maximum extent: warp reduction, then one atomic per CTA and quantity
  unsigned r[7] = {ordered_float(lo[0]), ordered_float(lo[1]), ordered_float(lo[2]), __float_as_uint(ext),
                   ordered_float(hi[0]), ordered_float(hi[1]), ordered_float(hi[2])};
#pragma unroll
  for (int d = 0; d < 3; ++d) r[d] = __reduce_min_sync(0xffffffffu, r[d]);
#pragma unroll
  for (int d = 3; d < 7; ++d) r[d] = __reduce_max_sync(0xffffffffu, r[d]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < 7; ++d) part[warp][d] = r[d];
  }
  __syncthreads();
  if (warp == 0) {
    unsigned v[7];
#pragma unroll
    for (int d = 0; d < 7; ++d) v[d] = lane < 8 ? part[lane][d] : (d < 3 ? 0xffffffffu : 0u);
#pragma unroll
    for (int d = 0; d < 3; ++d) v[d] = __reduce_min_sync(0xffffffffu, v[d]);
#pragma unroll
    for (int d = 3; d < 7; ++d) v[d] = __reduce_max_sync(0xffffffffu, v[d]);
    if (lane == 0) {
#pragma unroll
      for (int d = 0; d < 3; ++d) atomicMin(p.red + d, v[d]);
#pragma unroll
      for (int d = 3; d < 7; ++d) atomicMax(p.red + d, v[d]);
    }
  }
}

// pitch of the uniform grid: a little more than the largest box extent, so that rounding in the cell computation can
// never move a box corner across more than one cell boundary
__device__ __forceinline__ double
contact_pitch(const unsigned* red)
{
  const float ext = __uint_as_float(red[3]);
  return ext > 0.0f ? 1.0001 * (double)ext : 1.0;
}

__device__ __forceinline__ void
contact_cell(const float* lo, const unsigned* red, double pitch, long long cell[3])
{
#pragma unroll
  for (int d = 0; d < 3; ++d) cell[d] = (long long)floor(((double)lo[d] - (double)float_of_ordered(red[d])) / pitch);
}

__device__ __forceinline__ unsigned
contact_hash(long long cx, long long cy, long long cz)
{
  unsigned long long h = (unsigned long long)cx * 0x9E3779B97F4A7C15ull;
  h ^= (unsigned long long)cy * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
  h ^= (unsigned long long)cz * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
  return (unsigned)(h ^ (h >> 32));
}

__global__ void __launch_bounds__(256)
contact_bin_kernel(const ContactArgs p)
{
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < p.n_quads) {
    const float* tb = p.tri_box + 24 * q;
    float        lo[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) lo[d] = fminf(fminf(tb[d], tb[6 + d]), fminf(tb[12 + d], tb[18 + d]));
    const double pitch = contact_pitch(p.red);
    long long    cell[3];
    contact_cell(lo, p.red, pitch, cell);
    QuadBin& e = p.bin[q];
    e.cell[0] = cell[0], e.cell[1] = cell[1], e.cell[2] = cell[2];
    e.next = atomicExch(p.head + (contact_hash(cell[0], cell[1], cell[2]) & p.table_mask), (int)q);
  }
  // The contact nodes worth a search: those whose box meets the bounding box of all triangles (most skin nodes of a
  // body do not).  They are COMPACTED into a list, so that the pair kernel's CTAs hold eight working warps each: with
  // one warp per contact node of the whole skin, a CTA lived as long as its one or two interface nodes and kept six
  // idle warp slots occupied (r02p -> r02v).
  bool near = false;
  if (q < p.n_sec) {
    const int nd = p.sec_node[q];
    double    x[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) x[d] = p.X[d][nd] + p.u[d][nd];
    float box[6];
    inflate_to_float(x, x, p.sec_len[q], box);
    near = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) near = near && !(box[3 + d] < float_of_ordered(p.red[d]) || box[d] > float_of_ordered(p.red[4 + d]));
  }
  const unsigned vote = __ballot_sync(0xffffffffu, near);
  if (vote) {
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(p.counters + 4, (unsigned long long)__popc(vote));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (near) p.near_list[base + __popc(vote & ((1u << lane) - 1u))] = (int)q;
  }
}

__device__ __forceinline__ void
contact_cross(const double* u, const double* v, double* r)
{  // CrossProduct (src/nimble_utils.h:393-400)
  r[0] = u[1] * v[2] - u[2] * v[1];
  r[1] = u[2] * v[0] - u[0] * v[2];
  r[2] = u[0] * v[1] - u[1] * v[0];
}

__device__ __forceinline__ void
contact_add3(double* const fc[3], int node, double x, double y, double z)
{
  atomicAdd(fc[0] + node, x);
  atomicAdd(fc[1] + node, y);
  atomicAdd(fc[2] + node, z);
}

// One accepted-or-not (node, triangle) pair: ContactManager::Projection (src/nimble_contact_manager.cc:1549-1620,
// tolerance 1.e-8) and, for a node inside the facet that has not gone through it,
// PenaltyContactEnforcement::EnforceContact (src/nimble_contact_manager.h:94-128): facet first, then the node.
__device__ __forceinline__ bool
contact_pair(const ContactArgs& p, int64_t s, int nd, const double pt[3], int quad, int k)
{
  const double* q  = p.quad_xyz + 15 * (int64_t)quad;
  const int     kb = (k + 1) & 3;
  double        p1[3], p2[3], p3[3], u[3], v[3], w[3], n[3], cr[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) p1[d] = q[3 * k + d], p2[d] = q[3 * kb + d], p3[d] = q[12 + d];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    u[d] = p2[d] - p1[d];
    v[d] = p3[d] - p1[d];
    w[d] = pt[d] - p1[d];
  }
  contact_cross(u, v, n);
  const double n_squared = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
  contact_cross(u, w, cr);
  const double alpha3 = (cr[0] * n[0] + cr[1] * n[1] + cr[2] * n[2]) / n_squared;
  contact_cross(w, v, cr);
  const double alpha2 = (cr[0] * n[0] + cr[1] * n[1] + cr[2] * n[2]) / n_squared;
  const double alpha1 = 1.0 - alpha2 - alpha3;
  const double tol = 1.e-8, tol2 = 1.0 + tol;
  if (!((alpha1 > -tol && alpha1 < tol2) && (alpha2 > -tol && alpha2 < tol2) && (alpha3 > -tol && alpha3 < tol2))) return false;
  const double xp = alpha1 * p1[0] + alpha2 * p2[0] + alpha3 * p3[0];
  const double yp = alpha1 * p1[1] + alpha2 * p2[1] + alpha3 * p3[1];
  const double zp = alpha1 * p1[2] + alpha2 * p2[2] + alpha3 * p3[2];
  const double dx = pt[0] - xp, dy = pt[1] - yp, dz = pt[2] - zp;
  const double sc = 1.0 / sqrt(n_squared);
  const double nx = n[0] * sc, ny = n[1] * sc, nz = n[2] * sc;
  const double gap = dx * nx + dy * ny + dz * nz;
  if (!((gap < 0.0) && (gap > -p.quad_len[quad]))) return false;  // inside but not through
  p.status[4 * (int64_t)quad + k] = 1, p.status[4 * p.n_quads + s] = 1;
  const double scale = p.penalty * gap;
  const double cf[3] = {scale * nx, scale * ny, scale * nz};
  const int*   qn    = p.quad + 4 * (int64_t)quad;
  const double f3[3] = {(alpha3 * cf[0]) / 4.0, (alpha3 * cf[1]) / 4.0, (alpha3 * cf[2]) / 4.0};
  if (p.ordered) {
    const unsigned long long at = atomicAdd(p.counters + 5, 7ull);
    if ((long long)at + 7 <= p.contrib_cap) {
      const unsigned long long key = ((unsigned long long)s << 33) | ((unsigned long long)(4 * quad + k) << 3);
      const int    target[7] = {qn[k], qn[kb], qn[0], qn[1], qn[2], qn[3], nd};
      const double value[7][3] = {{alpha1 * cf[0], alpha1 * cf[1], alpha1 * cf[2]}, {alpha2 * cf[0], alpha2 * cf[1], alpha2 * cf[2]},
                                  {f3[0], f3[1], f3[2]}, {f3[0], f3[1], f3[2]}, {f3[0], f3[1], f3[2]}, {f3[0], f3[1], f3[2]},
                                  {-cf[0], -cf[1], -cf[2]}};
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        p.contrib_key[at + i]    = key | (unsigned long long)i;
        p.contrib_target[at + i] = target[i];
        p.contrib_val[3 * (at + i)] = value[i][0], p.contrib_val[3 * (at + i) + 1] = value[i][1], p.contrib_val[3 * (at + i) + 2] = value[i][2];
      }
      return true;
    }
    atomicAdd(p.counters + 6, 1ull);  // the lists are full: this pair goes the atomic way
  }
  contact_add3(p.fc, qn[k], alpha1 * cf[0], alpha1 * cf[1], alpha1 * cf[2]);
  contact_add3(p.fc, qn[kb], alpha2 * cf[0], alpha2 * cf[1], alpha2 * cf[2]);
#pragma unroll
  for (int i = 0; i < 4; ++i) contact_add3(p.fc, qn[i], f3[0], f3[1], f3[2]);
  contact_add3(p.fc, nd, -cf[0], -cf[1], -cf[2]);
  return true;
}

// One WARP per listed contact node, in rounds: (1) lane l < 27 advances along the hash chain of neighbour cell l until it
// holds a quad with at least one triangle whose float box meets the node's (ArborX::intersects, closed intervals);
// (2) the lanes' finds are compacted into the warp's list; (3) ALL 32 lanes share out the (quad, triangle) items of the
// list -- projection and enforcement run side by side instead of one triangle after another on the few lanes whose
// cells are occupied (a contact surface fills ~9 of a node's 27 neighbour cells).  (Filing the whole chains in one pass
// through a shared counter was measured slower, 116 vs 105 us: the overflow path costs the hot loop its registers.)
#ifndef NSM_CONTACT_MIN_BLOCKS
#define NSM_CONTACT_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(256, NSM_CONTACT_MIN_BLOCKS)
contact_pair_kernel(const ContactArgs p)
{
  constexpr int kCand = 64;  // candidate quads per round (a node's 27 cells normally hold ~20)
  __shared__ int      cand[8][kCand];
  __shared__ unsigned items[8][4 * kCand];  // quad << 2 | triangle, box test passed
  __shared__ int      n_cand[8], n_items[8];
  const int        warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t    n_near = (int64_t)p.counters[4], n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  unsigned         tested = 0, enforced = 0;
  for (int64_t slot = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; slot < n_near; slot += n_warps) {  // (whole warps)
  const int64_t s  = p.near_list[slot];
  const int     nd = p.sec_node[s];
  double    pt[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) pt[d] = p.X[d][nd] + p.u[d][nd];
  float box[6];
  inflate_to_float(pt, pt, p.sec_len[s], box);
  long long cell[3];
  contact_cell(box, p.red, contact_pitch(p.red), cell);
  int quad = -1;
  if (lane < 27) {
    cell[0] += lane % 3 - 1, cell[1] += (lane / 3) % 3 - 1, cell[2] += lane / 9 - 1;
    quad = p.head[contact_hash(cell[0], cell[1], cell[2]) & p.table_mask];
  }
  for (;;) {
    if (lane == 0) n_cand[warp] = 0, n_items[warp] = 0;
    __syncwarp();
    // (1) the divergent part, kept thin: each lane walks its chain and lists the quads of its cell
    while (quad >= 0) {
      const QuadBin e = p.bin[quad];
      if (e.cell[0] == cell[0] && e.cell[1] == cell[1] && e.cell[2] == cell[2]) {  // (else another cell of the same bucket)
        const int at = atomicAdd(&n_cand[warp], 1);
        if (at >= kCand) break;  // list full: this quad waits for the next round
        cand[warp][at] = quad;
      }
      quad = e.next;
    }
    __syncwarp();
    const int nc = min(n_cand[warp], kCand);
    __syncwarp();  // (everyone has read the count before lane 0 of the next node resets it)
    if (nc == 0) break;
    // (2) box tests, one candidate quad per lane
    for (int c = lane; c < nc; c += 32) {
      const int    at  = cand[warp][c];
      const float* tb  = p.tri_box + 24 * (int64_t)at;
      unsigned     hit = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float b0 = tb[6 * k], b1 = tb[6 * k + 1], b2 = tb[6 * k + 2], b3 = tb[6 * k + 3], b4 = tb[6 * k + 4], b5 = tb[6 * k + 5];
        if (!(box[3] < b0 || box[0] > b3 || box[4] < b1 || box[1] > b4 || box[5] < b2 || box[2] > b5)) hit |= 1u << k;
      }
      if (hit) {
        const int nh = __popc(hit);
        tested += nh;
        int at_item = atomicAdd(&n_items[warp], nh);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((hit >> k) & 1u) items[warp][at_item++] = ((unsigned)at << 2) | (unsigned)k;
      }
    }
    __syncwarp();
    // (3) all 32 lanes share out the (quad, triangle) items
    const int ni = n_items[warp];
    for (int item = lane; item < ni; item += 32) {
      const unsigned f = items[warp][item];
      if (contact_pair(p, s, nd, pt, (int)(f >> 2), (int)(f & 3u))) ++enforced;
    }
    __syncwarp();
  }
  }
  // counters: one pair of global atomics per warp that tested anything
  tested   = __reduce_add_sync(0xffffffffu, tested);
  enforced = __reduce_add_sync(0xffffffffu, enforced);
  if (lane == 0 && tested) {
    atomicAdd(p.counters + 1, (unsigned long long)tested);
    if (enforced) atomicAdd(p.counters, (unsigned long long)enforced);
  }
}

// ORDERED assembly, after the contributions have been sorted by (target node; contact node, triangle, slot): the thread
// at the head of a target's run adds the run in order -- PenaltyContactEnforcement::EnforceContact's scatter order inside a
// pair (facet node 1, facet node 2, the quad's four nodes, the contact node), pairs node-major with triangles ascending,
// which is the order of the serial walk (oracle/contact_oracle.c) -- and adds the sum to the (cleared) contact force.
__global__ void __launch_bounds__(256)
contact_gather_targets_kernel(long long n, const unsigned* __restrict__ order, const int* __restrict__ target, unsigned* out)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (unsigned)target[order[i]];
}

__global__ void __launch_bounds__(256)
contact_ordered_sum_kernel(long long n, const unsigned* __restrict__ sorted_target, const unsigned* __restrict__ order,
                           const double* __restrict__ val, double* f0, double* f1, double* f2)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned t = sorted_target[i];
  if (i > 0 && sorted_target[i - 1] == t) return;
  double a = 0.0, b = 0.0, c = 0.0;
  for (long long j = i; j < n && sorted_target[j] == t; ++j) {
    const double* v = val + 3 * (long long)order[j];
    a += v[0], b += v[1], c += v[2];
  }
  f0[t] = f0[t] + a, f1[t] = f1[t] + b, f2[t] = f2[t] + c;
}

__global__ void __launch_bounds__(256)
contact_iota_kernel(long long n, unsigned* out)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (unsigned)i;
}

// numActiveContactFaces / numActiveContactNodes (src/nimble_contact_manager.cc:692-714): entities whose contact_status
// was set by the last evaluation
__global__ void __launch_bounds__(256)
contact_count_kernel(int64_t n_tri, int64_t n_sec, const unsigned char* status, unsigned long long* out /* [2] */)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool    f = t < n_tri && status[t];
  const bool    n = t >= n_tri && t < n_tri + n_sec && status[t];
  const unsigned bf = __ballot_sync(0xffffffffu, f), bn = __ballot_sync(0xffffffffu, n);
  if ((threadIdx.x & 31) == 0) {
    if (bf) atomicAdd(out, (unsigned long long)__popc(bf));
    if (bn) atomicAdd(out + 1, (unsigned long long)__popc(bn));
  }
}

}  // namespace nsm
