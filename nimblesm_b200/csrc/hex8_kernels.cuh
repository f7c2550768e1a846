// nimblesm_b200/csrc/hex8_kernels.cuh — CUDA kernels of the hex8 explicit step (sm_100a, fp64, no FMA
// contraction).  See DESIGN.md §3 for the execution plan; in short:
//
//   element kernel  (hot; FP64-issue bound)   8 lanes = 1 element, lane q owns Gauss point q.
//       gather   lane j loads node j (conn coalesced, SoA coordinates), stages X / x through shared memory
//       compute  gradient operator a, b at the lane's point -> b^-1 -> F -> stress -> a^-1 -> per-node
//                Bt.sigma.detJ shares (24 values per lane), all in registers
//       reduce   shares are transposed through (per-warp) shared memory; lane n sums node n over the Gauss
//                points 0..7 in the reference's order (src/nimble_element.h:575-612)
//       assemble red.global.add.f64 into the SoA nodal force (ATOMIC) or a coalesced store of the
//                element's [8][3] forces for the ordered node-side gather (ORDERED)
//   node kernels    (HBM bound)   a = (1/m)(f+f_ext), v += dt/2 a | v += dt/2 a, BC, u += dt v, BC
//   setup kernels   lumped mass, critical time step, volume averages (thread per element; not hot)
#pragma once
#include <stdint.h>

#include "hex8_math.cuh"

namespace nsm {

#ifndef NSM_ELEM_THREADS
#define NSM_ELEM_THREADS 256
#endif
#ifndef NSM_ELEM_MIN_BLOCKS
#define NSM_ELEM_MIN_BLOCKS 2
#endif
constexpr int kElemThreads   = NSM_ELEM_THREADS;  // 8 warps; a warp owns 4 elements ("group") per pass
constexpr int kElemWarps     = kElemThreads / 32;
// CTA shape of the element kernel per material.  16 warps per SM either way (the register file is full at 128 registers
// per thread); the elastic instance, whose passes are half as long, runs 2 % faster as four CTAs of four warps than as
// two of eight (2.588 vs 2.642 ms at 8 M elements, profiles/r02b_kernel_variants.txt `t128b4`), the others do not care.
#ifndef NSM_ELEM_THREADS_ELASTIC
#define NSM_ELEM_THREADS_ELASTIC 128
#endif
template <int MAT>
struct ElemShape
{
  static constexpr int threads    = MAT == 0 ? NSM_ELEM_THREADS_ELASTIC : NSM_ELEM_THREADS;
  static constexpr int warps      = threads / 32;
  static constexpr int min_blocks = (NSM_ELEM_THREADS * NSM_ELEM_MIN_BLOCKS) / threads;
};
constexpr int kElemsPerWarp  = 4;
#ifndef NSM_TICKET_CHUNK
#define NSM_TICKET_CHUNK 8
#endif
constexpr int kTicketChunk   = NSM_TICKET_CHUNK;   // groups per work ticket
constexpr int kCoordStride   = 4;                    // [c][j][e_w], rows 4..7 of a component shifted by 2 doubles (coord_row)
constexpr int kCoordComp     = 8 * kCoordStride + 2; // doubles per component
constexpr int kCoordDoubles  = 3 * kCoordComp;       // one [3][8] coordinate set of the warp's 4 elements
constexpr int kStageDoubles  = 2 * kCoordDoubles;    // X and u of one group, filled by cp.async
constexpr int kShareStride   = 34;                   // even: the transposed read takes Gauss-point pairs as 16-byte loads; rows 3 apart land 12 banks apart, conflict free
constexpr int kShareDoubles  = 24 * kShareStride;
constexpr int kEfStride      = 3;                    // ORDERED: (element, node) slots of fx, fy, fz, contiguous: neighbouring nodes share sectors (32-byte padded slots measured no faster, r01B)
constexpr int kBinvGroupDoubles = 9 * 32;            // cached b^-1 of one group: [9][32 lanes]
constexpr int kConnSlotDoubles = 3 * 32 / 2 + 2;               // connectivity of three groups in flight: [3][32 lanes] int; + the next chunk's skip-mask word
constexpr int kWarpSmemBase   = 2 * kStageDoubles + 2 * kCoordDoubles + kShareDoubles + kConnSlotDoubles;  // stages, K, C, shares, conn (+ staged b^-1)
#ifndef NSM_BINV_STAGE
#define NSM_BINV_STAGE 1
#endif
#ifndef NSM_BINV_STAGE_ELASTIC
#define NSM_BINV_STAGE_ELASTIC 0
#endif

struct ElemArgs
{
  int64_t       n_elem;
  const int*    conn;        // [n_elem][8], in SCHEDULE order (== file order unless the block was reordered)
  const int*    orig;        // schedule position -> element index in file order (nullptr: identity)
  const double* X[3];        // reference coordinates, SoA
  const double* u[3];        // displacement, SoA
  double*       f[3];        // nodal internal force, SoA (ATOMIC)
  double*       ef;          // [n_elem][8][kEfStride] element nodal forces (ORDERED), already offset to the block
  double*       ipt;         // [n_elem][8][15 + n_state] or nullptr: the records this launch writes (N+1)
  const double* ipt_n;       // history-dependent materials: the previous records (N), same layout
  double        mat_a, mat_b;  // material-specific parameters (j2_plasticity: yield stress, hardening modulus)
  double*       binv_cache;  // [n_groups][9][32] or nullptr, already offset to the block
  double        bulk, shear;
  unsigned*     ticket;      // next unclaimed chunk of 4-element groups of this launch (zero when the launch starts)
  unsigned*     ticket_next; // the counter of the NEXT launch on this stream: zeroed here, so that no memset node sits
                             // between two element launches (counters alternate; launches of a context are stream-ordered)
  int           zero;        // always 0: keeps ptxas from proving the ticket address warp-uniform (see draw_ticket)
  // Element schedule of this launch (multi-GPU overlap, nsm_b200_step): kSchedAll walks every group;
  // kSchedList walks group_list[0 .. n_list) (the groups that touch a node shared with another rank, run first so
  // that their forces can travel while the rest computes); kSchedSkipFlagged walks every group whose bit in
  // chunk_mask is 0 (the rest).  Group indices are 32-bit (a block holds < 2^33 elements): the loop-carried state
  // of the persistent warp must stay in registers (profiles/r01m: a 20-byte spill of it cost a local-memory
  // round trip through L2 per pass, 6 % of the warp's time).
  int                  sched;
  const unsigned char* chunk_mask;  // [chunks of kTicketChunk groups] bit i: group chunk*8+i touches a shared node
  const int*           group_list;
  int                  n_list;
  // kSchedAll over a RANGE of groups [group_begin, group_begin + n_range) (n_range == 0: all groups): the pipelined
  // host step launches the elements whose nodes have already been uploaded (nsm_b200_step_host)
  int                  group_begin, n_range;
  int*          flags;       // [0] bit 0: non-positive Jacobian seen; [1]: integration points redone in IEEE mode
};

__device__ __forceinline__ void
cp_async8(double* smem_dst, const double* gsrc)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
// Same copy, ordered after the value `dep` is available (a register dependence the compiler cannot hoist
// the copy above): used to overwrite a staging slot only once its previous content has been consumed.
__device__ __forceinline__ void
cp_async8_after(double* smem_dst, const double* gsrc, double dep)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("{ .reg .b64 t; mov.b64 t, %2; cp.async.ca.shared.global [%0], [%1], 8; }" ::"r"(d), "l"(gsrc), "d"(dep) : "memory");
}
__device__ __forceinline__ void
cp_async4(int* smem_dst, const int* gsrc)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void
cp_async_commit()
{
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void
cp_async_wait()
{
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void
prefetch_l2(const void* g)
{
  asm volatile("prefetch.global.L2 [%0];" ::"l"(g));
}
__device__ __forceinline__ void
prefetch_l1(const void* g)
{
  asm volatile("prefetch.global.L1 [%0];" ::"l"(g));
}

// Offset of node row j within a component of a coordinate set.  Rows 4..7 sit two doubles further: when lane
// (q, ew) touches row q (gather copies, X / u reads, K / C stores) the 16 lanes of a half-warp then fall into 16
// distinct bank pairs (un-shifted, rows q and q + 4 collide: profiles/r01q, 4x / 2x excess wavefronts), while
// the broadcast read of one node by all lanes still touches four consecutive doubles.
__host__ __device__ constexpr int
coord_row(int j)
{
  return j * kCoordStride + 2 * (j >> 2);
}

template <int J>
__device__ __forceinline__ void
load_node(const double* sm, int ew, double& x0, double& x1, double& x2)
{
  x0 = sm[0 * kCoordComp + coord_row(J) + ew];
  x1 = sm[1 * kCoordComp + coord_row(J) + ew];
  x2 = sm[2 * kCoordComp + coord_row(J) + ew];
}

template <int J>
__device__ __forceinline__ void
accumulate_pair(const ShapeAtPoint& sh, const double* sX, const double* sC, int ew, double (&a)[3][3],
                double (&b)[3][3])
{
  double x0, x1, x2;
  load_node<J>(sC, ew, x0, x1, x2);
  grad_accumulate<J>(sh, x0, x1, x2, a);
  load_node<J>(sX, ew, x0, x1, x2);
  grad_accumulate<J>(sh, x0, x1, x2, b);
}

template <int J>
__device__ __forceinline__ void
accumulate_one(const ShapeAtPoint& sh, const double* sC, int ew, double (&a)[3][3])
{
  double x0, x1, x2;
  load_node<J>(sC, ew, x0, x1, x2);
  grad_accumulate<J>(sh, x0, x1, x2, a);
}

enum { kSchedAll = 0, kSchedList = 1, kSchedSkipFlagged = 2 };
static_assert(kTicketChunk == 8, "the skip mask holds the flag bits of one chunk of 8 groups in a byte");

// MAT: nsm_material_kind; ORDERED: store element forces instead of atomics; MODE bit0: store F/sigma (always set
// for a material with state variables: its records are its memory), bit1: read cached b^-1 (filled once by
// binv_cache_kernel).
enum { kModeStoreIpt = 1, kModeReadBinv = 2 };

// How the cached b^-1 reaches the lane.  Neohookean passes are long (~1130 DP instructions per lane), so the
// next group's values are staged in shared memory by cp.async while the stress is computed.  Elastic passes
// are half as long and the nine extra LDGSTS per lane saturate the LSU queue (profiles/r01g: mio_throttle 2.7
// stalls per issue), so there the values are loaded straight from the L2-prefetched cache line.
template <int MAT, int MODE>
struct BinvStaged
{
  static constexpr bool value = NSM_BINV_STAGE && (MODE & kModeReadBinv) && (MAT == 1 || NSM_BINV_STAGE_ELASTIC);
};

template <int MAT, int MODE>
__host__ __device__ constexpr int
warp_smem_doubles()
{
  return kWarpSmemBase + (BinvStaged<MAT, MODE>::value ? kBinvGroupDoubles : 0);
}

template <int N>
__device__ __forceinline__ void
store_share(const GradProducts& gp, double det, const double (&s)[6], double* share, int lane)
{
  double f1, f2, f3;
  node_force_at_point<N>(gp, det, s, f1, f2, f3);
  share[(N * 3 + 0) * kShareStride + lane] = f1;
  share[(N * 3 + 1) * kShareStride + lane] = f2;
  share[(N * 3 + 2) * kShareStride + lane] = f3;
}

// Everything one lane does for its integration point: gradient operators -> b^-1 -> F -> stress -> a^-1 ->
// the 24 nodal-force shares (stored to the warp's share buffer).  Returns bit 0: a fast-path operand left
// its window (FAST only; the caller reruns the point with FAST = false), bit 1: non-positive Jacobian.
//   FAST = true : branch-free arithmetic (hex8_math.cuh) in ONE basic block; the force-path Jacobian reuses
//                 the F-path one (the caller sends the warp to FAST = false when they differ)
//   FAST = false: plain IEEE operators, force-path Jacobian always rebuilt from the cur coordinates
template <int MAT, int MODE, bool FAST>
__device__ __forceinline__ unsigned
integration_point(const ShapeAtPoint& sh, const double* sX, const double* sK, const double* sC, int ew, int lane,
                  const double* binv_row, double* binv_slot, const double* binv_next, double bulk, double shear,
                  double* share, double (&F)[9], double (&sig)[6], const double* rec_n = nullptr, double mat_a = 0.0,
                  double mat_b = 0.0, double* state_out = nullptr, const double* rec_staged = nullptr)
{
  unsigned bad = 0u, jac = 0u;
  double   a[3][3], binv[3][3];
  zero33(a);
  if (MODE & kModeReadBinv) {
    // FAST: this lane's nine values were staged in shared memory by cp.async during the previous pass;
    // IEEE redo: the staging slot may already be refilling, read the cache itself
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        binv[i][k] = (FAST && BinvStaged<MAT, MODE>::value) ? binv_slot[(3 * i + k) * 32] : __ldcs(binv_row + (3 * i + k) * 32);
    accumulate_one<0>(sh, sK, ew, a);
    accumulate_one<1>(sh, sK, ew, a);
    accumulate_one<2>(sh, sK, ew, a);
    accumulate_one<3>(sh, sK, ew, a);
    accumulate_one<4>(sh, sK, ew, a);
    accumulate_one<5>(sh, sK, ew, a);
    accumulate_one<6>(sh, sK, ew, a);
    accumulate_one<7>(sh, sK, ew, a);
  } else {
    double b[3][3];
    zero33(b);
    accumulate_pair<0>(sh, sX, sK, ew, a, b);
    accumulate_pair<1>(sh, sX, sK, ew, a, b);
    accumulate_pair<2>(sh, sX, sK, ew, a, b);
    accumulate_pair<3>(sh, sX, sK, ew, a, b);
    accumulate_pair<4>(sh, sX, sK, ew, a, b);
    accumulate_pair<5>(sh, sX, sK, ew, a, b);
    accumulate_pair<6>(sh, sX, sK, ew, a, b);
    accumulate_pair<7>(sh, sX, sK, ew, a, b);
    const double detb = invert3x3<FAST>(b, binv, bad);
    jac |= !(detb > 0.0) ? 2u : 0u;
  }
  def_grad_from(a, binv, F);
  if (FAST && BinvStaged<MAT, MODE>::value) {
    // F is formed, so the staged values have been consumed: refill this lane's slots with the next group's
    // b^-1 (lane-private slots: no cross-lane hazard; the copy lands while the stress is computed)
    if (binv_next) {
#pragma unroll
      for (int i = 0; i < 9; ++i) cp_async8_after(binv_slot + i * 32, binv_next + i * 32, F[i]);
    }
    cp_async_commit();
  }

  if (!FAST) {  // force-path Jacobian from cur = ref + disp (src/nimble_element.cc:462-463)
    zero33(a);
    accumulate_one<0>(sh, sC, ew, a);
    accumulate_one<1>(sh, sC, ew, a);
    accumulate_one<2>(sh, sC, ew, a);
    accumulate_one<3>(sh, sC, ew, a);
    accumulate_one<4>(sh, sC, ew, a);
    accumulate_one<5>(sh, sC, ew, a);
    accumulate_one<6>(sh, sC, ew, a);
    accumulate_one<7>(sh, sC, ew, a);
  }
  double       ai[3][3];
  const double det = invert3x3<FAST>(a, ai, bad);
  jac |= !(det > 0.0) ? 2u : 0u;

  if (MAT == 0) {
    stress_elastic(bulk, shear, F, sig);
  } else if (MAT == 1) {
    stress_neohookean<FAST>(bulk, shear, F, sig, bad);
  } else {
    // history-dependent material: F_n, sigma_n, state_n of this point from the previous record.  FAST: the warp's
    // four records were copied into the (still unused) share buffer by coalesced cp.async at the top of the pass --
    // the copy group before the gather's, hence wait_group 1; the IEEE redo reads the record itself (the fast pass has
    // overwritten the buffer with its shares).  Lanes beyond the block's last element see a virgin point.
    double        Fn[9], sn[6], stn[kMaxStateVars], st[kMaxStateVars];
    const double* rec = rec_n;
    if (FAST) {
      cp_async_wait<1>();
      __syncwarp();
      if (rec_n) rec = rec_staged;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) Fn[i] = rec ? rec[i] : (i < 3 ? 1.0 : 0.0);
#pragma unroll
    for (int i = 0; i < 6; ++i) sn[i] = rec ? rec[9 + i] : 0.0;
#pragma unroll
    for (int i = 0; i < kMaxStateVars; ++i) stn[i] = rec ? rec[15 + i] : 0.0;
    if (FAST) __syncwarp();  // every lane holds its record before any lane's shares overwrite the buffer
    stress_j2<FAST>(bulk, shear, mat_a, mat_b, Fn, F, sn, stn, sig, st, bad);
#pragma unroll
    for (int i = 0; i < kMaxStateVars; ++i) state_out[i] = st[i];
  }

  GradProducts gp;
  gp.init(sh, ai);
  store_share<0>(gp, det, sig, share, lane);
  store_share<1>(gp, det, sig, share, lane);
  store_share<2>(gp, det, sig, share, lane);
  store_share<3>(gp, det, sig, share, lane);
  store_share<4>(gp, det, sig, share, lane);
  store_share<5>(gp, det, sig, share, lane);
  store_share<6>(gp, det, sig, share, lane);
  store_share<7>(gp, det, sig, share, lane);
  return bad | jac;
}

// One ticket from the grid-wide work counter, drawn by lane 0.  ptxas rewrites an atomic on a provably
// warp-uniform address into its warp-aggregated form, whose shuffle waits for the L2 round trip on the spot
// (profiles/r01f: 0.66 long-scoreboard stalls per issue); `lane * zero` (a kernel argument that is always 0)
// hides the uniformity, so a plain ATOMG is issued and its result is not touched until the caller needs it.
__device__ __forceinline__ unsigned*
ticket_address(unsigned* ticket, int lane, int zero)
{
  int opaque = lane;
  asm volatile("" : "+r"(opaque));  // the compiler must not fold `lane` to 0 under `if (lane == 0)`
  return ticket + opaque * zero;
}

__device__ __forceinline__ int
claim_group(unsigned* address, int lane)
{
  unsigned t = 0;
  if (lane == 0) t = atomicAdd(address, 1u);
  return (int)__shfl_sync(0xffffffffu, t, 0);
}

// Node id of this lane's (element, local node) in group g, or -1 beyond the block's last element.
__device__ __forceinline__ int
group_node(const ElemArgs& p, int g, int ew, int q)
{
  const int64_t e = (int64_t)g * kElemsPerWarp + ew;
  return e < p.n_elem ? __ldg(p.conn + e * 8 + q) : -1;
}

// The same lookup, asynchronous: global -> this lane's connectivity slot (cp.async, lands with the copy group it
// is committed in); -1 for a group beyond the schedule or a lane beyond the block's last element.
__device__ __forceinline__ void
stage_group_node(const ElemArgs& p, int* slot, int g, int n_groups, int ew, int q)
{
  const int64_t e = (int64_t)g * kElemsPerWarp + ew;
  if (g < n_groups && e < p.n_elem)
    cp_async4(slot, p.conn + e * 8 + q);
  else
    *slot = -1;
}

// Asynchronous gather of one group's X and u into a stage buffer: lane j <- node j of its element, global ->
// shared without passing through registers.  Lanes beyond the last element stage a unit cube at rest, so
// the tail of the last group computes finite values (and raises no Jacobian flag).
__device__ __forceinline__ void
stage_gather(const ElemArgs& p, double* st, int node, int q, int ew)
{
  double* sx = st + coord_row(q) + ew;
  double* su = sx + kCoordDoubles;
  if (node >= 0) {
    cp_async8(sx + 0 * kCoordComp, p.X[0] + node);
    cp_async8(sx + 1 * kCoordComp, p.X[1] + node);
    cp_async8(sx + 2 * kCoordComp, p.X[2] + node);
    cp_async8(su + 0 * kCoordComp, p.u[0] + node);
    cp_async8(su + 1 * kCoordComp, p.u[1] + node);
    cp_async8(su + 2 * kCoordComp, p.u[2] + node);
  } else {
    sx[0 * kCoordComp] = ((q & 3) == 1 || (q & 3) == 2) ? 1.0 : 0.0;
    sx[1 * kCoordComp] = (q & 2) ? 1.0 : 0.0;
    sx[2 * kCoordComp] = (q & 4) ? 1.0 : 0.0;
    su[0 * kCoordComp] = 0.0, su[1 * kCoordComp] = 0.0, su[2 * kCoordComp] = 0.0;
  }
}

// Persistent: every warp walks the block's groups with the grid-wide warp stride.  While group g is computed
// (FP64-pipe bound, ~10 k DP lane-ops per element) the gather of group g+W is in flight (cp.async, double
// buffered), the connectivity of group g+2W is being loaded and the cached b^-1 of group g+W is pulled into L2.
#ifdef NSM_ELEM_MAXREG  // A/B builds: an explicit register cap instead of the one __launch_bounds__ derives
#define NSM_ELEM_BOUNDS __maxnreg__(NSM_ELEM_MAXREG)
#else
#define NSM_ELEM_BOUNDS __launch_bounds__(ElemShape<MAT>::threads, ElemShape<MAT>::min_blocks)
#endif
template <int MAT, bool ORDERED, int MODE>
__global__ void NSM_ELEM_BOUNDS
element_force_kernel(const ElemArgs p)
{
  extern __shared__ double smem[];
  const int     lane = threadIdx.x & 31;
  const int     warp = threadIdx.x >> 5;
  const int     q    = lane & 7;   // Gauss point (compute) == local node (gather / assemble)
  const int     ew   = lane >> 3;  // element within the group
  double*       wsm  = smem + warp * warp_smem_doubles<MAT, MODE>();
  double*       sK    = wsm + 2 * kStageDoubles;  // ref + ((ref + disp) - ref): F-path coordinates
  double*       sC    = sK + kCoordDoubles;         // ref + disp: force-path coordinates
  double*       share = sC + kCoordDoubles;         // [24][kShareStride] nodal-force shares of the 32 points
  const int     n_groups = (int)((p.n_elem + kElemsPerWarp - 1) / kElemsPerWarp);

  ShapeAtPoint sh;
  sh.init(q);
  if (blockIdx.x == 0 && threadIdx.x == 0) *p.ticket_next = 0u;

  // Groups are claimed from a grid-wide ticket counter in chunks of kTicketChunk consecutive groups: warps of
  // one scheduler advance at different rates, and a static stride left the slow ones to finish alone
  // (profiles/r01e: 14 of 16 warps active on average).  The next chunk's ticket is drawn when the current chunk
  // is entered and first read kTicketChunk passes later, so the atomic's round trip (which queues behind every
  // other warp's on the one counter) never shows; groups are looked up two passes ahead (connectivity load).
  unsigned* const ticket_at = ticket_address(p.ticket, lane, p.zero);
  const int n_tickets  = p.sched == kSchedList ? p.n_list : (p.n_range > 0 ? p.n_range : n_groups);  // positions the counter hands out
  int       chunk_base = claim_group(ticket_at, lane) * kTicketChunk;  // chunk that holds the position two passes ahead
  int       chunk_off  = -1;
  unsigned  ticket     = 0;                                            // lane 0: the chunk after that one
  if (lane == 0) ticket = atomicAdd(ticket_at, 1u);
  unsigned skip_mask = 0;  // kSchedSkipFlagged: the 8 flag bits of the current chunk
  if (p.sched == kSchedSkipFlagged && chunk_base < n_tickets) skip_mask = __ldg(p.chunk_mask + chunk_base / kTicketChunk);
  // ... and the mask word of the next chunk, copied to shared memory half a chunk ahead (its ticket has arrived by
  // then), so that the L2 round trip does not show at the rollover either
  int* const sMask = reinterpret_cast<int*>(wsm + kWarpSmemBase - 2);
  auto next_group = [&]() -> int {
    for (;;) {
      if (++chunk_off == kTicketChunk) {
        chunk_base = (int)__shfl_sync(0xffffffffu, ticket, 0) * kTicketChunk;
        chunk_off  = 0;
        if (lane == 0 && chunk_base < n_tickets) ticket = atomicAdd(ticket_at, 1u);
        if (p.sched == kSchedSkipFlagged && chunk_base < n_tickets) {
          cp_async_commit();  // (a rollover inside a run of skipped positions follows the copy at once)
          cp_async_wait<0>();
          __syncwarp();
          skip_mask = ((unsigned)*sMask >> (8 * ((chunk_base / kTicketChunk) & 3))) & 0xffu;
        }
      } else if (p.sched == kSchedSkipFlagged && chunk_off == kTicketChunk / 2) {
        const int next_chunk = (int)__shfl_sync(0xffffffffu, ticket, 0);
        __syncwarp();
        if (lane == 0 && next_chunk * kTicketChunk < n_tickets)
          cp_async4(sMask, reinterpret_cast<const int*>(p.chunk_mask) + next_chunk / 4);
      }
      const int pos = chunk_base + chunk_off;
      if (pos >= n_tickets) return n_groups;
      if (p.sched == kSchedSkipFlagged && ((skip_mask >> chunk_off) & 1u)) continue;
      return p.sched == kSchedList ? __ldg(p.group_list + pos) : pos + p.group_begin;
    }
  };
  int g = next_group();
  if (g >= n_groups) return;
  int g_next = next_group();
  int g_nn   = g_next < n_groups ? next_group() : n_groups;
  // Connectivity travels three groups deep: slot k % 3 holds the node ids of the k-th group this warp works on.
  // The copy for group k+2 is issued at the top of pass k and read at the top of pass k+1 (to address the gather
  // of X and u) -- a whole pass later, so the DRAM latency of the streamed connectivity never shows
  // (profiles/r01o: a load issued at the end of a pass and consumed at the top of the next one stalled the warp
  // for 5 % of its time).  No node id is carried in a register across the pass.
  int* const    sN = reinterpret_cast<int*>(share + kShareDoubles) + lane;   // [3][32]: this lane's column
  double* const sB = share + kShareDoubles + kConnSlotDoubles + lane;        // [9][32] staged b^-1: this lane's column
  {
    const int n0 = group_node(p, g, ew, q);
    sN[0]        = n0;
    sN[32]       = g_next < n_groups ? group_node(p, g_next, ew, q) : -1;
    stage_gather(p, wsm, n0, q, ew);
    cp_async_commit();
  }
  if (BinvStaged<MAT, MODE>::value) {
#pragma unroll
    for (int i = 0; i < 9; ++i) cp_async8(sB + i * 32, p.binv_cache + (int64_t)g * kBinvGroupDoubles + lane + i * 32);
    cp_async_commit();
  }
  int stage = 0, slot = 0;

  while (g < n_groups) {
    // everything issued during the previous pass has had a pass to land: this group's X / u / b^-1 and the next
    // group's connectivity
    cp_async_wait<0>();
    __syncwarp();
    const bool has_next  = g_next < n_groups;
    const int  slot_next = slot == 2 ? 0 : slot + 1, slot_nn = slot == 0 ? 2 : slot - 1;
    // element data (F / sigma / state, ORDERED forces) live in FILE order: looked up where it is used, so that the
    // index does not occupy registers through the pass
    auto file_element = [&]() -> int64_t {
      const int64_t e_sched = (int64_t)g * kElemsPerWarp + ew;
      return p.orig ? (int64_t)__ldg(p.orig + e_sched) : e_sched;
    };
    constexpr int kRecord = 15 + MaterialState<MAT>::n;  // doubles per integration point (src/nimble_block.cc:84-108)
    const double* rec_n = nullptr;
    if (MaterialState<MAT>::n > 0) {
      // previous records of the warp's four elements -> share buffer (free until the shares are stored): the 8 lanes of
      // an element copy its 8 x kRecord contiguous doubles row by row, so every request is a full 64-byte run
      if (sN[slot * 32] >= 0) {
        const double* src = p.ipt_n + file_element() * (8 * kRecord);
        rec_n             = src + q * kRecord;
#pragma unroll
        for (int i = 0; i < kRecord; ++i) cp_async8(share + ew * (8 * kRecord) + i * 8 + q, src + i * 8 + q);
      }
      cp_async_commit();
    }
    if (has_next) stage_gather(p, wsm + (stage ^ 1) * kStageDoubles, sN[slot_next * 32], q, ew);
    stage_group_node(p, sN + slot_nn * 32, g_nn, n_groups, ew, q);
    cp_async_commit();
#ifdef NSM_BINV_PREFETCH_L1  // A/B: instances that load the cached b^-1 directly (elastic) pull the NEXT group's rows into L1
    if ((MODE & kModeReadBinv) && !BinvStaged<MAT, MODE>::value && has_next && lane < (kBinvGroupDoubles * 8) / 128)
      prefetch_l1(p.binv_cache + (int64_t)g_next * kBinvGroupDoubles + lane * 16);
#endif
#ifdef NSM_BINV_PREFETCH  // measured (r01y): the L2 prefetch a pass ahead of the copy costs 0.5 % instead of helping
    if ((MODE & kModeReadBinv) && g_nn < n_groups && lane < (kBinvGroupDoubles * 8) / 128)
      prefetch_l2(p.binv_cache + (int64_t)g_nn * kBinvGroupDoubles + lane * 16);  // DRAM -> L2 a full pass before the cp.async
#endif

    const double* sX   = wsm + stage * kStageDoubles;
    const double* sU   = sX + kCoordDoubles;
    const double* binv_row  = (MODE & kModeReadBinv) ? p.binv_cache + (int64_t)g * kBinvGroupDoubles + lane : nullptr;
    const double* binv_next = ((MODE & kModeReadBinv) && has_next) ? p.binv_cache + (int64_t)g_next * kBinvGroupDoubles + lane : nullptr;

    // current coordinates: the block functor forms cur = ref + disp (src/nimble_block.cc:309-316); the
    // serial F wrapper then passes disp' = cur - ref and the kernel re-adds it (src/nimble_element.cc:341-344,
    // src/nimble_element.h:455-457); the force wrapper uses cur itself (src/nimble_element.cc:462-463).
    const int    own = coord_row(q) + ew;  // this lane's node in a coordinate set
    const double X0 = sX[0 * kCoordComp + own], X1 = sX[1 * kCoordComp + own], X2 = sX[2 * kCoordComp + own];
    const double u0 = sU[0 * kCoordComp + own], u1 = sU[1 * kCoordComp + own], u2 = sU[2 * kCoordComp + own];
    const double c0 = X0 + u0, c1 = X1 + u1, c2 = X2 + u2;
    const double k0 = X0 + (c0 - X0), k1 = X1 + (c1 - X1), k2 = X2 + (c2 - X2);
    const bool   differs = (k0 != c0) || (k1 != c1) || (k2 != c2);
    sK[0 * kCoordComp + own] = k0;
    sK[1 * kCoordComp + own] = k1;
    sK[2 * kCoordComp + own] = k2;
#ifndef NSM_LAZY_SC  // A/B (scripts/build_variants.sh): the force-path coordinates are read by the cold path only
    sC[0 * kCoordComp + own] = c0;
    sC[1 * kCoordComp + own] = c1;
    sC[2 * kCoordComp + own] = c2;
#endif
    // The F-path and force-path Jacobians coincide unless ref + ((ref+d) - ref) != ref + d for some node of
    // the warp's elements (possible only when |d| is comparable to |ref|); decided warp-uniformly.
    const bool jacobians_differ = __any_sync(0xffffffffu, differs);
    __syncwarp();

    double   F[9], sig[6], state[kMaxStateVars];
    // (the fast pass always runs: it also refills the b^-1 staging slots and closes their copy group)
    unsigned st = integration_point<MAT, MODE, true>(sh, sX, sK, sC, ew, lane, binv_row, sB, binv_next, p.bulk, p.shear,
                                                     share, F, sig, rec_n, p.mat_a, p.mat_b, state, share + lane * kRecord);
    if (jacobians_differ) st |= 1u;
#ifdef NSM_LAZY_SC
    if (__any_sync(0xffffffffu, (st & 1u) != 0u)) {  // some lane goes cold: every lane files its node's coordinates
      sC[0 * kCoordComp + own] = c0;
      sC[1 * kCoordComp + own] = c1;
      sC[2 * kCoordComp + own] = c2;
      __syncwarp();
    }
#endif
    if (st & 1u) {  // cold: some operand outside the fast window (or distinct Jacobians): plain IEEE operators
      atomicAdd(p.flags + 1, 1);  // statistics: integration points redone (nsm_b200_cold_points)
      st = integration_point<MAT, MODE, false>(sh, sX, sK, sC, ew, lane, binv_row, sB, binv_next, p.bulk, p.shear, share,
                                               F, sig, rec_n, p.mat_a, p.mat_b, state);
    }
    const int  node = sN[slot * 32];
    const bool live = node >= 0;
    if (live && (st & 2u)) atomicOr(p.flags, 1);

    if ((MODE & kModeStoreIpt) && MaterialState<MAT>::n == 0 && live) {  // (records of a material with state: below)
      double* d = p.ipt + (file_element() * 8 + q) * kRecord;
#pragma unroll
      for (int i = 0; i < 9; ++i) d[i] = F[i];
#pragma unroll
      for (int i = 0; i < 6; ++i) d[9 + i] = sig[i];
#pragma unroll
      for (int i = 0; i < MaterialState<MAT>::n; ++i) d[15 + i] = state[i];
    }
    __syncwarp();

    // ---- lane n: node n, Gauss points in ascending order: force -= share (src/nimble_element.h:609-611)
    double         fx = 0.0, fy = 0.0, fz = 0.0;
    const double2* col = reinterpret_cast<const double2*>(share + (q * 3) * kShareStride + ew * 8);
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) {
      const double2 sx = col[0 * (kShareStride / 2) + gq], sy = col[1 * (kShareStride / 2) + gq], sz = col[2 * (kShareStride / 2) + gq];
      fx -= sx.x, fy -= sy.x, fz -= sz.x;
      fx -= sx.y, fy -= sy.y, fz -= sz.y;
    }
    if (live) {
      if (ORDERED) {
        double* o = p.ef + (file_element() * 8 + q) * kEfStride;
        o[0] = fx, o[1] = fy, o[2] = fz;
      } else {
        atomicAdd(p.f[0] + node, fx);
        atomicAdd(p.f[1] + node, fy);
        atomicAdd(p.f[2] + node, fz);
      }
    }
    if (MaterialState<MAT>::n > 0) {
      // the new records leave the way the old ones came: through the share buffer (the shares have been summed), as
      // 64-byte runs per element row instead of 8-byte stores 136 bytes apart
      __syncwarp();
      double* rs = share + lane * kRecord;
#pragma unroll
      for (int i = 0; i < 9; ++i) rs[i] = F[i];
#pragma unroll
      for (int i = 0; i < 6; ++i) rs[9 + i] = sig[i];
#pragma unroll
      for (int i = 0; i < MaterialState<MAT>::n; ++i) rs[15 + i] = state[i];
      __syncwarp();
      if (live) {
        double* dst = p.ipt + file_element() * (8 * kRecord);
#pragma unroll
        for (int i = 0; i < kRecord; ++i) dst[i * 8 + q] = share[ew * (8 * kRecord) + i * 8 + q];
      }
    }
    // roll the pipeline
    g      = g_next;
    g_next = g_nn;
    g_nn   = g_nn < n_groups ? next_group() : n_groups;
    slot   = slot_next;
    stage ^= 1;
  }
}

// ORDERED assembly: nodal force = sum of the node's element contributions in ascending (block, element) order
// (the serial reference's order).  Lists of up to 8 entries (every node of a hex mesh but the odd irregular one)
// load all slot ids first and then all 24 values, so that the loads overlap instead of forming a chain of 2 x 8
// dependent round trips; the additions keep the list order.
__device__ __forceinline__ void
ordered_node_sum(const double* __restrict__ ef, const int64_t* __restrict__ adj_off, const uint32_t* __restrict__ adj_slot,
                 int64_t node, double& f0, double& f1, double& f2)
{
  const int64_t b = adj_off[node], e = adj_off[node + 1];
  f0 = 0.0, f1 = 0.0, f2 = 0.0;
  if (e - b <= 8) {
    const int n = (int)(e - b);
    uint32_t  slot[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) slot[k] = k < n ? adj_slot[b + k] : 0u;
    double v[8][3];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double* s = ef + (int64_t)slot[k] * kEfStride;
      if (k < n) v[k][0] = s[0], v[k][1] = s[1], v[k][2] = s[2];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < n) f0 += v[k][0], f1 += v[k][1], f2 += v[k][2];
  } else {
    for (int64_t k = b; k < e; ++k) {
      const double* s = ef + (int64_t)adj_slot[k] * kEfStride;
      f0 += s[0], f1 += s[1], f2 += s[2];
    }
  }
}

// Boundary-first schedule: flag the groups that touch a node shared with another rank and list them.
__global__ void __launch_bounds__(256)
mark_shared_nodes_kernel(int64_t n_shared, const int* __restrict__ shared_node, unsigned char* node_flag)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_shared) node_flag[shared_node[i]] = 1;
}

__global__ void __launch_bounds__(256)
flag_groups_kernel(int64_t n_elem, const int* __restrict__ conn, const unsigned char* __restrict__ node_flag,
                   unsigned* group_bits, int* group_list, unsigned* n_list)
{
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (n_elem + kElemsPerWarp - 1) / kElemsPerWarp) return;
  bool touched = false;
  for (int k = 0; k < kElemsPerWarp * 8; ++k) {
    const int64_t s = g * kElemsPerWarp * 8 + k;
    if (s < n_elem * 8 && node_flag[conn[s]]) touched = true;
  }
  if (touched) {  // bit g of a little-endian bit set: byte c holds the flags of chunk c (kTicketChunk == 8)
    atomicOr(group_bits + (g >> 5), 1u << (g & 31));
    group_list[atomicAdd(n_list, 1u)] = (int)g;
  }
}

// ORDERED assembly with a peer exchange: nodal sums of the shared nodes only (the boundary groups are done)
__global__ void __launch_bounds__(256)
gather_shared_nodes_kernel(int64_t n_shared, const int* __restrict__ shared_node, const double* __restrict__ ef,
                           const int64_t* __restrict__ adj_off, const uint32_t* __restrict__ adj_slot, double* f0, double* f1,
                           double* f2)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_shared) return;
  const int nd = shared_node[i];
  double    a, b, c;
  ordered_node_sum(ef, adj_off, adj_slot, nd, a, b, c);
  f0[nd] = a, f1[nd] = b, f2[nd] = c;
}

// Inverse reference Jacobians b^-1 of every Gauss point, computed once (NSM_FLAG_CACHE_REF_JACOBIAN):
// the same arithmetic as the in-kernel path, so the cached values are the bits the step would recompute.
// Layout [group][9][32 lanes]: every load of the element kernel is one fully coalesced 256-byte row.
__global__ void __launch_bounds__(kElemThreads)
binv_cache_kernel(const ElemArgs p)
{
  const int64_t t    = (int64_t)blockIdx.x * kElemThreads + threadIdx.x;
  const int64_t g    = t >> 5;
  const int     lane = (int)(t & 31);
  const int     q = lane & 7, ew = lane >> 3;
  const int64_t e = g * kElemsPerWarp + ew;
  if (g >= (p.n_elem + kElemsPerWarp - 1) / kElemsPerWarp) return;
  ShapeAtPoint sh;
  sh.init(q);
  double b[3][3], binv[3][3], x[8][3];
  zero33(b);
  for (int j = 0; j < 8; ++j) {
    if (e < p.n_elem) {
      const int nd = p.conn[e * 8 + j];
      x[j][0] = p.X[0][nd], x[j][1] = p.X[1][nd], x[j][2] = p.X[2][nd];
    } else {
      x[j][0] = ((j & 3) == 1 || (j & 3) == 2) ? 1.0 : 0.0, x[j][1] = (j & 2) ? 1.0 : 0.0, x[j][2] = (j & 4) ? 1.0 : 0.0;
    }
  }
  grad_accumulate<0>(sh, x[0][0], x[0][1], x[0][2], b);
  grad_accumulate<1>(sh, x[1][0], x[1][1], x[1][2], b);
  grad_accumulate<2>(sh, x[2][0], x[2][1], x[2][2], b);
  grad_accumulate<3>(sh, x[3][0], x[3][1], x[3][2], b);
  grad_accumulate<4>(sh, x[4][0], x[4][1], x[4][2], b);
  grad_accumulate<5>(sh, x[5][0], x[5][1], x[5][2], b);
  grad_accumulate<6>(sh, x[6][0], x[6][1], x[6][2], b);
  grad_accumulate<7>(sh, x[7][0], x[7][1], x[7][2], b);
  unsigned     unused = 0u;
  const double detb   = invert3x3<false>(b, binv, unused);
  if (e < p.n_elem && !(detb > 0.0)) atomicOr(p.flags, 1);
  double* bc = p.binv_cache + g * kBinvGroupDoubles + lane;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) bc[(3 * i + k) * 32] = binv[i][k];
}

// ---------------------------------------------------------------------------------------------------
// stress seam: one thread per integration point, F[9] -> sigma[6]
// (BlockMaterialInterface::ComputeStress, src/nimble_kokkos_block_material_interface.cc:65-119)
// ---------------------------------------------------------------------------------------------------
template <int MAT>
__global__ void __launch_bounds__(128)
stress_kernel(int64_t n, const double* __restrict__ Fin, double* __restrict__ sout, double bulk, double shear)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double F[9], s[6];
#pragma unroll
  for (int k = 0; k < 9; ++k) F[k] = Fin[i * 9 + k];
  if (MAT == 0) {
    stress_elastic(bulk, shear, F, s);
  } else {  // the element kernel's pattern: branch-free fast path, IEEE redo when an operand left its window
    unsigned bad = 0u;
    stress_neohookean<true>(bulk, shear, F, s, bad);
    if (bad) stress_neohookean<false>(bulk, shear, F, s, bad);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) sout[i * 6 + k] = s[k];
}

// the same seam for a material with state variables: (F_n, F_np1, sigma_n, state_n) -> (sigma_np1, state_np1)
__global__ void __launch_bounds__(128)
stress_state_kernel(int64_t n, const double* __restrict__ Fn_in, const double* __restrict__ F_in, const double* __restrict__ sn_in,
                    const double* __restrict__ stn_in, double* __restrict__ s_out, double* __restrict__ st_out, double bulk,
                    double shear, double mat_a, double mat_b)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double Fn[9], F[9], sn[6], stn[kMaxStateVars], s[6], st[kMaxStateVars];
#pragma unroll
  for (int k = 0; k < 9; ++k) Fn[k] = Fn_in[i * 9 + k], F[k] = F_in[i * 9 + k];
#pragma unroll
  for (int k = 0; k < 6; ++k) sn[k] = sn_in[i * 6 + k];
#pragma unroll
  for (int k = 0; k < kMaxStateVars; ++k) stn[k] = stn_in[i * kMaxStateVars + k];
  unsigned bad = 0u;
  stress_j2<true>(bulk, shear, mat_a, mat_b, Fn, F, sn, stn, s, st, bad);
  if (bad) stress_j2<false>(bulk, shear, mat_a, mat_b, Fn, F, sn, stn, s, st, bad);
#pragma unroll
  for (int k = 0; k < 6; ++k) s_out[i * 6 + k] = s[k];
#pragma unroll
  for (int k = 0; k < kMaxStateVars; ++k) st_out[i * kMaxStateVars + k] = st[k];
}

// ---------------------------------------------------------------------------------------------------
// node kernels
// ---------------------------------------------------------------------------------------------------
struct NodeArgs
{
  int64_t         n_nodes;     // end of the node range this launch covers
  int64_t         node_begin;  // its start (0 except in the pipelined host step)
  double*         u[3];
  double*         v[3];
  double*         a[3];
  double*         f[3];
  const double*   fext[3];   // nullptr: external force is identically zero (src/nimble_model_data.cc:609-618)
  const double*   fcontact[3];  // nullptr: no contact; else a = (1/m)(f_int + f_ext + f_contact) (explicit_time_integrator.cc:243-249)
  const double*   mass;
  const int*      bc_of_dof[3];  // entry index or -1; nullptr when the deck has no kinematic BC
  const int*      bc_kind;
  const double*   bc_value;
  // ORDERED assembly
  const double*   ef;         // [slots][kEfStride]
  const int64_t*  adj_off;    // [n_nodes+1]
  const uint32_t* adj_slot;   // slot = global element * 8 + local node, ascending per node
};

__device__ __forceinline__ double
bc_velocity(int kind, double value, double u, double dt, double v_old)
{
  // BoundaryConditionManager::ApplyKinematicBC (src/nimble_boundary_condition_manager.h:146-201)
  if (kind == 0) return value;
  if (dt > 0.0) return (value - u) / dt;
  return v_old;
}

// first half of a step (src/integrators/explicit_time_integrator.cc:199-216):
//   v += (dt/2) a ; BC ; u += dt v ; BC ; and the nodal force is cleared for the coming assembly.
template <bool HAS_BC, bool ZERO_F>
__global__ void __launch_bounds__(256)
node_predict_kernel(const NodeArgs p, double hdt, double dt)
{
  const int64_t i = p.node_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_nodes) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double v = p.v[c][i];
    double u = p.u[c][i];
    v        = v + hdt * p.a[c][i];  // Viewify += AXPYResult: prod = alpha*1.0; data += prod*rhs (src/nimble_view.h:194-200)
    int bc   = -1;
    if (HAS_BC) {
      bc = p.bc_of_dof[c][i];
      if (bc >= 0) v = bc_velocity(p.bc_kind[bc], p.bc_value[bc], u, dt, v);
    }
    u = u + dt * v;
    if (HAS_BC) {
      if (bc >= 0) v = bc_velocity(p.bc_kind[bc], p.bc_value[bc], u, dt, v);
    }
    p.v[c][i] = v;
    p.u[c][i] = u;
    if (ZERO_F) p.f[c][i] = 0.0;
  }
}

// second half (src/integrators/explicit_time_integrator.cc:250-262): a = (1/m)(f_int + f_ext); v += (dt/2) a.
// ORDERED: f_int is first summed from the element forces in ascending (block, element) order.
template <bool ORDERED, bool HAS_FEXT>
__global__ void __launch_bounds__(256)
node_correct_kernel(const NodeArgs p, double hdt, int update_velocity)
{
  const int64_t i = p.node_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_nodes) return;
  double f0, f1, f2;
  if (ORDERED) {
    ordered_node_sum(p.ef, p.adj_off, p.adj_slot, i, f0, f1, f2);
    p.f[0][i] = f0, p.f[1][i] = f1, p.f[2][i] = f2;
  } else {
    f0 = p.f[0][i], f1 = p.f[1][i], f2 = p.f[2][i];
  }
  if (!update_velocity) return;
  const double rm = 1.0 / p.mass[i];
  double       s0 = f0 + (HAS_FEXT ? p.fext[0][i] : 0.0), s1 = f1 + (HAS_FEXT ? p.fext[1][i] : 0.0), s2 = f2 + (HAS_FEXT ? p.fext[2][i] : 0.0);
  if (p.fcontact[0]) s0 = s0 + p.fcontact[0][i], s1 = s1 + p.fcontact[1][i], s2 = s2 + p.fcontact[2][i];
  const double a0 = rm * s0;
  const double a1 = rm * s1;
  const double a2 = rm * s2;
  p.a[0][i] = a0, p.a[1][i] = a1, p.a[2][i] = a2;
  p.v[0][i] = p.v[0][i] + hdt * a0;
  p.v[1][i] = p.v[1][i] + hdt * a1;
  p.v[2][i] = p.v[2][i] + hdt * a2;
}

// Second half of step s fused with the first half of step s+1 (one pass over the nodes per interior step of a
// multi-step call): a = (1/m)(f_int + f_ext); v += (dt_s/2) a | v += (dt_{s+1}/2) a; BC; u += dt_{s+1} v; BC.
// Per node this is the operation sequence of node_correct_kernel followed by node_predict_kernel, so the
// results are bit-identical; a and the integer-time velocity are not written (nothing reads them before
// the next pass overwrites them), and the force field is cleared for the coming atomic assembly.
template <bool ORDERED, bool HAS_FEXT, bool HAS_BC, bool ZERO_F>
__global__ void __launch_bounds__(256)
node_fused_kernel(const NodeArgs p, double hdt, double hdt_next, double dt_next)
{
  const int64_t i = p.node_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_nodes) return;
  double f[3];
  if (ORDERED) {
    ordered_node_sum(p.ef, p.adj_off, p.adj_slot, i, f[0], f[1], f[2]);
  } else {
    f[0] = p.f[0][i], f[1] = p.f[1][i], f[2] = p.f[2][i];
  }
  const double rm = 1.0 / p.mass[i];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double sum = f[c] + (HAS_FEXT ? p.fext[c][i] : 0.0);
    if (p.fcontact[0]) sum = sum + p.fcontact[c][i];
    const double a = rm * sum;
    double       v = p.v[c][i];
    double       u = p.u[c][i];
    v              = v + hdt * a;
    v              = v + hdt_next * a;
    int bc         = -1;
    if (HAS_BC) {
      bc = p.bc_of_dof[c][i];
      if (bc >= 0) v = bc_velocity(p.bc_kind[bc], p.bc_value[bc], u, dt_next, v);
    }
    u = u + dt_next * v;
    if (HAS_BC) {
      if (bc >= 0) v = bc_velocity(p.bc_kind[bc], p.bc_value[bc], u, dt_next, v);
    }
    p.v[c][i] = v;
    p.u[c][i] = u;
    if (ZERO_F) p.f[c][i] = 0.0;
  }
}

// Boundary-condition programs (include/nsm_b200.h, nsm_bc_op): one thread per BC table entry evaluates its
// program at the entry's node for one step and writes the magnitude the node kernels read.  Every operation is
// IEEE-exact (no FMA contraction in this translation unit), so the value equals the host evaluation of the same
// expression tree bit for bit; sub-expressions of t alone arrive as host-evaluated slots.
struct BcProgramArgs
{
  int64_t       n_entries;
  const int*    program_of_entry;  // -1: keep the host magnitude
  const int*    node_of_entry;
  const int*    offsets;           // [n_programs + 1]
  const int*    code;
  const double* consts;
  const double* slots;             // this step's row
  const double* entry_consts;      // [n_entry_consts][n_entries]: host-evaluated functions of the entry's position
  const double* X[3];
  double*       value;             // [n_entries]
};

__global__ void __launch_bounds__(128)
bc_program_kernel(const BcProgramArgs p)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n_entries) return;
  const int prog = p.program_of_entry[k];
  if (prog < 0) return;
  const int nd = p.node_of_entry[k];
  double    st[16];
  int       sp = 0;
  for (int pc = p.offsets[prog]; pc < p.offsets[prog + 1]; ++pc) {
    const int word = p.code[pc], op = word & 0xff, arg = word >> 8;
    if (op <= 4 || op == 26) {  // pushes
      double v;
      switch (op) {
        case 0: v = p.consts[arg]; break;
        case 1: v = p.X[0][nd]; break;
        case 2: v = p.X[1][nd]; break;
        case 3: v = p.X[2][nd]; break;
        case 4: v = p.slots[arg]; break;
        default: v = p.entry_consts[(int64_t)arg * p.n_entries + k]; break;
      }
      st[sp++] = v;
    } else if (op == 10 || (op >= 11 && op <= 15) || op == 24) {  // unary
      const double a = st[sp - 1];
      double       r;
      switch (op) {
        case 10: r = -a; break;
        case 11: r = sqrt(a); break;
        case 12: r = fabs(a); break;
        case 13: r = floor(a); break;
        case 14: r = ceil(a); break;
        case 15: r = round(a); break;
        default: r = a != 0.0 ? 0.0 : 1.0; break;
      }
      st[sp - 1] = r;
    } else if (op == 25) {  // select
      const double c = st[sp - 1], b = st[sp - 2], a = st[sp - 3];
      sp -= 2;
      st[sp - 1] = a != 0.0 ? b : c;
    } else {  // binary
      const double b = st[sp - 1], a = st[sp - 2];
      double       r;
      --sp;
      switch (op) {
        case 5: r = a + b; break;
        case 6: r = a - b; break;
        case 7: r = a * b; break;
        case 8: r = a / b; break;
        case 9: r = fmod(a, b); break;
        case 16: r = a < b ? 1.0 : 0.0; break;
        case 17: r = a <= b ? 1.0 : 0.0; break;
        case 18: r = a > b ? 1.0 : 0.0; break;
        case 19: r = a >= b ? 1.0 : 0.0; break;
        case 20: r = a == b ? 1.0 : 0.0; break;
        case 21: r = ((a != 0.0) & (b != 0.0)) ? 1.0 : 0.0; break;
        case 22: r = ((a != 0.0) | (b != 0.0)) ? 1.0 : 0.0; break;
        default: r = ((a != 0.0) ^ (b != 0.0)) ? 1.0 : 0.0; break;
      }
      st[sp - 1] = r;
    }
  }
  p.value[k] = st[0];
}

// BoundaryConditionManager::ApplyKinematicBC alone (output steps, t = 0): table entries in deck order;
// duplicates of a dof were resolved to the last entry when the dof map was built.
__global__ void __launch_bounds__(256)
apply_bc_kernel(const NodeArgs p, double dt)
{
  const int64_t i = p.node_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_nodes) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int bc = p.bc_of_dof[c][i];
    if (bc >= 0) p.v[c][i] = bc_velocity(p.bc_kind[bc], p.bc_value[bc], p.u[c][i], dt, p.v[c][i]);
  }
}

// host AoS [n][3] staging (the caller's node order) <-> device SoA (internal node order: perm[caller id], or the
// same when perm == nullptr)
__global__ void __launch_bounds__(256)
aos_to_soa_kernel(int64_t n, const double* __restrict__ aos, double* x, double* y, double* z, const int* __restrict__ perm)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t j = perm ? perm[i] : i;
  x[j] = aos[3 * i], y[j] = aos[3 * i + 1], z[j] = aos[3 * i + 2];
}

__global__ void __launch_bounds__(256)
soa_to_aos_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                  double* aos, const int* __restrict__ perm)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t j = perm ? perm[i] : i;
  aos[3 * i] = x[j], aos[3 * i + 1] = y[j], aos[3 * i + 2] = z[j];
}

// the same for a node range [i0, i0 + n) without a node permutation: staging rows are the range's own (pipelined host step)
__global__ void __launch_bounds__(256)
aos_to_soa_range_kernel(int64_t i0, int64_t n, const double* __restrict__ aos, double* x, double* y, double* z)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  x[i0 + k] = aos[3 * (i0 + k)], y[i0 + k] = aos[3 * (i0 + k) + 1], z[i0 + k] = aos[3 * (i0 + k) + 2];
}

__global__ void __launch_bounds__(256)
soa_to_aos_range_kernel(int64_t i0, int64_t n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                        double* aos)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  aos[3 * (i0 + k)] = x[i0 + k], aos[3 * (i0 + k) + 1] = y[i0 + k], aos[3 * (i0 + k) + 2] = z[i0 + k];
}

// lowest and highest node id of every 4-element group (dependency ranges of the pipelined host step)
__global__ void __launch_bounds__(256)
group_node_range_kernel(int64_t n_elem, const int* __restrict__ conn, int* lo, int* hi)
{
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (n_elem + kElemsPerWarp - 1) / kElemsPerWarp) return;
  int a = 0x7fffffff, b = -1;
  for (int k = 0; k < kElemsPerWarp * 8; ++k) {
    const int64_t s = g * kElemsPerWarp * 8 + k;
    if (s < n_elem * 8) {
      const int nd = conn[s];
      a = nd < a ? nd : a, b = nd > b ? nd : b;
    }
  }
  lo[g] = a, hi[g] = b;
}

// scalar nodal field between the caller's order (staging) and the internal order
__global__ void __launch_bounds__(256)
permute_scalar_kernel(int64_t n, const double* __restrict__ in, double* __restrict__ out, const int* __restrict__ perm, int to_internal)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (to_internal)
    out[perm[i]] = in[i];
  else
    out[i] = in[perm[i]];
}

// ---------------------------------------------------------------------------------------------------
// setup / output kernels: one thread per element, plain loops over a constant shape table.
// ---------------------------------------------------------------------------------------------------
struct ShapeTables
{
  double N[64];    // [q][j]
  double dN[192];  // [q][j][3]
};
__constant__ ShapeTables c_shape;

__device__ __forceinline__ void
param_gradient_tbl(const double (&x)[24], int q, double (&J)[3][3])
{
  zero33(J);
  for (int j = 0; j < 8; ++j) {
    const double* d = &c_shape.dN[24 * q + 3 * j];
    for (int i = 0; i < 3; ++i) {
      J[i][0] += x[3 * j + i] * d[0];
      J[i][1] += x[3 * j + i] * d[1];
      J[i][2] += x[3 * j + i] * d[2];
    }
  }
}

struct SetupArgs
{
  int64_t         n_elem;
  const int*      conn;
  const double*   X[3];
  const double*   u[3];
  double          density, bulk;
  double*         mass;        // ATOMIC: nodal lumped mass (atomic add)
  double*         em;          // ORDERED: [n_elem][8] element lumped masses
  unsigned long long* min_dt_bits;
  int*            flags;
};

// HexElement::ComputeLumpedMass (src/nimble_element.cc:173-193 / src/nimble_element.h:266-309) and
// HexElement::ComputeCharacteristicLength (src/nimble_element.cc:221-260) + BlockBase::ComputeCriticalTimeStep
// (src/nimble_block_base.cc:51-84).
template <bool ORDERED>
__global__ void __launch_bounds__(128)
lumped_mass_kernel(const SetupArgs p)
{
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.n_elem) return;
  int    nd[8];
  double X[24], x[24];
  for (int j = 0; j < 8; ++j) {
    nd[j] = p.conn[e * 8 + j];
    for (int i = 0; i < 3; ++i) {
      X[3 * j + i] = p.X[i][nd[j]];
      x[3 * j + i] = X[3 * j + i] + p.u[i][nd[j]];
    }
  }
  double det[8];
  for (int q = 0; q < 8; ++q) {
    double a[3][3], ai[3][3];
    param_gradient_tbl(X, q, a);
    unsigned unused = 0u;
    det[q]          = invert3x3<false>(a, ai, unused);
    if (!(det[q] > 0.0)) atomicOr(p.flags, 1);
  }
  for (int i = 0; i < 8; ++i) {
    double mi = 0.0;
    for (int j = 0; j < 8; ++j) {
      double mij = 0.0;
      for (int q = 0; q < 8; ++q) mij += 1.0 * p.density * c_shape.N[8 * q + i] * c_shape.N[8 * q + j] * det[q];
      mi += mij;
    }
    if (ORDERED)
      p.em[e * 8 + i] = mi;
    else
      atomicAdd(p.mass + nd[i], mi);
  }
  // characteristic length on the current configuration, including the reference's quirk that the box
  // maxima start at 0.0 (src/nimble_element.cc:230)
  double xmax = 0.0, ymax = 0.0, zmax = 0.0;
  double xmin = 1.7976931348623157e308, ymin = xmin, zmin = xmin, dmin2 = xmin;
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
  const double dt = len / sqrt(p.bulk / p.density);
  // non-negative doubles order like their bit patterns
  if (dt >= 0.0) atomicMin(p.min_dt_bits, (unsigned long long)__double_as_longlong(dt));
}

// ordered nodal sum of element scalars (lumped mass), ascending (block, element)
__global__ void __launch_bounds__(256)
node_gather_scalar_kernel(int64_t n_nodes, const double* __restrict__ es, const int64_t* __restrict__ adj_off,
                          const uint32_t* __restrict__ adj_slot, double* out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  double s = 0.0;
  for (int64_t k = adj_off[i]; k < adj_off[i + 1]; ++k) s += es[adj_slot[k]];
  out[i] = s;
}

// Block::ComputeDerivedElementData (src/nimble_block.cc:438-497) -> HexElement::ComputeVolumeAverage
// (src/nimble_element.cc:262-282, src/nimble_element.h:343-392): volume = sum detJ (unit weights) on the
// current configuration; averages of the 15 integration-point fields.  out [16][n_elem].
__global__ void __launch_bounds__(128)
derived_kernel(int64_t n_elem, const int* __restrict__ conn, const double* X0, const double* X1, const double* X2,
               const double* u0, const double* u1, const double* u2, const double* __restrict__ ipt, double* out, int record)
{
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  const double* Xs[3] = {X0, X1, X2};
  const double* us[3] = {u0, u1, u2};
  double        x[24];
  for (int j = 0; j < 8; ++j) {
    const int nd = conn[e * 8 + j];
    for (int i = 0; i < 3; ++i) x[3 * j + i] = (Xs[i][nd] + us[i][nd]) + 0.0;
  }
  double vol = 0.0, avg[15 + kMaxStateVars];  // record = 15 + the material's state variables
  for (int i = 0; i < 15 + kMaxStateVars; ++i) avg[i] = 0.0;
  const double* qd = ipt + e * 8 * record;
  for (int g = 0; g < 8; ++g) {
    double a[3][3], ai[3][3];
    param_gradient_tbl(x, g, a);
    unsigned     unused = 0u;
    const double det    = invert3x3<false>(a, ai, unused);
    vol += det;
    for (int i = 0; i < 15 + kMaxStateVars; ++i)
      if (i < record) avg[i] += qd[g * record + i] * 1.0 * det;
  }
  out[e] = vol;
  for (int i = 0; i < record; ++i) out[(int64_t)(i + 1) * n_elem + e] = avg[i] / vol;
}

// Output: component split of the integration-point data, out[k][e - e0] = ipt[e][off[k]] for a range of elements
__global__ void __launch_bounds__(256)
select_ipt_components_kernel(int64_t e0, int64_t n, int n_comp, const int* __restrict__ off, const double* __restrict__ ipt,
                             double* __restrict__ out, int per_element)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * n_comp) return;
  const int64_t k = i / n, e = i - k * n;
  out[i]          = ipt[(e0 + e) * per_element + off[k]];
}

// Output / checks: full integration-point records of a list of elements, out[i] = ipt[elems[i]]
__global__ void __launch_bounds__(256)
gather_ipt_records_kernel(int64_t n, const int64_t* __restrict__ elems, const double* __restrict__ ipt, double* __restrict__ out,
                          int per_element)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * per_element) return;
  const int64_t k = i / per_element;
  out[i]          = ipt[elems[k] * per_element + (i - k * per_element)];
}

// F = identity, sigma = 0 (Block::InitializeElementData, src/nimble_block.cc:148-207)
// (state variables start from the material's initial values, 0 for the built-in history-dependent material)
__global__ void __launch_bounds__(256)
init_ipt_kernel(int64_t n_points, double* ipt, int record)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  double* d = ipt + i * record;
  for (int k = 0; k < record; ++k) d[k] = (k < 3) ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------------------------------
// node -> element adjacency (ORDERED assembly), built on the device
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adj_count_kernel(int64_t n_slots, const int* __restrict__ conn, unsigned long long* counts)
{
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  atomicAdd(counts + conn[s], 1ULL);
}

__global__ void __launch_bounds__(256)
adj_fill_kernel(int64_t n_slots, int64_t slot_base, const int* __restrict__ conn, const int64_t* __restrict__ off,
                unsigned long long* cursor, uint32_t* slots)
{
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  const int                n = conn[s];
  const unsigned long long k = atomicAdd(cursor + n, 1ULL);
  slots[off[n] + (int64_t)k] = (uint32_t)(slot_base + s);
}

__global__ void __launch_bounds__(256)
adj_sort_kernel(int64_t n_nodes, const int64_t* __restrict__ off, uint32_t* slots)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const int64_t b = off[i], e = off[i + 1];
  for (int64_t k = b + 1; k < e; ++k) {  // insertion sort; lists hold ~8 entries
    const uint32_t key = slots[k];
    int64_t        j   = k - 1;
    while (j >= b && slots[j] > key) {
      slots[j + 1] = slots[j];
      --j;
    }
    slots[j + 1] = key;
  }
}

// ---------------------------------------------------------------------------------------------------
// FP64 pipe micro-benchmark (roofline denominator for an FP64-issue-bound kernel)
// ---------------------------------------------------------------------------------------------------
template <bool FUSED>
__global__ void __launch_bounds__(256)
fp64_peak_kernel(double* out, int iters, double seed)
{
  double x0 = seed + threadIdx.x, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0;
  double x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
  const double m = 0.999999, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    if (FUSED) {
      x0 = fma(x0, m, c), x1 = fma(x1, m, c), x2 = fma(x2, m, c), x3 = fma(x3, m, c);
      x4 = fma(x4, m, c), x5 = fma(x5, m, c), x6 = fma(x6, m, c), x7 = fma(x7, m, c);
    } else {  // -fmad=false keeps these as DMUL + DADD
      x0 = x0 * m, x1 = x1 + c, x2 = x2 * m, x3 = x3 + c;
      x4 = x4 * m, x5 = x5 + c, x6 = x6 * m, x7 = x7 + c;
    }
  }
  out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace nsm
