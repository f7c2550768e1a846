// nimblesm_b200/csrc/nsm_b200.cu — context and C ABI (include/nsm_b200.h) of the B200 hex8 path.
//
// One context owns one GPU's share of the model: SoA fp64 nodal fields, per-block connectivity and
// material, optional integration-point storage, assembly tables, a CUDA stream and the peer-exchange
// state.  Host code (nimblesm_b200/host, C++; nimblesm_b200/capi.py, ctypes) reaches the kernels only
// through the extern "C" functions at the bottom.  No CPU fallback exists anywhere in this file.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <nvtx3/nvToolsExt.h>  // header-only; a no-op unless a profiler is attached

#include "../../include/nsm_b200.h"
#include "contact.cuh"
#include "hex8_kernels.cuh"
#include "peer_exchange.cuh"

using namespace nsm;

namespace {

struct Block
{
  int              id       = 0;
  int64_t          n_elem   = 0;
  int              material = 0;
  double           bulk = 0, shear = 0, density = 0;
  double           mat_a = 0, mat_b = 0;  // material-specific parameters (j2_plasticity: yield stress, hardening modulus)
  // history-dependent material: integration-point records N and N+1, [n_elem][8][15 + n_state] each, owned by the
  // block; rec[cur] is the one the element kernel writes.  roll_pending: ModelData::UpdateStates was requested, the
  // roles swap at the next force evaluation (src/nimble_model_data.h:104-107)
  int              n_state = 0;
  double*          rec[2]  = {nullptr, nullptr};
  int              cur     = 0;
  bool             roll_pending = false;
  std::vector<int> conn_host;  // dropped after finalize
  int*             conn      = nullptr;  // file order (mass, adjacency, derived data)
  int*             conn_sched = nullptr; // the element kernel's walk: == conn, or Morton-ordered (NSM_FLAG_REORDER_ELEMENTS)
  int*             orig      = nullptr;  // schedule position -> file-order element (nullptr: identity)
  int64_t          elem_base = 0;  // first global element (ascending block id order)
  int64_t          group_base = 0; // first 4-element group of the block in the b^-1 cache
  // boundary-first schedule (peer exchange attached): groups touching a node shared with another rank
  unsigned*        group_bits = nullptr;  // bit g: group g touches a node shared with another rank
  int*             group_list = nullptr;
  int              n_list     = 0;
};

thread_local std::string g_create_error;

// NVTX range over the host code that enqueues one phase of the step, named after the reference's own timer regions
// (src/integrators/explicit_time_integrator.cc:197-274: "Time Integration Scheme", "BC enforcement", "Force
// calculation", "Output"; the shared-node sum is timed as vector reduction, src/nimble_vector_communicator.h:144-157).
// A profiler's CUDA-launch correlation ties the kernels to the range they were launched from.
struct NvtxRange
{
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&)            = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

}  // namespace

struct nsm_b200_ctx
{
  int          device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t  ev_start = nullptr, ev_stop = nullptr;
  std::string  err;

  int64_t n_nodes = 0;
  bool    finalized = false;
  int     assembly  = NSM_ASSEMBLY_ATOMIC;
  unsigned flags_   = 0;

  std::vector<double> hx, hy, hz;  // host copies until finalize
  // NSM_FLAG_RENUMBER_NODES: caller's node id -> internal id (Morton order of the coordinates); empty = identity
  std::vector<int> node_perm_host;
  int*             node_perm = nullptr;
  std::map<int, Block> blocks;     // ascending id == reference processing order
  int64_t              n_elem_total = 0;

  // nodal fields, SoA
  double* X[3]    = {nullptr, nullptr, nullptr};
  double* u[3]    = {nullptr, nullptr, nullptr};
  double* v[3]    = {nullptr, nullptr, nullptr};
  double* a[3]    = {nullptr, nullptr, nullptr};
  double* f[3]    = {nullptr, nullptr, nullptr};
  double* fext[3] = {nullptr, nullptr, nullptr};
  double* mass    = nullptr;
  double* staging = nullptr;  // [n][3] AoS bounce buffer for host views
  // nsm_b200_step_host: the displacement is final after the first half of the step, so it travels back to the host
  // on a second stream (own bounce buffer) while the element kernels run
  double*      staging_u      = nullptr;
  cudaStream_t io_stream      = nullptr;
  cudaEvent_t  ev_u_staged    = nullptr;
  double*      early_u_host   = nullptr;  // set for the duration of one nsm_b200_step call
  bool    has_fext = false;

  // element data
  double* ipt  = nullptr;  // [n_elem_total][8][15], lazily allocated
  double* binv = nullptr;  // [n_elem_total][8][9]
  double* ef   = nullptr;  // ORDERED: [n_elem_total][8][kEfStride]
  int64_t*  adj_off  = nullptr;
  uint32_t* adj_slot = nullptr;

  // boundary conditions
  int64_t n_bc = 0;
  int64_t bc_rows = 1, bc_rows_cap = 0;  // per-step magnitude rows held in bc_value ([bc_rows][n_bc])
  int*    bc_of_dof[3] = {nullptr, nullptr, nullptr};
  int*    bc_kind  = nullptr;
  double* bc_value = nullptr;
  int*    bc_node  = nullptr;  // node of every table entry (boundary-condition programs read its coordinates)
  // boundary-condition programs (nsm_b200_set_bc_programs): magnitudes evaluated on the device each step
  int     bcp_programs = 0, bcp_slots = 0, bcp_rows = 0, bcp_rows_cap = 0;
  int*    bcp_offsets = nullptr;
  int*    bcp_code    = nullptr;
  double* bcp_consts  = nullptr;
  int*    bcp_of_entry = nullptr;
  double* bcp_slot_values = nullptr;  // [bcp_rows][bcp_slots]
  double* bcp_entry_consts = nullptr; // [bcp_n_entry_consts][n_bc] (NSM_BCOP_ENTRYCONST)
  int     bcp_n_entry_consts = 0, bcp_entry_consts_needed = 0;

  int*                d_flags  = nullptr;
  unsigned*        d_ticket = nullptr;  // element-kernel work counters [2]: launches alternate, each zeroes the other's
  int              ticket_parity = 0;
  unsigned long long* d_min_dt = nullptr;

  int64_t launches     = 0;
  int64_t device_bytes = 0;
  std::map<const void*, int64_t> alloc_bytes;  // what dev_alloc handed out, so that dev_release keeps device_bytes true

  // per-launch profiling (CUDA events on the stream)
  bool                     profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t                   ev_used = 0;
  double                   prof_elem_ms = 0, prof_node_ms = 0, prof_contact_ms = 0;
  int64_t                  prof_steps   = 0;

  // pipelined nsm_b200_step_host: node chunks travel up, are integrated, their elements run and the finished chunks
  // travel down while later chunks are still on their way up (PCIe is full duplex); see build_host_pipe
  struct HostPipe
  {
    int                  requested_chunks = -1;  // -1: automatic
    // element-kernel CTA slots left to the copy-side kernels of the pipelined STEP (NSM_B200_PIPE_RESERVE).  Measured
    // at 64 M elements (profiles/r02q_*): 0 -> 169 ms per step, 8 -> 157, 16 -> 152, 32 -> 148, 64 -> 147.  The
    // force seam moves a third of the bytes and is bound by the elements, so it reserves nothing (47.2 ms; 51 ms at 32).
    int                  reserve_ctas     = 32;
    int                  reserve_now      = 0;   // what the ranged launches of the call in flight leave free
    bool                 built            = false;
    int                  n_chunks         = 0;
    std::vector<int64_t> node_end;                 // [C] end of node chunk c
    std::map<int, std::vector<int>> up_end;        // block -> [C]: groups [0, up_end[c]) touch only nodes < node_end[c]
    std::vector<int>     done_after;               // [C]: node chunk k is complete once element chunk done_after[k] has run
    double*              stage[4] = {nullptr, nullptr, nullptr, nullptr};  // AoS bounce buffers of u, v, a, f
    cudaStream_t         up = nullptr, down = nullptr;
    std::vector<cudaEvent_t> ev_up, ev_elem;
    cudaEvent_t          ev_bc = nullptr;
  } pipe;

  // penalty contact (nsm_b200_set_contact): entities, per-evaluation scratch, the nodal contact force
  struct Contact
  {
    bool      active  = false;
    double    penalty = 0.0;
    int64_t   n_quads = 0, n_sec = 0;
    int*      quad = nullptr;
    double*   quad_len = nullptr;
    int*      sec_node = nullptr;
    double*   sec_len  = nullptr;
    int64_t   n_surf    = 0;
    int*      surf_node = nullptr;
    double*   quad_xyz = nullptr;
    float*    tri_box  = nullptr;
    QuadBin*  bin  = nullptr;
    int*      head = nullptr;
    unsigned  table_mask = 0;
    unsigned* red = nullptr;                  // [2][8], alternating by evaluation
    unsigned long long* counters = nullptr;   // [8]: enforced, box-tested, active faces, active nodes, near-list length
    int*      near_list = nullptr;
    unsigned char* status = nullptr;
    int       parity = 0;
    // ORDERED assembly: filed contributions and the scratch of their two stable sorts
    long long           contrib_cap = 0;
    unsigned long long *contrib_key = nullptr, *key_sorted = nullptr;
    int*                contrib_target = nullptr;
    double*             contrib_val    = nullptr;
    unsigned *          iota = nullptr, *order1 = nullptr, *order2 = nullptr, *target1 = nullptr, *target2 = nullptr;
    unsigned char*      sort_tmp = nullptr;
    size_t              sort_tmp_bytes = 0;
    long long           ordered_overflow_pairs = 0;  // of the last evaluation
  } contact;
  double* fc[3] = {nullptr, nullptr, nullptr};  // nodal contact force, SoA (allocated by nsm_b200_set_contact)

  // scratch of the stress seam (nsm_b200_compute_stress / _state): grow-only, so that a caller who crosses the seam
  // every step -- as the reference re-creates and calls it every step, src/nimble_kokkos_model_data.cc:1230-1234 --
  // pays two copies and a launch, not an allocation
  double* seam      = nullptr;
  size_t  seam_cap  = 0;  // doubles

  PeerExchange comm;
  // overlap of the shared-node exchange with the interior elements (nsm_b200_step)
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t  ev_boundary = nullptr, ev_packed = nullptr;
  bool         overlap     = false;
};

namespace {

int
fail(nsm_b200_ctx* c, int code, const char* fmt, ...)
{
  char    buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c)
    c->err = buf;
  else
    g_create_error = buf;
  return code;
}

#define NSM_CUDA(c, call)                                                                               \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      return fail((c), NSM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define NSM_REQUIRE(c, cond, msg) \
  do {                            \
    if (!(cond)) return fail((c), NSM_ERR_ARG, "%s", msg); \
  } while (0)

// Every entry point that touches the GPU runs with the context's device current and puts the caller's device back
// on return: a host thread may drive several contexts, or use another device through torch / its own CUDA code.
struct DeviceGuard
{
  int prev = -1, dev = -1;
  explicit DeviceGuard(int device) : dev(device)
  {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard()
  {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&)            = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define NSM_ENTER(c)                                   \
  NSM_REQUIRE((c), (c) != nullptr, "null context");    \
  DeviceGuard device_guard_((c)->device)

template <class T>
int
dev_alloc(nsm_b200_ctx* c, T** p, int64_t count)
{
  *p = nullptr;
  if (count <= 0) count = 1;
  NSM_CUDA(c, cudaMalloc((void**)p, (size_t)count * sizeof(T)));
  c->device_bytes += count * (int64_t)sizeof(T);
  c->alloc_bytes[*p] = count * (int64_t)sizeof(T);
  return NSM_OK;
}

// frees a dev_alloc buffer that is being replaced during the context's life (tables, per-step rows)
template <class T>
void
dev_release(nsm_b200_ctx* c, T*& p)
{
  if (!p) return;
  auto it = c->alloc_bytes.find(p);
  if (it != c->alloc_bytes.end()) {
    c->device_bytes -= it->second;
    c->alloc_bytes.erase(it);
  }
  cudaFree(p);
  p = nullptr;
}

inline unsigned
grid_for(int64_t n, int block)
{
  return (unsigned)((n + block - 1) / block);
}

ShapeTables
make_shape_tables()
{
  // HexElement::HexElement / ShapeFunctionValues / ShapeFunctionDerivatives (src/nimble_element.cc:55-171)
  ShapeTables  t;
  const double g = 0.577350269189626;
  const double c = 1.0 / 8.0;
  for (int q = 0; q < 8; ++q) {
    const double r = sgn_x(q) * g, s = sgn_y(q) * g, tt = sgn_z(q) * g;
    for (int j = 0; j < 8; ++j) {
      const double fr = 1.0 + sgn_x(j) * r, fs = 1.0 + sgn_y(j) * s, ft = 1.0 + sgn_z(j) * tt;
      t.N[8 * q + j]            = c * fr * fs * ft;
      t.dN[24 * q + 3 * j + 0] = (sgn_x(j) * c) * fs * ft;
      t.dN[24 * q + 3 * j + 1] = (sgn_y(j) * c) * fr * ft;
      t.dN[24 * q + 3 * j + 2] = (sgn_z(j) * c) * fr * fs;
    }
  }
  return t;
}

NodeArgs
node_args(nsm_b200_ctx* c, int64_t bc_row = 0)
{
  NodeArgs p{};
  p.n_nodes = c->n_nodes;
  for (int i = 0; i < 3; ++i) {
    p.u[i] = c->u[i], p.v[i] = c->v[i], p.a[i] = c->a[i], p.f[i] = c->f[i];
    p.fext[i]      = c->has_fext ? c->fext[i] : nullptr;
    p.fcontact[i]  = c->contact.active ? c->fc[i] : nullptr;
    p.bc_of_dof[i] = c->bc_of_dof[i];
  }
  p.mass     = c->mass;
  p.bc_kind  = c->bc_kind;
  // with boundary-condition programs the single row is rewritten on the device before each use
  p.bc_value = c->bc_value ? c->bc_value + ((c->bc_rows > 1 && c->bcp_programs == 0) ? bc_row : 0) * c->n_bc : nullptr;
  p.ef       = c->ef;
  p.adj_off  = c->adj_off;
  p.adj_slot = c->adj_slot;
  return p;
}

// Evaluates the boundary-condition programs for one step (slot row `row`) into row 0 of the magnitudes.
int
enqueue_bc_programs(nsm_b200_ctx* c, int64_t row)
{
  if (c->bcp_programs == 0 || c->n_bc == 0) return NSM_OK;
  BcProgramArgs p{};
  p.n_entries        = c->n_bc;
  p.program_of_entry = c->bcp_of_entry;
  p.node_of_entry    = c->bc_node;
  p.offsets          = c->bcp_offsets;
  p.code             = c->bcp_code;
  p.consts           = c->bcp_consts;
  p.slots            = c->bcp_slot_values + (c->bcp_rows > 1 ? row : 0) * c->bcp_slots;
  p.entry_consts     = c->bcp_entry_consts;
  if (c->bcp_n_entry_consts < c->bcp_entry_consts_needed)
    return fail(c, NSM_ERR_ARG, "boundary-condition programs name %d per-entry constants, %d supplied (nsm_b200_set_bc_entry_constants)",
                c->bcp_entry_consts_needed, c->bcp_n_entry_consts);
  for (int i = 0; i < 3; ++i) p.X[i] = c->X[i];
  p.value = c->bc_value;
  bc_program_kernel<<<grid_for(c->n_bc, 128), 128, 0, c->stream>>>(p);
  c->launches++;
  NSM_CUDA(c, cudaGetLastError());
  return NSM_OK;
}

void
free_bc_programs(nsm_b200_ctx* c)
{
  dev_release(c, c->bcp_offsets), dev_release(c, c->bcp_code), dev_release(c, c->bcp_consts), dev_release(c, c->bcp_of_entry);
  dev_release(c, c->bcp_slot_values), dev_release(c, c->bcp_entry_consts);
  c->bcp_programs = c->bcp_slots = c->bcp_rows = c->bcp_rows_cap = 0;
  c->bcp_n_entry_consts = c->bcp_entry_consts_needed = 0;
}

ElemArgs
elem_args(nsm_b200_ctx* c, const Block& b, int sched = kSchedAll)
{
  ElemArgs p{};
  p.sched      = sched;
  p.chunk_mask = (const unsigned char*)b.group_bits;
  p.group_list = b.group_list;
  p.n_list     = b.n_list;
  p.n_elem = b.n_elem;
  p.orig   = b.orig;
  p.conn   = b.conn_sched;
  for (int i = 0; i < 3; ++i) p.X[i] = c->X[i], p.u[i] = c->u[i], p.f[i] = c->f[i];
  p.ef         = c->ef ? c->ef + b.elem_base * 8 * kEfStride : nullptr;
  p.ipt        = b.n_state ? b.rec[b.cur] : (c->ipt ? c->ipt + b.elem_base * 120 : nullptr);
  p.ipt_n      = b.n_state ? b.rec[b.cur ^ 1] : nullptr;
  p.mat_a      = b.mat_a;
  p.mat_b      = b.mat_b;
  p.binv_cache = c->binv ? c->binv + b.group_base * kBinvGroupDoubles : nullptr;
  p.bulk       = b.bulk;
  p.shear      = b.shear;
  p.flags      = c->d_flags;
  p.ticket      = c->d_ticket + c->ticket_parity;  // (the launch sites flip the parity after each launch)
  p.ticket_next = c->d_ticket + (c->ticket_parity ^ 1);
  return p;
}

inline int64_t
groups_of(int64_t n_elem)
{
  return (n_elem + kElemsPerWarp - 1) / kElemsPerWarp;
}

// CTAs a ranged launch of the pipelined host paths leaves free (see enqueue_element_range); 0 everywhere else
thread_local int t_reserve_ctas = 0;

// Persistent launch: one wave of CTAs (SM count x resident CTAs per SM), fewer when the block is small.
template <int MAT, bool ORDERED, int MODE>
cudaError_t
launch_element(const ElemArgs& p, cudaStream_t s)
{
  // function attributes and occupancy are per device: a process may drive several GPUs (one thread each)
  static int wave_of_device[64] = {0};
  auto       k                  = element_force_kernel<MAT, ORDERED, MODE>;
  constexpr int    kThreads = ElemShape<MAT>::threads, kWarps = ElemShape<MAT>::warps;
  constexpr size_t kElemSmemBytes = (size_t)kWarps * warp_smem_doubles<MAT, MODE>() * sizeof(double);
  int        dev                = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  int wave = wave_of_device[dev];
  if (wave == 0) {
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kElemSmemBytes);
    if (e != cudaSuccess) return e;
    int sms = 0, per_sm = 0;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kThreads, kElemSmemBytes)) != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    wave = wave_of_device[dev] = sms * per_sm;
  }
  const int64_t positions = p.sched == kSchedList ? p.n_list : (p.n_range > 0 ? p.n_range : groups_of(p.n_elem));
  const int64_t need      = std::max<int64_t>((positions + kWarps * kTicketChunk - 1) / (kWarps * kTicketChunk), 1);
  // (the reservation is counted in CTAs of NSM_ELEM_THREADS threads)
  const int     reserve   = t_reserve_ctas * (kElemThreads / kThreads);
  k<<<(unsigned)std::min<int64_t>(need, std::max(wave - reserve, 1)), kThreads, kElemSmemBytes, s>>>(p);
  return cudaGetLastError();
}

template <int MAT, bool ORDERED>
cudaError_t
launch_element_mode(const ElemArgs& p, int mode, cudaStream_t s)
{
  switch (mode) {
    case 0: return launch_element<MAT, ORDERED, 0>(p, s);
    case 1: return launch_element<MAT, ORDERED, 1>(p, s);
    case 2: return launch_element<MAT, ORDERED, 2>(p, s);
    default: return launch_element<MAT, ORDERED, 3>(p, s);
  }
}

cudaError_t
launch_element_any(const ElemArgs& p, int material, bool ordered, int mode, cudaStream_t s)
{
  if (material == NSM_MAT_ELASTIC)
    return ordered ? launch_element_mode<0, true>(p, mode, s) : launch_element_mode<0, false>(p, mode, s);
  if (material == NSM_MAT_J2_PLASTICITY) {  // a material with state always writes its records
    if (mode & kModeReadBinv) return ordered ? launch_element<2, true, 3>(p, s) : launch_element<2, false, 3>(p, s);
    return ordered ? launch_element<2, true, 1>(p, s) : launch_element<2, false, 1>(p, s);
  }
  return ordered ? launch_element_mode<1, true>(p, mode, s) : launch_element_mode<1, false>(p, mode, s);
}

int
num_state_of(int material_kind)
{
  return material_kind == NSM_MAT_J2_PLASTICITY ? 2 : 0;
}

// records of a block: pointer, doubles per element
const double*
block_records(nsm_b200_ctx* c, const Block& b, int* per_element)
{
  *per_element = 8 * (15 + b.n_state);
  return b.n_state ? b.rec[b.cur] : c->ipt + b.elem_base * 120;
}

// ModelData::UpdateStates requested earlier: the records written last become the N records of this evaluation
void
roll_states(nsm_b200_ctx* c)
{
  for (auto& kv : c->blocks) {
    Block& b = kv.second;
    if (b.n_state && b.roll_pending) b.cur ^= 1;
    b.roll_pending = false;
  }
}

void
mark_states_for_roll(nsm_b200_ctx* c)
{
  for (auto& kv : c->blocks)
    if (kv.second.n_state) kv.second.roll_pending = true;
}

// one block's element kernel over the groups [g0, g1) on stream s (the pipelined host step)
int
enqueue_element_range(nsm_b200_ctx* c, const Block& b, bool store_ipt, int g0, int g1, cudaStream_t s)
{
  if (g1 <= g0) return NSM_OK;
  int mode = (store_ipt ? kModeStoreIpt : 0) | (c->binv ? kModeReadBinv : 0);
  ElemArgs p    = elem_args(c, b, kSchedAll);
  p.group_begin = g0, p.n_range = g1 - g0;
  // The element kernel is persistent and fills every SM (2 CTAs x 128 registers x 256 threads = the whole register
  // file): the small transpose / node kernels of the upload and download streams would wait for a whole ranged launch
  // to drain before they get an SM, and the copy engines would idle behind them.  Leave a few CTA slots free.
  t_reserve_ctas       = c->pipe.reserve_now;
  const cudaError_t le = launch_element_any(p, b.material, c->assembly == NSM_ASSEMBLY_ORDERED, mode, s);
  t_reserve_ctas       = 0;
  NSM_CUDA(c, le);
  c->ticket_parity ^= 1;
  c->launches++;
  return NSM_OK;
}

int
ensure_ipt(nsm_b200_ctx* c)
{
  if (c->ipt) return NSM_OK;
  bool any_stateless = false;  // blocks of a material with state variables own their records (Block::rec)
  for (auto& kv : c->blocks) any_stateless = any_stateless || kv.second.n_state == 0;
  if (!any_stateless) return NSM_OK;
  // stream-ordered allocation: this may run in the middle of a call whose earlier steps are already queued, and a
  // cudaMalloc would wait for the whole device -- including a peer's in-kernel wait for OUR next shared-node data
  // when ranks share one GPU (RankGroup lockstep mode)
  const int64_t count = std::max<int64_t>(c->n_elem_total * 120, 1);
  NSM_CUDA(c, cudaMallocAsync((void**)&c->ipt, (size_t)count * sizeof(double), c->stream));
  c->device_bytes += count * (int64_t)sizeof(double);
  const int64_t np = c->n_elem_total * 8;
  if (np > 0) {
    init_ipt_kernel<<<grid_for(np, 256), 256, 0, c->stream>>>(np, c->ipt, 15);
    c->launches++;
    NSM_CUDA(c, cudaGetLastError());
  }
  return NSM_OK;
}

// element kernels of all blocks, ascending block id (src/nimble_model_data.cc:636-659)
int
enqueue_element_kernels(nsm_b200_ctx* c, bool store_ipt, int sched = kSchedAll)
{
  const bool ordered = c->assembly == NSM_ASSEMBLY_ORDERED;
  int        mode    = 0;
  if (store_ipt) {
    int rc = ensure_ipt(c);
    if (rc) return rc;
    mode |= kModeStoreIpt;
  }
  if (c->binv) mode |= kModeReadBinv;
  for (auto& kv : c->blocks) {
    const Block& b = kv.second;
    if (b.n_elem == 0 || (sched == kSchedList && b.n_list == 0)) continue;
    NSM_CUDA(c, launch_element_any(elem_args(c, b, sched), b.material, ordered, mode, c->stream));
    c->ticket_parity ^= 1;
    c->launches++;
  }
  return NSM_OK;
}

int
check_flags(nsm_b200_ctx* c)
{
  int h = 0;
  NSM_CUDA(c, cudaMemcpyAsync(&h, c->d_flags, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  if (h & 1) {
    NSM_CUDA(c, cudaMemsetAsync(c->d_flags, 0, sizeof(int), c->stream));
    return fail(c, NSM_ERR_JACOBIAN, "non-positive Jacobian determinant in Invert3x3 (singular or inverted element)");
  }
  if (c->comm.poll_error(c->stream)) return fail(c, NSM_ERR_COMM, "%s", c->comm.error());
  return NSM_OK;
}

cudaEvent_t
prof_event(nsm_b200_ctx* c)
{
  if (c->ev_used == c->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    c->ev_pool.push_back(e);
  }
  cudaEvent_t e = c->ev_pool[c->ev_used++];
  cudaEventRecord(e, c->stream);
  return e;
}

void
prof_resolve(nsm_b200_ctx* c)
{
  // events come in fours: before the element kernels, after them, after the contact evaluation (nothing in between
  // without contact entities), after the node-side work of the step
  cudaStreamSynchronize(c->stream);
  for (size_t i = 0; i + 3 < c->ev_used; i += 4) {
    float m1 = 0, m2 = 0, m3 = 0;
    cudaEventElapsedTime(&m1, c->ev_pool[i], c->ev_pool[i + 1]);
    cudaEventElapsedTime(&m2, c->ev_pool[i + 1], c->ev_pool[i + 2]);
    cudaEventElapsedTime(&m3, c->ev_pool[i + 2], c->ev_pool[i + 3]);
    c->prof_elem_ms += m1;
    c->prof_contact_ms += m2;
    c->prof_node_ms += m3;
    c->prof_steps++;
  }
  c->ev_used = 0;
}

// ContactManager::ComputeContactForce on the device displacement (csrc/contact.cuh): three launches, no host sync
int
enqueue_contact(nsm_b200_ctx* c)
{
  auto& k = c->contact;
  if (!k.active) return NSM_OK;
  NvtxRange   range("Contact");
  ContactArgs p{};
  p.n_quads = k.n_quads, p.n_sec = k.n_sec;
  p.quad = k.quad, p.quad_len = k.quad_len, p.sec_node = k.sec_node, p.sec_len = k.sec_len;
  p.n_surf = k.n_surf, p.surf_node = k.surf_node;
  for (int i = 0; i < 3; ++i) p.X[i] = c->X[i], p.u[i] = c->u[i], p.fc[i] = c->fc[i];
  p.penalty  = k.penalty;
  p.quad_xyz = k.quad_xyz, p.tri_box = k.tri_box, p.bin = k.bin, p.head = k.head;
  p.table_mask = k.table_mask;
  p.red = k.red + 8 * k.parity, p.red_next = k.red + 8 * (k.parity ^ 1);
  p.counters = k.counters, p.status = k.status, p.near_list = k.near_list;
  p.ordered = k.contrib_cap > 0 ? 1 : 0, p.contrib_cap = k.contrib_cap;
  p.contrib_key = k.contrib_key, p.contrib_target = k.contrib_target, p.contrib_val = k.contrib_val;
  k.parity ^= 1;
  const int64_t n_update = std::max<int64_t>(std::max(k.n_quads, k.n_sec), 32);
  contact_update_kernel<<<grid_for(n_update, 256), 256, 0, c->stream>>>(p);
  c->launches++;
  if (k.n_quads > 0 && k.n_sec > 0) {
    contact_bin_kernel<<<grid_for(std::max(k.n_quads, k.n_sec), 256), 256, 0, c->stream>>>(p);
    // a grid that fills the device once; its warps stride over the list of near nodes (whose length only the device knows)
    contact_pair_kernel<<<(unsigned)std::min<int64_t>(grid_for(32 * k.n_sec, 256), 148 * 3 * 2), 256, 0, c->stream>>>(p);
    c->launches += 2;
    if (p.ordered) {
      // ORDERED assembly: sort the filed contributions by their place in the serial order, then (stably) by target node,
      // and add every target's run in order (csrc/contact.cuh).  The number of contributions is read back: one small
      // synchronisation per evaluation, in the mode whose point is reproducibility.
      unsigned long long filed[2] = {0, 0};
      NSM_CUDA(c, cudaMemcpyAsync(filed, k.counters + 5, sizeof filed, cudaMemcpyDeviceToHost, c->stream));
      NSM_CUDA(c, cudaStreamSynchronize(c->stream));
      k.ordered_overflow_pairs = (long long)filed[1];
      const long long n        = std::min<long long>((long long)filed[0], k.contrib_cap / 7 * 7);
      if (n > 0) {
        size_t tmp = k.sort_tmp_bytes;
        NSM_CUDA(c, cub::DeviceRadixSort::SortPairs(k.sort_tmp, tmp, k.contrib_key, k.key_sorted, k.iota, k.order1, (int)n, 0, 64, c->stream));
        contact_gather_targets_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(n, k.order1, k.contrib_target, k.target1);
        tmp = k.sort_tmp_bytes;
        NSM_CUDA(c, cub::DeviceRadixSort::SortPairs(k.sort_tmp, tmp, k.target1, k.target2, k.order1, k.order2, (int)n, 0, 32, c->stream));
        contact_ordered_sum_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(n, k.target2, k.order2, k.contrib_val, c->fc[0], c->fc[1], c->fc[2]);
        c->launches += 2;
      }
    }
  }
  NSM_CUDA(c, cudaGetLastError());
  return NSM_OK;
}

// the stress seam's device scratch, at least `doubles` long
int
seam_scratch(nsm_b200_ctx* c, size_t doubles, double** out)
{
  if (c->seam_cap < doubles) {
    NSM_CUDA(c, cudaStreamSynchronize(c->stream));
    dev_release(c, c->seam);
    c->seam_cap = 0;
    int rc = dev_alloc(c, &c->seam, (int64_t)doubles);
    if (rc) return rc;
    c->seam_cap = doubles;
  }
  *out = c->seam;
  return NSM_OK;
}

// internal force of the current device displacement into the device force field (+ shared-node sum)
int
enqueue_internal_force(nsm_b200_ctx* c, bool store_ipt)
{
  const bool ordered = c->assembly == NSM_ASSEMBLY_ORDERED;
  if (!ordered) {
    for (int i = 0; i < 3; ++i) NSM_CUDA(c, cudaMemsetAsync(c->f[i], 0, (size_t)c->n_nodes * sizeof(double), c->stream));
  }
  roll_states(c);
  int rc = enqueue_element_kernels(c, store_ipt);
  if (rc) return rc;
  if (ordered) {
    node_correct_kernel<true, false><<<grid_for(c->n_nodes, 256), 256, 0, c->stream>>>(node_args(c), 0.0, 0);
    c->launches++;
    NSM_CUDA(c, cudaGetLastError());
  }
  if (c->comm.active()) {
    rc = c->comm.reduce(c->stream, c->f, 3, &c->launches);
    if (rc) return fail(c, NSM_ERR_COMM, "%s", c->comm.error());
  }
  return NSM_OK;
}

int
field_ptrs(nsm_b200_ctx* c, int field, double** p, int* ncomp)
{
  *ncomp = 3;
  switch (field) {
    case NSM_FIELD_LUMPED_MASS:
      p[0]   = c->mass;
      *ncomp = 1;
      return NSM_OK;
    case NSM_FIELD_REFERENCE_COORDINATE: p[0] = c->X[0], p[1] = c->X[1], p[2] = c->X[2]; return NSM_OK;
    case NSM_FIELD_DISPLACEMENT: p[0] = c->u[0], p[1] = c->u[1], p[2] = c->u[2]; return NSM_OK;
    case NSM_FIELD_VELOCITY: p[0] = c->v[0], p[1] = c->v[1], p[2] = c->v[2]; return NSM_OK;
    case NSM_FIELD_ACCELERATION: p[0] = c->a[0], p[1] = c->a[1], p[2] = c->a[2]; return NSM_OK;
    case NSM_FIELD_INTERNAL_FORCE: p[0] = c->f[0], p[1] = c->f[1], p[2] = c->f[2]; return NSM_OK;
    case NSM_FIELD_EXTERNAL_FORCE: p[0] = c->fext[0], p[1] = c->fext[1], p[2] = c->fext[2]; return NSM_OK;
    case NSM_FIELD_CONTACT_FORCE:
      if (!c->fc[0]) return fail(c, NSM_ERR_ARG, "contact_force: no contact entities on this context (nsm_b200_set_contact)");
      p[0] = c->fc[0], p[1] = c->fc[1], p[2] = c->fc[2];
      return NSM_OK;
  }
  return fail(c, NSM_ERR_ARG, "unknown field id %d", field);
}

int
upload_field(nsm_b200_ctx* c, int field, const double* host, bool sync)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "upload_field: context not finalized");
  NSM_REQUIRE(c, host != nullptr, "upload_field: null host pointer");
  double* p[3];
  int     nc;
  int     rc = field_ptrs(c, field, p, &nc);
  if (rc) return rc;
  const int64_t n = c->n_nodes;
  if (nc == 1 && !c->node_perm) {
    NSM_CUDA(c, cudaMemcpyAsync(p[0], host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  } else if (nc == 1) {
    NSM_CUDA(c, cudaMemcpyAsync(c->staging, host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (n > 0) {
      permute_scalar_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(n, c->staging, p[0], c->node_perm, 1);
      c->launches++;
      NSM_CUDA(c, cudaGetLastError());
    }
  } else {
    NSM_CUDA(c, cudaMemcpyAsync(c->staging, host, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (n > 0) {
      aos_to_soa_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(n, c->staging, p[0], p[1], p[2], c->node_perm);
      c->launches++;
      NSM_CUDA(c, cudaGetLastError());
    }
    if (field == NSM_FIELD_EXTERNAL_FORCE) c->has_fext = true;
  }
  if (sync) NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  return NSM_OK;
}

int
download_field(nsm_b200_ctx* c, int field, double* host, bool sync)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "download_field: context not finalized");
  NSM_REQUIRE(c, host != nullptr, "download_field: null host pointer");
  double* p[3];
  int     nc;
  int     rc = field_ptrs(c, field, p, &nc);
  if (rc) return rc;
  const int64_t n = c->n_nodes;
  if (nc == 1 && !c->node_perm) {
    NSM_CUDA(c, cudaMemcpyAsync(host, p[0], (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  } else if (nc == 1) {
    if (n > 0) {
      permute_scalar_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(n, p[0], c->staging, c->node_perm, 0);
      c->launches++;
      NSM_CUDA(c, cudaGetLastError());
    }
    NSM_CUDA(c, cudaMemcpyAsync(host, c->staging, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  } else {
    if (n > 0) {
      soa_to_aos_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(n, p[0], p[1], p[2], c->staging, c->node_perm);
      c->launches++;
      NSM_CUDA(c, cudaGetLastError());
    }
    NSM_CUDA(c, cudaMemcpyAsync(host, c->staging, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  if (sync) NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  return NSM_OK;
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char*
nsm_b200_version(void)
{
  return "nsm_b200 0.1.0 sm_100a fp64 fmad=off"
#ifdef NSM_PLAIN_DIVISION
         " plain-division"
#endif
      ;
}

const char*
nsm_b200_last_error(const nsm_b200_ctx* ctx)
{
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int
nsm_b200_create(int device, nsm_b200_ctx** out)
{
  if (!out) return fail(nullptr, NSM_ERR_ARG, "nsm_b200_create: null out pointer");
  *out  = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, NSM_ERR_CUDA, "no CUDA device available (%s); this library has no CPU path",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= n) return fail(nullptr, NSM_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
  cudaDeviceProp prop;
  NSM_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, NSM_ERR_CUDA, "device %d is sm_%d%d; this library ships sm_100a code only", device, prop.major,
                prop.minor);
  DeviceGuard device_guard_(device);  // the caller's current device is put back on return
  auto* c   = new nsm_b200_ctx;
  c->device = device;
  cudaError_t ce = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaEventCreate(&c->ev_start);
  if (ce == cudaSuccess) ce = cudaEventCreate(&c->ev_stop);
  static const ShapeTables tables = make_shape_tables();
  if (ce == cudaSuccess) ce = cudaMemcpyToSymbol(c_shape, &tables, sizeof tables);
  if (ce != cudaSuccess) {  // nothing of a half-built context survives
    if (c->ev_stop) cudaEventDestroy(c->ev_stop);
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return fail(nullptr, NSM_ERR_CUDA, "nsm_b200_create: %s", cudaGetErrorString(ce));
  }
  *out = c;
  return NSM_OK;
}

void
nsm_b200_destroy(nsm_b200_ctx* c)
{
  if (!c) return;
  DeviceGuard device_guard_(c->device);
  cudaStreamSynchronize(c->stream);
  c->comm.destroy();
  auto fr = [](void* p) {
    if (p) cudaFree(p);
  };
  for (int i = 0; i < 3; ++i) {
    fr(c->X[i]), fr(c->u[i]), fr(c->v[i]), fr(c->a[i]), fr(c->f[i]), fr(c->fext[i]), fr(c->bc_of_dof[i]);
  }
  fr(c->seam);
  fr(c->node_perm), fr(c->mass), fr(c->staging), fr(c->staging_u), fr(c->ipt), fr(c->binv), fr(c->ef), fr(c->adj_off), fr(c->adj_slot);
  if (c->io_stream) cudaStreamDestroy(c->io_stream);
  if (c->ev_u_staged) cudaEventDestroy(c->ev_u_staged);
  fr(c->bc_kind), fr(c->bc_value), fr(c->bc_node), fr(c->d_flags), fr(c->d_min_dt), fr(c->d_ticket);
  free_bc_programs(c);
  for (auto& kv : c->blocks) {
    if (kv.second.conn_sched != kv.second.conn) fr(kv.second.conn_sched);
    fr(kv.second.conn), fr(kv.second.orig), fr(kv.second.group_bits), fr(kv.second.group_list);
    fr(kv.second.rec[0]), fr(kv.second.rec[1]);
  }
  {
    auto& k = c->contact;
    fr(k.quad), fr(k.quad_len), fr(k.sec_node), fr(k.sec_len), fr(k.quad_xyz), fr(k.tri_box), fr(k.bin), fr(k.head);
    fr(k.red), fr(k.counters), fr(k.status), fr(k.near_list), fr(k.surf_node);
    fr(k.contrib_key), fr(k.key_sorted), fr(k.contrib_target), fr(k.contrib_val), fr(k.iota), fr(k.order1), fr(k.order2), fr(k.target1),
        fr(k.target2), fr(k.sort_tmp);
    for (int i = 0; i < 3; ++i) fr(c->fc[i]);
  }
  for (double* p : c->pipe.stage) fr(p);
  if (c->pipe.up) cudaStreamDestroy(c->pipe.up);
  if (c->pipe.down) cudaStreamDestroy(c->pipe.down);
  if (c->pipe.ev_bc) cudaEventDestroy(c->pipe.ev_bc);
  for (auto e : c->pipe.ev_up) cudaEventDestroy(e);
  for (auto e : c->pipe.ev_elem) cudaEventDestroy(e);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  if (c->ev_boundary) cudaEventDestroy(c->ev_boundary);
  if (c->ev_packed) cudaEventDestroy(c->ev_packed);
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  cudaEventDestroy(c->ev_start);
  cudaEventDestroy(c->ev_stop);
  cudaStreamDestroy(c->stream);
  delete c;
}

int
nsm_b200_set_nodes(nsm_b200_ctx* c, int64_t n_nodes, const double* x, const double* y, const double* z)
{
  NSM_REQUIRE(c, c != nullptr, "null context");
  NSM_REQUIRE(c, !c->finalized, "set_nodes after finalize");
  NSM_REQUIRE(c, n_nodes >= 0 && n_nodes < (int64_t)2147483647, "node count must fit int32 local ids");
  NSM_REQUIRE(c, n_nodes == 0 || (x && y && z), "null coordinate pointer");
  c->n_nodes = n_nodes;
  c->hx.assign(x, x + n_nodes);
  c->hy.assign(y, y + n_nodes);
  c->hz.assign(z, z + n_nodes);
  return NSM_OK;
}

int
nsm_b200_add_block(nsm_b200_ctx* c, int block_id, int64_t n_elem, const int32_t* conn, int material_kind,
                   double bulk_modulus, double shear_modulus, double density)
{
  const double params[3] = {bulk_modulus, shear_modulus, density};
  return nsm_b200_add_block_params(c, block_id, n_elem, conn, material_kind, 3, params);
}

int
nsm_b200_material_num_state(int material_kind)
{
  return (material_kind >= 0 && material_kind < NSM_MAT_COUNT) ? num_state_of(material_kind) : -1;
}

int
nsm_b200_material_num_params(int material_kind)
{
  if (material_kind < 0 || material_kind >= NSM_MAT_COUNT) return -1;
  return material_kind == NSM_MAT_J2_PLASTICITY ? 5 : 3;
}

const char*
nsm_b200_material_state_label(int material_kind, int index)
{
  static const char* const j2[2] = {"equivalent_plastic_strain", "von_mises_stress"};
  if (material_kind == NSM_MAT_J2_PLASTICITY && index >= 0 && index < 2) return j2[index];
  return nullptr;
}

double
nsm_b200_material_state_initial_value(int, int)
{
  return 0.0;
}

int
nsm_b200_add_block_params(nsm_b200_ctx* c, int block_id, int64_t n_elem, const int32_t* conn, int material_kind, int n_params,
                          const double* params)
{
  NSM_REQUIRE(c, c != nullptr, "null context");
  if (material_kind < 0 || material_kind >= NSM_MAT_COUNT)
    return fail(c, NSM_ERR_MATERIAL, "unknown material kind %d (0 = elastic, 1 = neohookean, 2 = j2_plasticity)", material_kind);
  if (n_params != nsm_b200_material_num_params(material_kind) || !params)
    return fail(c, NSM_ERR_MATERIAL, "material kind %d takes %d parameters, %d given", material_kind,
                nsm_b200_material_num_params(material_kind), n_params);
  const double bulk_modulus = params[0], shear_modulus = params[1], density = params[2];
  NSM_REQUIRE(c, !c->finalized, "add_block after finalize");
  NSM_REQUIRE(c, n_elem >= 0 && (n_elem == 0 || conn), "bad element count / null connectivity");
  NSM_REQUIRE(c, n_elem <= ((int64_t)1 << 32), "a block holds at most 2^32 elements (32-bit group indices)");
  NSM_REQUIRE(c, c->blocks.find(block_id) == c->blocks.end(), "duplicate block id");
  for (int64_t i = 0; i < n_elem * 8; ++i)
    if (conn[i] < 0 || conn[i] >= c->n_nodes)
      return fail(c, NSM_ERR_ARG, "block %d: connectivity entry %lld = %d outside [0, %lld)", block_id, (long long)i,
                  conn[i], (long long)c->n_nodes);
  Block b;
  b.id = block_id, b.n_elem = n_elem, b.material = material_kind;
  b.bulk = bulk_modulus, b.shear = shear_modulus, b.density = density;
  b.n_state = num_state_of(material_kind);
  if (material_kind == NSM_MAT_J2_PLASTICITY) b.mat_a = params[3], b.mat_b = params[4];
  b.conn_host.assign(conn, conn + n_elem * 8);
  c->blocks[block_id] = std::move(b);
  return NSM_OK;
}

int
nsm_b200_finalize(nsm_b200_ctx* c, int assembly, unsigned flags)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, !c->finalized, "finalize called twice");
  NSM_REQUIRE(c, assembly == NSM_ASSEMBLY_ATOMIC || assembly == NSM_ASSEMBLY_ORDERED, "unknown assembly mode");
  c->assembly = assembly;
  c->flags_   = flags;
  const int64_t n = c->n_nodes;
  int           rc;
  {  // size limits first: nothing is allocated for a model this build cannot index
    int64_t total = 0;
    for (auto& kv : c->blocks) total += kv.second.n_elem;
    NSM_REQUIRE(c, total * 8 < (int64_t)4294967295LL, "too many elements for 32-bit assembly slots on one GPU");
  }
  for (int i = 0; i < 3; ++i) {
    if ((rc = dev_alloc(c, &c->X[i], n))) return rc;
    if ((rc = dev_alloc(c, &c->u[i], n))) return rc;
    if ((rc = dev_alloc(c, &c->v[i], n))) return rc;
    if ((rc = dev_alloc(c, &c->a[i], n))) return rc;
    if ((rc = dev_alloc(c, &c->f[i], n))) return rc;
    if ((rc = dev_alloc(c, &c->fext[i], n))) return rc;
    for (double* p : {c->u[i], c->v[i], c->a[i], c->f[i], c->fext[i]})
      NSM_CUDA(c, cudaMemsetAsync(p, 0, (size_t)std::max<int64_t>(n, 1) * sizeof(double), c->stream));
  }
  if ((rc = dev_alloc(c, &c->mass, n))) return rc;
  NSM_CUDA(c, cudaMemsetAsync(c->mass, 0, (size_t)std::max<int64_t>(n, 1) * sizeof(double), c->stream));
  if ((rc = dev_alloc(c, &c->staging, n * 3))) return rc;
  if ((rc = dev_alloc(c, &c->d_flags, 2))) return rc;
  if ((rc = dev_alloc(c, &c->d_min_dt, 1))) return rc;
  if ((rc = dev_alloc(c, &c->d_ticket, 2))) return rc;
  NSM_CUDA(c, cudaMemsetAsync(c->d_ticket, 0, 2 * sizeof(unsigned), c->stream));
  NSM_CUDA(c, cudaMemsetAsync(c->d_flags, 0, 2 * sizeof(int), c->stream));
  if ((flags & NSM_FLAG_RENUMBER_NODES) && n > 1) {
    // internal node order = Morton order of the coordinates (21 bits per axis of the bounding box); the caller's
    // ids are mapped at every entry point that takes or returns nodal data
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    const std::vector<double>* xyz[3] = {&c->hx, &c->hy, &c->hz};
    for (int d = 0; d < 3; ++d)
      for (int64_t i = 0; i < n; ++i) lo[d] = std::min(lo[d], (*xyz[d])[i]), hi[d] = std::max(hi[d], (*xyz[d])[i]);
    auto spread = [](uint64_t v) {
      v &= 0x1fffffULL;
      v = (v | v << 32) & 0x1f00000000ffffULL;
      v = (v | v << 16) & 0x1f0000ff0000ffULL;
      v = (v | v << 8) & 0x100f00f00f00f00fULL;
      v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
      v = (v | v << 2) & 0x1249249249249249ULL;
      return v;
    };
    std::vector<std::pair<uint64_t, int>> key((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
      uint64_t k = 0;
      for (int d = 0; d < 3; ++d) {
        const double w = hi[d] > lo[d] ? ((*xyz[d])[i] - lo[d]) / (hi[d] - lo[d]) : 0.0;
        k |= spread((uint64_t)(w * 2097151.0)) << d;
      }
      key[i] = {k, (int)i};
    }
    std::sort(key.begin(), key.end());
    c->node_perm_host.resize((size_t)n);
    std::vector<double> nx((size_t)n), ny((size_t)n), nz((size_t)n);
    for (int64_t k = 0; k < n; ++k) {
      const int i                = key[k].second;
      c->node_perm_host[i]       = (int)k;
      nx[k] = c->hx[i], ny[k] = c->hy[i], nz[k] = c->hz[i];
    }
    c->hx.swap(nx), c->hy.swap(ny), c->hz.swap(nz);
    for (auto& kv : c->blocks)
      for (int& nd : kv.second.conn_host) nd = c->node_perm_host[nd];
    if ((rc = dev_alloc(c, &c->node_perm, n))) return rc;
    NSM_CUDA(c, cudaMemcpyAsync(c->node_perm, c->node_perm_host.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  }
  NSM_CUDA(c, cudaMemcpyAsync(c->X[0], c->hx.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NSM_CUDA(c, cudaMemcpyAsync(c->X[1], c->hy.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NSM_CUDA(c, cudaMemcpyAsync(c->X[2], c->hz.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));

  int64_t base = 0, gbase = 0;
  for (auto& kv : c->blocks) {
    Block& b    = kv.second;
    b.elem_base = base;
    b.group_base = gbase;
    base += b.n_elem;
    gbase += groups_of(b.n_elem);
    if ((rc = dev_alloc(c, &b.conn, b.n_elem * 8))) return rc;
    NSM_CUDA(c, cudaMemcpyAsync(b.conn, b.conn_host.data(), (size_t)b.n_elem * 8 * sizeof(int), cudaMemcpyHostToDevice,
                                c->stream));
    b.conn_sched = b.conn;
    if ((flags & NSM_FLAG_REORDER_ELEMENTS) && b.n_elem > 1) {
      // schedule = elements sorted along a Morton curve of their centroids (21 bits per axis of the block's box):
      // consecutive groups touch neighbouring nodes again, whatever the file order was
      const int64_t ne = b.n_elem;
      double        lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      std::vector<double> cen((size_t)ne * 3);
      const double* xyz[3] = {c->hx.data(), c->hy.data(), c->hz.data()};
      for (int64_t e = 0; e < ne; ++e)
        for (int d = 0; d < 3; ++d) {
          double s = 0.0;
          for (int j = 0; j < 8; ++j) s += xyz[d][b.conn_host[e * 8 + j]];
          cen[e * 3 + d] = s;
          lo[d] = std::min(lo[d], s), hi[d] = std::max(hi[d], s);
        }
      auto spread = [](uint64_t v) {  // 21 bits -> every third bit
        v &= 0x1fffffULL;
        v = (v | v << 32) & 0x1f00000000ffffULL;
        v = (v | v << 16) & 0x1f0000ff0000ffULL;
        v = (v | v << 8) & 0x100f00f00f00f00fULL;
        v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
        v = (v | v << 2) & 0x1249249249249249ULL;
        return v;
      };
      std::vector<std::pair<uint64_t, int>> key((size_t)ne);
      for (int64_t e = 0; e < ne; ++e) {
        uint64_t k = 0;
        for (int d = 0; d < 3; ++d) {
          const double w = hi[d] > lo[d] ? (cen[e * 3 + d] - lo[d]) / (hi[d] - lo[d]) : 0.0;
          k |= spread((uint64_t)(w * 2097151.0)) << d;
        }
        key[e] = {k, (int)e};
      }
      std::sort(key.begin(), key.end());
      std::vector<int> orig((size_t)ne), conn_s((size_t)ne * 8);
      for (int64_t s = 0; s < ne; ++s) {
        orig[s] = key[s].second;
        for (int j = 0; j < 8; ++j) conn_s[s * 8 + j] = b.conn_host[(int64_t)key[s].second * 8 + j];
      }
      if ((rc = dev_alloc(c, &b.conn_sched, ne * 8))) return rc;
      if ((rc = dev_alloc(c, &b.orig, ne))) return rc;
      NSM_CUDA(c, cudaMemcpy(b.conn_sched, conn_s.data(), (size_t)ne * 8 * sizeof(int), cudaMemcpyHostToDevice));
      NSM_CUDA(c, cudaMemcpy(b.orig, orig.data(), (size_t)ne * sizeof(int), cudaMemcpyHostToDevice));
    }
  }
  c->n_elem_total = base;
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  for (auto& kv : c->blocks) std::vector<int>().swap(kv.second.conn_host);
  std::vector<double>().swap(c->hx), std::vector<double>().swap(c->hy), std::vector<double>().swap(c->hz);

  if (assembly == NSM_ASSEMBLY_ORDERED) {
    // node -> (element, local node) adjacency, ascending slot == ascending (block id, element)
    if ((rc = dev_alloc(c, &c->ef, c->n_elem_total * 8 * kEfStride))) return rc;
    NSM_CUDA(c, cudaMemsetAsync(c->ef, 0, (size_t)std::max<int64_t>(c->n_elem_total * 8 * kEfStride, 1) * sizeof(double), c->stream));
    if ((rc = dev_alloc(c, &c->adj_off, n + 1))) return rc;
    if ((rc = dev_alloc(c, &c->adj_slot, c->n_elem_total * 8))) return rc;
    unsigned long long* counts = nullptr;
    void*               tmp    = nullptr;
    struct Scratch  // freed on every return path
    {
      unsigned long long*& a;
      void*&               b;
      ~Scratch()
      {
        if (a) cudaFree(a);
        if (b) cudaFree(b);
      }
    } scratch{counts, tmp};
    NSM_CUDA(c, cudaMalloc((void**)&counts, (size_t)(n + 1) * sizeof(unsigned long long)));
    NSM_CUDA(c, cudaMemsetAsync(counts, 0, (size_t)(n + 1) * sizeof(unsigned long long), c->stream));
    for (auto& kv : c->blocks) {
      const Block& b = kv.second;
      if (b.n_elem == 0) continue;
      adj_count_kernel<<<grid_for(b.n_elem * 8, 256), 256, 0, c->stream>>>(b.n_elem * 8, b.conn, counts);
      c->launches++;
    }
    size_t tmp_size = 0;
    NSM_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tmp_size, (const long long*)counts, (long long*)c->adj_off,
                                               (int)(n + 1), c->stream));
    NSM_CUDA(c, cudaMalloc(&tmp, tmp_size ? tmp_size : 1));
    NSM_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp, tmp_size, (const long long*)counts, (long long*)c->adj_off,
                                               (int)(n + 1), c->stream));
    c->launches++;
    NSM_CUDA(c, cudaMemsetAsync(counts, 0, (size_t)(n + 1) * sizeof(unsigned long long), c->stream));
    for (auto& kv : c->blocks) {
      const Block& b = kv.second;
      if (b.n_elem == 0) continue;
      adj_fill_kernel<<<grid_for(b.n_elem * 8, 256), 256, 0, c->stream>>>(b.n_elem * 8, b.elem_base * 8, b.conn,
                                                                          c->adj_off, counts, c->adj_slot);
      c->launches++;
    }
    if (n > 0) {
      adj_sort_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(n, c->adj_off, c->adj_slot);
      c->launches++;
    }
    NSM_CUDA(c, cudaStreamSynchronize(c->stream));
    NSM_CUDA(c, cudaGetLastError());
  }

  for (auto& kv : c->blocks) {  // Block::InitializeElementData (src/nimble_block.cc:148-207): N and N+1 start alike
    Block& b = kv.second;
    if (!b.n_state) continue;
    const int     record = 15 + b.n_state;
    const int64_t np     = b.n_elem * 8;
    for (int k = 0; k < 2; ++k) {
      if ((rc = dev_alloc(c, &b.rec[k], np * record))) return rc;
      if (np > 0) {
        init_ipt_kernel<<<grid_for(np, 256), 256, 0, c->stream>>>(np, b.rec[k], record);
        c->launches++;
      }
    }
    NSM_CUDA(c, cudaGetLastError());
  }
  c->finalized = true;
  if (flags & NSM_FLAG_STORE_IPT_EVERY_STEP) {
    if ((rc = ensure_ipt(c))) return rc;
  }
  if (flags & NSM_FLAG_CACHE_REF_JACOBIAN) {
    int64_t n_groups = 0;
    for (auto& kv : c->blocks) n_groups += groups_of(kv.second.n_elem);
    // The cache is a bytes-for-flops trade (576 B per element for 16 % fewer FP64 instructions) and the FIRST thing to
    // give way when memory is short: it must leave room for the integration-point records of an output step
    // (960 B per element, allocated on first use) next to what is already resident -- e.g. the N / N+1 records of a
    // material with state variables, 2 x 1088 B per element.  64 M elements: 36.9 GB cache + 61.4 GB records fit beside
    // 14 GB of mesh and fields; with 139 GB of state records they do not, and the reference Jacobians are recomputed.
    size_t free_b = 0, total_b = 0;
    NSM_CUDA(c, cudaMemGetInfo(&free_b, &total_b));
    int64_t stateless = 0;
    for (auto& kv : c->blocks) stateless += kv.second.n_state ? 0 : kv.second.n_elem;
    const int64_t need    = n_groups * kBinvGroupDoubles * (int64_t)sizeof(double);
    const int64_t reserve = (c->ipt ? 0 : stateless * 120 * (int64_t)sizeof(double)) + (int64_t)(total_b / 20);
    if (need + reserve > (int64_t)free_b) {
      flags &= ~(unsigned)NSM_FLAG_CACHE_REF_JACOBIAN;
      c->flags_ = flags;
    }
  }
  if (flags & NSM_FLAG_CACHE_REF_JACOBIAN) {
    int64_t n_groups = 0;
    for (auto& kv : c->blocks) n_groups += groups_of(kv.second.n_elem);
    if ((rc = dev_alloc(c, &c->binv, n_groups * kBinvGroupDoubles))) return rc;
    for (auto& kv : c->blocks) {
      const Block& b = kv.second;
      if (b.n_elem == 0) continue;
      binv_cache_kernel<<<grid_for(groups_of(b.n_elem) * 32, kElemThreads), kElemThreads, 0, c->stream>>>(elem_args(c, b));
      c->launches++;
      NSM_CUDA(c, cudaGetLastError());
    }
    if ((rc = check_flags(c))) return rc;
  }
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  return NSM_OK;
}

int64_t
nsm_b200_num_nodes(const nsm_b200_ctx* c)
{
  return c ? c->n_nodes : -1;
}

int64_t
nsm_b200_num_elements(const nsm_b200_ctx* c, int block_id)
{
  if (!c) return -1;
  if (block_id < 0) {
    int64_t n = 0;
    for (auto& kv : c->blocks) n += kv.second.n_elem;
    return n;
  }
  auto it = c->blocks.find(block_id);
  return it == c->blocks.end() ? -1 : it->second.n_elem;
}

int64_t
nsm_b200_device_bytes(const nsm_b200_ctx* c)
{
  return c ? c->device_bytes : -1;
}

unsigned
nsm_b200_effective_flags(const nsm_b200_ctx* c)
{
  return c ? c->flags_ : 0u;
}

int
nsm_b200_upload_field(nsm_b200_ctx* c, int field, const double* host)
{
  return upload_field(c, field, host, true);
}
int
nsm_b200_download_field(nsm_b200_ctx* c, int field, double* host)
{
  return download_field(c, field, host, true);
}
int
nsm_b200_upload_field_async(nsm_b200_ctx* c, int field, const double* host)
{
  return upload_field(c, field, host, false);
}
int
nsm_b200_download_field_async(nsm_b200_ctx* c, int field, double* host)
{
  // the AoS bounce buffer is shared: serialise with earlier async transfers through the stream order
  return download_field(c, field, host, false);
}

int
nsm_b200_sync(nsm_b200_ctx* c)
{
  NSM_ENTER(c);
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  return NSM_OK;
}

void*
nsm_b200_host_alloc(int64_t bytes)
{
  void* p = nullptr;
  if (cudaMallocHost(&p, (size_t)(bytes > 0 ? bytes : 1)) != cudaSuccess) return nullptr;
  return p;
}

void
nsm_b200_host_free(void* p)
{
  if (p) cudaFreeHost(p);
}

int
nsm_b200_compute_lumped_mass(nsm_b200_ctx* c, double* critical_dt)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "compute_lumped_mass: context not finalized");
  const bool ordered = c->assembly == NSM_ASSEMBLY_ORDERED;
  const int64_t n    = c->n_nodes;
  const unsigned long long inf_bits = 0x7ff0000000000000ULL;
  NSM_CUDA(c, cudaMemcpyAsync(c->d_min_dt, &inf_bits, sizeof inf_bits, cudaMemcpyHostToDevice, c->stream));
  NSM_CUDA(c, cudaMemsetAsync(c->mass, 0, (size_t)std::max<int64_t>(n, 1) * sizeof(double), c->stream));
  double* em = nullptr;  // (stream-ordered: see ensure_ipt)
  if (ordered) NSM_CUDA(c, cudaMallocAsync((void**)&em, (size_t)std::max<int64_t>(c->n_elem_total * 8, 1) * sizeof(double), c->stream));
  for (auto& kv : c->blocks) {
    const Block& b = kv.second;
    if (b.n_elem == 0) continue;
    SetupArgs p{};
    p.n_elem = b.n_elem, p.conn = b.conn;
    for (int i = 0; i < 3; ++i) p.X[i] = c->X[i], p.u[i] = c->u[i];
    p.density = b.density, p.bulk = b.bulk;
    p.mass        = c->mass;
    p.em          = ordered ? em + b.elem_base * 8 : nullptr;
    p.min_dt_bits = c->d_min_dt;
    p.flags       = c->d_flags;
    if (ordered)
      lumped_mass_kernel<true><<<grid_for(b.n_elem, 128), 128, 0, c->stream>>>(p);
    else
      lumped_mass_kernel<false><<<grid_for(b.n_elem, 128), 128, 0, c->stream>>>(p);
    c->launches++;
    NSM_CUDA(c, cudaGetLastError());
  }
  if (ordered && n > 0) {
    node_gather_scalar_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(n, em, c->adj_off, c->adj_slot, c->mass);
    c->launches++;
    NSM_CUDA(c, cudaGetLastError());
  }
  if (c->comm.active()) {
    double* m1[3] = {c->mass, nullptr, nullptr};
    if (c->comm.reduce(c->stream, m1, 1, &c->launches)) return fail(c, NSM_ERR_COMM, "%s", c->comm.error());
  }
  unsigned long long bits = 0;
  NSM_CUDA(c, cudaMemcpyAsync(&bits, c->d_min_dt, sizeof bits, cudaMemcpyDeviceToHost, c->stream));
  if (em) cudaFreeAsync(em, c->stream);
  int rc = check_flags(c);
  if (rc) return rc;
  if (critical_dt) memcpy(critical_dt, &bits, sizeof bits);
  return NSM_OK;
}

int
nsm_b200_internal_force(nsm_b200_ctx* c, int store_ipt)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "internal_force: context not finalized");
  NvtxRange range("Force calculation");
  int rc = enqueue_internal_force(c, store_ipt != 0 || (c->flags_ & NSM_FLAG_STORE_IPT_EVERY_STEP));
  if (rc) return rc;
  return check_flags(c);
}

int
nsm_b200_compute_stress(nsm_b200_ctx* c, int material_kind, double bulk, double shear, int64_t n_points,
                        const double* def_grad, double* stress)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, n_points >= 0 && (n_points == 0 || (def_grad && stress)), "compute_stress: bad arguments");
  if (material_kind != NSM_MAT_ELASTIC && material_kind != NSM_MAT_NEOHOOKEAN)
    return fail(c, NSM_ERR_MATERIAL, "unknown material kind %d", material_kind);
  if (n_points == 0) return NSM_OK;
  double* dF = nullptr;
  int     rc = seam_scratch(c, (size_t)n_points * 15, &dF);
  if (rc) return rc;
  double* dS = dF + (size_t)n_points * 9;
  NSM_CUDA(c, cudaMemcpyAsync(dF, def_grad, (size_t)n_points * 9 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (material_kind == NSM_MAT_ELASTIC)
    stress_kernel<0><<<grid_for(n_points, 128), 128, 0, c->stream>>>(n_points, dF, dS, bulk, shear);
  else
    stress_kernel<1><<<grid_for(n_points, 128), 128, 0, c->stream>>>(n_points, dF, dS, bulk, shear);
  c->launches++;
  NSM_CUDA(c, cudaGetLastError());
  NSM_CUDA(c, cudaMemcpyAsync(stress, dS, (size_t)n_points * 6 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  return NSM_OK;
}

int
nsm_b200_compute_stress_state(nsm_b200_ctx* c, int material_kind, int n_params, const double* params, int64_t n_points,
                              const double* def_grad_n, const double* def_grad_np1, const double* stress_n, const double* state_n,
                              double* stress_np1, double* state_np1)
{
  NSM_ENTER(c);
  if (material_kind < 0 || material_kind >= NSM_MAT_COUNT) return fail(c, NSM_ERR_MATERIAL, "unknown material kind %d", material_kind);
  if (n_params != nsm_b200_material_num_params(material_kind) || !params)
    return fail(c, NSM_ERR_MATERIAL, "material kind %d takes %d parameters, %d given", material_kind,
                nsm_b200_material_num_params(material_kind), n_params);
  const int ns = num_state_of(material_kind);
  if (ns == 0)  // F_n, sigma_n and the state arrays are not read by a material without state
    return nsm_b200_compute_stress(c, material_kind, params[0], params[1], n_points, def_grad_np1, stress_np1);
  NSM_REQUIRE(c, n_points >= 0 && (n_points == 0 || (def_grad_n && def_grad_np1 && stress_n && state_n && stress_np1 && state_np1)),
              "compute_stress_state: bad arguments");
  if (n_points == 0) return NSM_OK;
  // one device buffer: F_n 9, F_np1 9, sigma_n 6, state_n ns | sigma_np1 6, state_np1 ns
  const size_t n = (size_t)n_points;
  double*      d = nullptr;
  int          rc_s = seam_scratch(c, n * (size_t)(30 + 2 * ns), &d);
  if (rc_s) return rc_s;
  double *dFn = d, *dF = dFn + 9 * n, *dsn = dF + 9 * n, *dstn = dsn + 6 * n, *ds = dstn + ns * n, *dst = ds + 6 * n;
  cudaError_t e = cudaMemcpyAsync(dFn, def_grad_n, 9 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dF, def_grad_np1, 9 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dsn, stress_n, 6 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dstn, state_n, ns * n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    stress_state_kernel<<<grid_for(n_points, 128), 128, 0, c->stream>>>(n_points, dFn, dF, dsn, dstn, ds, dst, params[0], params[1],
                                                                        params[3], params[4]);
    c->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(stress_np1, ds, 6 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(state_np1, dst, ns * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) return fail(c, NSM_ERR_CUDA, "compute_stress_state: %s", cudaGetErrorString(e));
  return NSM_OK;
}

int
nsm_b200_set_bc_table(nsm_b200_ctx* c, int64_t n, const int32_t* node, const int32_t* comp, const int32_t* kind)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "set_bc_table: context not finalized");
  NSM_REQUIRE(c, n >= 0 && (n == 0 || (node && comp && kind)), "set_bc_table: bad arguments");
  for (int i = 0; i < 3; ++i) dev_release(c, c->bc_of_dof[i]);
  dev_release(c, c->bc_kind), dev_release(c, c->bc_value), dev_release(c, c->bc_node);
  free_bc_programs(c);
  c->n_bc = n;
  if (n == 0) return NSM_OK;
  // dof -> last table entry constraining it (later entries win, like the reference's sequential loop)
  std::vector<int> map[3];
  for (int i = 0; i < 3; ++i) map[i].assign((size_t)c->n_nodes, -1);
  for (int64_t k = 0; k < n; ++k) {
    if (node[k] < 0 || node[k] >= c->n_nodes || comp[k] < 0 || comp[k] > 2 ||
        (kind[k] != NSM_BC_PRESCRIBED_VELOCITY && kind[k] != NSM_BC_PRESCRIBED_DISPLACEMENT))
      return fail(c, NSM_ERR_ARG, "set_bc_table: entry %lld invalid (node %d comp %d kind %d)", (long long)k, node[k],
                  comp[k], kind[k]);
    map[comp[k]][c->node_perm_host.empty() ? node[k] : c->node_perm_host[node[k]]] = (int)k;
  }
  std::vector<int> node_internal(node, node + n);
  if (!c->node_perm_host.empty())
    for (int& nd : node_internal) nd = c->node_perm_host[nd];
  int rc;
  for (int i = 0; i < 3; ++i) {
    if ((rc = dev_alloc(c, &c->bc_of_dof[i], c->n_nodes))) return rc;
    NSM_CUDA(c, cudaMemcpyAsync(c->bc_of_dof[i], map[i].data(), (size_t)c->n_nodes * sizeof(int), cudaMemcpyHostToDevice,
                                c->stream));
  }
  if ((rc = dev_alloc(c, &c->bc_kind, n))) return rc;
  if ((rc = dev_alloc(c, &c->bc_value, n))) return rc;
  if ((rc = dev_alloc(c, &c->bc_node, n))) return rc;
  NSM_CUDA(c, cudaMemcpyAsync(c->bc_node, node_internal.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  c->bc_rows = 1, c->bc_rows_cap = 1;
  NSM_CUDA(c, cudaMemcpyAsync(c->bc_kind, kind, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  NSM_CUDA(c, cudaMemsetAsync(c->bc_value, 0, (size_t)n * sizeof(double), c->stream));
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  return NSM_OK;
}

int
nsm_b200_set_bc_values(nsm_b200_ctx* c, int64_t n, const double* value)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "set_bc_values: context not finalized");
  NSM_REQUIRE(c, n == c->n_bc, "set_bc_values: length differs from the BC table");
  if (n == 0) return NSM_OK;
  c->bc_rows = 1;
  NSM_CUDA(c, cudaMemcpyAsync(c->bc_value, value, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));  // `value` may be reused by the caller right away
  return NSM_OK;
}

int
nsm_b200_set_bc_values_steps(nsm_b200_ctx* c, int n_rows, int64_t n, const double* value)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "set_bc_values_steps: context not finalized");
  NSM_REQUIRE(c, n == c->n_bc && n_rows >= 1, "set_bc_values_steps: bad arguments");
  if (n == 0) return NSM_OK;
  if (n_rows > c->bc_rows_cap) {
    NSM_CUDA(c, cudaStreamSynchronize(c->stream));
    dev_release(c, c->bc_value);
    int rc      = dev_alloc(c, &c->bc_value, (int64_t)n_rows * n);
    if (rc) return rc;
    c->bc_rows_cap = n_rows;
  }
  c->bc_rows = n_rows;
  NSM_CUDA(c, cudaMemcpyAsync(c->bc_value, value, (size_t)n_rows * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  return NSM_OK;
}

int
nsm_b200_set_bc_programs(nsm_b200_ctx* c, int n_programs, const int32_t* program_offsets, const int32_t* code, int n_consts,
                         const double* consts, int n_slots, int64_t n_entries, const int32_t* program_of_entry)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "set_bc_programs: context not finalized");
  NSM_REQUIRE(c, n_programs >= 0 && n_consts >= 0 && n_slots >= 0, "set_bc_programs: negative count");
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  free_bc_programs(c);
  if (n_programs == 0) return NSM_OK;
  NSM_REQUIRE(c, program_offsets && code && program_of_entry && (n_consts == 0 || consts), "set_bc_programs: null argument");
  NSM_REQUIRE(c, n_entries == c->n_bc && c->n_bc > 0, "set_bc_programs: length differs from the BC table");
  NSM_REQUIRE(c, program_offsets[0] == 0, "set_bc_programs: offsets must start at 0");
  // validate: every program leaves exactly one value, never underflows or exceeds the device stack, and only
  // names constants / slots that exist
  int entry_consts_needed = 0;
  for (int p = 0; p < n_programs; ++p) {
    NSM_REQUIRE(c, program_offsets[p + 1] > program_offsets[p], "set_bc_programs: empty program");
    int depth = 0;
    for (int pc = program_offsets[p]; pc < program_offsets[p + 1]; ++pc) {
      const int op = code[pc] & 0xff, arg = code[pc] >> 8;
      int       pops = 2;
      if (op == NSM_BCOP_ENTRYCONST) {
        pops = 0;
        if (arg < 0) return fail(c, NSM_ERR_ARG, "set_bc_programs: negative per-entry constant index");
        entry_consts_needed = std::max(entry_consts_needed, arg + 1);
      } else if (op <= NSM_BCOP_SLOT) {
        pops = 0;
        if (op == NSM_BCOP_CONST && (arg < 0 || arg >= n_consts)) return fail(c, NSM_ERR_ARG, "set_bc_programs: constant %d out of range", arg);
        if (op == NSM_BCOP_SLOT && (arg < 0 || arg >= n_slots)) return fail(c, NSM_ERR_ARG, "set_bc_programs: slot %d out of range", arg);
      } else if (op == NSM_BCOP_NEG || (op >= NSM_BCOP_SQRT && op <= NSM_BCOP_ROUND) || op == NSM_BCOP_NOT)
        pops = 1;
      else if (op == NSM_BCOP_SELECT)
        pops = 3;
      else if (op >= NSM_BCOP_COUNT)
        return fail(c, NSM_ERR_ARG, "set_bc_programs: unknown operation %d", op);
      if (depth < pops) return fail(c, NSM_ERR_ARG, "set_bc_programs: program %d underflows its stack", p);
      depth += 1 - pops;
      if (depth > NSM_BC_STACK_DEPTH) return fail(c, NSM_ERR_ARG, "set_bc_programs: program %d needs more than %d stack entries", p, NSM_BC_STACK_DEPTH);
    }
    if (depth != 1) return fail(c, NSM_ERR_ARG, "set_bc_programs: program %d leaves %d values", p, depth);
  }
  for (int64_t k = 0; k < n_entries; ++k)
    if (program_of_entry[k] < -1 || program_of_entry[k] >= n_programs)
      return fail(c, NSM_ERR_ARG, "set_bc_programs: entry %lld names program %d", (long long)k, program_of_entry[k]);
  const int n_code = program_offsets[n_programs];
  int       rc;
  if ((rc = dev_alloc(c, &c->bcp_offsets, n_programs + 1))) return rc;
  if ((rc = dev_alloc(c, &c->bcp_code, n_code))) return rc;
  if ((rc = dev_alloc(c, &c->bcp_consts, std::max(n_consts, 1)))) return rc;
  if ((rc = dev_alloc(c, &c->bcp_of_entry, n_entries))) return rc;
  if ((rc = dev_alloc(c, &c->bcp_slot_values, std::max(n_slots, 1)))) return rc;
  NSM_CUDA(c, cudaMemcpyAsync(c->bcp_offsets, program_offsets, (size_t)(n_programs + 1) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  NSM_CUDA(c, cudaMemcpyAsync(c->bcp_code, code, (size_t)n_code * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  if (n_consts) NSM_CUDA(c, cudaMemcpyAsync(c->bcp_consts, consts, (size_t)n_consts * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NSM_CUDA(c, cudaMemcpyAsync(c->bcp_of_entry, program_of_entry, (size_t)n_entries * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  NSM_CUDA(c, cudaMemsetAsync(c->bcp_slot_values, 0, (size_t)std::max(n_slots, 1) * sizeof(double), c->stream));
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  c->bcp_programs = n_programs, c->bcp_slots = n_slots, c->bcp_rows = 1, c->bcp_rows_cap = 1;
  c->bcp_entry_consts_needed = entry_consts_needed;
  return NSM_OK;
}

int
nsm_b200_set_bc_entry_constants(nsm_b200_ctx* c, int n_constants, int64_t n_entries, const double* values)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "set_bc_entry_constants: context not finalized");
  NSM_REQUIRE(c, c->bcp_programs > 0, "set_bc_entry_constants: no boundary-condition programs set");
  NSM_REQUIRE(c, n_constants >= 0 && n_entries == c->n_bc && (n_constants == 0 || values), "set_bc_entry_constants: bad arguments");
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  dev_release(c, c->bcp_entry_consts);
  c->bcp_n_entry_consts = 0;
  if (n_constants == 0) return NSM_OK;
  int rc = dev_alloc(c, &c->bcp_entry_consts, (int64_t)n_constants * n_entries);
  if (rc) return rc;
  NSM_CUDA(c, cudaMemcpyAsync(c->bcp_entry_consts, values, (size_t)n_constants * n_entries * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  c->bcp_n_entry_consts = n_constants;
  return NSM_OK;
}

int
nsm_b200_set_bc_slots_steps(nsm_b200_ctx* c, int n_rows, int n_slots, const double* slots)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "set_bc_slots_steps: context not finalized");
  NSM_REQUIRE(c, c->bcp_programs > 0, "set_bc_slots_steps: no boundary-condition programs set");
  NSM_REQUIRE(c, n_rows >= 1 && n_slots == c->bcp_slots && (n_slots == 0 || slots), "set_bc_slots_steps: bad arguments");
  if (n_slots == 0) {
    c->bcp_rows = n_rows;
    return NSM_OK;
  }
  if (n_rows > c->bcp_rows_cap) {
    NSM_CUDA(c, cudaStreamSynchronize(c->stream));
    dev_release(c, c->bcp_slot_values);
    int rc             = dev_alloc(c, &c->bcp_slot_values, (int64_t)n_rows * n_slots);
    if (rc) return rc;
    c->bcp_rows_cap = n_rows;
  }
  c->bcp_rows = n_rows;
  NSM_CUDA(c, cudaMemcpyAsync(c->bcp_slot_values, slots, (size_t)n_rows * n_slots * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  return NSM_OK;
}

int
nsm_b200_apply_kinematic_bc(nsm_b200_ctx* c, double time_current, double time_previous)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "apply_kinematic_bc: context not finalized");
  if (c->n_bc == 0 || c->n_nodes == 0) return NSM_OK;
  NvtxRange range("BC enforcement");
  int rc_p = enqueue_bc_programs(c, 0);
  if (rc_p) return rc_p;
  const double dt = time_current - time_previous;
  apply_bc_kernel<<<grid_for(c->n_nodes, 256), 256, 0, c->stream>>>(node_args(c), dt);
  c->launches++;
  NSM_CUDA(c, cudaGetLastError());
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  return NSM_OK;
}

int
nsm_b200_step(nsm_b200_ctx* c, int n_steps, double* time, double dt_user, int store_ipt_last)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "step: context not finalized");
  NSM_REQUIRE(c, time != nullptr && n_steps >= 0, "step: bad arguments");
  NSM_REQUIRE(c, c->bc_rows <= 1 || n_steps <= c->bc_rows, "step: more steps than per-step boundary-condition rows");
  NSM_REQUIRE(c, c->bcp_programs == 0 || c->bcp_rows <= 1 || n_steps <= c->bcp_rows, "step: more steps than per-step boundary-condition slot rows");
  const bool     ordered = c->assembly == NSM_ASSEMBLY_ORDERED;
  const bool     has_bc  = c->n_bc > 0;
  const int64_t  n       = c->n_nodes;
  const unsigned ngrid   = grid_for(n, 256);
  double         t       = *time;
  double         dt      = 0.0;
  const bool     gather_in_node_kernel = ordered && !c->comm.active();
  for (int s = 0; s < n_steps; ++s) {
    // explicit_time_integrator.cc:192-195: dt is re-derived from the accumulated time every step
    const double t_prev = t;
    t += dt_user;
    dt               = t - t_prev;
    const double hdt = 0.5 * dt;
    const bool   store =
        (store_ipt_last && s == n_steps - 1) || (c->flags_ & NSM_FLAG_STORE_IPT_EVERY_STEP);
    if (n > 0 && s == 0) {  // later first halves ride in the previous step's fused node pass
      NvtxRange range("Time Integration Scheme");
      int rc_p = enqueue_bc_programs(c, 0);
      if (rc_p) return rc_p;
      const NodeArgs na = node_args(c, 0);
      if (has_bc) {
        if (ordered)
          node_predict_kernel<true, false><<<ngrid, 256, 0, c->stream>>>(na, hdt, dt);
        else
          node_predict_kernel<true, true><<<ngrid, 256, 0, c->stream>>>(na, hdt, dt);
      } else {
        if (ordered)
          node_predict_kernel<false, false><<<ngrid, 256, 0, c->stream>>>(na, hdt, dt);
        else
          node_predict_kernel<false, true><<<ngrid, 256, 0, c->stream>>>(na, hdt, dt);
      }
      c->launches++;
      if (c->early_u_host) {  // nsm_b200_step_host: u is final from here on; send it home while the elements compute
        soa_to_aos_kernel<<<ngrid, 256, 0, c->stream>>>(n, c->u[0], c->u[1], c->u[2], c->staging_u, c->node_perm);
        c->launches++;
        NSM_CUDA(c, cudaEventRecord(c->ev_u_staged, c->stream));
        NSM_CUDA(c, cudaStreamWaitEvent(c->io_stream, c->ev_u_staged, 0));
        NSM_CUDA(c, cudaMemcpyAsync(c->early_u_host, c->staging_u, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->io_stream));
        c->early_u_host = nullptr;
      }
    }
    if (c->profiling) prof_event(c);
    roll_states(c);  // UpdateStates of the previous step (explicit_time_integrator.cc:277)
    int rc;
    {
    NvtxRange range("Force calculation");
    if (c->overlap) {
      // boundary-first: the groups that touch shared nodes run first, their nodal forces travel to the peers on
      // the exchange stream while the interior groups compute on this one
      if ((rc = enqueue_element_kernels(c, store, kSchedList))) return rc;
      if (ordered && c->comm.num_shared_nodes() > 0) {
        const int64_t ns = c->comm.num_shared_nodes();
        gather_shared_nodes_kernel<<<grid_for(ns, 256), 256, 0, c->stream>>>(ns, c->comm.shared_nodes_device(), c->ef, c->adj_off,
                                                                             c->adj_slot, c->f[0], c->f[1], c->f[2]);
        c->launches++;
      }
      NSM_CUDA(c, cudaEventRecord(c->ev_boundary, c->stream));
      NSM_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->ev_boundary, 0));
      if (c->comm.pack(c->comm_stream, c->f, 3, &c->launches)) return fail(c, NSM_ERR_COMM, "%s", c->comm.error());
      NSM_CUDA(c, cudaEventRecord(c->ev_packed, c->comm_stream));
      if ((rc = enqueue_element_kernels(c, store, kSchedSkipFlagged))) return rc;
    } else {
      if ((rc = enqueue_element_kernels(c, store))) return rc;
    }
    }
    if (c->profiling) prof_event(c);
    mark_states_for_roll(c);
    if ((rc = enqueue_contact(c))) return rc;  // explicit_time_integrator.cc:232-239, on the displacement of this step
    if (c->profiling) prof_event(c);
    if (n > 0) {
      const NodeArgs na = node_args(c, s + 1);  // boundary-condition magnitudes of the step the fused pass opens
      if (c->comm.active()) {
        NvtxRange range("Vector Reduction");
        if (c->overlap) NSM_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_packed, 0));  // the pack has read f before it changes
        if (ordered) {
          node_correct_kernel<true, false><<<ngrid, 256, 0, c->stream>>>(na, 0.0, 0);
          c->launches++;
        }
        if (c->overlap) {
          if (c->comm.finish(c->stream, c->f, 3, &c->launches)) return fail(c, NSM_ERR_COMM, "%s", c->comm.error());
        } else if (c->comm.reduce(c->stream, c->f, 3, &c->launches)) {
          return fail(c, NSM_ERR_COMM, "%s", c->comm.error());
        }
      }
      NvtxRange range("Time Integration Scheme");
      if (s + 1 < n_steps) {
        int rc_p = enqueue_bc_programs(c, s + 1);  // magnitudes of the step the fused pass opens
        if (rc_p) return rc_p;
        const double dt_next  = (t + dt_user) - t;
        const double hdt_next = 0.5 * dt_next;
        const int    variant  = (gather_in_node_kernel ? 4 : 0) | (c->has_fext ? 2 : 0) | (has_bc ? 1 : 0);
#define NSM_FUSED(O, FE, BC, Z) node_fused_kernel<O, FE, BC, Z><<<ngrid, 256, 0, c->stream>>>(na, hdt, hdt_next, dt_next)
        if (ordered) {  // element forces are gathered, nothing to clear
          switch (variant) {
            case 0: NSM_FUSED(false, false, false, false); break;
            case 1: NSM_FUSED(false, false, true, false); break;
            case 2: NSM_FUSED(false, true, false, false); break;
            case 3: NSM_FUSED(false, true, true, false); break;
            case 4: NSM_FUSED(true, false, false, false); break;
            case 5: NSM_FUSED(true, false, true, false); break;
            case 6: NSM_FUSED(true, true, false, false); break;
            default: NSM_FUSED(true, true, true, false); break;
          }
        } else {
          switch (variant) {
            case 0: NSM_FUSED(false, false, false, true); break;
            case 1: NSM_FUSED(false, false, true, true); break;
            case 2: NSM_FUSED(false, true, false, true); break;
            default: NSM_FUSED(false, true, true, true); break;
          }
        }
#undef NSM_FUSED
      } else if (gather_in_node_kernel) {
        if (c->has_fext)
          node_correct_kernel<true, true><<<ngrid, 256, 0, c->stream>>>(na, hdt, 1);
        else
          node_correct_kernel<true, false><<<ngrid, 256, 0, c->stream>>>(na, hdt, 1);
      } else {
        if (c->has_fext)
          node_correct_kernel<false, true><<<ngrid, 256, 0, c->stream>>>(na, hdt, 1);
        else
          node_correct_kernel<false, false><<<ngrid, 256, 0, c->stream>>>(na, hdt, 1);
      }
      c->launches++;
    }
    if (c->profiling) {
      prof_event(c);
      if (c->ev_used >= 4 * 256) prof_resolve(c);
    }
    NSM_CUDA(c, cudaGetLastError());
  }
  if (store_ipt_last && n_steps > 0 && has_bc && n > 0) {
    // output step: ApplyKinematicConditions once more before the data is read (explicit_time_integrator.cc:266-269)
    apply_bc_kernel<<<ngrid, 256, 0, c->stream>>>(node_args(c, n_steps - 1), dt);
    c->launches++;
  }
  *time = t;
  if (c->profiling) prof_resolve(c);
  return check_flags(c);
}

}  // extern "C"

namespace {

// Dependency ranges of the pipelined host step.  Nodes travel in C chunks of consecutive ids.  Per block, with lo(g) /
// hi(g) the lowest / highest node of group g: up_end[c] = number of leading groups whose nodes all lie below the end of
// chunk c (prefix maximum of hi), so those groups can run as soon as chunk c has been uploaded and integrated; a node
// chunk k is final once every group with a node below its end has run, i.e. after the element chunk done_after[k].
// Meshes numbered with locality (lattice order, Morton order) give a diagonal pipeline; a randomly numbered mesh
// degenerates gracefully into "everything after the last upload", which is the unpipelined schedule.
int
build_host_pipe(nsm_b200_ctx* c)
{
  auto&         P = c->pipe;
  const int64_t n = c->n_nodes;
  // automatic: 32 chunks from a million nodes on (64 M elements, same box, profiles/r02z_*: 8 / 16 / 24 / 32 / 48 chunks ->
  // 186 / 175 / 154 / 151 / 153 ms per step; the force seam 59 / 49 / 47 / 46 / 46 ms)
  int           C = P.requested_chunks >= 0 ? P.requested_chunks : (n >= (int64_t)1 << 20 ? 32 : 1);
  C               = (int)std::min<int64_t>(C, std::max<int64_t>(n, 1));
  P.built         = true;
  P.n_chunks      = 0;
  if (C < 2 || c->comm.active() || c->node_perm || n == 0) return NSM_OK;  // the plain schedule
  P.node_end.resize(C);
  for (int k = 0; k < C; ++k) P.node_end[k] = n * (k + 1) / C;
  std::vector<int> need(C, 0);  // over all blocks: the element chunk after which node chunk k is final
  for (auto& kv : c->blocks) {
    const Block&  b  = kv.second;
    const int64_t ng = groups_of(b.n_elem);
    std::vector<int> up(C, 0);
    if (ng > 0) {
      int *d_lo = nullptr, *d_hi = nullptr;
      NSM_CUDA(c, cudaMalloc((void**)&d_lo, (size_t)ng * sizeof(int)));
      NSM_CUDA(c, cudaMalloc((void**)&d_hi, (size_t)ng * sizeof(int)));
      group_node_range_kernel<<<grid_for(ng, 256), 256, 0, c->stream>>>(b.n_elem, b.conn_sched, d_lo, d_hi);
      c->launches++;
      std::vector<int> lo((size_t)ng), hi((size_t)ng);
      NSM_CUDA(c, cudaMemcpyAsync(lo.data(), d_lo, (size_t)ng * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      NSM_CUDA(c, cudaMemcpyAsync(hi.data(), d_hi, (size_t)ng * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      NSM_CUDA(c, cudaStreamSynchronize(c->stream));
      cudaFree(d_lo), cudaFree(d_hi);
      int64_t g = 0;
      int     k = 0;
      for (; k < C; ++k) {  // prefix maximum of hi: leading groups entirely below node_end[k]
        while (g < ng && hi[g] < P.node_end[k]) ++g;
        up[k] = (int)g;
      }
      up[C - 1] = (int)ng;
      // last group with a node below node_end[k] -> the element chunk that contains it
      std::vector<int64_t> last(C, -1);
      for (int64_t gg = 0; gg < ng; ++gg) {
        const int kk = (int)std::min<int64_t>(((int64_t)lo[gg] * C) / std::max<int64_t>(n, 1), C - 1);
        // lo[gg] lies in node chunk >= kk' where node_end[kk'] > lo[gg]; mark every chunk from there on
        int first = kk;
        while (first > 0 && P.node_end[first - 1] > lo[gg]) --first;
        while (P.node_end[first] <= lo[gg]) ++first;
        if (gg > last[first]) last[first] = gg;
      }
      int64_t run = -1;  // chunk k needs every group whose lowest node is below node_end[k]: running maximum over k' <= k
      for (int kk = 0; kk < C; ++kk) {
        run = std::max(run, last[kk]);
        if (run >= 0) {
          int cc = 0;
          while (up[cc] <= run) ++cc;  // the element chunk [up[cc-1], up[cc]) that holds group `run`
          need[kk] = std::max(need[kk], cc);
        }
      }
    }
    P.up_end[kv.first] = up;
  }
  for (int k = 1; k < C; ++k) need[k] = std::max(need[k], need[k - 1]);  // chunks finish in order (one download queue)
  for (int k = 0; k < C; ++k) need[k] = std::max(need[k], k);            // ... and never before their own upload
  P.done_after = need;
  for (double*& p : P.stage) {
    int rc = dev_alloc(c, &p, std::max<int64_t>(n * 3, 1));
    if (rc) return rc;
  }
  // the copy-side streams outrank the element stream: their CTAs are dispatched first whenever a slot is free
  int prio_least = 0, prio_greatest = 0;
  NSM_CUDA(c, cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  NSM_CUDA(c, cudaStreamCreateWithPriority(&P.up, cudaStreamNonBlocking, prio_greatest));
  NSM_CUDA(c, cudaStreamCreateWithPriority(&P.down, cudaStreamNonBlocking, prio_greatest));
  if (const char* e = getenv("NSM_B200_PIPE_RESERVE")) P.reserve_ctas = std::max(0, atoi(e));
  NSM_CUDA(c, cudaEventCreateWithFlags(&P.ev_bc, cudaEventDisableTiming));
  P.ev_up.resize(C), P.ev_elem.resize(C);
  for (int k = 0; k < C; ++k) {
    NSM_CUDA(c, cudaEventCreateWithFlags(&P.ev_up[k], cudaEventDisableTiming));
    NSM_CUDA(c, cudaEventCreateWithFlags(&P.ev_elem[k], cudaEventDisableTiming));
  }
  P.n_chunks = C;
  return NSM_OK;
}

// One explicit step on host-resident state, pipelined over node chunks (see build_host_pipe).  Same kernels and the same
// per-node / per-element operation order as nsm_b200_step: the results are bit-identical (ORDERED) to the plain path.
int
step_host_pipelined(nsm_b200_ctx* c, double* time, double dt_user, double* displacement, double* velocity, double* acceleration,
                    double* internal_force)
{
  auto&         P       = c->pipe;
  const int     C       = P.n_chunks;
  const bool    ordered = c->assembly == NSM_ASSEMBLY_ORDERED;
  const bool    has_bc  = c->n_bc > 0;
  struct Reserve  // the ranged element launches of this call leave CTA slots to the copy-side kernels
  {
    int& now;
    Reserve(int& n, int v) : now(n) { now = v; }
    ~Reserve() { now = 0; }
  } reserve(P.reserve_now, P.reserve_ctas);
  const bool    store   = (c->flags_ & NSM_FLAG_STORE_IPT_EVERY_STEP) != 0;
  const double  t_prev  = *time;
  const double  t       = t_prev + dt_user;
  const double  dt      = t - t_prev, hdt = 0.5 * dt;
  double* const host[4] = {displacement, velocity, acceleration, internal_force};
  double* const* const dev[4] = {c->u, c->v, c->a, c->f};
  if (store) {
    int rc = ensure_ipt(c);
    if (rc) return rc;
  }
  {
    NvtxRange range("BC enforcement");
    int rc = enqueue_bc_programs(c, 0);  // this step's magnitudes, before any chunk is integrated
    if (rc) return rc;
    NSM_CUDA(c, cudaEventRecord(P.ev_bc, c->stream));
    NSM_CUDA(c, cudaStreamWaitEvent(P.up, P.ev_bc, 0));
  }
  roll_states(c);
  int next_done = 0;  // node chunks [0, next_done) have been sent home
  for (int k = 0; k < C; ++k) {
    const int64_t  i0 = k ? P.node_end[k - 1] : 0, i1 = P.node_end[k], m = i1 - i0;
    const unsigned grid = grid_for(m, 256);
    {  // ---- up: u, v, a of the chunk, AoS -> SoA, first half of the step
      NvtxRange range("Time Integration Scheme");
      for (int f = 0; f < 3; ++f) {
        NSM_CUDA(c, cudaMemcpyAsync(P.stage[f] + 3 * i0, host[f] + 3 * i0, (size_t)m * 3 * sizeof(double), cudaMemcpyHostToDevice, P.up));
        aos_to_soa_range_kernel<<<grid, 256, 0, P.up>>>(i0, m, P.stage[f], dev[f][0], dev[f][1], dev[f][2]);
      }
      NodeArgs na   = node_args(c, 0);
      na.node_begin = i0, na.n_nodes = i1;
      if (has_bc) {
        if (ordered)
          node_predict_kernel<true, false><<<grid, 256, 0, P.up>>>(na, hdt, dt);
        else
          node_predict_kernel<true, true><<<grid, 256, 0, P.up>>>(na, hdt, dt);
      } else {
        if (ordered)
          node_predict_kernel<false, false><<<grid, 256, 0, P.up>>>(na, hdt, dt);
        else
          node_predict_kernel<false, true><<<grid, 256, 0, P.up>>>(na, hdt, dt);
      }
      c->launches += 4;
      NSM_CUDA(c, cudaEventRecord(P.ev_up[k], P.up));
    }
    {  // ---- elements whose nodes are all on the device now
      NvtxRange range("Force calculation");
      NSM_CUDA(c, cudaStreamWaitEvent(c->stream, P.ev_up[k], 0));
      for (auto& kv : c->blocks) {
        const std::vector<int>& up = P.up_end.at(kv.first);
        int rc = enqueue_element_range(c, kv.second, store, k ? up[k - 1] : 0, up[k], c->stream);
        if (rc) return rc;
      }
      NSM_CUDA(c, cudaEventRecord(P.ev_elem[k], c->stream));
    }
    {  // ---- down: the displacement of this chunk is final; so is everything of the chunks whose elements have all run
      NvtxRange range("Time Integration Scheme");
      NSM_CUDA(c, cudaStreamWaitEvent(P.down, P.ev_up[k], 0));
      soa_to_aos_range_kernel<<<grid, 256, 0, P.down>>>(i0, m, c->u[0], c->u[1], c->u[2], P.stage[0]);
      c->launches++;
      NSM_CUDA(c, cudaMemcpyAsync(host[0] + 3 * i0, P.stage[0] + 3 * i0, (size_t)m * 3 * sizeof(double), cudaMemcpyDeviceToHost, P.down));
      for (; next_done < C && P.done_after[next_done] <= k; ++next_done) {
        const int64_t  j0 = next_done ? P.node_end[next_done - 1] : 0, j1 = P.node_end[next_done], mm = j1 - j0;
        const unsigned g2 = grid_for(mm, 256);
        NSM_CUDA(c, cudaStreamWaitEvent(P.down, P.ev_elem[k], 0));
        NodeArgs na   = node_args(c, 0);
        na.node_begin = j0, na.n_nodes = j1;
        if (ordered) {
          if (c->has_fext)
            node_correct_kernel<true, true><<<g2, 256, 0, P.down>>>(na, hdt, 1);
          else
            node_correct_kernel<true, false><<<g2, 256, 0, P.down>>>(na, hdt, 1);
        } else {
          if (c->has_fext)
            node_correct_kernel<false, true><<<g2, 256, 0, P.down>>>(na, hdt, 1);
          else
            node_correct_kernel<false, false><<<g2, 256, 0, P.down>>>(na, hdt, 1);
        }
        // internal force, velocity, acceleration of the finished chunk (staging rows of v and a are free again: their
        // upload was consumed on the up stream before ev_up of the chunk)
        const int fields[3] = {3, 1, 2};
        for (int f : fields) {
          if (!host[f]) continue;  // (internal_force == NULL: the caller does not want the force back this step)
          soa_to_aos_range_kernel<<<g2, 256, 0, P.down>>>(j0, mm, dev[f][0], dev[f][1], dev[f][2], P.stage[f]);
          NSM_CUDA(c, cudaMemcpyAsync(host[f] + 3 * j0, P.stage[f] + 3 * j0, (size_t)mm * 3 * sizeof(double), cudaMemcpyDeviceToHost, P.down));
        }
        c->launches += 4;
      }
    }
    NSM_CUDA(c, cudaGetLastError());
  }
  mark_states_for_roll(c);
  *time = t;
  NSM_CUDA(c, cudaStreamSynchronize(P.down));
  NSM_CUDA(c, cudaStreamSynchronize(P.up));
  return check_flags(c);
}

// ModelData::ComputeInternalForce on host views (displacement in, internal force out), pipelined over the same node
// chunks: chunk k of u travels up while the elements below it run, and the force of every node chunk whose elements have
// all run travels down while later chunks are still on their way up.  Same element kernels and (ORDERED) the same nodal
// summation as nsm_b200_internal_force, hence the same bits.
int
internal_force_host_pipelined(nsm_b200_ctx* c, const double* displacement, double* internal_force, bool store)
{
  auto&      P       = c->pipe;
  const int  C       = P.n_chunks;
  const bool ordered = c->assembly == NSM_ASSEMBLY_ORDERED;
  if (store) {
    int rc = ensure_ipt(c);
    if (rc) return rc;
  }
  roll_states(c);
  NvtxRange range("Force calculation");
  // whatever the caller queued on the context stream (uploads, an earlier step) comes first
  NSM_CUDA(c, cudaEventRecord(P.ev_bc, c->stream));
  NSM_CUDA(c, cudaStreamWaitEvent(P.up, P.ev_bc, 0));
  NSM_CUDA(c, cudaStreamWaitEvent(P.down, P.ev_bc, 0));
  int next_done = 0;
  for (int k = 0; k < C; ++k) {
    const int64_t  i0 = k ? P.node_end[k - 1] : 0, i1 = P.node_end[k], m = i1 - i0;
    const unsigned grid = grid_for(m, 256);
    NSM_CUDA(c, cudaMemcpyAsync(P.stage[0] + 3 * i0, displacement + 3 * i0, (size_t)m * 3 * sizeof(double), cudaMemcpyHostToDevice, P.up));
    aos_to_soa_range_kernel<<<grid, 256, 0, P.up>>>(i0, m, P.stage[0], c->u[0], c->u[1], c->u[2]);
    c->launches++;
    if (!ordered)  // the atomic assembly adds into a cleared force
      for (int i = 0; i < 3; ++i) NSM_CUDA(c, cudaMemsetAsync(c->f[i] + i0, 0, (size_t)m * sizeof(double), P.up));
    NSM_CUDA(c, cudaEventRecord(P.ev_up[k], P.up));
    NSM_CUDA(c, cudaStreamWaitEvent(c->stream, P.ev_up[k], 0));
    for (auto& kv : c->blocks) {
      const std::vector<int>& up = P.up_end.at(kv.first);
      int rc = enqueue_element_range(c, kv.second, store, k ? up[k - 1] : 0, up[k], c->stream);
      if (rc) return rc;
    }
    NSM_CUDA(c, cudaEventRecord(P.ev_elem[k], c->stream));
    for (; next_done < C && P.done_after[next_done] <= k; ++next_done) {
      const int64_t  j0 = next_done ? P.node_end[next_done - 1] : 0, j1 = P.node_end[next_done], mm = j1 - j0;
      const unsigned g2 = grid_for(mm, 256);
      NSM_CUDA(c, cudaStreamWaitEvent(P.down, P.ev_elem[k], 0));
      if (ordered) {
        NodeArgs na   = node_args(c, 0);
        na.node_begin = j0, na.n_nodes = j1;
        node_correct_kernel<true, false><<<g2, 256, 0, P.down>>>(na, 0.0, 0);
        c->launches++;
      }
      soa_to_aos_range_kernel<<<g2, 256, 0, P.down>>>(j0, mm, c->f[0], c->f[1], c->f[2], P.stage[3]);
      c->launches++;
      NSM_CUDA(c, cudaMemcpyAsync(internal_force + 3 * j0, P.stage[3] + 3 * j0, (size_t)mm * 3 * sizeof(double), cudaMemcpyDeviceToHost, P.down));
    }
    NSM_CUDA(c, cudaGetLastError());
  }
  // later work on the context stream sees the complete force field
  NSM_CUDA(c, cudaEventRecord(P.ev_bc, P.down));
  NSM_CUDA(c, cudaStreamWaitEvent(c->stream, P.ev_bc, 0));
  NSM_CUDA(c, cudaStreamSynchronize(P.down));
  NSM_CUDA(c, cudaStreamSynchronize(P.up));
  return check_flags(c);
}

}  // namespace

extern "C" {

int
nsm_b200_internal_force_host(nsm_b200_ctx* c, const double* displacement, double* internal_force, int store_ipt)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "internal_force_host: context not finalized");
  NSM_REQUIRE(c, displacement && internal_force, "internal_force_host: null argument");
  const bool store = store_ipt != 0 || (c->flags_ & NSM_FLAG_STORE_IPT_EVERY_STEP);
  if (!c->pipe.built) {
    int rc = build_host_pipe(c);
    if (rc) return rc;
  }
  if (c->pipe.n_chunks >= 2 && !c->comm.active()) return internal_force_host_pipelined(c, displacement, internal_force, store);
  int rc = upload_field(c, NSM_FIELD_DISPLACEMENT, displacement, false);
  if (rc) return rc;
  rc = enqueue_internal_force(c, store);
  if (rc) return rc;
  rc = download_field(c, NSM_FIELD_INTERNAL_FORCE, internal_force, false);
  if (rc) return rc;
  return check_flags(c);
}

int
nsm_b200_set_host_step_chunks(nsm_b200_ctx* c, int n_chunks)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, n_chunks >= -1 && n_chunks <= 4096, "set_host_step_chunks: chunk count out of range");
  NSM_REQUIRE(c, !c->pipe.built, "set_host_step_chunks: call before the first nsm_b200_step_host");
  c->pipe.requested_chunks = n_chunks;
  return NSM_OK;
}

int
nsm_b200_step_host(nsm_b200_ctx* c, double* time, double dt_user, double* displacement, double* velocity, double* acceleration,
                   double* internal_force)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "step_host: context not finalized");
  NSM_REQUIRE(c, time && displacement && velocity && acceleration, "step_host: null argument");
  if (!c->pipe.built) {
    int rc = build_host_pipe(c);
    if (rc) return rc;
  }
  if (c->pipe.n_chunks >= 2 && c->bc_rows <= 1 && !c->comm.active() && !c->contact.active)  // (contact needs the whole displacement)
    return step_host_pipelined(c, time, dt_user, displacement, velocity, acceleration, internal_force);
  const int64_t n = c->n_nodes;
  if (!c->io_stream) {
    int rc = dev_alloc(c, &c->staging_u, std::max<int64_t>(n * 3, 1));
    if (rc) return rc;
    NSM_CUDA(c, cudaStreamCreateWithFlags(&c->io_stream, cudaStreamNonBlocking));
    NSM_CUDA(c, cudaEventCreateWithFlags(&c->ev_u_staged, cudaEventDisableTiming));
  }
  int rc;
  if ((rc = upload_field(c, NSM_FIELD_DISPLACEMENT, displacement, false))) return rc;
  if ((rc = upload_field(c, NSM_FIELD_VELOCITY, velocity, false))) return rc;
  if ((rc = upload_field(c, NSM_FIELD_ACCELERATION, acceleration, false))) return rc;
  c->early_u_host = n > 0 ? displacement : nullptr;
  rc              = nsm_b200_step(c, 1, time, dt_user, 0);
  c->early_u_host = nullptr;
  if (rc) {
    cudaStreamSynchronize(c->io_stream);
    return rc;
  }
  if (internal_force && (rc = download_field(c, NSM_FIELD_INTERNAL_FORCE, internal_force, false))) return rc;
  if ((rc = download_field(c, NSM_FIELD_VELOCITY, velocity, false))) return rc;
  if ((rc = download_field(c, NSM_FIELD_ACCELERATION, acceleration, false))) return rc;
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  NSM_CUDA(c, cudaStreamSynchronize(c->io_stream));
  return NSM_OK;
}

int
nsm_b200_get_element_data(nsm_b200_ctx* c, int block_id, double* out)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "get_element_data: context not finalized");
  auto it = c->blocks.find(block_id);
  NSM_REQUIRE(c, it != c->blocks.end(), "get_element_data: unknown block id");
  int rc = ensure_ipt(c);
  if (rc) return rc;
  const Block& b = it->second;
  int           per_element = 0;
  const double* rec         = block_records(c, b, &per_element);
  NSM_CUDA(c, cudaMemcpyAsync(out, rec, (size_t)b.n_elem * per_element * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  return NSM_OK;
}

int
nsm_b200_element_data_stride(const nsm_b200_ctx* c, int block_id)
{
  if (!c) return -1;
  auto it = c->blocks.find(block_id);
  return it == c->blocks.end() ? -1 : 15 + it->second.n_state;
}

int
nsm_b200_update_states(nsm_b200_ctx* c)
{
  NSM_REQUIRE(c, c != nullptr && c->finalized, "update_states: context not finalized");
  mark_states_for_roll(c);
  return NSM_OK;
}

int
nsm_b200_set_element_data(nsm_b200_ctx* c, int block_id, int previous, const double* in)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "set_element_data: context not finalized");
  auto it = c->blocks.find(block_id);
  NSM_REQUIRE(c, it != c->blocks.end() && in, "set_element_data: unknown block id / null data");
  Block& b = it->second;
  NSM_REQUIRE(c, b.n_state > 0 || !previous, "set_element_data: the block's material carries no state (no N records are kept)");
  roll_states(c);  // a pending UpdateStates is applied first, so that `previous` names the records the next evaluation reads
  double* dst = nullptr;
  if (b.n_state) {
    dst = b.rec[previous ? (b.cur ^ 1) : b.cur];
  } else {
    int rc = ensure_ipt(c);
    if (rc) return rc;
    dst = c->ipt + b.elem_base * 120;
  }
  NSM_CUDA(c, cudaMemcpyAsync(dst, in, (size_t)b.n_elem * 8 * (15 + b.n_state) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  return NSM_OK;
}

int
nsm_b200_get_element_data_previous(nsm_b200_ctx* c, int block_id, double* out)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "get_element_data_previous: context not finalized");
  auto it = c->blocks.find(block_id);
  NSM_REQUIRE(c, it != c->blocks.end(), "get_element_data_previous: unknown block id");
  const Block& b = it->second;
  NSM_REQUIRE(c, b.n_state > 0, "get_element_data_previous: the block's material carries no state (no N records are kept)");
  NSM_CUDA(c, cudaMemcpyAsync(out, b.rec[b.cur ^ 1], (size_t)b.n_elem * 8 * (15 + b.n_state) * sizeof(double), cudaMemcpyDeviceToHost,
                              c->stream));
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  return NSM_OK;
}

int
nsm_b200_derived_element_data(nsm_b200_ctx* c, int block_id, double* out)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "derived_element_data: context not finalized");
  auto it = c->blocks.find(block_id);
  NSM_REQUIRE(c, it != c->blocks.end(), "derived_element_data: unknown block id");
  NvtxRange nvtx_range("Output");
  int rc = ensure_ipt(c);
  if (rc) return rc;
  const Block& b = it->second;
  if (b.n_elem == 0) return NSM_OK;
  double*       d           = nullptr;
  int           per_element = 0;
  const double* rec         = block_records(c, b, &per_element);
  const int     record      = per_element / 8;
  NSM_CUDA(c, cudaMalloc((void**)&d, (size_t)b.n_elem * (1 + record) * sizeof(double)));
  derived_kernel<<<grid_for(b.n_elem, 128), 128, 0, c->stream>>>(b.n_elem, b.conn, c->X[0], c->X[1], c->X[2], c->u[0],
                                                                 c->u[1], c->u[2], rec, d, record);
  c->launches++;
  NSM_CUDA(c, cudaGetLastError());
  NSM_CUDA(c, cudaMemcpyAsync(out, d, (size_t)b.n_elem * (1 + record) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaFree(d);
  return NSM_OK;
}

int
nsm_b200_get_element_components(nsm_b200_ctx* c, int block_id, int n_components, const int32_t* offsets, double* out)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "get_element_components: context not finalized");
  auto it = c->blocks.find(block_id);
  NSM_REQUIRE(c, it != c->blocks.end(), "get_element_components: unknown block id");
  NvtxRange nvtx_range("Output");
  NSM_REQUIRE(c, n_components >= 0 && (n_components == 0 || (offsets && out)), "get_element_components: bad arguments");
  const Block& b = it->second;
  for (int k = 0; k < n_components; ++k)
    if (offsets[k] < 0 || offsets[k] >= 8 * (15 + b.n_state))
      return fail(c, NSM_ERR_ARG, "get_element_components: offset %d out of 0..%d", offsets[k], 8 * (15 + b.n_state) - 1);
  int rc = ensure_ipt(c);
  if (rc) return rc;
  int           per_element = 0;
  const double* rec         = block_records(c, b, &per_element);
  if (b.n_elem == 0 || n_components == 0) return NSM_OK;
  // bounded staging: ranges of elements, [n_components][range] on the device, one strided copy per range
  const int64_t range = std::min<int64_t>(b.n_elem, std::max<int64_t>(((int64_t)32 << 20) / n_components, 1024));
  int*          d_off = nullptr;
  double*       d_out = nullptr;
  NSM_CUDA(c, cudaMalloc((void**)&d_off, (size_t)n_components * sizeof(int)));
  cudaError_t e = cudaMalloc((void**)&d_out, (size_t)range * n_components * sizeof(double));
  if (e != cudaSuccess) {
    cudaFree(d_off);
    return fail(c, NSM_ERR_CUDA, "get_element_components: %s", cudaGetErrorString(e));
  }
  cudaMemcpyAsync(d_off, offsets, (size_t)n_components * sizeof(int), cudaMemcpyHostToDevice, c->stream);
  for (int64_t e0 = 0; e0 < b.n_elem && e == cudaSuccess; e0 += range) {
    const int64_t n = std::min(range, b.n_elem - e0);
    select_ipt_components_kernel<<<grid_for(n * n_components, 256), 256, 0, c->stream>>>(e0, n, n_components, d_off,
                                                                                         rec, d_out, per_element);
    c->launches++;
    e = cudaMemcpy2DAsync(out + e0, (size_t)b.n_elem * sizeof(double), d_out, (size_t)n * sizeof(double), (size_t)n * sizeof(double),
                          (size_t)n_components, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);  // d_out is reused by the next range
  }
  cudaFree(d_off);
  cudaFree(d_out);
  if (e != cudaSuccess) return fail(c, NSM_ERR_CUDA, "get_element_components: %s", cudaGetErrorString(e));
  return NSM_OK;
}

int
nsm_b200_get_element_data_subset(nsm_b200_ctx* c, int block_id, int64_t n, const int64_t* elements, double* out)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "get_element_data_subset: context not finalized");
  auto it = c->blocks.find(block_id);
  NSM_REQUIRE(c, it != c->blocks.end(), "get_element_data_subset: unknown block id");
  NSM_REQUIRE(c, n >= 0 && (n == 0 || (elements && out)), "get_element_data_subset: bad arguments");
  const Block& b = it->second;
  for (int64_t i = 0; i < n; ++i)
    if (elements[i] < 0 || elements[i] >= b.n_elem)
      return fail(c, NSM_ERR_ARG, "get_element_data_subset: element %lld outside [0, %lld)", (long long)elements[i], (long long)b.n_elem);
  int rc = ensure_ipt(c);
  if (rc) return rc;
  if (n == 0) return NSM_OK;
  int64_t*      d_el        = nullptr;
  double*       d_out       = nullptr;
  int           per_element = 0;
  const double* rec         = block_records(c, b, &per_element);
  NSM_CUDA(c, cudaMalloc((void**)&d_el, (size_t)n * sizeof(int64_t)));
  cudaError_t e = cudaMalloc((void**)&d_out, (size_t)n * per_element * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_el, elements, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    gather_ipt_records_kernel<<<grid_for(n * per_element, 256), 256, 0, c->stream>>>(n, d_el, rec, d_out, per_element);
    c->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)n * per_element * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d_el);
  cudaFree(d_out);
  if (e != cudaSuccess) return fail(c, NSM_ERR_CUDA, "get_element_data_subset: %s", cudaGetErrorString(e));
  return NSM_OK;
}

// ---- peer exchange ---------------------------------------------------------------------------------
int
nsm_b200_comm_init(nsm_b200_ctx* c, int rank, int world_size, int n_peers, const int32_t* peer_ranks,
                   const int64_t* pair_offsets, const int32_t* pair_local_nodes)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "comm_init: context not finalized");
  NSM_REQUIRE(c, n_peers >= 0 && (n_peers == 0 || (peer_ranks && pair_offsets)), "comm_init: bad arguments");
  static const int64_t zero = 0;
  if (n_peers == 0) pair_offsets = &zero;
  for (int64_t k = 0; k < pair_offsets[n_peers]; ++k)
    if (pair_local_nodes[k] < 0 || pair_local_nodes[k] >= c->n_nodes)
      return fail(c, NSM_ERR_ARG, "comm_init: shared node id out of range");
  std::vector<int32_t> internal(pair_local_nodes, pair_local_nodes + pair_offsets[n_peers]);
  if (!c->node_perm_host.empty())
    for (int32_t& nd : internal) nd = c->node_perm_host[nd];
  if (c->comm.init(c->device, rank, world_size, n_peers, peer_ranks, pair_offsets, internal.data()))
    return fail(c, NSM_ERR_COMM, "%s", c->comm.error());
  return NSM_OK;
}

int
nsm_b200_comm_export(nsm_b200_ctx* c, unsigned char handle[NSM_COMM_HANDLE_BYTES])
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "comm_export: context not finalized");
  if (c->comm.export_handle(handle)) return fail(c, NSM_ERR_COMM, "%s", c->comm.error());
  return NSM_OK;
}

int
nsm_b200_comm_attach(nsm_b200_ctx* c, int peer_rank, const unsigned char handle[NSM_COMM_HANDLE_BYTES])
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "comm_attach: context not finalized");
  if (c->comm.attach(peer_rank, handle)) return fail(c, NSM_ERR_COMM, "%s", c->comm.error());
  return NSM_OK;
}

int
nsm_b200_comm_ready(nsm_b200_ctx* c)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "comm_ready: context not finalized");
  if (c->comm.ready(c->stream)) return fail(c, NSM_ERR_COMM, "%s", c->comm.error());
  // boundary-first element schedule: flag and list the groups that touch a shared node
  const int64_t ns = c->comm.num_shared_nodes();
  if (ns > 0 && c->n_nodes > 0) {
    unsigned char*      node_flag = nullptr;
    unsigned*           d_count   = nullptr;
    NSM_CUDA(c, cudaMalloc((void**)&node_flag, (size_t)c->n_nodes));
    NSM_CUDA(c, cudaMalloc((void**)&d_count, sizeof(unsigned)));
    NSM_CUDA(c, cudaMemsetAsync(node_flag, 0, (size_t)c->n_nodes, c->stream));
    mark_shared_nodes_kernel<<<grid_for(ns, 256), 256, 0, c->stream>>>(ns, c->comm.shared_nodes_device(), node_flag);
    c->launches++;
    for (auto& kv : c->blocks) {
      Block& b = kv.second;
      if (b.n_elem == 0) continue;
      const int64_t ng     = groups_of(b.n_elem);
      const int64_t words = (ng + 31) / 32;
      int           rc;
      if ((rc = dev_alloc(c, &b.group_bits, words))) return rc;
      if ((rc = dev_alloc(c, &b.group_list, ng))) return rc;
      NSM_CUDA(c, cudaMemsetAsync(b.group_bits, 0, (size_t)words * sizeof(unsigned), c->stream));
      NSM_CUDA(c, cudaMemsetAsync(d_count, 0, sizeof(unsigned), c->stream));
      flag_groups_kernel<<<grid_for(ng, 256), 256, 0, c->stream>>>(b.n_elem, b.conn_sched, node_flag, b.group_bits, b.group_list, d_count);
      c->launches++;
      unsigned h = 0;
      NSM_CUDA(c, cudaMemcpyAsync(&h, d_count, sizeof h, cudaMemcpyDeviceToHost, c->stream));
      NSM_CUDA(c, cudaStreamSynchronize(c->stream));
      b.n_list = (int)h;
    }
    NSM_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(node_flag);
    cudaFree(d_count);
    NSM_CUDA(c, cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
    NSM_CUDA(c, cudaEventCreateWithFlags(&c->ev_boundary, cudaEventDisableTiming));
    NSM_CUDA(c, cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming));
    c->overlap = true;
  }
  return NSM_OK;

}

int
nsm_b200_comm_set_host_barrier(nsm_b200_ctx* c, void (*barrier)(void*), void* arg)
{
  NSM_ENTER(c);
  if (c->comm.set_host_barrier(barrier, arg)) return fail(c, NSM_ERR_COMM, "%s", c->comm.error());
  return NSM_OK;
}

int
nsm_b200_comm_set_timeout(nsm_b200_ctx* c, double seconds)
{
  NSM_REQUIRE(c, c != nullptr, "null context");
  NSM_REQUIRE(c, seconds > 0.0, "comm_set_timeout: the timeout must be positive");
  c->comm.set_timeout_seconds(seconds);
  return NSM_OK;
}

// ---- measurement ------------------------------------------------------------------------------------
int
nsm_b200_set_contact(nsm_b200_ctx* c, double penalty, int64_t n_faces, const int32_t* face_nodes, const double* face_len, int64_t n_cn,
                     const int32_t* cn_ids, const double* cn_len)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized, "set_contact: context not finalized");
  NSM_REQUIRE(c, n_faces >= 0 && n_cn >= 0 && n_faces < ((int64_t)1 << 28) && n_cn < ((int64_t)1 << 31), "set_contact: entity count out of range");
  NSM_REQUIRE(c, (n_faces == 0 || (face_nodes && face_len)) && (n_cn == 0 || (cn_ids && cn_len)), "set_contact: null argument");
  NSM_REQUIRE(c, !c->comm.active(), "set_contact: contexts with a peer exchange are not supported (contact across partitions)");
  // every argument is checked before the entities in place are touched: a refused call leaves the context as it was
  // (ComputeContactForce throws on a non-positive penalty, src/nimble_contact_manager.cc:398-400)
  NSM_REQUIRE(c, (n_faces == 0 && n_cn == 0) || penalty > 0.0, "Error in ComputeContactForce(), invalid penalty_parameter.");
  std::vector<int> quads((size_t)n_faces * 4), nodes((size_t)n_cn);
  for (int64_t i = 0; i < n_faces * 4; ++i) {
    NSM_REQUIRE(c, face_nodes[i] >= 0 && face_nodes[i] < c->n_nodes, "set_contact: face node id out of range");
    quads[(size_t)i] = c->node_perm_host.empty() ? face_nodes[i] : c->node_perm_host[face_nodes[i]];
  }
  for (int64_t i = 0; i < n_cn; ++i) {
    NSM_REQUIRE(c, cn_ids[i] >= 0 && cn_ids[i] < c->n_nodes, "set_contact: contact node id out of range");
    nodes[(size_t)i] = c->node_perm_host.empty() ? cn_ids[i] : c->node_perm_host[cn_ids[i]];
  }
  auto& k = c->contact;
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  dev_release(c, k.quad), dev_release(c, k.quad_len), dev_release(c, k.sec_node), dev_release(c, k.sec_len), dev_release(c, k.quad_xyz);
  dev_release(c, k.tri_box), dev_release(c, k.bin), dev_release(c, k.head), dev_release(c, k.status), dev_release(c, k.near_list);
  dev_release(c, k.surf_node);
  dev_release(c, k.contrib_key), dev_release(c, k.key_sorted), dev_release(c, k.contrib_target), dev_release(c, k.contrib_val);
  dev_release(c, k.iota), dev_release(c, k.order1), dev_release(c, k.order2), dev_release(c, k.target1), dev_release(c, k.target2);
  dev_release(c, k.sort_tmp);
  k.contrib_cap = 0, k.sort_tmp_bytes = 0, k.ordered_overflow_pairs = 0;
  k.n_surf = 0;
  k.active = false, k.n_quads = k.n_sec = 0;
  if (n_faces == 0 && n_cn == 0) {
    if (c->fc[0])
      for (int i = 0; i < 3; ++i) NSM_CUDA(c, cudaMemsetAsync(c->fc[i], 0, (size_t)std::max<int64_t>(c->n_nodes, 1) * sizeof(double), c->stream));
    return NSM_OK;
  }
  // nodes of the contact sub-model, each once (the update kernel clears their contact force)
  std::vector<int> surf(quads);
  surf.insert(surf.end(), nodes.begin(), nodes.end());
  std::sort(surf.begin(), surf.end());
  surf.erase(std::unique(surf.begin(), surf.end()), surf.end());
  int rc;
  const int64_t n_tri = 4 * n_faces;
  unsigned      table = 1024;
  while ((int64_t)table < 2 * n_faces) table <<= 1;
  if ((rc = dev_alloc(c, &k.quad, n_faces * 4)) || (rc = dev_alloc(c, &k.quad_len, n_faces)) || (rc = dev_alloc(c, &k.sec_node, n_cn)) ||
      (rc = dev_alloc(c, &k.sec_len, n_cn)) || (rc = dev_alloc(c, &k.quad_xyz, n_faces * 15)) || (rc = dev_alloc(c, &k.tri_box, n_tri * 6)) ||
      (rc = dev_alloc(c, &k.bin, n_faces)) || (rc = dev_alloc(c, &k.head, (int64_t)table)) ||
      (rc = dev_alloc(c, &k.status, n_tri + n_cn)) || (rc = dev_alloc(c, &k.near_list, n_cn)) ||
      (rc = dev_alloc(c, &k.surf_node, (int64_t)surf.size())))
    return rc;
  if (!k.red && ((rc = dev_alloc(c, &k.red, 16)) || (rc = dev_alloc(c, &k.counters, 8)))) return rc;
  if (c->assembly == NSM_ASSEMBLY_ORDERED && n_cn > 0 && n_faces > 0) {
    // room for four accepted pairs per contact node (a node on a facet vertex meets the facets around it) + slack; pairs
    // beyond it are added atomically and counted (nsm_b200_contact_stats)
    const long long cap = std::min<long long>(7 * (4 * n_cn + 4096), 2147483647LL / 7 * 7);
    if ((rc = dev_alloc(c, &k.contrib_key, cap)) || (rc = dev_alloc(c, &k.key_sorted, cap)) || (rc = dev_alloc(c, &k.contrib_target, cap)) ||
        (rc = dev_alloc(c, &k.contrib_val, 3 * cap)) || (rc = dev_alloc(c, &k.iota, cap)) || (rc = dev_alloc(c, &k.order1, cap)) ||
        (rc = dev_alloc(c, &k.order2, cap)) || (rc = dev_alloc(c, &k.target1, cap)) || (rc = dev_alloc(c, &k.target2, cap)))
      return rc;
    size_t t64 = 0, t32 = 0;
    NSM_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, t64, k.contrib_key, k.key_sorted, k.iota, k.order1, (int)cap, 0, 64, c->stream));
    NSM_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, t32, k.target1, k.target2, k.order1, k.order2, (int)cap, 0, 32, c->stream));
    k.sort_tmp_bytes = std::max(t64, t32);
    if ((rc = dev_alloc(c, &k.sort_tmp, (int64_t)k.sort_tmp_bytes))) return rc;
    contact_iota_kernel<<<grid_for(cap, 256), 256, 0, c->stream>>>(cap, k.iota);
    c->launches++;
    k.contrib_cap = cap;
  }
  for (int i = 0; i < 3; ++i)
    if (!c->fc[i] && (rc = dev_alloc(c, &c->fc[i], c->n_nodes))) return rc;
  for (int i = 0; i < 3; ++i) NSM_CUDA(c, cudaMemsetAsync(c->fc[i], 0, (size_t)std::max<int64_t>(c->n_nodes, 1) * sizeof(double), c->stream));
  const unsigned red0[16] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u};
  NSM_CUDA(c, cudaMemcpyAsync(k.red, red0, sizeof red0, cudaMemcpyHostToDevice, c->stream));
  NSM_CUDA(c, cudaMemsetAsync(k.counters, 0, 8 * sizeof(unsigned long long), c->stream));
  NSM_CUDA(c, cudaMemsetAsync(k.status, 0, (size_t)std::max<int64_t>(n_tri + n_cn, 1), c->stream));
  if (n_faces) {
    NSM_CUDA(c, cudaMemcpyAsync(k.quad, quads.data(), quads.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    NSM_CUDA(c, cudaMemcpyAsync(k.quad_len, face_len, (size_t)n_faces * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  if (n_cn) {
    NSM_CUDA(c, cudaMemcpyAsync(k.sec_node, nodes.data(), nodes.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    NSM_CUDA(c, cudaMemcpyAsync(k.sec_len, cn_len, (size_t)n_cn * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  if (!surf.empty()) NSM_CUDA(c, cudaMemcpyAsync(k.surf_node, surf.data(), surf.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  k.n_surf  = (int64_t)surf.size();
  k.penalty = penalty, k.n_quads = n_faces, k.n_sec = n_cn, k.table_mask = table - 1, k.parity = 0;
  k.active  = true;
  return NSM_OK;
}

int
nsm_b200_contact_force(nsm_b200_ctx* c)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized && c->contact.active, "contact_force: no contact entities on this context (nsm_b200_set_contact)");
  int rc = enqueue_contact(c);
  if (rc) return rc;
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  return NSM_OK;
}

int
nsm_b200_contact_force_host(nsm_b200_ctx* c, const double* displacement, double* contact_force)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, c->finalized && c->contact.active, "contact_force_host: no contact entities on this context (nsm_b200_set_contact)");
  NSM_REQUIRE(c, contact_force != nullptr, "contact_force_host: null argument");
  int rc;
  if (displacement && (rc = upload_field(c, NSM_FIELD_DISPLACEMENT, displacement, false))) return rc;
  if ((rc = enqueue_contact(c))) return rc;
  return download_field(c, NSM_FIELD_CONTACT_FORCE, contact_force, true);
}

int
nsm_b200_contact_stats(nsm_b200_ctx* c, int64_t stats[5])
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, stats != nullptr, "contact_stats: null argument");
  stats[0] = stats[1] = stats[2] = stats[3] = stats[4] = 0;
  auto& k = c->contact;
  if (!k.active) return NSM_OK;
  const int64_t n_tri = 4 * k.n_quads;
  NSM_CUDA(c, cudaMemsetAsync(k.counters + 2, 0, 2 * sizeof(unsigned long long), c->stream));
  contact_count_kernel<<<grid_for(std::max<int64_t>(n_tri + k.n_sec, 1), 256), 256, 0, c->stream>>>(n_tri, k.n_sec, k.status, k.counters + 2);
  c->launches++;
  unsigned long long h[4];
  NSM_CUDA(c, cudaMemcpyAsync(h, k.counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 4; ++i) stats[i] = (int64_t)h[i];
  stats[4] = (int64_t)k.ordered_overflow_pairs;
  return NSM_OK;
}

int
nsm_b200_contact_status(nsm_b200_ctx* c, unsigned char* face_status, unsigned char* node_status)
{
  NSM_ENTER(c);
  auto& k = c->contact;
  NSM_REQUIRE(c, c->finalized && k.active, "contact_status: no contact entities on this context (nsm_b200_set_contact)");
  const int64_t n_tri = 4 * k.n_quads;
  if (face_status && n_tri > 0) NSM_CUDA(c, cudaMemcpyAsync(face_status, k.status, (size_t)n_tri, cudaMemcpyDeviceToHost, c->stream));
  if (node_status && k.n_sec > 0) NSM_CUDA(c, cudaMemcpyAsync(node_status, k.status + n_tri, (size_t)k.n_sec, cudaMemcpyDeviceToHost, c->stream));
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  return NSM_OK;
}

int
nsm_b200_timer_start(nsm_b200_ctx* c)
{
  NSM_ENTER(c);
  NSM_CUDA(c, cudaStreamSynchronize(c->stream));
  NSM_CUDA(c, cudaEventRecord(c->ev_start, c->stream));
  return NSM_OK;
}

int
nsm_b200_timer_stop(nsm_b200_ctx* c, float* ms)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, ms != nullptr, "timer_stop: bad arguments");
  NSM_CUDA(c, cudaEventRecord(c->ev_stop, c->stream));
  NSM_CUDA(c, cudaEventSynchronize(c->ev_stop));
  NSM_CUDA(c, cudaEventElapsedTime(ms, c->ev_start, c->ev_stop));
  return NSM_OK;
}

int64_t
nsm_b200_launch_count(const nsm_b200_ctx* c)
{
  return c ? c->launches : -1;
}

int
nsm_b200_profile(nsm_b200_ctx* c, int enable)
{
  NSM_REQUIRE(c, c != nullptr, "null context");
  c->profiling    = enable != 0;
  c->prof_elem_ms = c->prof_node_ms = c->prof_contact_ms = 0;
  c->prof_steps   = 0;
  c->ev_used      = 0;
  return NSM_OK;
}

int
nsm_b200_profile_read(nsm_b200_ctx* c, double* elem_ms, double* node_ms, int64_t* n)
{
  NSM_ENTER(c);
  prof_resolve(c);
  const double k = c->prof_steps ? 1.0 / (double)c->prof_steps : 0.0;
  if (elem_ms) *elem_ms = c->prof_elem_ms * k;
  if (node_ms) *node_ms = c->prof_node_ms * k;
  if (n) *n = c->prof_steps;
  return NSM_OK;
}

int
nsm_b200_profile_read_contact(nsm_b200_ctx* c, double* contact_ms)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, contact_ms != nullptr, "profile_read_contact: null argument");
  prof_resolve(c);
  *contact_ms = c->prof_steps ? c->prof_contact_ms / (double)c->prof_steps : 0.0;
  return NSM_OK;
}

int64_t
nsm_b200_cold_points(nsm_b200_ctx* c)
{
  if (!c || !c->finalized) return -1;
  DeviceGuard device_guard_(c->device);
  int h = 0;
  if (cudaMemcpyAsync(&h, c->d_flags + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -1;
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
  return (int64_t)(unsigned)h;
}

int
nsm_b200_fp64_peak(nsm_b200_ctx* c, double* dadd_dmul_tops, double* dfma_tops)
{
  NSM_ENTER(c);
  cudaDeviceProp prop;
  NSM_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
  double*   out = nullptr;
  NSM_CUDA(c, cudaMalloc((void**)&out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  double res[2] = {0, 0};
  for (int fused = 0; fused < 2; ++fused) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0, c->stream);
      if (fused)
        fp64_peak_kernel<true><<<blocks, threads, 0, c->stream>>>(out, iters, 1.0);
      else
        fp64_peak_kernel<false><<<blocks, threads, 0, c->stream>>>(out, iters, 1.0);
      cudaEventRecord(e1, c->stream);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
      c->launches++;
    }
    res[fused] = (double)blocks * threads * (double)iters * 8.0 / (best * 1e-3) / 1e12;
  }
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  cudaFree(out);
  NSM_CUDA(c, cudaGetLastError());
  if (dadd_dmul_tops) *dadd_dmul_tops = res[0];
  if (dfma_tops) *dfma_tops = res[1];
  return NSM_OK;
}

int
nsm_b200_fp64_peak_sustained(nsm_b200_ctx* c, double seconds, double* dadd_dmul_tops)
{
  NSM_ENTER(c);
  NSM_REQUIRE(c, seconds > 0.0 && seconds <= 30.0 && dadd_dmul_tops, "fp64_peak_sustained: bad arguments");
  cudaDeviceProp prop;
  NSM_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
  double*   out = nullptr;
  NSM_CUDA(c, cudaMalloc((void**)&out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  // back-to-back launches for `seconds`; the rate of the last quarter is what a long FP64-bound run can sustain
  // (power capping lowers the SM clock after the first few hundred milliseconds)
  std::vector<float> ms;
  double             elapsed = 0.0;
  while (elapsed < seconds * 1e3) {
    cudaEventRecord(e0, c->stream);
    for (int k = 0; k < 4; ++k) fp64_peak_kernel<false><<<blocks, threads, 0, c->stream>>>(out, iters, 1.0);
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1);
    ms.push_back(t / 4.0f);
    elapsed += t;
    c->launches += 4;
  }
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  cudaFree(out);
  NSM_CUDA(c, cudaGetLastError());
  double sum = 0.0;
  size_t n0  = ms.size() - std::max<size_t>(ms.size() / 4, 1);
  for (size_t i = n0; i < ms.size(); ++i) sum += ms[i];
  const double avg = sum / (double)(ms.size() - n0);
  *dadd_dmul_tops  = (double)blocks * threads * (double)iters * 8.0 / (avg * 1e-3) / 1e12;
  return NSM_OK;
}

}  // extern "C"
