// nimblesm_b200/csrc/peer_exchange.cuh — shared-node sum across GPUs over NVLink peer memory.
//
// Replaces VectorCommunicator::VectorReduction -> ReductionInfo::PerformReduction -> one MPI_Iallreduce per
// rank clique on host buffers (src/nimble_vector_communicator.h:144-157, src/nimble.mpi.reduction.h:163-175,
// src/nimble.mpi.rank_clique_reducer.h:130-257).  Here the partial nodal sums never leave the devices:
//
//   pack    one kernel stores every shared node's partial value straight into each co-holder's receive
//           buffer (st.global on a peer-mapped pointer => NVLink/NVSwitch write); the last CTA to finish
//           publishes a sequence number in each peer's flag word (fence.sys + store)
//   wait    a one-warp kernel spins (bounded) until every peer's sequence number has arrived
//   unpack  per shared node the contributions of all holders INCLUDING self are added in ascending rank
//           order, so every replica of the node ends with the bit-identical total (MPI_SUM gives no such
//           guarantee; replicas of a node must not drift apart in an explicit code)
//
// Receive buffers are double-buffered by call parity; a rank can be at most one call ahead of a peer
// because its next pack is stream-ordered after its own unpack, which needs the peer's previous pack.
// Buffers are exported as cudaIpc handles (process-per-GPU) or raw pointers (thread-per-GPU, same pid).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

namespace nsm {

constexpr int kMaxPeers   = 15;
constexpr int kFlagWords  = 16;  // one uint64 per possible sender rank
constexpr int kCommComps  = 3;

struct CommBlob  // fits NSM_COMM_HANDLE_BYTES (192)
{
  cudaIpcMemHandle_t ipc;          // 64
  uint64_t           raw_ptr;      // same-process attach
  int32_t            pid, device;  //
  int32_t            rank, n_peers;
  int64_t            total_pairs;
  int32_t            peer_rank[kMaxPeers > 7 ? 7 : kMaxPeers];
  int32_t            peer_off[kMaxPeers > 7 ? 7 : kMaxPeers];
};
static_assert(sizeof(CommBlob) <= 192, "comm blob must fit the ABI handle size");

struct PackArgs
{
  int64_t         total;       // send entries
  int             ncomp;
  const int*      entry_node;  // local node id of each entry
  const int*      entry_peer;  // peer index of each entry
  const int64_t*  entry_k;     // position inside that peer's region
  double*         peer_data[8];  // peer receive data base (already offset to my region and parity)
  unsigned long long* peer_flag[8];
  int             n_peers;
  unsigned long long seq;
  const double*   f[3];
  unsigned int*   done_counter;
};

__global__ void __launch_bounds__(256)
comm_pack_kernel(const PackArgs p)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < p.total) {
    const int     nd = p.entry_node[t];
    const int     pi = p.entry_peer[t];
    double*       d  = p.peer_data[pi] + p.entry_k[t] * kCommComps;
    for (int c = 0; c < p.ncomp; ++c) d[c] = p.f[c][nd];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(p.done_counter, 1u);
    if (prev == gridDim.x - 1) {
      __threadfence_system();
      for (int i = 0; i < p.n_peers; ++i) *(volatile unsigned long long*)p.peer_flag[i] = p.seq;
      *p.done_counter = 0u;
      __threadfence_system();
    }
  }
}

__global__ void
comm_wait_kernel(const volatile unsigned long long* flags, const int* peer_ranks, int n_peers, unsigned long long seq,
                 long long timeout_ns, int* err_flags)
{
  const int i = threadIdx.x;
  if (i < n_peers) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (flags[peer_ranks[i]] < seq) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if ((long long)(t1 - t0) > timeout_ns) {
        atomicOr(err_flags, 2);
        break;
      }
      __nanosleep(200);
    }
  }
  __threadfence_system();
}

struct UnpackArgs
{
  int64_t        n_shared;
  int            ncomp;
  const int*     node;     // local id of each shared node
  const int64_t* src_off;  // [n_shared+1]
  const int64_t* src;      // ascending holder rank; -1 = this rank's own partial, else receive entry index
  const double*  recv;     // my receive data of this parity
  double*        f[3];
  const int*     err;      // set by a wait that timed out: the receive data is stale, leave the partial sums alone
};

__global__ void __launch_bounds__(256)
comm_unpack_kernel(const UnpackArgs p)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_shared || *p.err != 0) return;
  const int nd = p.node[i];
  for (int c = 0; c < p.ncomp; ++c) {
    double  s     = 0.0;
    bool    first = true;
    for (int64_t k = p.src_off[i]; k < p.src_off[i + 1]; ++k) {
      const int64_t e = p.src[k];
      const double  x = e < 0 ? p.f[c][nd] : __ldcg(p.recv + e * kCommComps + c);
      s     = first ? x : s + x;
      first = false;
    }
    p.f[c][nd] = s;
  }
}

class PeerExchange
{
 public:
  bool
  active() const
  {
    return ready_;
  }
  const char*
  error() const
  {
    return err_.c_str();
  }

  // peer_ranks[n_peers]; pair_off[n_peers+1]; pair_node: for each peer, the local ids of the nodes shared with
  // it in ascending GLOBAL id order (both sides sort alike, src/nimble.mpi.reduction.cc:114-120).
  int
  init(int device, int rank, int world, int n_peers, const int32_t* peer_ranks, const int64_t* pair_off,
       const int32_t* pair_node)
  {
    device_ = device, rank_ = rank, world_ = world;
    if (world > kFlagWords) return set_err("world size above 16 is not supported by the peer exchange");
    if (n_peers > 7) return set_err("more than 7 peers per rank is not supported (8 GPUs per box)");
    peers_.assign(peer_ranks, peer_ranks + n_peers);
    off_.assign(pair_off, pair_off + n_peers + 1);
    total_ = off_.empty() ? 0 : off_.back();
    nodes_.assign(pair_node, pair_node + total_);
    for (int i = 0; i < n_peers; ++i)
      if (peers_[i] < 0 || peers_[i] >= world || peers_[i] == rank) return set_err("invalid peer rank");
    const size_t bytes = kFlagWords * sizeof(unsigned long long) + 2 * (size_t)std::max<int64_t>(total_, 1) * kCommComps * sizeof(double);
    if (cudaMalloc(&buf_, bytes) != cudaSuccess) return set_err("cudaMalloc of the receive buffer failed");
    cudaMemset(buf_, 0, bytes);
    cudaDeviceSynchronize();
    attached_.assign(n_peers, false);
    peer_base_.assign(n_peers, nullptr);
    peer_total_.assign(n_peers, 0);
    peer_off_.assign(n_peers, 0);
    inited_ = true;
    return 0;
  }

  int
  export_handle(unsigned char* out)
  {
    if (!inited_) return set_err("comm_export before comm_init");
    CommBlob b;
    memset(&b, 0, sizeof b);
    if (cudaIpcGetMemHandle(&b.ipc, buf_) != cudaSuccess) {
      cudaGetLastError();  // IPC may be unavailable (e.g. same-process use); raw pointer still works there
      memset(&b.ipc, 0, sizeof b.ipc);
    }
    b.raw_ptr = (uint64_t)(uintptr_t)buf_;
    b.pid     = (int32_t)getpid();
    b.device  = device_;
    b.rank    = rank_;
    b.n_peers = (int32_t)peers_.size();
    b.total_pairs = total_;
    for (size_t i = 0; i < peers_.size(); ++i) {
      b.peer_rank[i] = peers_[i];
      b.peer_off[i]  = (int32_t)off_[i];
    }
    memset(out, 0, 192);
    memcpy(out, &b, sizeof b);
    return 0;
  }

  int
  attach(int peer_rank, const unsigned char* blob)
  {
    if (!inited_) return set_err("comm_attach before comm_init");
    CommBlob b;
    memcpy(&b, blob, sizeof b);
    if (b.rank != peer_rank) return set_err("handle does not belong to the named peer rank");
    int pi = -1;
    for (size_t i = 0; i < peers_.size(); ++i)
      if (peers_[i] == peer_rank) pi = (int)i;
    if (pi < 0) return set_err("attach: rank is not one of this rank's peers");
    int my_off = -1;
    for (int i = 0; i < b.n_peers; ++i)
      if (b.peer_rank[i] == rank_) my_off = b.peer_off[i];
    if (my_off < 0) return set_err("attach: the peer does not list this rank as a peer (asymmetric sharing)");
    void* base = nullptr;
    if (b.pid == (int32_t)getpid()) {
      if (b.device != device_) {
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return set_err("cudaDeviceEnablePeerAccess failed");
        cudaGetLastError();
      }
      base = (void*)(uintptr_t)b.raw_ptr;
    } else {
      if (cudaIpcOpenMemHandle(&base, b.ipc, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
        return set_err("cudaIpcOpenMemHandle failed (peer not reachable over NVLink/PCIe P2P?)");
      ipc_opened_.push_back(base);
    }
    peer_base_[pi]  = (unsigned char*)base;
    peer_total_[pi] = b.total_pairs;
    peer_off_[pi]   = my_off;
    attached_[pi]   = true;
    return 0;
  }

  int
  ready(cudaStream_t stream)
  {
    if (!inited_) return set_err("comm_ready before comm_init");
    for (size_t i = 0; i < attached_.size(); ++i)
      if (!attached_[i]) return set_err("comm_ready: a peer has not been attached");
    // send entry tables
    std::vector<int>     e_node(total_), e_peer(total_);
    std::vector<int64_t> e_k(total_);
    for (size_t pi = 0; pi < peers_.size(); ++pi)
      for (int64_t k = off_[pi]; k < off_[pi + 1]; ++k) e_node[k] = nodes_[k], e_peer[k] = (int)pi, e_k[k] = k - off_[pi];
    // unique shared nodes and their contributions in ascending rank order
    std::map<int, std::vector<std::pair<int, int64_t>>> by_node;  // node -> (rank, recv entry)
    for (size_t pi = 0; pi < peers_.size(); ++pi)
      for (int64_t k = off_[pi]; k < off_[pi + 1]; ++k) by_node[nodes_[k]].push_back({peers_[pi], k});
    std::vector<int>     u_node;
    std::vector<int64_t> s_off{0}, s_src;
    for (auto& kv : by_node) {
      auto v = kv.second;
      v.push_back({rank_, -1});
      std::sort(v.begin(), v.end());
      u_node.push_back(kv.first);
      for (auto& pr : v) s_src.push_back(pr.second);
      s_off.push_back((int64_t)s_src.size());
    }
    n_shared_ = (int64_t)u_node.size();
    if (upload(&d_entry_node_, e_node) || upload(&d_entry_peer_, e_peer) || upload(&d_entry_k_, e_k) ||
        upload(&d_node_, u_node) || upload(&d_src_off_, s_off) || upload(&d_src_, s_src) || upload(&d_peer_ranks_, peers_))
      return set_err("device allocation of the exchange tables failed");
    if (cudaMalloc((void**)&d_counter_, sizeof(unsigned)) != cudaSuccess) return set_err("cudaMalloc failed");
    cudaMemsetAsync(d_counter_, 0, sizeof(unsigned), stream);
    if (cudaMalloc((void**)&d_err_, sizeof(int)) != cudaSuccess) return set_err("cudaMalloc failed");
    cudaMemsetAsync(d_err_, 0, sizeof(int), stream);
    cudaStreamSynchronize(stream);
    ready_ = true;
    return 0;
  }

  // f[c][node] <- sum over holders, for every shared node.  All ranks must call in the same order.
  // Two halves so that the caller can put independent work between them (interior elements):
  //   pack    reads this rank's partial values of the shared nodes and stores them into the peers' buffers
  //   finish  waits for every peer's values of the same call and forms the rank-ordered sums in place
  int
  pack(cudaStream_t stream, double* const* f, int ncomp, int64_t* launches)
  {
    if (!ready_) return set_err("reduce before comm_ready");
    if (peers_.empty()) return 0;
    ++seq_;
    const int parity = (int)(seq_ & 1);
    PackArgs  p{};
    p.total = total_, p.ncomp = ncomp;
    p.entry_node = d_entry_node_, p.entry_peer = d_entry_peer_, p.entry_k = d_entry_k_;
    p.n_peers = (int)peers_.size();
    for (size_t pi = 0; pi < peers_.size(); ++pi) {
      unsigned char* base = peer_base_[pi];
      double*        data = (double*)(base + kFlagWords * sizeof(unsigned long long));
      p.peer_data[pi]     = data + ((int64_t)parity * std::max<int64_t>(peer_total_[pi], 1) + peer_off_[pi]) * kCommComps;
      p.peer_flag[pi]     = (unsigned long long*)base + rank_;
    }
    p.seq = seq_;
    for (int c = 0; c < 3; ++c) p.f[c] = c < ncomp ? f[c] : nullptr;
    p.done_counter = d_counter_;
    const unsigned grid = (unsigned)std::max<int64_t>((total_ + 255) / 256, 1);
    comm_pack_kernel<<<grid, 256, 0, stream>>>(p);
    if (host_barrier_) cudaEventRecord(packed_, stream);
    if (launches) *launches += 1;
    if (cudaGetLastError() != cudaSuccess) return set_err("peer exchange kernel launch failed");
    return 0;
  }

  int
  finish(cudaStream_t stream, double* const* f, int ncomp, int64_t* launches)
  {
    if (!ready_) return set_err("reduce before comm_ready");
    if (peers_.empty()) return 0;
    const int parity = (int)(seq_ & 1);
    if (host_barrier_) {
      // ranks that share one GPU must not wait for each other inside a kernel (the waiting kernel can sit in front of
      // the peer's pack in a hardware queue): the host waits for this rank's pack, meets the peers, then unpacks
      if (cudaEventSynchronize(packed_) != cudaSuccess) return set_err("peer exchange: waiting for the pack failed");
      host_barrier_(host_barrier_arg_);
    } else {
      comm_wait_kernel<<<1, 32, 0, stream>>>((const volatile unsigned long long*)buf_, d_peer_ranks_, (int)peers_.size(),
                                             seq_, timeout_ns_, d_err_);
    }
    UnpackArgs u{};
    u.n_shared = n_shared_, u.ncomp = ncomp, u.node = d_node_, u.src_off = d_src_off_, u.src = d_src_;
    u.recv = (const double*)((unsigned char*)buf_ + kFlagWords * sizeof(unsigned long long)) +
             (int64_t)parity * std::max<int64_t>(total_, 1) * kCommComps;
    for (int c = 0; c < 3; ++c) u.f[c] = c < ncomp ? f[c] : nullptr;
    u.err = d_err_;
    if (n_shared_ > 0) comm_unpack_kernel<<<(unsigned)((n_shared_ + 255) / 256), 256, 0, stream>>>(u);
    if (launches) *launches += 2;
    if (cudaGetLastError() != cudaSuccess) return set_err("peer exchange kernel launch failed");
    return 0;
  }

  int
  reduce(cudaStream_t stream, double* const* f, int ncomp, int64_t* launches)
  {
    if (pack(stream, f, ncomp, launches)) return 1;
    return finish(stream, f, ncomp, launches);
  }

  // shared nodes of this rank (unique local ids, device / host) for the boundary-first element schedule
  const int*
  shared_nodes_device() const
  {
    return d_node_;
  }
  int64_t
  num_shared_nodes() const
  {
    return n_shared_;
  }

  // non-zero when a wait timed out (checked by the caller together with the Jacobian flag)
  int
  poll_error(cudaStream_t stream)
  {
    if (!ready_) return 0;
    int h = 0;
    cudaMemcpyAsync(&h, d_err_, sizeof(int), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    if (h) {
      set_err("peer exchange timed out waiting for a peer's shared-node data");
      cudaMemsetAsync(d_err_, 0, sizeof(int), stream);
    }
    return h;
  }

  void
  destroy()
  {
    for (void* p : ipc_opened_) cudaIpcCloseMemHandle(p);
    ipc_opened_.clear();
    for (void* p : {(void*)buf_, (void*)d_entry_node_, (void*)d_entry_peer_, (void*)d_entry_k_, (void*)d_node_,
                    (void*)d_src_off_, (void*)d_src_, (void*)d_peer_ranks_, (void*)d_counter_, (void*)d_err_})
      if (p) cudaFree(p);
    buf_ = nullptr;
    if (packed_) cudaEventDestroy(packed_);
    packed_ = nullptr;
    ready_ = inited_ = false;
  }

  void
  set_timeout_seconds(double s)
  {
    timeout_ns_ = (long long)(s * 1e9);
  }
  // see finish(): replaces the in-kernel wait by a host rendezvous of all ranks (called once per exchange by every rank)
  int
  set_host_barrier(void (*barrier)(void*), void* arg)
  {
    host_barrier_ = barrier, host_barrier_arg_ = arg;
    if (barrier && !packed_ && cudaEventCreateWithFlags(&packed_, cudaEventDisableTiming) != cudaSuccess) return set_err("cudaEventCreate failed");
    return 0;
  }

 private:
  template <class T>
  int
  upload(T** d, const std::vector<T>& h)
  {
    if (cudaMalloc((void**)d, std::max<size_t>(h.size(), 1) * sizeof(T)) != cudaSuccess) return 1;
    if (!h.empty() && cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return 1;
    return 0;
  }
  int
  set_err(const char* m)
  {
    err_ = m;
    return 1;
  }

  int                  device_ = 0, rank_ = 0, world_ = 1;
  bool                 inited_ = false, ready_ = false;
  std::string          err_;
  std::vector<int>     peers_;
  std::vector<int64_t> off_;
  std::vector<int>     nodes_;
  int64_t              total_ = 0, n_shared_ = 0;
  void*                buf_ = nullptr;
  std::vector<bool>    attached_;
  std::vector<unsigned char*> peer_base_;
  std::vector<int64_t> peer_total_, peer_off_;
  std::vector<void*>   ipc_opened_;
  int *                d_entry_node_ = nullptr, *d_entry_peer_ = nullptr, *d_node_ = nullptr, *d_peer_ranks_ = nullptr;
  int64_t *            d_entry_k_ = nullptr, *d_src_off_ = nullptr, *d_src_ = nullptr;
  unsigned*            d_counter_ = nullptr;
  int*                 d_err_     = nullptr;
  unsigned long long   seq_       = 0;
  long long            timeout_ns_ = 20LL * 1000000000LL;
  void (*host_barrier_)(void*)     = nullptr;
  void*                host_barrier_arg_ = nullptr;
  cudaEvent_t          packed_     = nullptr;
};

}  // namespace nsm
