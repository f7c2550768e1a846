// nimblesm_b200/host/data_manager.h — nimble::DataManager (src/nimble_data_manager.{h,cc}) and
// nimble::VectorCommunicator (src/nimble_vector_communicator.h:104-171) for the B200 build.
//
// DataManager owns the model data, the field ids, the boundary-condition manager, the Exodus writer and the
// vector communicator, and allocates the nodal fields in the reference's order (src/nimble_data_manager.cc:
// 135-158).  VectorCommunicator keeps the reference's role -- find the nodes this rank shares with other
// ranks from the global node ids and sum nodal fields over the holders -- but the sum itself runs on the
// devices over NVLink peer memory (csrc/peer_exchange.cuh): Initialize() only derives the shared-node tables,
// which ModelData passes to nsm_b200_comm_init.  Ranks are threads of one process (one GPU each), joined by a
// RankGroup; a serial run has a group of one and no exchange.
#pragma once
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "boundary_condition.h"
#include "exodus_output.h"
#include "genesis_mesh.h"
#include "model_data.h"
#include "parser.h"

namespace nimble_b200 {

// Rendezvous of the rank threads of one process (the MPI_COMM_WORLD of this build): barrier + all-gather of
// byte blobs.  All members are called collectively by every rank.
class RankGroup
{
 public:
  explicit RankGroup(int num_ranks) : num_ranks_(num_ranks), slots_(num_ranks) {}
  int
  NumRanks() const
  {
    return num_ranks_;
  }
  void
  Barrier();
  // every rank contributes `mine`; returns all contributions indexed by rank
  std::vector<std::vector<char>>
  AllGather(int rank, const std::vector<char>& mine);
  double
  MinAll(int rank, double value);
  // Ranks that share one GPU (NimbleSM_b200 --devices 0,0: the multi-rank path on a single-GPU box) must enter
  // every device call that contains a shared-node exchange together: a rank's in-kernel wait for its peer's data
  // would otherwise sit in front of a device-synchronising call (cudaFree, a pageable copy) of that peer's thread.
  void
  SetLockstep(bool on)
  {
    lockstep_ = on;
  }
  bool
  Lockstep() const
  {
    return lockstep_;
  }

 private:
  bool                           lockstep_ = false;
  int                            num_ranks_;
  std::mutex                     mutex_;
  std::condition_variable        cv_;
  int                            waiting_ = 0;
  long                           generation_ = 0;
  std::vector<std::vector<char>> slots_;
};

class VectorCommunicator
{
 public:
  VectorCommunicator(int dim, unsigned int num_nodes, std::shared_ptr<RankGroup> group = nullptr, int rank = 0)
      : dim_(dim), num_nodes_(num_nodes), group_(std::move(group)), rank_(rank)
  {
  }
  // GenerateReductionInfo (src/nimble.mpi.reduction.cc:50-123): for every other rank, the LOCAL ids of the
  // nodes shared with it, sorted by GLOBAL id so that both sides enumerate them alike (:114-120)
  void
  Initialize(std::vector<int> const& global_node_ids);
  int
  Rank() const
  {
    return rank_;
  }
  int
  NumRanks() const
  {
    return group_ ? group_->NumRanks() : 1;
  }
  std::shared_ptr<RankGroup>
  Group() const
  {
    return group_;
  }
  const std::vector<int>&
  PeerRanks() const
  {
    return peer_ranks_;
  }
  const std::vector<int64_t>&
  PairOffsets() const
  {
    return pair_offsets_;
  }
  const std::vector<int>&
  PairLocalNodes() const
  {
    return pair_local_nodes_;
  }
  // attaches the device exchange of `device` to its peers (collective)
  void
  ConnectDevices(DeviceContext& device);

 private:
  int                        dim_;
  unsigned int               num_nodes_;
  std::shared_ptr<RankGroup> group_;
  int                        rank_;
  std::vector<int>           peer_ranks_;
  std::vector<int64_t>       pair_offsets_{0};
  std::vector<int>           pair_local_nodes_;
};

class DataManager
{
 public:
  // device / assembly / flags configure the nimble_b200::ModelData this manager creates
  DataManager(const Parser& parser, const GenesisMesh& mesh, int device = 0, int assembly = NSM_ASSEMBLY_ORDERED,
              unsigned flags = NSM_FLAG_CACHE_REF_JACOBIAN, std::shared_ptr<RankGroup> group = nullptr);
  ~DataManager() = default;
  void
  InitializeOutput(const std::string& filename);
  const Parser&
  GetParser() const
  {
    return parser_;
  }
  const GenesisMesh&
  GetMesh() const
  {
    return mesh_;
  }
  std::shared_ptr<ModelDataBase>
  GetModelData()
  {
    return model_data_;
  }
  const FieldIds&
  GetFieldIDs() const
  {
    return field_ids_;
  }
  FieldIds&
  GetFieldIDs()
  {
    return field_ids_;
  }
  std::shared_ptr<VectorCommunicator>
  GetVectorCommunicator()
  {
    return vector_communicator_;
  }
  void
  WriteOutput(double time_current);
  std::shared_ptr<ExodusOutput>
  GetExodusOutput()
  {
    return exodus_output_;
  }
  void
  SetBlockMaterialInterfaceFactory(const std::shared_ptr<BlockMaterialInterfaceFactoryBase>& block_material_factory)
  {
    block_material_factory_ = block_material_factory;
  }
  const std::shared_ptr<BlockMaterialInterfaceFactoryBase>&
  GetBlockMaterialInterfaceFactory() const
  {
    return block_material_factory_;
  }
  std::shared_ptr<BoundaryConditionManager>
  GetBoundaryConditionManager()
  {
    return boundary_condition_;
  }

 protected:
  void
  Initialize(int device, int assembly, unsigned flags, std::shared_ptr<RankGroup> group);
  const Parser&                                      parser_;
  const GenesisMesh&                                 mesh_;
  std::shared_ptr<ModelDataBase>                     model_data_;
  FieldIds                                           field_ids_;
  std::shared_ptr<VectorCommunicator>                vector_communicator_;
  std::shared_ptr<ExodusOutput>                      exodus_output_;
  std::shared_ptr<BlockMaterialInterfaceFactoryBase> block_material_factory_;
  std::shared_ptr<BoundaryConditionManager>          boundary_condition_;
};

}  // namespace nimble_b200
