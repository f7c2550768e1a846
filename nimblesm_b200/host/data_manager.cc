// nimblesm_b200/host/data_manager.cc — see data_manager.h.
#include "data_manager.h"

#include <algorithm>
#include <cstring>
#include <stdexcept>

namespace nimble_b200 {

// ---- RankGroup --------------------------------------------------------------------------------------------
void
RankGroup::Barrier()
{
  std::unique_lock<std::mutex> lock(mutex_);
  const long                   gen = generation_;
  if (++waiting_ == num_ranks_) {
    waiting_ = 0;
    ++generation_;
    cv_.notify_all();
  } else {
    cv_.wait(lock, [&] { return generation_ != gen; });
  }
}

std::vector<std::vector<char>>
RankGroup::AllGather(int rank, const std::vector<char>& mine)
{
  {
    std::lock_guard<std::mutex> lock(mutex_);
    slots_[rank] = mine;
  }
  Barrier();
  std::vector<std::vector<char>> all;
  {
    std::lock_guard<std::mutex> lock(mutex_);
    all = slots_;
  }
  Barrier();  // nobody overwrites a slot before everybody has copied
  return all;
}

double
RankGroup::MinAll(int rank, double value)
{
  std::vector<char> blob(sizeof(double));
  memcpy(blob.data(), &value, sizeof(double));
  double m = value;
  for (auto const& b : AllGather(rank, blob)) {
    double v;
    memcpy(&v, b.data(), sizeof(double));
    m = std::min(m, v);
  }
  return m;
}

// ---- VectorCommunicator -----------------------------------------------------------------------------------
void
VectorCommunicator::Initialize(std::vector<int> const& global_node_ids)
{
  peer_ranks_.clear();
  pair_offsets_.assign(1, 0);
  pair_local_nodes_.clear();
  if (!group_ || group_->NumRanks() <= 1) return;
  std::vector<char> blob(global_node_ids.size() * sizeof(int));
  memcpy(blob.data(), global_node_ids.data(), blob.size());
  const auto all = group_->AllGather(rank_, blob);
  // my (global id -> local id), then for every other rank the ids it also holds
  std::vector<std::pair<int, int>> mine(global_node_ids.size());
  for (size_t i = 0; i < global_node_ids.size(); ++i) mine[i] = {global_node_ids[i], (int)i};
  std::sort(mine.begin(), mine.end());
  for (int r = 0; r < group_->NumRanks(); ++r) {
    if (r == rank_) continue;
    std::vector<int> theirs(all[r].size() / sizeof(int));
    memcpy(theirs.data(), all[r].data(), all[r].size());
    std::sort(theirs.begin(), theirs.end());
    std::vector<int> shared_local;  // ascending global id
    size_t           a = 0, b = 0;
    while (a < mine.size() && b < theirs.size()) {
      if (mine[a].first < theirs[b])
        ++a;
      else if (theirs[b] < mine[a].first)
        ++b;
      else {
        shared_local.push_back(mine[a].second);
        ++a, ++b;
      }
    }
    if (shared_local.empty()) continue;
    peer_ranks_.push_back(r);
    pair_local_nodes_.insert(pair_local_nodes_.end(), shared_local.begin(), shared_local.end());
    pair_offsets_.push_back((int64_t)pair_local_nodes_.size());
  }
}

void
VectorCommunicator::ConnectDevices(DeviceContext& device)
{
  if (!group_ || group_->NumRanks() <= 1) return;
  device.check(nsm_b200_comm_init(device.get(), rank_, group_->NumRanks(), (int)peer_ranks_.size(), peer_ranks_.data(),
                                  pair_offsets_.data(), pair_local_nodes_.data()),
               "VectorCommunicator: comm_init");
  std::vector<char> handle(NSM_COMM_HANDLE_BYTES);
  device.check(nsm_b200_comm_export(device.get(), (unsigned char*)handle.data()), "VectorCommunicator: comm_export");
  const auto all = group_->AllGather(rank_, handle);
  for (int p : peer_ranks_)
    device.check(nsm_b200_comm_attach(device.get(), p, (const unsigned char*)all[p].data()), "VectorCommunicator: comm_attach");
  device.check(nsm_b200_comm_ready(device.get()), "VectorCommunicator: comm_ready");
  if (group_->Lockstep())  // ranks share a GPU: rendezvous on the host instead of waiting inside a kernel
    device.check(nsm_b200_comm_set_host_barrier(device.get(), [](void* g) { static_cast<RankGroup*>(g)->Barrier(); }, group_.get()),
                 "VectorCommunicator: host barrier");
  group_->Barrier();  // every rank is attached before any rank exchanges
}

// ---- DataManager ------------------------------------------------------------------------------------------
DataManager::DataManager(const Parser& parser, const GenesisMesh& mesh, int device, int assembly, unsigned flags,
                         std::shared_ptr<RankGroup> group)
    : parser_(parser), mesh_(mesh), boundary_condition_(new BoundaryConditionManager())
{
  Initialize(device, assembly, flags, std::move(group));
}

void
DataManager::Initialize(int device, int assembly, unsigned flags, std::shared_ptr<RankGroup> group)
{
  const int dim       = mesh_.GetDim();
  const int num_nodes = (int)mesh_.GetNumNodes();
  vector_communicator_ = std::make_shared<VectorCommunicator>(dim, num_nodes, group, parser_.GetRankID());
  std::vector<int> global_node_ids(mesh_.GetNodeGlobalIds(), mesh_.GetNodeGlobalIds() + num_nodes);
  vector_communicator_->Initialize(global_node_ids);

  model_data_ = std::make_shared<ModelData>(device, assembly, flags);
  model_data_->SetDimension(dim);

  const std::string scheme = parser_.TimeIntegrationScheme();
  if (scheme != "explicit")
    throw std::invalid_argument("\nError: the B200 build implements the explicit time integration scheme only (deck asks for " + scheme + ")\n");
  boundary_condition_->Initialize(mesh_.GetNodeSetNames(), mesh_.GetNodeSets(), {}, {}, parser_.GetBoundaryConditionStrings(), dim,
                                  scheme);
  // nodal fields in the reference's order (src/nimble_data_manager.cc:135-147)
  field_ids_.lumped_mass           = model_data_->AllocateNodeData(SCALAR, "lumped_mass", num_nodes);
  field_ids_.reference_coordinates = model_data_->AllocateNodeData(VECTOR, "reference_coordinate", num_nodes);
  field_ids_.displacement          = model_data_->AllocateNodeData(VECTOR, "displacement", num_nodes);
  field_ids_.velocity              = model_data_->AllocateNodeData(VECTOR, "velocity", num_nodes);
  field_ids_.acceleration          = model_data_->AllocateNodeData(VECTOR, "acceleration", num_nodes);
  field_ids_.internal_force        = model_data_->AllocateNodeData(VECTOR, "internal_force", num_nodes);
  field_ids_.external_force        = model_data_->AllocateNodeData(VECTOR, "external_force", num_nodes);
  field_ids_.contact_force         = model_data_->AllocateNodeData(VECTOR, "contact_force", num_nodes);
  model_data_->SetReferenceCoordinates(mesh_);
}

void
DataManager::InitializeOutput(const std::string& filename)
{
  std::vector<std::string> global_data_labels;
  exodus_output_ = std::make_shared<ExodusOutput>();
  exodus_output_->Initialize(filename, mesh_);
  model_data_->InitializeExodusOutput(*this);
  exodus_output_->InitializeDatabase(mesh_, global_data_labels, model_data_->GetNodeDataLabelsForOutput(),
                                     model_data_->GetElementDataLabelsForOutput(),
                                     model_data_->GetDerivedElementDataLabelsForOutput());
}

void
DataManager::WriteOutput(double time_current)
{
  model_data_->WriteExodusOutput(*this, time_current);
}

}  // namespace nimble_b200
