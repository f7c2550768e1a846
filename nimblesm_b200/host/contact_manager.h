// nimblesm_b200/host/contact_manager.h — nimble::ContactManager for the B200 build (src/nimble_contact_manager.{h,cc},
// src/contact/serial/arborx_serial_contact_manager.cc): penalty contact between the skin of the primary blocks
// (triangulated faces) and the skin nodes of the secondary blocks.
//
// Host side = what the reference does once, on the host, at start-up: parse the `contact:` deck line, skin the
// blocks, list the contact nodes, measure the characteristic lengths (CreateContactEntities, :184-393).  The per-step
// part -- coordinates, bounding boxes, search, projection, enforcement, scatter -- runs on the device behind
// nsm_b200_set_contact / nsm_b200_contact_force (csrc/contact.cuh): a uniform hashed grid in place of the ArborX BVH,
// the same accepted pairs, the same per-pair arithmetic.  Inside the fused step (ModelData::AdvanceOnDevice) the
// contact force never leaves the device; ComputeContactForce keeps the reference's call shape for the call-by-call
// sequence.  Several ranks (mesh partitions): the contact surface is replicated -- every rank learns the whole skin
// once, rank 0 evaluates the contact force of the whole surface on its GPU each step from the pooled displacements and
// every rank picks its nodes' entries (see BuildReplicatedSubModel); such runs take the call-by-call sequence.
#pragma once
#include <cstddef>
#include <memory>
#include <string>
#include <vector>

#include "device.h"
#include "exodus_output.h"
#include "genesis_mesh.h"
#include "view.h"

namespace nimble_b200 {

class DataManager;
class RankGroup;
class VectorCommunicator;

// `contact:` line -> block names and penalty parameter; throws std::invalid_argument with the reference's messages
// (ParseContactCommand, src/nimble_contact_manager.cc:95-149)
void
ParseContactCommand(std::string const& command, std::vector<std::string>& primary_block_names,
                    std::vector<std::string>& secondary_block_names, double& penalty_parameter);

// what CreateContactEntities builds on the host and sends to the device (mesh node ids)
struct ContactEntityLists
{
  std::vector<int>    primary_face_nodes;  // [n_faces][4], Exodus face order
  std::vector<int>    primary_face_entity_ids;
  std::vector<double> primary_face_char_len;
  std::vector<int>    contact_node_ids;
  std::vector<double> contact_node_char_len;
};

// the replicated contact sub-model of a multi-rank run (ContactManager::BuildReplicatedSubModel): identical on every
// rank except for the `held_*` lists
struct ReplicatedContactSubModel
{
  ContactEntityLists  lists;          // node ids = surface indices
  std::vector<double> surface_xyz;    // [n_surface][3] model coordinates
  std::vector<int>    surface_gid;    // [n_surface] global node id, ascending
  std::vector<int>    held_local, held_surface;  // surface nodes this rank holds: local node id, surface index
};

// The contact visualisation database (InitializeContactVisualization / WriteVisualizationData,
// src/nimble_contact_manager.cc:431-678), device-free: a second Exodus file with one triangle element per contact facet
// (its own three nodes: two face nodes and the fictitious node at the face centre), one single-node element per contact
// node, the entities' displacement and contact_status as nodal variables and num_contacts as a global one.  Entity ids
// as in the reference: element id = contact_entity_global_id_ (facet: skin entity id | triangle ordinal; contact node:
// global node id + 1), facet nodes 3 id + max id + 9 / 10 / 11.
class ContactVisualizationDatabase
{
 public:
  ContactVisualizationDatabase(GenesisMesh const& mesh, ContactEntityLists const& lists, std::string const& exodus_file_name);
  // displacement [n_nodes][3] of the mesh nodes (nullptr = entities at their model coordinates), status flags per
  // triangle and per contact node (nullptr = 0)
  void
  WriteStep(double t, const double* displacement, const unsigned char* face_status, const unsigned char* node_status);
  // coordinates of the visualisation nodes for nodal coordinates model + displacement: per facet node 1, node 2,
  // (c0 + c1 + c2 + c3) / 4 of its face (ContactEntity::SetCoordinates, src/nimble_contact_entity.h:236-258), then the
  // contact nodes
  void
  EntityVertices(const double* displacement, std::vector<double>& x, std::vector<double>& y, std::vector<double>& z) const;
  void
  Close()
  {
    out_.Close();
  }

 private:
  GenesisMesh const&        model_mesh_;
  ContactEntityLists const& lists_;
  GenesisMesh               mesh_;  // genesis_mesh_for_contact_visualization_
  ExodusOutput              out_;   // exodus_output_for_contact_visualization_
};

class ContactManager
{
 public:
  explicit ContactManager(DataManager& data_manager) : data_manager_(data_manager) {}
  virtual ~ContactManager() = default;

  bool
  ContactEnabled() const
  {
    return contact_enabled_;
  }
  void
  SetPenaltyParameter(double penalty_parameter)
  {
    penalty_parameter_ = penalty_parameter;
  }
  double
  GetPenaltyForceParam() const noexcept
  {
    return penalty_parameter_;
  }

  // Skin faces of the listed blocks: faces met exactly once, each in the Exodus face order of its element, sorted by
  // their sorted node lists (the reference's std::map order); entity id = (element global id + 1 + offset) << 5 |
  // face ordinal << 2 (SkinBlocks, :788-934)
  static void
  SkinBlocks(GenesisMesh const& mesh, std::vector<int> const& block_ids, int entity_id_offset, std::vector<std::vector<int>>& skin_faces,
             std::vector<int>& entity_ids);

  // CreateContactEntities (:184-393): contact entities of this rank, sent to the device
  void
  CreateContactEntities(GenesisMesh const& mesh, VectorCommunicator& vector_communicator, std::vector<int> const& primary_block_ids,
                        std::vector<int> const& secondary_block_ids);

  // ComputeContactForce (:395-429): contact force of the displacement the device holds, into the caller's view
  void
  ComputeContactForce(int step, bool debug_output, Viewify<2> contact_force);

  std::size_t
  numContactFaces() const
  {
    return 4 * lists_.primary_face_char_len.size();
  }
  std::size_t
  numContactNodes() const
  {
    return lists_.contact_node_ids.size();
  }
  std::size_t
  numActiveContactFaces() const;
  std::size_t
  numActiveContactNodes() const;

  // InitializeContactVisualization / ContactVisualizationWriteStep (src/nimble_contact_manager.cc:431-602): the database
  // above, fed with the displacement of the model data and the contact_status flags of the device.  Written by
  // single-rank runs (with the replicated sub-model of a multi-rank run the request is reported and skipped).
  void
  InitializeContactVisualization(std::string const& contact_visualization_exodus_file_name);
  // `evaluated` = false writes the entities at their model coordinates (the initial write, before any evaluation)
  void
  ContactVisualizationWriteStep(double time_current, bool evaluated = true);

  ContactEntityLists const&
  EntityLists() const
  {
    return lists_;
  }
  // contact across partitions: the replicated sub-model is in use (the integrator then sequences the steps call by call)
  bool
  Replicated() const
  {
    return replicated_;
  }
  // several ranks: the replicated contact sub-model (surface numbering = ascending global node id); collective over the
  // ranks of the communicator's group, no device involved
  static void
  BuildReplicatedSubModel(GenesisMesh const& mesh, VectorCommunicator& vector_communicator, std::vector<int> const& primary_block_ids,
                          std::vector<int> const& secondary_block_ids, ReplicatedContactSubModel& out);
  // everything CreateContactEntities does before the upload (no device involved)
  static void
  BuildEntityLists(GenesisMesh const& mesh, std::vector<int> const& primary_block_ids, std::vector<int> const& secondary_block_ids,
                   ContactEntityLists& lists);

 protected:
  DataManager&       data_manager_;
  bool               contact_enabled_   = false;
  double             penalty_parameter_ = 0.0;
  ContactEntityLists lists_;
  bool                           replicated_ = false;
  std::shared_ptr<RankGroup>     group_;
  int                            rank_ = 0;
  std::vector<double>            surface_xyz_;                  // [n_surface][3] model coordinates
  std::vector<int>               held_local_, held_surface_;    // surface nodes this rank holds: local id, surface index
  std::unique_ptr<DeviceContext> sub_model_;                    // rank 0: the element-free context of the surface nodes
  // contact visualisation
  const GenesisMesh*                            mesh_ = nullptr;  // the mesh the entity lists refer to (single rank)
  std::unique_ptr<ContactVisualizationDatabase> visualization_;
};

// the reference's factory (GetContactManager, :151-171): nullptr when the deck has no `contact:` line
std::shared_ptr<ContactManager>
GetContactManager(DataManager& data_manager);

}  // namespace nimble_b200
