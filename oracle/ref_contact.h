// oracle/ref_contact.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The reference's contact manager (src/nimble_contact_manager.{h,cc}, src/contact/serial/arborx_serial_contact_manager.cc)
// cannot be compiled in this container: its header declares Kokkos views unconditionally and the search is ArborX.  Its
// contact ENTITIES can (src/nimble_contact_entity.{h,cc} need neither).  RefContact therefore keeps the reference's own
// unmodified nimble::ContactEntity objects -- constructor, SetCoordinates, SetBoundingBox, SetNodalContactForces,
// ScatterForceToContactManagerForceVector are the reference's code -- and restates only what lives in the uncompilable
// files: the skinning and entity creation glue (nimble_contact_manager.cc:184-393, 788-934, 1043-1205), the projection
// (:1549-1620), the penalty force (nimble_contact_manager.h:80-128) and the pair loop (an all-pairs walk in place of the
// ArborX BVH query; boxes narrowed to float as ArborX::Box holds them, src/contact/arborx_utils.h:85-90).
#pragma once
#include <string>
#include <vector>

#include "nimble_contact_entity.h"
#include "nimble_genesis_mesh.h"

namespace nsm_oracle {

class RefContact
{
 public:
  // `contact:` deck line -> block names and penalty (ParseContactCommand, nimble_contact_manager.cc:95-149)
  static void
  ParseCommand(const std::string& command, std::vector<std::string>& primary, std::vector<std::string>& secondary, double& penalty);

  void
  Create(const nimble::GenesisMesh& mesh, const std::vector<int>& primary_block_ids, const std::vector<int>& secondary_block_ids,
         double penalty);

  // displacement, contact_force: [n_nodes][3] of the parent mesh.  Returns the number of enforced node-face pairs.
  long
  Compute(const double* displacement, double* contact_force, int n_mesh_nodes);

  int
  NumFaces() const
  {
    return (int)faces_.size();
  }
  int
  NumNodes() const
  {
    return (int)nodes_.size();
  }
  // what Create built, for pinning the product's host-side entity creation: quads [n][4] and contact nodes as MESH node ids
  std::vector<int>    primary_quads, secondary_quads, contact_node_ids;
  std::vector<double> primary_char_len, contact_node_char_len;

 private:
  double                             penalty_ = 0.0;
  std::vector<int>                   node_ids_;  // contact sub-model node -> mesh node
  std::vector<double>                model_coord_, coord_, force_;
  std::vector<nimble::ContactEntity> faces_, nodes_;
};

}  // namespace nsm_oracle
