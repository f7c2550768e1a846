// nimblesm_b200/host/contact_manager.cc — see contact_manager.h.
#include "contact_manager.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>

#include "data_manager.h"
#include "genesis_mesh.h"
#include "model_data.h"

namespace nimble_b200 {

void
ParseContactCommand(std::string const& command, std::vector<std::string>& primary_block_names,
                    std::vector<std::string>& secondary_block_names, double& penalty_parameter)
{
  std::istringstream in(command);
  std::string        word;
  in >> word;
  if (word != "primary_blocks" && word != "master_blocks")
    throw std::invalid_argument("\n**** Error processing contact command, unknown key: " + word + "\n");
  enum { PRIMARY, SECONDARY, DONE } reading = PRIMARY;
  while (reading != DONE && (in >> word)) {
    if (reading == PRIMARY && (word == "secondary_blocks" || word == "slave_blocks"))
      reading = SECONDARY;
    else if (reading == SECONDARY && word == "penalty_parameter")
      reading = DONE;
    else
      (reading == PRIMARY ? primary_block_names : secondary_block_names).push_back(word);
  }
  if (reading == PRIMARY)
    throw std::invalid_argument("\n**** Error processing contact command, expected \"secondary_blocks\" or \"slave_blocks\" (deprectated).\n");
  if (reading == SECONDARY) throw std::invalid_argument("\n**** Error processing contact command, expected \"penalty_parameter\".\n");
  in >> penalty_parameter;
}

void
ContactManager::SkinBlocks(GenesisMesh const& mesh, std::vector<int> const& block_ids, int entity_id_offset,
                           std::vector<std::vector<int>>& skin_faces, std::vector<int>& entity_ids)
{
  // Exodus hex8 face ordinal -> local nodes
  static const int kFaceNodes[6][4] = {{0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {0, 4, 7, 3}, {0, 3, 2, 1}, {4, 5, 6, 7}};
  // Every face of every element with its node ids sorted: faces that occur once are skin.  The reference counts them in
  // a std::map keyed by the sorted ids and walks the map (src/nimble_contact_manager.cc:797-933), so its skin comes out
  // in the lexicographic order of those keys, each face listed as the FIRST element that showed it lists it.  A sort of
  // a flat array gives the same order without a node allocation per face (48 M faces for an 8 M-element block); the
  // position in the array breaks ties so that "first" keeps its meaning.
  struct Face
  {
    std::array<int, 4> key, nodes;
    int                element_id, ordinal;
    long long          position;
  };
  std::vector<Face> faces;
  long long         total = 0;
  for (int block_id : block_ids) total += 6LL * mesh.GetNumElementsInBlock(block_id);
  faces.reserve((size_t)total);
  for (int block_id : block_ids) {
    const int        n_elem = mesh.GetNumElementsInBlock(block_id);
    const int        npe    = mesh.GetNumNodesPerElement(block_id);
    const int* const conn   = mesh.GetConnectivity(block_id);
    const auto&      gid    = mesh.GetElementGlobalIdsInBlock(block_id);
    if (n_elem > 0 && npe != 8) throw std::invalid_argument("\nError in ContactManager::SkinBlocks(), contact blocks must be hex8 blocks.\n");
    for (int e = 0; e < n_elem; ++e)
      for (int ordinal = 0; ordinal < 6; ++ordinal) {
        Face f;
        for (int k = 0; k < 4; ++k) f.nodes[k] = conn[8 * e + kFaceNodes[ordinal][k]];
        f.key = f.nodes;
        std::sort(f.key.begin(), f.key.end());
        f.element_id = gid[e] + 1;  // 1-based: a valid Exodus id in the contact visualisation output
        f.ordinal    = ordinal;
        f.position   = (long long)faces.size();
        faces.push_back(f);
      }
  }
  std::sort(faces.begin(), faces.end(), [](const Face& a, const Face& b) { return a.key != b.key ? a.key < b.key : a.position < b.position; });
  skin_faces.clear();
  entity_ids.clear();
  for (size_t i = 0; i < faces.size();) {
    size_t j = i + 1;
    while (j < faces.size() && faces[j].key == faces[i].key) ++j;
    const Face& o = faces[i];
    if (j - i == 1) {
      skin_faces.emplace_back(o.nodes.begin(), o.nodes.end());
      entity_ids.push_back(((o.element_id + entity_id_offset) << 5) | (o.ordinal << 2));  // 2 low bits: triangle ordinal, set later
    } else if (j - i != 2) {
      throw std::runtime_error("Error in mesh skinning routine, face found more than two times!\n");
    }
    i = j;
  }
}

void
ContactManager::BuildEntityLists(GenesisMesh const& mesh, std::vector<int> const& primary_block_ids, std::vector<int> const& secondary_block_ids,
                                 ContactEntityLists& lists)
{
  auto& primary_face_nodes_      = lists.primary_face_nodes;
  auto& primary_face_entity_ids_ = lists.primary_face_entity_ids;
  auto& primary_face_char_len_   = lists.primary_face_char_len;
  auto& contact_node_ids_        = lists.contact_node_ids;
  auto& contact_node_char_len_   = lists.contact_node_char_len;
  const double* x = mesh.GetCoordinatesX();
  const double* y = mesh.GetCoordinatesY();
  const double* z = mesh.GetCoordinatesZ();
  std::vector<std::vector<int>> primary, secondary;
  std::vector<int>              secondary_entity_ids;
  const int                     offset = mesh.GetMaxNodeGlobalId();  // no entity id is shared by a node and a face
  SkinBlocks(mesh, primary_block_ids, offset, primary, primary_face_entity_ids_);
  SkinBlocks(mesh, secondary_block_ids, offset, secondary, secondary_entity_ids);
  // squared edge length in the model configuration, summed x, y, z as the reference does (:295-302, :1066-1073)
  auto edge2 = [&](int a, int b) { return (x[b] - x[a]) * (x[b] - x[a]) + (y[b] - y[a]) * (y[b] - y[a]) + (z[b] - z[a]) * (z[b] - z[a]); };
  primary_face_nodes_.clear(), primary_face_char_len_.clear(), contact_node_ids_.clear(), contact_node_char_len_.clear();
  for (auto const& face : primary) {
    double longest = std::numeric_limits<double>::lowest();
    for (int i = 0; i < 4; ++i) longest = std::max(longest, std::sqrt(edge2(face[i], face[(i + 1) % 4])));
    primary_face_nodes_.insert(primary_face_nodes_.end(), face.begin(), face.end());
    primary_face_char_len_.push_back(longest);
  }
  // contact nodes in the order the secondary faces first show them; a node keeps the largest length of its faces
  std::map<int, std::size_t> position;
  for (auto const& face : secondary) {
    double longest2 = std::numeric_limits<double>::lowest();
    for (int i = 0; i < 4; ++i) longest2 = std::max(longest2, edge2(face[i], face[(i + 1) % 4]));
    const double len = std::sqrt(longest2);
    for (int node : face) {
      auto at = position.find(node);
      if (at == position.end()) {
        position[node] = contact_node_ids_.size();
        contact_node_ids_.push_back(node);
        contact_node_char_len_.push_back(len);
      } else if (contact_node_char_len_[at->second] < len) {
        contact_node_char_len_[at->second] = len;
      }
    }
  }
}

void
ContactManager::CreateContactEntities(GenesisMesh const& mesh, VectorCommunicator& vector_communicator, std::vector<int> const& primary_block_ids,
                                      std::vector<int> const& secondary_block_ids)
{
  if (vector_communicator.NumRanks() > 1)
    throw std::invalid_argument(
        "\nError: contact across mesh partitions (the reference's ghost-face exchange, src/contact/parallel) is outside the B200 "
        "hex8 path; run decks with a `contact:` line on one GPU.\n");
  BuildEntityLists(mesh, primary_block_ids, secondary_block_ids, lists_);
  contact_enabled_ = true;
  auto* model_data = dynamic_cast<ModelData*>(data_manager_.GetModelData().get());
  if (!model_data) throw std::runtime_error("ContactManager needs a nimble_b200::ModelData");
  DeviceContext& d = model_data->Device();
  d.check(nsm_b200_set_contact(d.get(), penalty_parameter_, (int64_t)lists_.primary_face_char_len.size(), lists_.primary_face_nodes.data(),
                               lists_.primary_face_char_len.data(), (int64_t)lists_.contact_node_ids.size(), lists_.contact_node_ids.data(),
                               lists_.contact_node_char_len.data()),
          "ContactManager::CreateContactEntities");
  model_data->SetContactOnDevice(true);
  if (data_manager_.GetParser().GetRankID() == 0) {
    std::cout << "Contact initialization:" << std::endl;
    std::cout << "  number of triangular contact facets (primary blocks): " << numContactFaces() << std::endl;
    std::cout << "  number of contact nodes (secondary blocks): " << numContactNodes() << "\n" << std::endl;
  }
}

void
ContactManager::ComputeContactForce(int, bool, Viewify<2> contact_force)
{
  if (penalty_parameter_ <= 0.0) throw std::invalid_argument("\nError in ComputeContactForce(), invalid penalty_parameter.\n");
  auto*          model_data = dynamic_cast<ModelData*>(data_manager_.GetModelData().get());
  DeviceContext& d          = model_data->Device();
  // the displacement reached the device through ModelData::UpdateWithNewDisplacement, as in the reference
  d.check(nsm_b200_contact_force_host(d.get(), nullptr, contact_force.data()), "ContactManager::ComputeContactForce");
}

std::size_t
ContactManager::numActiveContactFaces() const
{
  auto*   model_data = dynamic_cast<ModelData*>(data_manager_.GetModelData().get());
  int64_t st[4]      = {0, 0, 0, 0};
  model_data->Device().check(nsm_b200_contact_stats(model_data->Device().get(), st), "ContactManager::numActiveContactFaces");
  return (std::size_t)st[2];
}

std::size_t
ContactManager::numActiveContactNodes() const
{
  auto*   model_data = dynamic_cast<ModelData*>(data_manager_.GetModelData().get());
  int64_t st[4]      = {0, 0, 0, 0};
  model_data->Device().check(nsm_b200_contact_stats(model_data->Device().get(), st), "ContactManager::numActiveContactNodes");
  return (std::size_t)st[3];
}

std::shared_ptr<ContactManager>
GetContactManager(DataManager& data_manager)
{
  if (!data_manager.GetParser().HasContact()) return nullptr;
  return std::make_shared<ContactManager>(data_manager);
}

}  // namespace nimble_b200
