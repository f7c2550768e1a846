// nimblesm_b200/host/contact_manager.cc — see contact_manager.h.
#include "contact_manager.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
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

namespace {

// Entity lists from skin faces (node ids of any numbering) and a coordinate lookup: characteristic lengths in the
// reference's operation order (src/nimble_contact_manager.cc:288-330, 1062-1077), contact nodes in the order the
// secondary faces first show them, each with the largest length of its faces.
template <class Coord>
void
lists_from_faces(std::vector<std::vector<int>> const& primary, std::vector<std::vector<int>> const& secondary, Coord const& xyz,
                 ContactEntityLists& lists)
{
  auto edge2 = [&](int a, int b) {
    const double *pa = xyz(a), *pb = xyz(b);
    return (pb[0] - pa[0]) * (pb[0] - pa[0]) + (pb[1] - pa[1]) * (pb[1] - pa[1]) + (pb[2] - pa[2]) * (pb[2] - pa[2]);
  };
  lists.primary_face_nodes.clear(), lists.primary_face_char_len.clear(), lists.contact_node_ids.clear(), lists.contact_node_char_len.clear();
  for (auto const& face : primary) {
    double longest = std::numeric_limits<double>::lowest();
    for (int i = 0; i < 4; ++i) longest = std::max(longest, std::sqrt(edge2(face[i], face[(i + 1) % 4])));
    lists.primary_face_nodes.insert(lists.primary_face_nodes.end(), face.begin(), face.end());
    lists.primary_face_char_len.push_back(longest);
  }
  std::map<int, std::size_t> position;
  for (auto const& face : secondary) {
    double longest2 = std::numeric_limits<double>::lowest();
    for (int i = 0; i < 4; ++i) longest2 = std::max(longest2, edge2(face[i], face[(i + 1) % 4]));
    const double len = std::sqrt(longest2);
    for (int node : face) {
      auto at = position.find(node);
      if (at == position.end()) {
        position[node] = lists.contact_node_ids.size();
        lists.contact_node_ids.push_back(node);
        lists.contact_node_char_len.push_back(len);
      } else if (lists.contact_node_char_len[at->second] < len) {
        lists.contact_node_char_len[at->second] = len;
      }
    }
  }
}

}  // namespace

void
ContactManager::BuildEntityLists(GenesisMesh const& mesh, std::vector<int> const& primary_block_ids, std::vector<int> const& secondary_block_ids,
                                 ContactEntityLists& lists)
{
  const double* x = mesh.GetCoordinatesX();
  const double* y = mesh.GetCoordinatesY();
  const double* z = mesh.GetCoordinatesZ();
  std::vector<std::vector<int>> primary, secondary;
  std::vector<int>              secondary_entity_ids;
  const int                     offset = mesh.GetMaxNodeGlobalId();  // no entity id is shared by a node and a face
  SkinBlocks(mesh, primary_block_ids, offset, primary, lists.primary_face_entity_ids);
  SkinBlocks(mesh, secondary_block_ids, offset, secondary, secondary_entity_ids);
  std::vector<double> point(3 * (size_t)mesh.GetNumNodes());
  for (size_t i = 0; i < (size_t)mesh.GetNumNodes(); ++i) point[3 * i] = x[i], point[3 * i + 1] = y[i], point[3 * i + 2] = z[i];
  lists_from_faces(primary, secondary, [&](int n) { return point.data() + 3 * (size_t)n; }, lists);
}

// Contact across mesh partitions: the REPLICATED sub-model.  The reference exchanges ghost faces between ranks and
// searches a distributed tree (src/contact/parallel/arborx_parallel_contact_manager.cc).  A contact surface is small
// (two-dimensional), so here every rank learns the whole of it once -- its own skin faces with global node ids and
// coordinates travel through the rank group, faces that two ranks list are partition cuts and drop out
// (RemoveInternalSkinFaces, src/nimble_contact_manager.cc:937-1041), the rest is sorted by global ids, which is the
// order a serial run's skin has -- and rank 0 owns a second, element-free device context whose nodes are the surface
// nodes.  Per step the ranks pool the displacement of the surface nodes they hold, rank 0 evaluates the contact force
// of the whole surface on its GPU, and every rank picks the entries of its own nodes: all holders of a shared node get
// the SAME bits (one evaluation), which is what the reference's clique all-reduce of contact_force guarantees.
void
ContactManager::BuildReplicatedSubModel(GenesisMesh const& mesh, VectorCommunicator& vc, std::vector<int> const& primary_block_ids,
                                        std::vector<int> const& secondary_block_ids, ReplicatedContactSubModel& out)
{
  auto& surface_xyz_  = out.surface_xyz;
  auto& held_local_   = out.held_local;
  auto& held_surface_ = out.held_surface;
  auto& lists_        = out.lists;
  struct Wire  // one skin face on the wire
  {
    int    gid[4];
    int    secondary;
    double xyz[12];
  };
  const int*    gid = mesh.GetNodeGlobalIds();
  const double* x   = mesh.GetCoordinatesX();
  const double* y   = mesh.GetCoordinatesY();
  const double* z   = mesh.GetCoordinatesZ();
  std::vector<Wire> mine;
  for (int side = 0; side < 2; ++side) {
    std::vector<std::vector<int>> faces;
    std::vector<int>              ids;
    SkinBlocks(mesh, side ? secondary_block_ids : primary_block_ids, 0, faces, ids);
    for (auto const& f : faces) {
      Wire w;
      w.secondary = side;
      for (int k = 0; k < 4; ++k) {
        w.gid[k] = gid[f[k]];
        w.xyz[3 * k] = x[f[k]], w.xyz[3 * k + 1] = y[f[k]], w.xyz[3 * k + 2] = z[f[k]];
      }
      mine.push_back(w);
    }
  }
  std::vector<char> blob(mine.size() * sizeof(Wire));
  if (!mine.empty()) memcpy(blob.data(), mine.data(), blob.size());
  const auto all = vc.Group()->AllGather(vc.Rank(), blob);
  struct Merged
  {
    std::array<int, 4> key;
    Wire               w;
  };
  std::vector<Merged> merged;
  for (auto const& b : all) {
    const size_t n = b.size() / sizeof(Wire);
    for (size_t i = 0; i < n; ++i) {
      Merged m;
      memcpy(&m.w, b.data() + i * sizeof(Wire), sizeof(Wire));
      for (int k = 0; k < 4; ++k) m.key[k] = m.w.gid[k];
      std::sort(m.key.begin(), m.key.end());
      merged.push_back(m);
    }
  }
  std::stable_sort(merged.begin(), merged.end(), [](const Merged& a, const Merged& b) {
    return a.w.secondary != b.w.secondary ? a.w.secondary < b.w.secondary : a.key < b.key;
  });
  // surface nodes = nodes of the faces that survive, ascending global id
  std::map<int, std::array<double, 3>> node_xyz;
  std::vector<const Merged*>           kept;
  for (size_t i = 0; i < merged.size();) {
    size_t j = i + 1;
    while (j < merged.size() && merged[j].w.secondary == merged[i].w.secondary && merged[j].key == merged[i].key) ++j;
    if (j - i == 1) {
      kept.push_back(&merged[i]);
      for (int k = 0; k < 4; ++k) node_xyz[merged[i].w.gid[k]] = {merged[i].w.xyz[3 * k], merged[i].w.xyz[3 * k + 1], merged[i].w.xyz[3 * k + 2]};
    } else if (j - i != 2) {
      throw std::runtime_error("Error in mesh skinning routine, face found more than two times!\n");
    }
    i = j;
  }
  std::map<int, int> surface_of_gid;
  surface_xyz_.clear();
  out.surface_gid.clear();
  for (auto const& kv : node_xyz) {
    out.surface_gid.push_back(kv.first);
    surface_of_gid[kv.first] = (int)surface_of_gid.size();
    surface_xyz_.insert(surface_xyz_.end(), kv.second.begin(), kv.second.end());
  }
  std::vector<std::vector<int>> primary, secondary;
  for (const Merged* m : kept) {
    std::vector<int> f(4);
    for (int k = 0; k < 4; ++k) f[k] = surface_of_gid.at(m->w.gid[k]);
    (m->w.secondary ? secondary : primary).push_back(f);
  }
  lists_from_faces(primary, secondary, [&](int n) { return surface_xyz_.data() + 3 * (size_t)n; }, lists_);
  lists_.primary_face_entity_ids.assign(primary.size(), 0);
  // the surface nodes this rank holds: (local node, surface index)
  held_local_.clear(), held_surface_.clear();
  for (int i = 0; i < (int)mesh.GetNumNodes(); ++i) {
    auto at = surface_of_gid.find(gid[i]);
    if (at != surface_of_gid.end()) held_local_.push_back(i), held_surface_.push_back(at->second);
  }
}

void
ContactManager::CreateContactEntities(GenesisMesh const& mesh, VectorCommunicator& vector_communicator, std::vector<int> const& primary_block_ids,
                                      std::vector<int> const& secondary_block_ids)
{
  auto* model_data = dynamic_cast<ModelData*>(data_manager_.GetModelData().get());
  if (!model_data) throw std::runtime_error("ContactManager needs a nimble_b200::ModelData");
  if (vector_communicator.NumRanks() > 1) {
    group_ = vector_communicator.Group();
    rank_  = vector_communicator.Rank();
    ReplicatedContactSubModel sub;
    BuildReplicatedSubModel(mesh, vector_communicator, primary_block_ids, secondary_block_ids, sub);
    lists_        = std::move(sub.lists);
    surface_xyz_  = std::move(sub.surface_xyz);
    held_local_   = std::move(sub.held_local);
    held_surface_ = std::move(sub.held_surface);
    replicated_      = true;
    contact_enabled_ = true;
    if (rank_ == 0) {
      // the contact sub-model on rank 0's GPU: surface nodes, no elements
      const size_t        n = surface_xyz_.size() / 3;
      std::vector<double> sx(n), sy(n), sz(n);
      for (size_t i = 0; i < n; ++i) sx[i] = surface_xyz_[3 * i], sy[i] = surface_xyz_[3 * i + 1], sz[i] = surface_xyz_[3 * i + 2];
      sub_model_.reset(new DeviceContext(model_data->DeviceIndex()));
      DeviceContext& d = *sub_model_;
      d.check(nsm_b200_set_nodes(d.get(), (int64_t)n, sx.data(), sy.data(), sz.data()), "ContactManager (contact sub-model nodes)");
      // (the model's assembly mode: ORDERED sums the contact force in the serial order, bit-reproducible)
      d.check(nsm_b200_finalize(d.get(), model_data->Assembly(), 0), "ContactManager (contact sub-model)");
      d.check(nsm_b200_set_contact(d.get(), penalty_parameter_, (int64_t)lists_.primary_face_char_len.size(), lists_.primary_face_nodes.data(),
                                   lists_.primary_face_char_len.data(), (int64_t)lists_.contact_node_ids.size(), lists_.contact_node_ids.data(),
                                   lists_.contact_node_char_len.data()),
              "ContactManager::CreateContactEntities (contact sub-model)");
      std::cout << "Contact initialization:" << std::endl;
      std::cout << "  number of triangular contact facets (primary blocks): " << numContactFaces() << std::endl;
      std::cout << "  number of contact nodes (secondary blocks): " << numContactNodes() << std::endl;
      std::cout << "  " << vector_communicator.NumRanks() << " ranks: the contact surface (" << n
                << " nodes) is replicated and evaluated on rank 0's GPU\n"
                << std::endl;
    }
    return;
  }
  BuildEntityLists(mesh, primary_block_ids, secondary_block_ids, lists_);
  mesh_            = &mesh;
  contact_enabled_ = true;
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
  auto* model_data = dynamic_cast<ModelData*>(data_manager_.GetModelData().get());
  if (replicated_) {
    // pool the displacement of the surface nodes (replicas of a shared node carry the same bits: any holder may write)
    Viewify<2>        displacement = model_data->GetVectorNodeData("displacement");
    const size_t      n_held = held_local_.size(), n_surface = surface_xyz_.size() / 3;
    std::vector<char> blob(n_held * (sizeof(int) + 3 * sizeof(double)));
    for (size_t i = 0; i < n_held; ++i) {
      char* at = blob.data() + i * (sizeof(int) + 3 * sizeof(double));
      memcpy(at, &held_surface_[i], sizeof(int));
      const double u[3] = {displacement(held_local_[i], 0), displacement(held_local_[i], 1), displacement(held_local_[i], 2)};
      memcpy(at + sizeof(int), u, sizeof u);
    }
    const auto        pooled = group_->AllGather(rank_, blob);
    std::vector<char> force_blob;
    if (rank_ == 0) {
      std::vector<double> u(3 * n_surface, 0.0);
      for (auto const& b : pooled)
        for (size_t at = 0; at + sizeof(int) + 3 * sizeof(double) <= b.size(); at += sizeof(int) + 3 * sizeof(double)) {
          int idx;
          memcpy(&idx, b.data() + at, sizeof(int));
          memcpy(&u[3 * (size_t)idx], b.data() + at + sizeof(int), 3 * sizeof(double));
        }
      force_blob.resize(3 * n_surface * sizeof(double));
      sub_model_->check(nsm_b200_contact_force_host(sub_model_->get(), u.data(), reinterpret_cast<double*>(force_blob.data())),
                        "ContactManager::ComputeContactForce (contact sub-model)");
    }
    const auto    forces = group_->AllGather(rank_, force_blob);
    const double* fc     = reinterpret_cast<const double*>(forces[0].data());
    const int     n_local = contact_force.size()[0];
    for (int i = 0; i < n_local; ++i) contact_force(i, 0) = contact_force(i, 1) = contact_force(i, 2) = 0.0;
    for (size_t i = 0; i < n_held; ++i)
      for (int c = 0; c < 3; ++c) contact_force(held_local_[i], c) = fc[3 * (size_t)held_surface_[i] + c];
    return;
  }
  DeviceContext& d = model_data->Device();
  // the displacement reached the device through ModelData::UpdateWithNewDisplacement, as in the reference
  d.check(nsm_b200_contact_force_host(d.get(), nullptr, contact_force.data()), "ContactManager::ComputeContactForce");
}

void
ContactVisualizationDatabase::EntityVertices(const double* displacement, std::vector<double>& x, std::vector<double>& y, std::vector<double>& z) const
{
  const double* const mx[3] = {model_mesh_.GetCoordinatesX(), model_mesh_.GetCoordinatesY(), model_mesh_.GetCoordinatesZ()};
  const size_t        n_faces = lists_.primary_face_char_len.size(), n_cn = lists_.contact_node_ids.size();
  std::vector<double>* const out[3] = {&x, &y, &z};
  for (int d = 0; d < 3; ++d) {
    out[d]->resize(12 * n_faces + n_cn);
    auto at = [&](int node) { return displacement ? mx[d][node] + displacement[3 * (size_t)node + d] : mx[d][node]; };
    size_t k = 0;
    for (size_t f = 0; f < n_faces; ++f) {
      const int*   q      = &lists_.primary_face_nodes[4 * f];
      const double centre = (at(q[0]) + at(q[1]) + at(q[2]) + at(q[3])) / 4.0;
      for (int t = 0; t < 4; ++t) {
        (*out[d])[k++] = at(q[t]);
        (*out[d])[k++] = at(q[(t + 1) % 4]);
        (*out[d])[k++] = centre;
      }
    }
    for (size_t i = 0; i < n_cn; ++i) (*out[d])[k++] = at(lists_.contact_node_ids[i]);
  }
}

void
ContactManager::InitializeContactVisualization(std::string const& contact_visualization_exodus_file_name)
{
  if (replicated_ || !mesh_) {
    if (rank_ == 0)
      std::cout << "(contact visualization: written by single-rank runs only; skipped for the replicated contact surface)" << std::endl;
    return;
  }
  visualization_.reset(new ContactVisualizationDatabase(*mesh_, lists_, contact_visualization_exodus_file_name));
}

ContactVisualizationDatabase::ContactVisualizationDatabase(GenesisMesh const& model_mesh, ContactEntityLists const& lists,
                                                           std::string const& contact_visualization_exodus_file_name)
    : model_mesh_(model_mesh), lists_(lists)
{
  const size_t n_faces = lists_.primary_face_char_len.size(), n_cn = lists_.contact_node_ids.size();
  // entity ids (src/nimble_contact_manager.cc:924-927, 1106-1176, 321): facet = skin entity id | triangle ordinal,
  // contact node = global node id + 1
  std::vector<int> face_ids(4 * n_faces), node_ids(n_cn);
  int              max_contact_entity_id = 0;
  for (size_t f = 0; f < n_faces; ++f)
    for (int t = 0; t < 4; ++t) {
      face_ids[4 * f + t]   = lists_.primary_face_entity_ids[f] | t;
      max_contact_entity_id = std::max(max_contact_entity_id, face_ids[4 * f + t]);
    }
  const int* const gid = model_mesh_.GetNodeGlobalIds();
  for (size_t i = 0; i < n_cn; ++i) {
    node_ids[i]           = gid[lists_.contact_node_ids[i]] + 1;
    max_contact_entity_id = std::max(max_contact_entity_id, node_ids[i]);
  }
  std::vector<int>                node_global_id, elem_global_id, block_ids = {1, 2};
  std::vector<double>             node_x, node_y, node_z;
  std::map<int, std::string>      block_names              = {{1, "contact_faces"}, {2, "contact_nodes"}};
  std::map<int, std::vector<int>> block_elem_global_ids    = {{1, {}}, {2, {}}};
  std::map<int, int>              block_num_nodes_per_elem = {{1, 3}, {2, 1}};
  std::map<int, std::vector<int>> block_elem_connectivity  = {{1, {}}, {2, {}}};
  EntityVertices(nullptr, node_x, node_y, node_z);
  int node_index = 0;
  for (size_t i = 0; i < 4 * n_faces; ++i) {
    for (int v = 0; v < 3; ++v) {
      node_global_id.push_back(3 * face_ids[i] + max_contact_entity_id + 9 + v);  // :525-535
      block_elem_connectivity[1].push_back(node_index++);
    }
    elem_global_id.push_back(face_ids[i]);
  }
  for (size_t i = 0; i < n_cn; ++i) {
    node_global_id.push_back(node_ids[i]);
    block_elem_connectivity[2].push_back(node_index++);
    elem_global_id.push_back(node_ids[i]);
  }
  GenesisMesh&  mesh = mesh_;
  ExodusOutput& out  = out_;
  mesh.Initialize("contact_visualization", node_global_id, node_x, node_y, node_z, elem_global_id, block_ids, block_names, block_elem_global_ids,
                  block_num_nodes_per_elem, block_elem_connectivity);
  out.Initialize(contact_visualization_exodus_file_name, mesh);
  std::map<int, std::vector<std::string>> no_elem_data = {{1, {}}, {2, {}}};
  out.InitializeDatabase(mesh, {"num_contacts"}, {"displacement_x", "displacement_y", "displacement_z", "contact_status"}, no_elem_data, no_elem_data);
}

void
ContactVisualizationDatabase::WriteStep(double t, const double* displacement, const unsigned char* face_status, const unsigned char* node_status)
{
  const GenesisMesh&  mesh    = mesh_;
  const size_t        n_faces = lists_.primary_face_char_len.size(), n_cn = lists_.contact_node_ids.size();
  const double* const model[3] = {mesh.GetCoordinatesX(), mesh.GetCoordinatesY(), mesh.GetCoordinatesZ()};
  std::vector<std::vector<double>> node_data(4);
  EntityVertices(displacement, node_data[0], node_data[1], node_data[2]);
  for (int d = 0; d < 3; ++d)
    for (size_t i = 0; i < node_data[d].size(); ++i) node_data[d][i] -= model[d][i];  // (:644-668)
  node_data[3].assign(12 * n_faces + n_cn, 0.0);
  double num_contacts = 0.0;
  for (size_t i = 0; i < 4 * n_faces; ++i) {
    const double status = face_status && face_status[i] ? 1.0 : 0.0;
    num_contacts += status;  // numActiveContactFaces (:692-702)
    for (int v = 0; v < 3; ++v) node_data[3][3 * i + v] = status;
  }
  for (size_t i = 0; i < n_cn; ++i) node_data[3][12 * n_faces + i] = node_status && node_status[i] ? 1.0 : 0.0;
  std::map<int, std::vector<std::string>>         no_labels = {{1, {}}, {2, {}}};
  std::map<int, std::vector<std::vector<double>>> no_data   = {{1, {}}, {2, {}}};
  out_.WriteStep(t, {num_contacts}, node_data, no_labels, no_data, no_labels, no_data);
}

void
ContactManager::ContactVisualizationWriteStep(double time_current, bool evaluated)
{
  if (!visualization_) return;
  if (!evaluated) {
    visualization_->WriteStep(time_current, nullptr, nullptr, nullptr);
    return;
  }
  auto*                      model_data = dynamic_cast<ModelData*>(data_manager_.GetModelData().get());
  Viewify<2>                 displacement = model_data->GetVectorNodeData("displacement");
  std::vector<unsigned char> face_status(numContactFaces()), node_status(numContactNodes());
  DeviceContext&             d = model_data->Device();
  d.check(nsm_b200_contact_status(d.get(), face_status.data(), node_status.data()), "ContactManager::ContactVisualizationWriteStep");
  visualization_->WriteStep(time_current, displacement.data(), face_status.data(), node_status.data());
}

std::size_t
ContactManager::numActiveContactFaces() const
{
  auto*   model_data = dynamic_cast<ModelData*>(data_manager_.GetModelData().get());
  int64_t st[5]      = {0, 0, 0, 0, 0};
  if (replicated_) {
    if (sub_model_) sub_model_->check(nsm_b200_contact_stats(sub_model_->get(), st), "ContactManager::numActiveContactFaces");
    return (std::size_t)st[2];
  }
  model_data->Device().check(nsm_b200_contact_stats(model_data->Device().get(), st), "ContactManager::numActiveContactFaces");
  return (std::size_t)st[2];
}

std::size_t
ContactManager::numActiveContactNodes() const
{
  auto*   model_data = dynamic_cast<ModelData*>(data_manager_.GetModelData().get());
  int64_t st[5]      = {0, 0, 0, 0, 0};
  if (replicated_) {
    if (sub_model_) sub_model_->check(nsm_b200_contact_stats(sub_model_->get(), st), "ContactManager::numActiveContactNodes");
    return (std::size_t)st[3];
  }
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
