// oracle/ref_contact.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See ref_contact.h.
#include "ref_contact.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <limits>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>

#include "nimble_utils.h"

namespace nsm_oracle {

namespace {

using nimble::ContactEntity;

struct SkinFace
{
  std::array<int, 4> nodes;      // Exodus face order of the element that owns the face
  int                entity_id;  // (element global id + 1 + offset) << 5 | face ordinal << 2
};

// ContactManager::SkinBlocks (nimble_contact_manager.cc:788-934): faces that occur once in the listed blocks, in the
// lexicographic order of their sorted node lists (the reference's std::map iteration order).
std::vector<SkinFace>
skin_blocks(const nimble::GenesisMesh& mesh, const std::vector<int>& block_ids, int entity_id_offset)
{
  static const int face_nodes[6][4] = {{0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {0, 4, 7, 3}, {0, 3, 2, 1}, {4, 5, 6, 7}};
  struct Seen
  {
    int                count;
    std::array<int, 4> nodes;
    int                elem, ordinal;
  };
  std::map<std::array<int, 4>, Seen> table;
  for (int block_id : block_ids) {
    const int         n_elem = mesh.GetNumElementsInBlock(block_id);
    const int         npe    = mesh.GetNumNodesPerElement(block_id);
    const int* const  conn   = mesh.GetConnectivity(block_id);
    const auto&       gids   = mesh.GetElementGlobalIdsInBlock(block_id);
    for (int e = 0; e < n_elem; ++e)
      for (int f = 0; f < 6; ++f) {
        std::array<int, 4> nodes, key;
        for (int k = 0; k < 4; ++k) key[k] = nodes[k] = conn[e * npe + face_nodes[f][k]];
        std::sort(key.begin(), key.end());
        auto it = table.find(key);
        if (it == table.end())
          table[key] = Seen{1, nodes, gids[e] + 1, f};
        else
          it->second.count += 1;
      }
  }
  std::vector<SkinFace> skin;
  for (const auto& kv : table) {
    if (kv.second.count == 1) {
      int id = (kv.second.elem + entity_id_offset) << 5;
      id |= kv.second.ordinal << 2;
      skin.push_back(SkinFace{kv.second.nodes, id});
    } else if (kv.second.count != 2) {
      throw std::runtime_error("Error in mesh skinning routine, face found more than two times!");
    }
  }
  return skin;
}

// ContactManager::Projection (nimble_contact_manager.cc:1549-1620), tolerance 1.e-8 (nimble_contact_manager.h:297)
void
projection(const ContactEntity& node, const ContactEntity& tri, bool& in, double& gap, double* normal, double* barycentric_coordinates)
{
  const double tol  = 1.e-8;
  const double p[3] = {node.coord_1_x_, node.coord_1_y_, node.coord_1_z_};
  const double p1[3] = {tri.coord_1_x_, tri.coord_1_y_, tri.coord_1_z_};
  const double p2[3] = {tri.coord_2_x_, tri.coord_2_y_, tri.coord_2_z_};
  const double p3[3] = {tri.coord_3_x_, tri.coord_3_y_, tri.coord_3_z_};
  double       u[3], v[3], w[3], n[3], cross[3];
  for (int i = 0; i < 3; i++) {
    u[i] = p2[i] - p1[i];
    v[i] = p3[i] - p1[i];
    w[i] = p[i] - p1[i];
  }
  ::CrossProduct(u, v, n);
  const double n_squared = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
  ::CrossProduct(u, w, cross);
  const double alpha3 = (cross[0] * n[0] + cross[1] * n[1] + cross[2] * n[2]) / n_squared;
  ::CrossProduct(w, v, cross);
  const double alpha2 = (cross[0] * n[0] + cross[1] * n[1] + cross[2] * n[2]) / n_squared;
  const double alpha1 = 1.0 - alpha2 - alpha3;
  const double tol2   = 1.0 + tol;
  const bool   a1 = (alpha1 > -tol && alpha1 < tol2), a2 = (alpha2 > -tol && alpha2 < tol2), a3 = (alpha3 > -tol && alpha3 < tol2);
  in = false;
  if (a1 && a2 && a3) {
    const double xp = alpha1 * p1[0] + alpha2 * p2[0] + alpha3 * p3[0];
    const double yp = alpha1 * p1[1] + alpha2 * p2[1] + alpha3 * p3[1];
    const double zp = alpha1 * p1[2] + alpha2 * p2[2] + alpha3 * p3[2];
    const double dx = node.coord_1_x_ - xp, dy = node.coord_1_y_ - yp, dz = node.coord_1_z_ - zp;
    const double s  = 1.0 / std::sqrt(n_squared);
    normal[0] = n[0] * s, normal[1] = n[1] * s, normal[2] = n[2] * s;
    gap       = dx * normal[0] + dy * normal[1] + dz * normal[2];
    barycentric_coordinates[0] = alpha1, barycentric_coordinates[1] = alpha2, barycentric_coordinates[2] = alpha3;
    if ((gap < 0.0) && (gap > -tri.char_len_)) in = true;
  }
}

// ArborX::intersects on the entities' boxes, which ArborX::Point narrows to float (src/contact/arborx_utils.h:85-90, 106-121)
bool
boxes_intersect(const ContactEntity& a, const ContactEntity& b)
{
  const float alo[3] = {(float)a.bounding_box_x_min_, (float)a.bounding_box_y_min_, (float)a.bounding_box_z_min_};
  const float ahi[3] = {(float)a.bounding_box_x_max_, (float)a.bounding_box_y_max_, (float)a.bounding_box_z_max_};
  const float blo[3] = {(float)b.bounding_box_x_min_, (float)b.bounding_box_y_min_, (float)b.bounding_box_z_min_};
  const float bhi[3] = {(float)b.bounding_box_x_max_, (float)b.bounding_box_y_max_, (float)b.bounding_box_z_max_};
  for (int d = 0; d < 3; ++d)
    if (ahi[d] < blo[d] || alo[d] > bhi[d]) return false;
  return true;
}

}  // namespace

void
RefContact::ParseCommand(const std::string& command, std::vector<std::string>& primary, std::vector<std::string>& secondary, double& penalty)
{
  std::stringstream ss(command);
  std::string       tok;
  ss >> tok;
  if (tok != "primary_blocks" && tok != "master_blocks") throw std::invalid_argument("contact command: unknown key " + tok);
  int stage = 0;  // 0: primary names, 1: secondary names, 2: penalty read
  while (stage < 2 && (ss >> tok)) {
    if (stage == 0 && (tok == "secondary_blocks" || tok == "slave_blocks"))
      stage = 1;
    else if (stage == 1 && tok == "penalty_parameter") {
      ss >> penalty;
      stage = 2;
    } else
      (stage == 0 ? primary : secondary).push_back(tok);
  }
  if (stage != 2) throw std::invalid_argument("contact command: expected secondary_blocks ... penalty_parameter <value>");
}

// ContactManager::CreateContactEntities + CreateContactNodesAndFaces for one rank (nimble_contact_manager.cc:184-393,
// 1043-1205): no partition-boundary faces, no ghosted nodes.
void
RefContact::Create(const nimble::GenesisMesh& mesh, const std::vector<int>& primary_block_ids, const std::vector<int>& secondary_block_ids,
                   double penalty)
{
  penalty_              = penalty;
  const double* cx      = mesh.GetCoordinatesX();
  const double* cy      = mesh.GetCoordinatesY();
  const double* cz      = mesh.GetCoordinatesZ();
  const int     offset  = mesh.GetMaxNodeGlobalId();
  auto          primary = skin_blocks(mesh, primary_block_ids, offset), secondary = skin_blocks(mesh, secondary_block_ids, offset);
  std::set<int> node_set;
  for (auto& f : primary) node_set.insert(f.nodes.begin(), f.nodes.end());
  for (auto& f : secondary) node_set.insert(f.nodes.begin(), f.nodes.end());
  node_ids_.assign(node_set.begin(), node_set.end());
  std::map<int, int> sub_id;
  for (size_t i = 0; i < node_ids_.size(); ++i) sub_id[node_ids_[i]] = (int)i;
  for (auto& f : primary) {
    primary_quads.insert(primary_quads.end(), f.nodes.begin(), f.nodes.end());
    for (int& n : f.nodes) n = sub_id.at(n);
  }
  for (auto& f : secondary) {
    secondary_quads.insert(secondary_quads.end(), f.nodes.begin(), f.nodes.end());
    for (int& n : f.nodes) n = sub_id.at(n);
  }
  const size_t n_sub = node_ids_.size();
  model_coord_.resize(3 * n_sub), coord_.resize(3 * n_sub), force_.assign(3 * n_sub, 0.0);
  for (size_t i = 0; i < n_sub; ++i) {
    model_coord_[3 * i] = coord_[3 * i] = cx[node_ids_[i]];
    model_coord_[3 * i + 1] = coord_[3 * i + 1] = cy[node_ids_[i]];
    model_coord_[3 * i + 2] = coord_[3 * i + 2] = cz[node_ids_[i]];
  }
  auto edge_sq = [&](int a, int b) {
    return (coord_[3 * b] - coord_[3 * a]) * (coord_[3 * b] - coord_[3 * a]) + (coord_[3 * b + 1] - coord_[3 * a + 1]) * (coord_[3 * b + 1] - coord_[3 * a + 1]) +
           (coord_[3 * b + 2] - coord_[3 * a + 2]) * (coord_[3 * b + 2] - coord_[3 * a + 2]);
  };
  // contact nodes of the secondary faces with their characteristic lengths (:281-330)
  const int*            gid = mesh.GetNodeGlobalIds();
  std::vector<int>      sec_nodes, sec_entity;
  std::map<int, double> sec_len;
  for (auto& f : secondary) {
    double max_sq = std::numeric_limits<double>::lowest();
    for (int i = 0; i < 4; ++i) max_sq = std::max(max_sq, edge_sq(f.nodes[i], f.nodes[(i + 1) % 4]));
    const double len = std::sqrt(max_sq);
    for (int n : f.nodes) {
      if (std::find(sec_nodes.begin(), sec_nodes.end(), n) == sec_nodes.end()) {
        sec_nodes.push_back(n);
        sec_entity.push_back(gid[node_ids_[n]] + 1);
        sec_len[n] = len;
      } else if (sec_len[n] < len)
        sec_len[n] = len;
    }
  }
  // four triangles per primary face around a fictitious centre node (:1043-1190)
  faces_.clear(), nodes_.clear();
  for (auto& f : primary) {
    double max_len = std::numeric_limits<double>::lowest();
    for (int i = 0; i < 4; ++i) max_len = std::max(max_len, std::sqrt(edge_sq(f.nodes[i], f.nodes[(i + 1) % 4])));
    double centre[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 3; ++j) centre[j] += coord_[3 * f.nodes[i] + j];
    for (double& j : centre) j /= 4;
    int fict[4] = {f.nodes[0], f.nodes[1], f.nodes[2], f.nodes[3]};
    for (int k = 0; k < 4; ++k) {
      const int n1 = f.nodes[k], n2 = f.nodes[(k + 1) % 4];
      double    mc[9];
      for (int i = 0; i < 3; ++i) mc[i] = coord_[3 * n1 + i], mc[3 + i] = coord_[3 * n2 + i], mc[6 + i] = centre[i];
      faces_.push_back(ContactEntity(ContactEntity::TRIANGLE, f.entity_id | k, (int)faces_.size(), mc, max_len, n1, n2, fict));
    }
    primary_char_len.push_back(max_len);
  }
  for (size_t i = 0; i < sec_nodes.size(); ++i) {
    const int n = sec_nodes[i];
    double    mc[3] = {coord_[3 * n], coord_[3 * n + 1], coord_[3 * n + 2]};
    nodes_.push_back(ContactEntity(ContactEntity::NODE, sec_entity[i], (int)i, mc, sec_len.at(n), n));
    contact_node_ids.push_back(node_ids_[n]);
    contact_node_char_len.push_back(sec_len.at(n));
  }
}

// ArborXSerialContactManager::ComputeSerialContactForce (src/contact/serial/arborx_serial_contact_manager.cc:147-196)
// with the BVH query replaced by a walk over all node-face pairs.
long
RefContact::Compute(const double* displacement, double* contact_force, int n_mesh_nodes)
{
  std::fill(contact_force, contact_force + 3 * (size_t)n_mesh_nodes, 0.0);
  // ContactManager::ApplyDisplacements (nimble_contact_manager.cc:750-786)
  for (size_t i = 0; i < node_ids_.size(); ++i)
    for (int d = 0; d < 3; ++d) coord_[3 * i + d] = model_coord_[3 * i + d] + displacement[3 * node_ids_[i] + d];
  for (auto& e : nodes_) e.SetCoordinates(coord_);
  for (auto& e : faces_) e.SetCoordinates(coord_);
  std::fill(force_.begin(), force_.end(), 0.0);  // ZeroContactForce
  for (auto& e : faces_) e.set_contact_status(false);
  for (auto& e : nodes_) e.set_contact_status(false);
  long pairs = 0;
  for (auto& node : nodes_)
    for (auto& face : faces_) {
      if (!boxes_intersect(node, face)) continue;
      double gap = 0.0, normal[3] = {0., 0., 0.}, facet_coordinates[3] = {0., 0., 0.};
      bool   inside = false;
      projection(node, face, inside, gap, normal, facet_coordinates);
      if (!inside) continue;
      face.set_contact_status(true);
      node.set_contact_status(true);
      ++pairs;
      // PenaltyContactEnforcement::EnforceContact (nimble_contact_manager.h:94-128) on scratch entities
      double       cf[3];
      const double scale = penalty_ * gap;  // details::getContactForce (:80-85)
      for (int i = 0; i < 3; ++i) cf[i] = scale * normal[i];
      ContactEntity tmp_face;
      tmp_face.entity_type_                   = face.entity_type_;
      tmp_face.node_id_for_node_1_            = face.node_id_for_node_1_;
      tmp_face.node_id_for_node_2_            = face.node_id_for_node_2_;
      tmp_face.node_id_1_for_fictitious_node_ = face.node_id_1_for_fictitious_node_;
      tmp_face.node_id_2_for_fictitious_node_ = face.node_id_2_for_fictitious_node_;
      tmp_face.node_id_3_for_fictitious_node_ = face.node_id_3_for_fictitious_node_;
      tmp_face.node_id_4_for_fictitious_node_ = face.node_id_4_for_fictitious_node_;
      tmp_face.SetNodalContactForces(cf, facet_coordinates);
      tmp_face.ScatterForceToContactManagerForceVector(force_);
      ContactEntity tmp_node;
      tmp_node.entity_type_        = ContactEntity::NODE;
      tmp_node.node_id_for_node_1_ = node.node_id_for_node_1_;
      tmp_node.SetNodalContactForces(cf);
      tmp_node.ScatterForceToContactManagerForceVector(force_);
    }
  // ContactManager::GetForces (nimble_contact_manager.cc:732-748)
  for (size_t i = 0; i < node_ids_.size(); ++i)
    for (int d = 0; d < 3; ++d) contact_force[3 * node_ids_[i] + d] = force_[3 * i + d];
  return pairs;
}

}  // namespace nsm_oracle
