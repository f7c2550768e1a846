// nimblesm_b200/host/genesis_mesh.cc — see genesis_mesh.h.  Variable names follow the Exodus II NetCDF
// schema found in the reference's test meshes.
#include "genesis_mesh.h"

#include <algorithm>

#include <array>

#include <iostream>
#include <sstream>
#include <stdexcept>

#include "netcdf3.h"

namespace nimble_b200 {

namespace {
std::string
numbered(const char* stem, int k)
{
  std::ostringstream s;
  s << stem << k;
  return s.str();
}
}  // namespace

void
GenesisMesh::ReadFile(std::string file_name)
{
  file_name_ = file_name;
  if (!IsValid()) return;
  nc3::Reader f(file_name);
  dim_                    = (int)f.dim("num_dim");
  const int num_nodes     = (int)f.dim("num_nodes");
  const int num_elem      = (int)f.dim_or("num_elem", 0);
  const int num_blocks    = (int)f.dim_or("num_el_blk", 0);
  const int num_node_sets = (int)f.dim_or("num_node_sets", 0);

  // coordinates: coordx/coordy/coordz, or the older single "coord" [num_dim][num_nodes]
  if (f.has_var("coordx")) {
    node_x_ = f.read_double("coordx");
    node_y_ = dim_ > 1 && f.has_var("coordy") ? f.read_double("coordy") : std::vector<double>(num_nodes, 0.0);
    node_z_ = dim_ > 2 && f.has_var("coordz") ? f.read_double("coordz") : std::vector<double>();
  } else if (f.has_var("coord")) {
    std::vector<double> c = f.read_double("coord");
    node_x_.assign(c.begin(), c.begin() + num_nodes);
    node_y_.assign(c.begin() + num_nodes, c.begin() + 2 * num_nodes);
    if (dim_ > 2) node_z_.assign(c.begin() + 2 * num_nodes, c.begin() + 3 * num_nodes);
  } else {
    node_x_.assign(num_nodes, 0.0), node_y_.assign(num_nodes, 0.0);
    if (dim_ > 2) node_z_.assign(num_nodes, 0.0);
  }

  // global ids, 1-based on file (src/nimble_genesis_mesh.cc:97-111); identity when the map is absent
  node_global_id_.resize(num_nodes);
  if (f.has_var("node_num_map")) {
    node_global_id_ = f.read_int("node_num_map");
    for (int& g : node_global_id_) g -= 1;
  } else {
    for (int i = 0; i < num_nodes; ++i) node_global_id_[i] = i;
  }
  elem_global_id_.resize(num_elem);
  if (f.has_var("elem_num_map")) {
    elem_global_id_ = f.read_int("elem_num_map");
    for (int& g : elem_global_id_) g -= 1;
  } else {
    for (int i = 0; i < num_elem; ++i) elem_global_id_[i] = i;
  }
  // decomp's auxiliary maps named "original_global_id_map" win (src/nimble_genesis_mesh.cc:113-149)
  const int num_node_maps = (int)f.dim_or("num_node_maps", 0), num_elem_maps = (int)f.dim_or("num_elem_maps", 0);
  if (num_node_maps > 1 || num_elem_maps > 1)
    throw std::runtime_error("GenesisMesh::ReadFile(), multiple auxiliary node/element maps not supported!");
  if (num_node_maps > 0) {
    std::vector<std::string> nm = f.has_var("nmap_names") ? f.read_strings("nmap_names") : std::vector<std::string>();
    if (nm.empty() || nm[0] != "original_global_id_map")
      throw std::runtime_error("GenesisMesh::ReadFile(), unsupported auxiliary node map!");
    node_global_id_ = f.read_int("node_map1");
    for (int& g : node_global_id_) g -= 1;
  }
  if (num_elem_maps > 0) {
    std::vector<std::string> nm = f.has_var("emap_names") ? f.read_strings("emap_names") : std::vector<std::string>();
    if (nm.empty() || nm[0] != "original_global_id_map")
      throw std::runtime_error("GenesisMesh::ReadFile(), unsupported auxiliary element map!");
    elem_global_id_ = f.read_int("elem_map1");
    for (int& g : elem_global_id_) g -= 1;
  }

  // node sets (src/nimble_genesis_mesh.cc:151-191)
  if (num_node_sets > 0) {
    node_set_ids_ = f.read_int("ns_prop1");
    std::vector<std::string> names = f.has_var("ns_names") ? f.read_strings("ns_names") : std::vector<std::string>();
    for (int i = 0; i < num_node_sets; ++i) {
      const int   id   = node_set_ids_[i];
      std::string name = i < (int)names.size() ? names[i] : std::string();
      if (name.empty()) name = numbered("nodelist_", id);
      node_set_names_[id]          = name;
      node_sets_[id]               = std::vector<int>();
      ns_distribution_factors_[id] = std::vector<double>();
      const std::string nv = numbered("node_ns", i + 1), dv = numbered("dist_fact_ns", i + 1);
      if (!f.has_var(nv)) continue;  // a set with no local nodes keeps its id
      std::vector<int> nodes = f.read_int(nv);
      const size_t     ndf   = f.has_var(dv) ? f.read_double(dv).size() : 0;
      // the reference loads a set only when #dist-factors == #nodes (:178-184); files written without
      // distribution factors (df count 0) are accepted too -- ex_get_set_param reports 0 there and the
      // reference would drop the set, which no shipped deck relies on
      if (!nodes.empty() && (ndf == nodes.size() || ndf == 0)) {
        for (int& n : nodes) n -= 1;
        node_sets_[id] = nodes;
      }
      if (ndf) ns_distribution_factors_[id] = f.read_double(dv);
    }
  }

  // element blocks (src/nimble_genesis_mesh.cc:236-312)
  std::vector<int>         all_ids = num_blocks ? f.read_int("eb_prop1") : std::vector<int>();
  std::vector<std::string> eb_names = f.has_var("eb_names") ? f.read_strings("eb_names") : std::vector<std::string>();
  std::map<int, int>       file_index;
  for (int i = 0; i < num_blocks; ++i) {
    const int         id  = all_ids[i];
    const std::string cv  = numbered("connect", i + 1);
    const int         nel = f.has_var(cv) ? (int)f.dim_or(numbered("num_el_in_blk", i + 1), 0) : 0;
    std::string       name = i < (int)eb_names.size() ? eb_names[i] : std::string();
    if (name.empty()) name = numbered("block_", id);
    file_index[id] = i;
    if (nel > 0) {
      block_ids_.push_back(id);
      block_names_[id] = name;
    }
    all_block_ids_.push_back(id);
    all_block_names_[id] = name;
  }
  int elem_local_index = 0;
  for (int id : block_ids_) {
    const int         i   = file_index[id];
    const std::string cv  = numbered("connect", i + 1);
    const int         nel = (int)f.dim(numbered("num_el_in_blk", i + 1));
    const int         npe = (int)f.dim(numbered("num_nod_per_el", i + 1));
    block_num_nodes_per_elem_[id] = npe;
    std::vector<int>& gids        = block_elem_global_ids_[id];
    gids.resize(nel);
    for (int e = 0; e < nel; ++e) gids[e] = elem_global_id_[elem_local_index++];
    std::vector<int> conn = f.read_int(cv);
    if ((int)conn.size() != nel * npe) throw std::runtime_error("GenesisMesh::ReadFile(), bad connectivity size in " + cv);
    for (int& c : conn) c -= 1;
    block_elem_connectivity_[id] = conn;
  }
}

void
GenesisMesh::Initialize(
    std::string const&                     file_name,
    std::vector<int> const&                node_global_id,
    std::vector<double> const&             node_x,
    std::vector<double> const&             node_y,
    std::vector<double> const&             node_z,
    std::vector<int> const&                elem_global_id,
    std::vector<int> const&                block_ids,
    std::map<int, std::string> const&      block_names,
    std::map<int, std::vector<int>> const& block_elem_global_ids,
    std::map<int, int> const&              block_num_nodes_per_elem,
    std::map<int, std::vector<int>> const& block_elem_connectivity,
    std::map<int, std::string> const&      node_set_names,
    std::map<int, std::vector<int>> const& node_sets)
{
  file_name_ = file_name;
  dim_       = node_z.empty() ? 2 : 3;
  node_global_id_ = node_global_id, node_x_ = node_x, node_y_ = node_y, node_z_ = node_z;
  elem_global_id_ = elem_global_id;
  block_ids_ = all_block_ids_ = block_ids;
  block_names_ = all_block_names_ = block_names;
  block_elem_global_ids_          = block_elem_global_ids;
  block_num_nodes_per_elem_       = block_num_nodes_per_elem;
  block_elem_connectivity_        = block_elem_connectivity;
  node_set_names_                 = node_set_names;
  node_sets_                      = node_sets;
  node_set_ids_.clear();
  for (auto const& kv : node_sets_) node_set_ids_.push_back(kv.first);
}

GenesisMesh
GenesisMesh::StructuredCube(int n)
{
  const int           nn = n + 1;
  const long long     num_nodes = (long long)nn * nn * nn, num_elem = (long long)n * n * n;
  if (num_nodes > 2147483647LL) throw std::invalid_argument("StructuredCube: node count exceeds int32");
  std::vector<double> x(num_nodes), y(num_nodes), z(num_nodes);
  std::vector<int>    ngid(num_nodes), egid(num_elem), conn(num_elem * 8), all, face;
  all.reserve(num_nodes);
  for (int k = 0; k < nn; ++k)
    for (int j = 0; j < nn; ++j)
      for (int i = 0; i < nn; ++i) {
        const int id = i + nn * (j + nn * k);
        x[id] = (double)i / n, y[id] = (double)j / n, z[id] = (double)k / n;
        ngid[id] = id;
        all.push_back(id);
        if (i == 0) face.push_back(id);
      }
  static const int corner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
  for (int k = 0; k < n; ++k)
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        const long long e = i + (long long)n * (j + (long long)n * k);
        egid[e]           = (int)e;
        for (int c = 0; c < 8; ++c) conn[e * 8 + c] = (i + corner[c][0]) + nn * ((j + corner[c][1]) + nn * (k + corner[c][2]));
      }
  GenesisMesh m;
  m.Initialize("synthetic_cube", ngid, x, y, z, egid, {1}, {{1, "block_1"}}, {{1, egid}}, {{1, 8}}, {{1, conn}},
               {{1, "nodelist_1"}, {2, "nodelist_2"}}, {{1, all}, {2, face}});
  return m;
}

int
GenesisMesh::GetMaxNodeGlobalId() const
{
  int m = -1;
  for (int id : node_global_id_)
    if (id > m) m = id;
  return m;
}

bool
GenesisMesh::HasBlock(std::string const& block_name) const
{
  for (auto const& kv : block_names_)
    if (kv.second == block_name) return true;
  return false;
}

int
GenesisMesh::GetNumElementsInBlock(int block_id) const
{
  const int npe = block_num_nodes_per_elem_.at(block_id);
  return npe ? (int)block_elem_connectivity_.at(block_id).size() / npe : 0;
}

std::map<int, int>
GenesisMesh::GetNumElementsInBlock() const
{
  std::map<int, int> out;
  for (int id : block_ids_) out[id] = GetNumElementsInBlock(id);
  return out;
}

std::string
GenesisMesh::GetElementType(int block_id) const
{
  // src/nimble_genesis_mesh.cc:480-505: the type string of the file is ignored, nodes per element decide
  // (SPHERE and TRIANGLE are the blocks of the contact visualisation database; "TRIANGLE:" in 2D is the reference's spelling)
  const int npe = block_num_nodes_per_elem_.at(block_id);
  if (dim_ == 2 && npe == 3) return "TRIANGLE:";
  if (dim_ == 2 && npe == 4) return "QUAD";
  if (dim_ == 3 && npe == 1) return "SPHERE";
  if (dim_ == 3 && npe == 3) return "TRIANGLE";
  if (dim_ == 3 && npe == 4) return "TETRA";
  if (dim_ == 3 && npe == 8) return "HEX";
  throw std::invalid_argument("Error processing input mesh, unknown element type.");
}

void
GenesisMesh::BlockNamesToOnProcessorBlockIds(std::vector<std::string> const& block_names, std::vector<int>& block_ids) const
{
  block_ids.clear();
  for (auto const& name : block_names)
    if (HasBlock(name)) block_ids.push_back(GetBlockId(name));
}

int
GenesisMesh::GetBlockId(std::string const& block_name) const
{
  for (auto const& kv : all_block_names_)
    if (kv.second == block_name) return kv.first;
  return -1;
}

void
GenesisMesh::Print(bool verbose, int my_rank) const
{
  std::cout << "\n--Genesis Mesh [" << my_rank << "] " << file_name_ << "\n  dimension " << dim_ << ", nodes " << GetNumNodes()
            << ", elements " << GetNumElements() << ", blocks " << GetNumBlocks() << ", node sets " << GetNumNodeSets() << "\n";
  if (verbose) {
    for (int id : block_ids_)
      std::cout << "  block " << id << " \"" << block_names_.at(id) << "\": " << GetNumElementsInBlock(id) << " "
                << GetElementType(id) << "\n";
    for (int id : node_set_ids_) std::cout << "  node set " << id << " \"" << node_set_names_.at(id) << "\": " << node_sets_.at(id).size() << " nodes\n";
  }
}

std::vector<int>
GenesisMesh::RcbElementPartition(int n_parts) const
{
  std::vector<std::array<double, 3>> cent;
  for (int id : block_ids_) {
    const int               npe  = block_num_nodes_per_elem_.at(id);
    const std::vector<int>& conn = block_elem_connectivity_.at(id);
    const size_t            nel  = npe ? conn.size() / npe : 0;
    for (size_t e = 0; e < nel; ++e) {
      std::array<double, 3> s{0.0, 0.0, 0.0};
      const double*         xyz[3] = {node_x_.data(), node_y_.data(), dim_ == 3 ? node_z_.data() : nullptr};
      for (int d = 0; d < 3; ++d) {
        if (!xyz[d]) continue;
        const int* n = &conn[e * npe];
        if (npe == 8)  // balanced tree: the summation order of the Python mirror (nimblesm_b200/mesh.py, numpy's 8-way reduction)
          s[d] = ((xyz[d][n[0]] + xyz[d][n[1]]) + (xyz[d][n[2]] + xyz[d][n[3]])) + ((xyz[d][n[4]] + xyz[d][n[5]]) + (xyz[d][n[6]] + xyz[d][n[7]]));
        else
          for (int j = 0; j < npe; ++j) s[d] += xyz[d][n[j]];
        s[d] /= npe;
      }
      cent.push_back(s);
    }
  }
  std::vector<int> part(cent.size(), 0), ids(cent.size());
  for (size_t i = 0; i < ids.size(); ++i) ids[i] = (int)i;
  // explicit stack of (begin, end, first part, number of parts) over `ids`; ties keep the order of the level above
  // (ascending element order at the root)
  struct Job { size_t b, e; int p0, np; };
  std::vector<Job> jobs{{0, ids.size(), 0, n_parts < 1 ? 1 : n_parts}};
  while (!jobs.empty()) {
    const Job j = jobs.back();
    jobs.pop_back();
    if (j.np == 1) {
      for (size_t i = j.b; i < j.e; ++i) part[ids[i]] = j.p0;
      continue;
    }
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (size_t i = j.b; i < j.e; ++i)
      for (int d = 0; d < 3; ++d) {
        lo[d] = std::min(lo[d], cent[ids[i]][d]);
        hi[d] = std::max(hi[d], cent[ids[i]][d]);
      }
    int d = 0;
    for (int k = 1; k < 3; ++k)
      if (hi[k] - lo[k] > hi[d] - lo[d]) d = k;
    std::stable_sort(ids.begin() + j.b, ids.begin() + j.e, [&](int a, int b) { return cent[a][d] < cent[b][d]; });
    const int    left = j.np / 2;
    const size_t cut  = j.b + ((j.e - j.b) * (size_t)left) / (size_t)j.np;
    jobs.push_back({j.b, cut, j.p0, left});
    jobs.push_back({cut, j.e, j.p0 + left, j.np - left});
  }
  return part;
}

void
GenesisMesh::KeepPart(std::vector<int> const& part_of_element, int part)
{
  std::vector<char> used(node_x_.size(), 0);
  size_t            flat = 0;
  std::vector<int>  kept_blocks, new_elem_gid;
  for (int id : block_ids_) {
    const int         npe  = block_num_nodes_per_elem_.at(id);
    std::vector<int>& conn = block_elem_connectivity_.at(id);
    std::vector<int>& gids = block_elem_global_ids_.at(id);
    const size_t      nel  = npe ? conn.size() / npe : 0;
    std::vector<int>  c2, g2;
    for (size_t e = 0; e < nel; ++e, ++flat) {
      if (flat >= part_of_element.size()) throw std::invalid_argument("GenesisMesh::KeepPart: partition vector too short");
      if (part_of_element[flat] != part) continue;
      for (int j = 0; j < npe; ++j) {
        c2.push_back(conn[e * npe + j]);
        used[conn[e * npe + j]] = 1;
      }
      g2.push_back(gids[e]);
    }
    conn.swap(c2);
    gids.swap(g2);
    if (!gids.empty()) {
      kept_blocks.push_back(id);
      new_elem_gid.insert(new_elem_gid.end(), gids.begin(), gids.end());
    }
  }
  // blocks without an element in this part disappear from the local list (as in a Nemesis piece); names stay global
  for (int id : block_ids_)
    if (std::find(kept_blocks.begin(), kept_blocks.end(), id) == kept_blocks.end()) {
      block_elem_connectivity_.erase(id), block_elem_global_ids_.erase(id), block_num_nodes_per_elem_.erase(id);
      block_names_.erase(id);
    }
  block_ids_      = kept_blocks;
  elem_global_id_ = new_elem_gid;
  std::vector<int> remap(node_x_.size(), -1);
  std::vector<int>    gid2;
  std::vector<double> x2, y2, z2;
  for (size_t n = 0; n < used.size(); ++n)
    if (used[n]) {
      remap[n] = (int)gid2.size();
      gid2.push_back(node_global_id_[n]);
      x2.push_back(node_x_[n]), y2.push_back(node_y_[n]);
      if (n < node_z_.size()) z2.push_back(node_z_[n]);
    }
  node_global_id_.swap(gid2), node_x_.swap(x2), node_y_.swap(y2), node_z_.swap(z2);
  for (int id : block_ids_)
    for (int& n : block_elem_connectivity_.at(id)) n = remap[n];
  for (auto& kv : node_sets_) {
    std::vector<int>     keep;
    std::vector<double>  df;
    std::vector<double>& old_df = ns_distribution_factors_[kv.first];
    for (size_t k = 0; k < kv.second.size(); ++k)
      if (remap[kv.second[k]] >= 0) {
        keep.push_back(remap[kv.second[k]]);
        if (k < old_df.size()) df.push_back(old_df[k]);
      }
    kv.second.swap(keep);
    old_df.swap(df);
  }
}

}  // namespace nimble_b200
