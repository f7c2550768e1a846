// nimblesm_b200/host/exodus_output.cc — see exodus_output.h.
#include "exodus_output.h"

#include <set>
#include <sstream>
#include <stdexcept>

#include "genesis_mesh.h"
#include "netcdf3.h"

namespace nimble_b200 {

namespace {
std::string
numbered(const char* stem, int k, const char* tail = "")
{
  std::ostringstream s;
  s << stem << k << tail;
  return s.str();
}
constexpr int kLenName = 256;  // the reference writes MAX_NAME_LENGTH = 255 + 1 (Appendix A: len_name 256 in .e files)
}  // namespace

ExodusOutput::ExodusOutput()  = default;
ExodusOutput::~ExodusOutput()
{
  if (writer_.joinable()) writer_.join();
}

void
ExodusOutput::Initialize(std::string const& filename, GenesisMesh const& mesh)
{
  filename_          = filename;
  dim_               = mesh.GetDim();
  num_nodes_         = (int)mesh.GetNumNodes();
  num_elements_      = (int)mesh.GetNumElements();
  num_blocks_        = (int)mesh.GetNumBlocks();
  num_global_blocks_ = (int)mesh.GetNumGlobalBlocks();
  block_ids_         = mesh.GetBlockIds();
  all_block_ids_     = mesh.GetAllBlockIds();
  num_node_sets_     = mesh.GetNumNodeSets();
  for (int i = 0; i < num_global_blocks_; ++i) block_file_index_[all_block_ids_[i]] = i + 1;
}

void
ExodusOutput::InitializeDatabase(GenesisMesh const& mesh, std::vector<std::string> const& global_data_names,
                                 std::vector<std::string> const& node_data_names,
                                 std::map<int, std::vector<std::string>> const& elem_data_names,
                                 std::map<int, std::vector<std::string>> const& derived_elem_data_names)
{
  file_.reset(new nc3::Writer(filename_));
  nc3::Writer& f = *file_;
  num_global_vars_ = (int)global_data_names.size();
  num_node_vars_   = num_nodes_ > 0 ? (int)node_data_names.size() : 0;

  std::set<std::string> unique;
  for (int id : block_ids_) {
    for (auto const& n : elem_data_names.at(id)) unique.insert(n);
    for (auto const& n : derived_elem_data_names.at(id)) unique.insert(n);
  }
  std::vector<std::string> elem_var_names(unique.begin(), unique.end());
  for (size_t i = 0; i < elem_var_names.size(); ++i) elem_data_index_[elem_var_names[i]] = (int)i + 1;

  // ---- global attributes and dimensions (ex_create / ex_put_init) --------------------------------------
  f.put_global_float("api_version", 7.17f);
  f.put_global_float("version", 7.17f);
  f.put_global_int("floating_point_word_size", 8);
  f.put_global_int("file_size", 1);
  f.put_global_int("maximum_name_length", 32);
  f.put_global_int("int64_status", 0);
  f.put_global_text("title", "NimbleSM");
  f.def_dim("len_string", 33);
  f.def_dim("len_line", 81);
  f.def_dim("four", 4);
  f.def_dim("len_name", kLenName);
  f.def_dim("time_step", 0);
  f.def_dim("num_dim", dim_);
  f.def_dim("num_nodes", num_nodes_);
  f.def_dim("num_elem", num_elements_);
  f.def_dim("num_el_blk", num_global_blocks_);
  f.def_dim("num_qa_rec", 1);
  if (num_node_sets_ > 0) f.def_dim("num_node_sets", num_node_sets_);
  std::map<int, std::vector<int>> node_sets = mesh.GetNodeSets();
  std::vector<int>                ns_ids    = mesh.GetNodeSetIds();
  for (int i = 0; i < num_global_blocks_; ++i) {
    const int  id    = all_block_ids_[i];
    const bool local = mesh.GetNumElementsInBlock().count(id) != 0 && mesh.GetNumElementsInBlock(id) > 0;
    if (local) {
      f.def_dim(numbered("num_el_in_blk", i + 1), mesh.GetNumElementsInBlock(id));
      f.def_dim(numbered("num_nod_per_el", i + 1), mesh.GetNumNodesPerElement(id));
    }
  }
  for (int i = 0; i < num_node_sets_; ++i)
    if (!node_sets[ns_ids[i]].empty()) f.def_dim(numbered("num_nod_ns", i + 1), (int64_t)node_sets[ns_ids[i]].size());
  if (num_global_vars_ > 0) f.def_dim("num_glo_var", num_global_vars_);
  if (num_node_vars_ > 0) f.def_dim("num_nod_var", num_node_vars_);
  if (!elem_var_names.empty()) f.def_dim("num_elem_var", (int64_t)elem_var_names.size());

  // ---- variables -----------------------------------------------------------------------------------------
  f.def_var("time_whole", nc3::DOUBLE, {"time_step"});
  f.def_var("qa_records", nc3::CHAR, {"num_qa_rec", "four", "len_string"});
  f.def_var("coor_names", nc3::CHAR, {"num_dim", "len_name"});
  f.def_var("eb_names", nc3::CHAR, {"num_el_blk", "len_name"});
  f.def_var("eb_status", nc3::INT, {"num_el_blk"});
  f.def_var("eb_prop1", nc3::INT, {"num_el_blk"});
  f.put_var_text_attribute("eb_prop1", "name", "ID");
  if (num_node_sets_ > 0) {
    f.def_var("ns_status", nc3::INT, {"num_node_sets"});
    f.def_var("ns_prop1", nc3::INT, {"num_node_sets"});
    f.put_var_text_attribute("ns_prop1", "name", "ID");
    f.def_var("ns_names", nc3::CHAR, {"num_node_sets", "len_name"});
  }
  static const char* const coord_vars[3] = {"coordx", "coordy", "coordz"};
  if (num_nodes_ > 0)
    for (int d = 0; d < dim_; ++d) f.def_var(coord_vars[d], nc3::DOUBLE, {"num_nodes"});
  if (num_nodes_ > 0) f.def_var("node_num_map", nc3::INT, {"num_nodes"});
  if (num_elements_ > 0) f.def_var("elem_num_map", nc3::INT, {"num_elem"});
  for (int i = 0; i < num_global_blocks_; ++i) {
    const int id = all_block_ids_[i];
    if (mesh.GetNumElementsInBlock().count(id) == 0 || mesh.GetNumElementsInBlock(id) == 0) continue;
    const std::string cv = numbered("connect", i + 1);
    f.def_var(cv, nc3::INT, {numbered("num_el_in_blk", i + 1), numbered("num_nod_per_el", i + 1)});
    f.put_var_text_attribute(cv, "elem_type", mesh.GetElementType(id));
  }
  for (int i = 0; i < num_node_sets_; ++i)
    if (!node_sets[ns_ids[i]].empty()) f.def_var(numbered("node_ns", i + 1), nc3::INT, {numbered("num_nod_ns", i + 1)});
  if (num_global_vars_ > 0) {
    f.def_var("name_glo_var", nc3::CHAR, {"num_glo_var", "len_name"});
    f.def_var("vals_glo_var", nc3::DOUBLE, {"time_step", "num_glo_var"});
  }
  if (num_node_vars_ > 0) {
    f.def_var("name_nod_var", nc3::CHAR, {"num_nod_var", "len_name"});
    for (int k = 0; k < num_node_vars_; ++k) f.def_var(numbered("vals_nod_var", k + 1), nc3::DOUBLE, {"time_step", "num_nodes"});
  }
  if (!elem_var_names.empty()) {
    f.def_var("name_elem_var", nc3::CHAR, {"num_elem_var", "len_name"});
    for (int id : block_ids_) {
      const int b = block_file_index_.at(id);
      std::set<std::string> on_block(elem_data_names.at(id).begin(), elem_data_names.at(id).end());
      on_block.insert(derived_elem_data_names.at(id).begin(), derived_elem_data_names.at(id).end());
      for (auto const& n : on_block)
        f.def_var(numbered("vals_elem_var", elem_data_index_.at(n), numbered("eb", b).c_str()), nc3::DOUBLE,
                  {"time_step", numbered("num_el_in_blk", b)});
    }
  }
  f.end_define();

  // ---- mesh payload --------------------------------------------------------------------------------------
  f.put_strings("qa_records", {"NimbleSM", "b200", "", ""});
  {
    std::vector<std::string> coor{"x", "y", "z"};
    coor.resize((size_t)dim_);
    f.put_strings("coor_names", coor);
  }
  std::vector<std::string> eb_names;
  std::vector<int>         eb_status;
  for (int id : all_block_ids_) {
    eb_names.push_back(mesh.GetBlockName(id));
    eb_status.push_back(1);
  }
  f.put_strings("eb_names", eb_names);
  if (num_global_blocks_ > 0) {
    f.put_int("eb_status", eb_status.data(), num_global_blocks_);
    f.put_int("eb_prop1", all_block_ids_.data(), num_global_blocks_);
  }
  if (num_nodes_ > 0) {
    const double* xyz[3] = {mesh.GetCoordinatesX(), mesh.GetCoordinatesY(), mesh.GetCoordinatesZ()};
    for (int d = 0; d < dim_; ++d) f.put_double(coord_vars[d], xyz[d], num_nodes_);
    std::vector<int> gid(num_nodes_);
    for (int i = 0; i < num_nodes_; ++i) gid[i] = mesh.GetNodeGlobalIds()[i] + 1;
    f.put_int("node_num_map", gid.data(), num_nodes_);
  }
  if (num_elements_ > 0) {
    std::vector<int> gid(num_elements_);
    for (int i = 0; i < num_elements_; ++i) gid[i] = mesh.GetElementGlobalIds()[i] + 1;
    f.put_int("elem_num_map", gid.data(), num_elements_);
  }
  for (int i = 0; i < num_global_blocks_; ++i) {
    const int id = all_block_ids_[i];
    if (mesh.GetNumElementsInBlock().count(id) == 0 || mesh.GetNumElementsInBlock(id) == 0) continue;
    const int64_t    n = (int64_t)mesh.GetNumElementsInBlock(id) * mesh.GetNumNodesPerElement(id);
    std::vector<int> conn(mesh.GetConnectivity(id), mesh.GetConnectivity(id) + n);
    for (int& c : conn) c += 1;
    f.put_int(numbered("connect", i + 1), conn.data(), n);
  }
  if (num_node_sets_ > 0) {
    std::vector<int>           status(num_node_sets_, 1);
    std::map<int, std::string> names = mesh.GetNodeSetNames();
    std::vector<std::string>   ns_names;
    for (int id : ns_ids) ns_names.push_back(names[id]);
    f.put_int("ns_status", status.data(), num_node_sets_);
    f.put_int("ns_prop1", ns_ids.data(), num_node_sets_);
    f.put_strings("ns_names", ns_names);
    for (int i = 0; i < num_node_sets_; ++i) {
      std::vector<int> nodes = node_sets[ns_ids[i]];
      if (nodes.empty()) continue;
      for (int& n : nodes) n += 1;
      f.put_int(numbered("node_ns", i + 1), nodes.data(), (int64_t)nodes.size());
    }
  }
  if (num_global_vars_ > 0) f.put_strings("name_glo_var", global_data_names);
  if (num_node_vars_ > 0) f.put_strings("name_nod_var", node_data_names);
  if (!elem_var_names.empty()) f.put_strings("name_elem_var", elem_var_names);
  f.flush();
}

void
ExodusOutput::WriteStep(double time, std::vector<double> const& global_data, std::vector<std::vector<double>> const& node_data,
                        std::map<int, std::vector<std::string>> const& elem_data_names,
                        std::map<int, std::vector<std::vector<double>>> const& elem_data,
                        std::map<int, std::vector<std::string>> const& derived_elem_data_names,
                        std::map<int, std::vector<std::vector<double>>> const& derived_elem_data)
{
  if (!file_) throw std::runtime_error("ExodusOutput::WriteStep before InitializeDatabase");
  Wait();
  WritePlane(time, global_data, node_data, elem_data_names, elem_data, derived_elem_data_names, derived_elem_data);
}

void
ExodusOutput::WritePlane(double time, std::vector<double> const& global_data, std::vector<std::vector<double>> const& node_data,
                         std::map<int, std::vector<std::string>> const& elem_data_names,
                         std::map<int, std::vector<std::vector<double>>> const& elem_data,
                         std::map<int, std::vector<std::string>> const& derived_elem_data_names,
                         std::map<int, std::vector<std::vector<double>>> const& derived_elem_data)
{
  nc3::Writer&  f   = *file_;
  const int64_t rec = exodus_write_count_;
  exodus_write_count_ += 1;
  f.put_record_double("time_whole", rec, &time, 1);
  if (num_global_vars_ > 0) f.put_record_double("vals_glo_var", rec, global_data.data(), num_global_vars_);
  for (int k = 0; k < num_node_vars_; ++k) f.put_record_double(numbered("vals_nod_var", k + 1), rec, node_data[k].data(), num_nodes_);
  for (int id : block_ids_) {
    const int b = block_file_index_.at(id);
    auto      put = [&](std::vector<std::string> const& names, std::vector<std::vector<double>> const& data) {
      for (size_t v = 0; v < names.size(); ++v)
        f.put_record_double(numbered("vals_elem_var", elem_data_index_.at(names[v]), numbered("eb", b).c_str()), rec, data[v].data(),
                            (int64_t)data[v].size());
    };
    put(elem_data_names.at(id), elem_data.at(id));
    put(derived_elem_data_names.at(id), derived_elem_data.at(id));
  }
  f.flush();
}

void
ExodusOutput::WriteStepAsync(double time, std::vector<double> global_data, std::vector<std::vector<double>> node_data,
                             std::map<int, std::vector<std::string>> elem_data_names,
                             std::map<int, std::vector<std::vector<double>>> elem_data,
                             std::map<int, std::vector<std::string>> derived_elem_data_names,
                             std::map<int, std::vector<std::vector<double>>> derived_elem_data)
{
  if (!file_) throw std::runtime_error("ExodusOutput::WriteStepAsync before InitializeDatabase");
  Wait();
  writer_ = std::thread([this, time, g = std::move(global_data), n = std::move(node_data), en = std::move(elem_data_names),
                         e = std::move(elem_data), dn = std::move(derived_elem_data_names), d = std::move(derived_elem_data)] {
    try {
      WritePlane(time, g, n, en, e, dn, d);
    } catch (...) {
      writer_error_ = std::current_exception();
    }
  });
}

void
ExodusOutput::Wait()
{
  if (writer_.joinable()) writer_.join();
  if (writer_error_) {
    std::exception_ptr e = writer_error_;
    writer_error_        = nullptr;
    std::rethrow_exception(e);
  }
}

void
ExodusOutput::Close()
{
  Wait();
  if (file_) file_->close();
}

}  // namespace nimble_b200
