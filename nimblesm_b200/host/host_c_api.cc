// nimblesm_b200/host/host_c_api.cc — plain-C entry points over the host classes, bound with ctypes by the CPU
// tests (tests/test_host_cpp.py).  Nothing here computes on the device; every function returns 0 on success
// and writes a message into `err` otherwise (exceptions never cross the boundary).
#include <cstdio>
#include <cstring>
#include <sstream>
#include <thread>

#include "boundary_condition.h"
#include "contact_manager.h"
#include "data_manager.h"
#include "exodus_output.h"
#include "expression.h"
#include "genesis_mesh.h"
#include "material.h"
#include "parser.h"

using namespace nimble_b200;

namespace {
int
fail(char* err, int errlen, const std::string& m)
{
  if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", m.c_str());
  return 1;
}
int
put(char* out, int outlen, const std::string& s, char* err, int errlen)
{
  if ((int)s.size() + 1 > outlen) return fail(err, errlen, "output buffer too small");
  memcpy(out, s.c_str(), s.size() + 1);
  return 0;
}
}  // namespace

extern "C" {

int
nsmh_expression_eval(const char* text, double x, double y, double z, double t, double* out, char* err, int errlen)
{
  try {
    *out = Expression(text).eval(x, y, z, t);
    return 0;
  } catch (std::exception const& e) {
    return fail(err, errlen, e.what());
  }
}

// Expression::compile for one expression: the device program, its constants, and its slots evaluated at time t.
// Returns 0 = compiled, 2 = no bit-exact device form (the host evaluates it), 1 = parse error.
int
nsmh_expression_compile(const char* text, double t, int max_words, int* n_words, int* code, int max_consts, int* n_consts,
                        double* consts, int max_slots, int* n_slots, double* slot_values, char* err, int errlen)
{
  try {
    Expression              e(text);
    std::vector<int32_t>    c;
    std::vector<double>     k;
    std::vector<Expression> s, ec;
    if (!e.compile(c, k, s, ec, 16)) return 2;
    if ((int)c.size() > max_words || (int)k.size() > max_consts || (int)s.size() > max_slots) return fail(err, errlen, "buffers too small");
    *n_words = (int)c.size(), *n_consts = (int)k.size(), *n_slots = (int)s.size();
    for (size_t i = 0; i < c.size(); ++i) code[i] = c[i];
    for (size_t i = 0; i < k.size(); ++i) consts[i] = k[i];
    for (size_t i = 0; i < s.size(); ++i) slot_values[i] = s[i].eval(0.0, 0.0, 0.0, t);
    return 0;
  } catch (std::exception const& e) {
    return fail(err, errlen, e.what());
  }
}

// The entry constants of the same compilation (sub-trees of the position alone that need libm / pow, NSM_BCOP_ENTRYCONST)
// evaluated at one point.  Returns 0 and their count, 2 when the expression has no device form, 1 on a parse error.
int
nsmh_expression_entry_constants(const char* text, double x, double y, double z, int max_values, int* n_values, double* values,
                                char* err, int errlen)
{
  try {
    Expression              e(text);
    std::vector<int32_t>    c;
    std::vector<double>     k;
    std::vector<Expression> s, ec;
    if (!e.compile(c, k, s, ec, 16)) return 2;
    if ((int)ec.size() > max_values) return fail(err, errlen, "buffers too small");
    *n_values = (int)ec.size();
    for (size_t i = 0; i < ec.size(); ++i) values[i] = ec[i].eval(x, y, z, 0.0);
    return 0;
  } catch (std::exception const& e) {
    return fail(err, errlen, e.what());
  }
}

int
nsmh_io_file_name(const char* serial, const char* ext, const char* label, int rank, int nranks, char* out, int outlen)
{
  return put(out, outlen, IOFileName(serial, ext, label, rank, nranks), nullptr, 0);
}

// deck text -> JSON summary of what the Parser holds
int
nsmh_deck_summary(const char* deck_text, const int* block_ids, int n_blocks, char* out, int outlen, char* err, int errlen)
{
  try {
    Parser p;
    p.InitializeFromString(deck_text);
    std::ostringstream j;
    j.precision(17);
    j << "{\"genesis\":\"" << p.GenesisFileName() << "\",\"exodus\":\"" << p.ExodusFileName() << "\",\"scheme\":\""
      << p.TimeIntegrationScheme() << "\",\"initial_time\":" << p.InitialTime() << ",\"final_time\":" << p.FinalTime()
      << ",\"num_load_steps\":" << p.NumLoadSteps() << ",\"output_frequency\":" << p.OutputFrequency() << ",\"output_fields\":\""
      << p.GetOutputFieldString() << "\",\"n_bc\":" << p.GetBoundaryConditionStrings().size() << ",\"materials\":{";
    MaterialFactory factory;
    for (int i = 0; i < n_blocks; ++i) {
      const std::string mp = p.GetModelMaterialParameters(block_ids[i]);
      j << (i ? "," : "") << "\"" << block_ids[i] << "\":";
      if (mp == "none") {
        j << "null";
        continue;
      }
      factory.parse_and_create(mp);
      auto m = factory.get_material();
      j << "{\"model\":\"" << m->Parameters().GetMaterialName() << "\",\"density\":" << m->GetDensity() << ",\"bulk_modulus\":"
        << m->GetBulkModulus() << ",\"shear_modulus\":" << m->GetShearModulus() << ",\"num_state\":" << m->NumStateVariables() << "}";
    }
    j << "}}";
    return put(out, outlen, j.str(), err, errlen);
  } catch (std::exception const& e) {
    return fail(err, errlen, e.what());
  }
}

// MaterialFactoryBase::ParseMaterialParametersString through a factory that registers extra parameter names, as the
// TestMaterialFactory of the reference's unit test does (unit_tests/test_nimble_material_params.cc:63-90):
// material string -> JSON {"name", "doubles": {...}, "strings": {...}}; names are blank-separated lists (may be empty)
int
nsmh_material_params(const char* material_string, const char* extra_double_names, const char* extra_string_names, char* out, int outlen,
                     char* err, int errlen)
{
  struct TestFactory : MaterialFactoryBase
  {
    std::shared_ptr<MaterialParameters>
    parse_string(const char* params) const
    {
      return ParseMaterialParametersString(params);
    }
    void
    create() override
    {
    }
  };
  try {
    TestFactory        factory;
    std::istringstream dn(extra_double_names ? extra_double_names : ""), sn(extra_string_names ? extra_string_names : "");
    for (std::string w; dn >> w;) factory.add_valid_double_parameter_name(w.c_str());
    for (std::string w; sn >> w;) factory.add_valid_string_parameter_name(w.c_str());
    auto               params = factory.parse_string(material_string);
    std::ostringstream j;
    j.precision(17);
    j << "{\"name\":\"" << params->GetMaterialName(false) << "\",\"upper\":\"" << params->GetMaterialName(true)
      << "\",\"n_doubles\":" << params->GetNumParameters() << ",\"n_strings\":" << params->GetNumStringParameters() << ",\"doubles\":{";
    bool first = true;
    for (auto const& kv : params->GetParameters()) {
      if (!params->IsParameter(kv.first.c_str())) return fail(err, errlen, "IsParameter denies a listed parameter");
      j << (first ? "" : ",") << "\"" << kv.first << "\":" << params->GetParameterValue(kv.first.c_str());
      first = false;
    }
    j << "},\"strings\":{";
    first = true;
    for (auto const& kv : params->GetStringParameters()) {
      if (!params->IsStringParameter(kv.first.c_str())) return fail(err, errlen, "IsStringParameter denies a listed parameter");
      j << (first ? "" : ",") << "\"" << kv.first << "\":\"" << params->GetStringParameterValue(kv.first.c_str()) << "\"";
      first = false;
    }
    j << "}}";
    return put(out, outlen, j.str(), err, errlen);
  } catch (std::exception const& e) {
    return fail(err, errlen, e.what());
  }
}

// Genesis file -> JSON summary (counts, id maps, checksums) for comparison with an independent reader
int
nsmh_mesh_summary(const char* path, char* out, int outlen, char* err, int errlen)
{
  try {
    GenesisMesh m;
    m.ReadFile(path);
    std::ostringstream j;
    j.precision(17);
    double sx = 0, sy = 0, sz = 0;
    for (unsigned i = 0; i < m.GetNumNodes(); ++i) sx += m.GetCoordinatesX()[i], sy += m.GetCoordinatesY()[i], sz += m.GetCoordinatesZ()[i];
    long long gsum = 0;
    for (unsigned i = 0; i < m.GetNumNodes(); ++i) gsum += m.GetNodeGlobalIds()[i];
    j << "{\"dim\":" << m.GetDim() << ",\"num_nodes\":" << m.GetNumNodes() << ",\"num_elements\":" << m.GetNumElements()
      << ",\"sum_x\":" << sx << ",\"sum_y\":" << sy << ",\"sum_z\":" << sz << ",\"node_gid_sum\":" << gsum << ",\"blocks\":{";
    bool first = true;
    for (int id : m.GetAllBlockIds()) {
      long long csum = 0;
      int       ne   = 0;
      bool      local = false;
      for (int b : m.GetBlockIds()) local |= b == id;
      if (local) {
        ne = m.GetNumElementsInBlock(id);
        for (long i = 0; i < (long)ne * m.GetNumNodesPerElement(id); ++i) csum += (long long)(i % 7 + 1) * m.GetConnectivity(id)[i];
      }
      j << (first ? "" : ",") << "\"" << id << "\":{\"name\":\"" << m.GetBlockName(id) << "\",\"num_elements\":" << ne << ",\"conn_checksum\":" << csum
        << "}";
      first = false;
    }
    j << "},\"node_sets\":{";
    first      = true;
    auto sets  = m.GetNodeSets();
    auto names = m.GetNodeSetNames();
    for (int id : m.GetNodeSetIds()) {
      long long s = 0;
      for (int n : sets[id]) s += n;
      j << (first ? "" : ",") << "\"" << id << "\":{\"name\":\"" << names[id] << "\",\"size\":" << sets[id].size() << ",\"sum\":" << s << "}";
      first = false;
    }
    j << "}}";
    return put(out, outlen, j.str(), err, errlen);
  } catch (std::exception const& e) {
    return fail(err, errlen, e.what());
  }
}

// GenesisMesh::RcbElementPartition + KeepPart: JSON summary of one part of a serial mesh
int
nsmh_mesh_part(const char* path, int n_parts, int part, char* out, int outlen, char* err, int errlen)
{
  try {
    GenesisMesh m;
    m.ReadFile(path);
    m.KeepPart(m.RcbElementPartition(n_parts), part);
    std::ostringstream j;
    j.precision(17);
    auto list = [&](const int* p, size_t n) {
      j << "[";
      for (size_t i = 0; i < n; ++i) j << (i ? "," : "") << p[i];
      j << "]";
    };
    j << "{\"node_gid\":";
    list(m.GetNodeGlobalIds(), m.GetNumNodes());
    j << ",\"x\":[";
    for (unsigned i = 0; i < m.GetNumNodes(); ++i) j << (i ? "," : "") << m.GetCoordinatesX()[i];
    j << "],\"block_ids\":";
    auto ids = m.GetBlockIds();
    list(ids.data(), ids.size());
    j << ",\"all_block_ids\":";
    auto all = m.GetAllBlockIds();
    list(all.data(), all.size());
    j << ",\"elem_gid\":{";
    for (size_t b = 0; b < ids.size(); ++b) {
      j << (b ? "," : "") << "\"" << ids[b] << "\":";
      list(m.GetElementGlobalIdsInBlock(ids[b]).data(), m.GetElementGlobalIdsInBlock(ids[b]).size());
    }
    j << "},\"conn\":{";
    for (size_t b = 0; b < ids.size(); ++b) {
      j << (b ? "," : "") << "\"" << ids[b] << "\":";
      list(m.GetConnectivity(ids[b]), (size_t)m.GetNumElementsInBlock(ids[b]) * m.GetNumNodesPerElement(ids[b]));
    }
    j << "},\"node_sets\":{";
    bool first = true;
    for (auto const& kv : m.GetNodeSets()) {
      j << (first ? "" : ",") << "\"" << kv.first << "\":";
      list(kv.second.data(), kv.second.size());
      first = false;
    }
    j << "}}";
    return put(out, outlen, j.str(), err, errlen);
  } catch (std::exception const& e) {
    return fail(err, errlen, e.what());
  }
}

// reads a Genesis file and writes an Exodus file with `n_steps` planes of synthetic data through ExodusOutput:
// nodal displacement_{x,y,z} = (step+1) * coordinate, element "volume" = element index + step, per-point
// "ipt01_stress_xx" = 10 * element index + step.  The test reads the file back with scipy.
int
nsmh_exodus_roundtrip(const char* genesis_path, const char* out_path, int n_steps, char* err, int errlen)
{
  try {
    GenesisMesh m;
    m.ReadFile(genesis_path);
    ExodusOutput ex;
    ex.Initialize(out_path, m);
    std::map<int, std::vector<std::string>> elem_names, derived_names;
    for (int id : m.GetBlockIds()) {
      elem_names[id]    = {"ipt01_stress_xx"};
      derived_names[id] = {"volume"};
    }
    ex.InitializeDatabase(m, {}, {"displacement_x", "displacement_y", "displacement_z"}, elem_names, derived_names);
    const int n = (int)m.GetNumNodes();
    for (int s = 0; s < n_steps; ++s) {
      std::vector<std::vector<double>> node(3, std::vector<double>(n));
      for (int i = 0; i < n; ++i) {
        node[0][i] = (s + 1) * m.GetCoordinatesX()[i];
        node[1][i] = (s + 1) * m.GetCoordinatesY()[i];
        node[2][i] = (s + 1) * m.GetCoordinatesZ()[i];
      }
      std::map<int, std::vector<std::vector<double>>> elem, derived;
      for (int id : m.GetBlockIds()) {
        const int ne = m.GetNumElementsInBlock(id);
        elem[id].assign(1, std::vector<double>(ne));
        derived[id].assign(1, std::vector<double>(ne));
        for (int e = 0; e < ne; ++e) elem[id][0][e] = 10.0 * e + s, derived[id][0][e] = e + s;
      }
      if (s % 2)  // both entries: the synchronous one and the background writer (copies are moved into the job)
        ex.WriteStepAsync(0.25 * s, {}, node, elem_names, elem, derived_names, derived);
      else
        ex.WriteStep(0.25 * s, {}, node, elem_names, elem, derived_names, derived);
    }
    ex.Close();
    return 0;
  } catch (std::exception const& e) {
    return fail(err, errlen, e.what());
  }
}

// boundary-condition manager on a Genesis file: device table size, time dependence, magnitudes at time t
int
nsmh_bc_table(const char* genesis_path, const char* deck_text, double t, int max_entries, int* n_entries, int* node, int* comp,
              int* kind, double* value, int* time_dependent, char* err, int errlen)
{
  try {
    GenesisMesh m;
    m.ReadFile(genesis_path);
    Parser p;
    p.InitializeFromString(deck_text);
    BoundaryConditionManager bc;
    bc.Initialize(m.GetNodeSetNames(), m.GetNodeSets(), {}, {}, p.GetBoundaryConditionStrings(), m.GetDim(), p.TimeIntegrationScheme());
    const auto& tb = bc.GetDeviceTable();
    *n_entries     = (int)tb.node.size();
    *time_dependent = bc.HasTimeDependentMagnitudes() ? 1 : 0;
    if (*n_entries > max_entries) return fail(err, errlen, "too many entries");
    std::vector<double> X((size_t)m.GetNumNodes() * 3);
    for (unsigned i = 0; i < m.GetNumNodes(); ++i)
      X[3 * i] = m.GetCoordinatesX()[i], X[3 * i + 1] = m.GetCoordinatesY()[i], X[3 * i + 2] = m.GetCoordinatesZ()[i];
    Viewify<2> Xv(X.data(), {(int)m.GetNumNodes(), 3}, {3, 1});
    bc.EvaluateMagnitudes(t, Xv, value);
    for (int k = 0; k < *n_entries; ++k) node[k] = tb.node[k], comp[k] = tb.comp[k], kind[k] = tb.kind[k];
    return 0;
  } catch (std::exception const& e) {
    return fail(err, errlen, e.what());
  }
}

// BoundaryConditionManager::GetDevicePrograms on a deck: JSON {active, n_programs, n_slots, n_code, entries_with_program,
// slots_at_t: [...]}; the magnitudes the device would compute are checked on the GPU (tests/test_gpu_*), this exposes
// the host-side decision (all time-dependent expressions compile -> device; otherwise or with NSM_B200_HOST_BC=1 -> host)
int
nsmh_bc_programs(const char* genesis_path, const char* deck_text, double t, char* out, int outlen, char* err, int errlen)
{
  try {
    GenesisMesh m;
    m.ReadFile(genesis_path);
    Parser p;
    p.InitializeFromString(deck_text);
    BoundaryConditionManager bc;
    bc.Initialize(m.GetNodeSetNames(), m.GetNodeSets(), {}, {}, p.GetBoundaryConditionStrings(), m.GetDim(), p.TimeIntegrationScheme());
    const auto&        pr = bc.GetDevicePrograms();
    std::ostringstream j;
    j.precision(17);
    int with = 0;
    for (int v : pr.program_of_entry) with += v >= 0;
    j << "{\"active\":" << (pr.active ? "true" : "false") << ",\"time_dependent\":" << (bc.HasTimeDependentMagnitudes() ? "true" : "false")
      << ",\"n_programs\":" << (int)pr.offsets.size() - 1 << ",\"n_slots\":" << pr.slots.size() << ",\"n_code\":" << pr.code.size()
      << ",\"n_entries\":" << bc.GetDeviceTable().node.size() << ",\"entries_with_program\":" << with << ",\"slots_at_t\":[";
    std::vector<double> sv(pr.slots.size());
    bc.EvaluateSlots(t, sv.data());
    for (size_t k = 0; k < sv.size(); ++k) j << (k ? "," : "") << sv[k];
    j << "]}";
    return put(out, outlen, j.str(), err, errlen);
  } catch (std::exception const& e) {
    return fail(err, errlen, e.what());
  }
}

// ContactManager's host side on a Genesis file + deck: the entity lists CreateContactEntities sends to the device
// (no device involved).  Arrays may be null to query the counts: n[0] primary faces, n[1] contact nodes.
int
nsmh_contact_entities(const char* genesis_path, const char* deck_text, long long n[2], double* penalty, int* face_nodes, int* face_entity_ids,
                      double* face_len, int* contact_nodes, double* contact_len, char* err, int errlen)
{
  try {
    GenesisMesh m;
    m.ReadFile(genesis_path);
    Parser p;
    p.InitializeFromString(deck_text);
    if (!p.HasContact()) return fail(err, errlen, "the deck has no contact line");
    std::vector<std::string> pn, sn;
    ParseContactCommand(p.ContactString(), pn, sn, *penalty);
    std::vector<int> pi, si;
    m.BlockNamesToOnProcessorBlockIds(pn, pi);
    m.BlockNamesToOnProcessorBlockIds(sn, si);
    ContactEntityLists l;
    ContactManager::BuildEntityLists(m, pi, si, l);
    n[0] = (long long)l.primary_face_char_len.size(), n[1] = (long long)l.contact_node_ids.size();
    if (face_nodes) std::copy(l.primary_face_nodes.begin(), l.primary_face_nodes.end(), face_nodes);
    if (face_entity_ids) std::copy(l.primary_face_entity_ids.begin(), l.primary_face_entity_ids.end(), face_entity_ids);
    if (face_len) std::copy(l.primary_face_char_len.begin(), l.primary_face_char_len.end(), face_len);
    if (contact_nodes) std::copy(l.contact_node_ids.begin(), l.contact_node_ids.end(), contact_nodes);
    if (contact_len) std::copy(l.contact_node_char_len.begin(), l.contact_node_char_len.end(), contact_len);
    return 0;
  } catch (std::exception const& e) {
    return fail(err, errlen, e.what());
  }
}

// The contact visualisation database of a Genesis file + deck (ContactVisualizationDatabase, no device involved):
// n_times time planes; displacement [n_times][n_nodes][3] of the mesh nodes; plane k is written "unevaluated" (entities
// at their model coordinates) when evaluated[k] == 0; status flags all zero.
int
nsmh_contact_visualization(const char* genesis_path, const char* deck_text, const char* out_path, int n_times, const double* times,
                           const double* displacement, const int* evaluated, char* err, int errlen)
{
  try {
    GenesisMesh m;
    m.ReadFile(genesis_path);
    Parser p;
    p.InitializeFromString(deck_text);
    if (!p.HasContact() || !p.ContactVisualization()) return fail(err, errlen, "the deck asks for no contact visualization");
    std::vector<std::string> pn, sn;
    double                   penalty = 0.0;
    ParseContactCommand(p.ContactString(), pn, sn, penalty);
    std::vector<int> pi, si;
    m.BlockNamesToOnProcessorBlockIds(pn, pi);
    m.BlockNamesToOnProcessorBlockIds(sn, si);
    ContactEntityLists l;
    ContactManager::BuildEntityLists(m, pi, si, l);
    ContactVisualizationDatabase db(m, l, out_path);
    const size_t                 plane = 3 * m.GetNumNodes();
    for (int k = 0; k < n_times; ++k) db.WriteStep(times[k], evaluated[k] ? displacement + plane * (size_t)k : nullptr, nullptr, nullptr);
    db.Close();
    return 0;
  } catch (std::exception const& e) {
    return fail(err, errlen, e.what());
  }
}

// ContactManager::BuildReplicatedSubModel on `n_ranks` threads, one Genesis piece each (paths separated by '\n'): the
// sub-model every rank ends up with, as JSON {n_surface, surface_gid, primary_quads_gid, contact_nodes_gid,
// primary_char_len, contact_node_char_len, held: [per rank count], identical: true when all ranks built the same lists}
int
nsmh_contact_replicated(const char* piece_paths, const char* deck_text, int n_ranks, char* out, int outlen, char* err, int errlen)
{
  try {
    std::vector<std::string> paths;
    {
      std::istringstream in(piece_paths);
      std::string        line;
      while (std::getline(in, line))
        if (!line.empty()) paths.push_back(line);
    }
    if ((int)paths.size() != n_ranks) return fail(err, errlen, "one piece path per rank expected");
    Parser p;
    p.InitializeFromString(deck_text);
    std::vector<std::string> pn, sn;
    double                   penalty = 0.0;
    ParseContactCommand(p.ContactString(), pn, sn, penalty);
    auto                                   group = std::make_shared<RankGroup>(n_ranks);
    std::vector<ReplicatedContactSubModel> subs(n_ranks);
    std::vector<std::string>               errors(n_ranks);
    std::vector<std::thread>               threads;
    for (int r = 0; r < n_ranks; ++r)
      threads.emplace_back([&, r] {
        try {
          GenesisMesh m;
          m.ReadFile(paths[r]);
          VectorCommunicator vc(m.GetDim(), m.GetNumNodes(), group, r);
          std::vector<int>   gids(m.GetNodeGlobalIds(), m.GetNodeGlobalIds() + m.GetNumNodes());
          vc.Initialize(gids);
          std::vector<int> pi, si;
          m.BlockNamesToOnProcessorBlockIds(pn, pi);
          m.BlockNamesToOnProcessorBlockIds(sn, si);
          ContactManager::BuildReplicatedSubModel(m, vc, pi, si, subs[r]);
        } catch (std::exception const& e) {
          errors[r] = e.what();
        }
      });
    for (auto& t : threads) t.join();
    for (auto const& e : errors)
      if (!e.empty()) return fail(err, errlen, e);
    bool same = true;
    for (int r = 1; r < n_ranks; ++r)
      same = same && subs[r].surface_gid == subs[0].surface_gid && subs[r].lists.primary_face_nodes == subs[0].lists.primary_face_nodes &&
             subs[r].lists.contact_node_ids == subs[0].lists.contact_node_ids &&
             subs[r].lists.primary_face_char_len == subs[0].lists.primary_face_char_len &&
             subs[r].lists.contact_node_char_len == subs[0].lists.contact_node_char_len && subs[r].surface_xyz == subs[0].surface_xyz;
    const auto&        s0 = subs[0];
    std::ostringstream j;
    j.precision(17);
    auto ints = [&](const char* key, const std::vector<int>& v, bool to_gid) {
      j << "\"" << key << "\":[";
      for (size_t i = 0; i < v.size(); ++i) j << (i ? "," : "") << (to_gid ? s0.surface_gid[v[i]] : v[i]);
      j << "],";
    };
    auto dbls = [&](const char* key, const std::vector<double>& v) {
      j << "\"" << key << "\":[";
      for (size_t i = 0; i < v.size(); ++i) j << (i ? "," : "") << v[i];
      j << "],";
    };
    j << "{\"n_surface\":" << s0.surface_gid.size() << ",";
    ints("surface_gid", s0.surface_gid, false);
    ints("primary_quads_gid", s0.lists.primary_face_nodes, true);
    ints("contact_nodes_gid", s0.lists.contact_node_ids, true);
    dbls("primary_char_len", s0.lists.primary_face_char_len);
    dbls("contact_node_char_len", s0.lists.contact_node_char_len);
    j << "\"held\":[";
    for (int r = 0; r < n_ranks; ++r) j << (r ? "," : "") << subs[r].held_local.size();
    j << "],\"identical\":" << (same ? "true" : "false") << "}";
    return put(out, outlen, j.str(), err, errlen);
  } catch (std::exception const& e) {
    return fail(err, errlen, e.what());
  }
}

// VectorCommunicator::Initialize on `n_ranks` threads: rank r holds global ids gids[off[r] .. off[r+1]).
// Returns, for rank `query_rank`, its peers and shared local node lists (flattened).
int
nsmh_shared_node_tables(int n_ranks, const int* gids, const int* off, int query_rank, int max_out, int* n_peers, int* peer_ranks,
                        long long* pair_off, int* pair_nodes, char* err, int errlen)
{
  try {
    auto group = std::make_shared<RankGroup>(n_ranks);
    std::vector<std::shared_ptr<VectorCommunicator>> comms(n_ranks);
    std::vector<std::thread>                         th;
    for (int r = 0; r < n_ranks; ++r)
      th.emplace_back([&, r] {
        comms[r] = std::make_shared<VectorCommunicator>(3, (unsigned)(off[r + 1] - off[r]), group, r);
        comms[r]->Initialize(std::vector<int>(gids + off[r], gids + off[r + 1]));
      });
    for (auto& t : th) t.join();
    const VectorCommunicator& c = *comms[query_rank];
    *n_peers = (int)c.PeerRanks().size();
    if ((int)c.PairLocalNodes().size() > max_out) return fail(err, errlen, "output too small");
    for (int i = 0; i < *n_peers; ++i) peer_ranks[i] = c.PeerRanks()[i];
    for (int i = 0; i <= *n_peers; ++i) pair_off[i] = c.PairOffsets()[i];
    for (size_t i = 0; i < c.PairLocalNodes().size(); ++i) pair_nodes[i] = c.PairLocalNodes()[i];
    return 0;
  } catch (std::exception const& e) {
    return fail(err, errlen, e.what());
  }
}

}  // extern "C"
