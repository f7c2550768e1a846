// oracle/ref_glue.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin C-ABI shim that links the UNMODIFIED NimbleSM serial sources (compiled in place from
// /root/reference/src by oracle/Makefile into oracle/_ref/libnimble_ref.so) so that tests and the
// cpu_baseline / --impl reference legs of bench.py can drive the reference's own CPU path:
//   * a deck + in-memory mesh  ->  nimble::Parser / GenesisMesh / DataManager / serial ModelData
//   * the explicit central-difference loop of src/integrators/explicit_time_integrator.cc:123-278
//     (re-stated here because that TU pulls in the contact manager and cannot be compiled)
//   * the six out-of-line DataManager members of src/nimble_data_manager.cc:70-198 (re-stated because
//     that TU hard-wires nimble_kokkos::ModelData, :113)
// All arithmetic (element, material, utils, block, BC manager, expression parser) is the reference's.
// Nothing under nimblesm_b200/ may include or link this file.
#include "nimble_expression_parser.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "nimble_block.h"
#include "nimble_boundary_condition_manager.h"
#include "nimble_data_manager.h"
#include "nimble_genesis_mesh.h"
#include "nimble_material_factory.h"
#include "nimble_model_data.h"
#include "nimble_parser.h"
#include "nimble_vector_communicator.h"
#include "nimble_view.h"
#include "ref_contact.h"
#include "ref_state_material.h"
#ifdef NSM_REF_BINDING  // tests/ref_binding: the same glue with the B200 binding class as the model data
#include "b200_model_data.h"
#endif

// ---------------------------------------------------------------------------------------------
// DataManager out-of-line members (follows src/nimble_data_manager.cc:70-198; serial ModelData)
// ---------------------------------------------------------------------------------------------
namespace nimble {

DataManager::DataManager(const nimble::Parser& parser, const nimble::GenesisMesh& mesh)
    : parser_(parser),
      mesh_(mesh),
      model_data_(),
      field_ids_(),
      vector_communicator_(nullptr),
      boundary_condition_(new nimble::BoundaryConditionManager())
{
  Initialize();
}

void
DataManager::Initialize()
{
  const auto dim       = static_cast<int>(mesh_.GetDim());
  const auto num_nodes = static_cast<int>(mesh_.GetNumNodes());
  int        comm      = 0;
  vector_communicator_ = std::make_shared<nimble::VectorCommunicator>(dim, num_nodes, comm);
  std::vector<int> global_node_ids(num_nodes);
  int const* const gids = mesh_.GetNodeGlobalIds();
  for (int n = 0; n < num_nodes; ++n) global_node_ids[n] = gids[n];
  vector_communicator_->Initialize(global_node_ids);

#ifdef NSM_REF_BINDING
  model_data_ = std::make_shared<nsm_binding::B200ModelData>();  // the one-line selection of INTEGRATION.md §2
#else
  model_data_ = std::make_shared<nimble::ModelData>();  // serial path (reference: nimble_kokkos::ModelData)
#endif
  model_data_->SetDimension(dim);

  boundary_condition_->Initialize(
      mesh_.GetNodeSetNames(),
      mesh_.GetNodeSets(),
      mesh_.GetSideSetNames(),
      mesh_.GetSideSets(),
      parser_.GetBoundaryConditionStrings(),
      dim,
      parser_.TimeIntegrationScheme());

  field_ids_.lumped_mass           = model_data_->AllocateNodeData(nimble::SCALAR, "lumped_mass", num_nodes);
  field_ids_.reference_coordinates = model_data_->AllocateNodeData(nimble::VECTOR, "reference_coordinate", num_nodes);
  field_ids_.displacement          = model_data_->AllocateNodeData(nimble::VECTOR, "displacement", num_nodes);
  field_ids_.velocity              = model_data_->AllocateNodeData(nimble::VECTOR, "velocity", num_nodes);
  field_ids_.acceleration          = model_data_->AllocateNodeData(nimble::VECTOR, "acceleration", num_nodes);
  field_ids_.internal_force        = model_data_->AllocateNodeData(nimble::VECTOR, "internal_force", num_nodes);
  field_ids_.external_force        = model_data_->AllocateNodeData(nimble::VECTOR, "external_force", num_nodes);
  field_ids_.contact_force         = model_data_->AllocateNodeData(nimble::VECTOR, "contact_force", num_nodes);
  model_data_->SetReferenceCoordinates(mesh_);
}

void
DataManager::InitializeOutput(const std::string& filename)
{
  std::vector<std::string> global_data_labels;
  exodus_output_ = std::shared_ptr<nimble::ExodusOutput>(new nimble::ExodusOutput);
  exodus_output_->Initialize(filename, mesh_);
  model_data_->InitializeExodusOutput(*this);
  exodus_output_->InitializeDatabase(
      mesh_,
      global_data_labels,
      model_data_->GetNodeDataLabelsForOutput(),
      model_data_->GetElementDataLabelsForOutput(),
      model_data_->GetDerivedElementDataLabelsForOutput());
}

void
DataManager::WriteOutput(double time_current)
{
  model_data_->WriteExodusOutput(*this, time_current);
}

void
DataManager::SetBlockMaterialInterfaceFactory(
    const std::shared_ptr<nimble::BlockMaterialInterfaceFactoryBase>& block_material_factory)
{
  block_material_factory_ = block_material_factory;
}

const std::shared_ptr<nimble::BlockMaterialInterfaceFactoryBase>&
DataManager::GetBlockMaterialInterfaceFactory() const
{
  return block_material_factory_;
}

}  // namespace nimble

namespace {

// The reference's array-initialised mesh has no node-set setter; members are protected
// (src/nimble_genesis_mesh.h:328-355), so a subclass fills them.
struct ArrayMesh : nimble::GenesisMesh
{
  void
  SetNodeSets(int n_sets, const int* ids, const int* sizes, const int* nodes)
  {
    long off = 0;
    for (int s = 0; s < n_sets; ++s) {
      int id = ids[s];
      node_set_ids_.push_back(id);
      node_set_names_[id] = "nodelist_" + std::to_string(id);
      node_sets_[id]      = std::vector<int>(nodes + off, nodes + off + sizes[s]);
      off += sizes[s];
    }
  }
};

struct Snapshot
{
  double                                          time;
  std::map<std::string, std::vector<double>>      node;     // label -> AoS copy
  std::map<int, std::vector<double>>              elem;     // block -> [nelem][ndata] np1 copy
  std::map<int, std::vector<std::vector<double>>> derived;  // block -> [label][elem]
};

struct RefRun
{
  nimble::Parser                           parser;
  ArrayMesh                                mesh;
  std::unique_ptr<nimble::DataManager>     dm;
  std::shared_ptr<nimble::MaterialFactory> factory;
  nimble::ModelData*                       md = nullptr;
  // loop state (explicit_time_integrator.cc:126-136)
  double                time_current = 0.0, time_previous = 0.0, dt_user = 0.0;
  int                   step = 0, num_load_steps = 0, output_frequency = 0;
  bool                  keep_snapshots = false;
  bool                  write_output   = false;  // also through the reference's own ExodusOutput (text form in this build)
  std::vector<Snapshot> snaps;
  std::string           err;
  // penalty contact (decks with a `contact:` line): the reference's own ContactEntity objects, see ref_contact.h
  std::unique_ptr<nsm_oracle::RefContact> contact;
  long                                    contact_pairs_last = 0;
};

void
take_snapshot(RefRun& r)
{
  Snapshot s;
  s.time = r.time_current;
  for (const char* lbl :
       {"lumped_mass", "reference_coordinate", "displacement", "velocity", "acceleration", "internal_force",
        "external_force", "contact_force"}) {
    int id = r.md->GetFieldId(lbl);
    if (id < 0) continue;
    bool   scalar = std::string(lbl) == "lumped_mass";
    int    n      = static_cast<int>(r.mesh.GetNumNodes()) * (scalar ? 1 : 3);
    double* p     = r.md->GetNodeData(id);
    s.node[lbl]   = std::vector<double>(p, p + n);
  }
  auto ref  = r.md->GetVectorNodeData("reference_coordinate");
  auto disp = r.md->GetVectorNodeData("displacement");
  for (auto& kv : r.md->GetBlocks()) {
    int   bid = kv.first;
    auto& np1 = r.md->GetElementDataNew(bid);
    s.elem[bid] = np1;
    // derived data: volume + volume averages of every ipt field (F 9 + sigma 6)
    // (Block::ComputeDerivedElementData, src/nimble_block.cc:438-497) — driven by the deck's output fields.
    auto const& dlabels = r.md->GetDerivedElementDataLabelsForOutput().at(bid);
    std::vector<std::vector<double>> derived;
    kv.second->ComputeDerivedElementData(
        ref.data(),
        disp.data(),
        r.mesh.GetNumElementsInBlock(bid),
        r.mesh.GetConnectivity(bid),
        static_cast<int>(r.md->GetElementDataLabels().at(bid).size()),
        np1,
        static_cast<int>(dlabels.size()),
        derived);
    s.derived[bid] = derived;
  }
  r.snaps.push_back(std::move(s));
  if (r.write_output) r.dm->WriteOutput(r.time_current);  // data_manager.WriteOutput (explicit_time_integrator.cc:149, 270)
}

}  // namespace

extern "C" {

const char*
nsmref_last_error(void* h)
{
  return static_cast<RefRun*>(h)->err.c_str();
}

// conn: concatenated per block, 0-based local node ids, 8 per element; elem_gid 0-based; node sets 0-based.
void*
nsmref_open(
    const char*   deck_path,
    int           n_nodes,
    const int*    node_gid,
    const double* x,
    const double* y,
    const double* z,
    int           n_blocks,
    const int*    block_ids,
    const int*    block_nelem,
    const int*    conn,
    const int*    elem_gid,
    int           n_nodesets,
    const int*    ns_ids,
    const int*    ns_sizes,
    const int*    ns_nodes,
    int           keep_snapshots)
{
  auto* r = new RefRun;
  try {
    r->parser.SetInputFilename(deck_path);
    r->parser.Initialize();
    std::vector<int>                gid(node_gid, node_gid + n_nodes);
    std::vector<double>             vx(x, x + n_nodes), vy(y, y + n_nodes), vz(z, z + n_nodes);
    std::vector<int>                bids(block_ids, block_ids + n_blocks);
    std::map<int, std::string>      bnames;
    std::map<int, std::vector<int>> belem, bconn;
    std::map<int, int>              bnpe;
    std::vector<int>                egid;
    long                            eoff = 0;
    for (int b = 0; b < n_blocks; ++b) {
      int id     = block_ids[b];
      bnames[id] = "block_" + std::to_string(id);
      bnpe[id]   = 8;
      belem[id]  = std::vector<int>(elem_gid + eoff, elem_gid + eoff + block_nelem[b]);
      bconn[id]  = std::vector<int>(conn + 8 * eoff, conn + 8 * (eoff + block_nelem[b]));
      egid.insert(egid.end(), belem[id].begin(), belem[id].end());
      eoff += block_nelem[b];
    }
    r->mesh.Initialize("in_memory.g", gid, vx, vy, vz, egid, bids, bnames, belem, bnpe, bconn);
    r->mesh.SetNodeSets(n_nodesets, ns_ids, ns_sizes, ns_nodes);
    r->dm.reset(new nimble::DataManager(r->parser, r->mesh));
    r->factory = std::make_shared<nsm_oracle::StateMaterialFactory>();  // nimble::MaterialFactory + the test-only state material
    r->md      = dynamic_cast<nimble::ModelData*>(r->dm->GetModelData().get());
    r->md->InitializeBlocks(*r->dm, r->factory);  // src/nimble.cc:362
    r->md->InitializeExodusOutput(*r->dm);
    r->keep_snapshots = keep_snapshots != 0;
    if (r->parser.HasContact()) {
      // explicit_time_integrator.cc:76-92: block names of the `contact:` line -> ids on this rank -> contact entities
      std::vector<std::string> primary_names, secondary_names;
      double                   penalty = 0.0;
      nsm_oracle::RefContact::ParseCommand(r->parser.ContactString(), primary_names, secondary_names, penalty);
      std::vector<int> primary_ids, secondary_ids;
      r->mesh.BlockNamesToOnProcessorBlockIds(primary_names, primary_ids);
      r->mesh.BlockNamesToOnProcessorBlockIds(secondary_names, secondary_ids);
      r->contact.reset(new nsm_oracle::RefContact);
      r->contact->Create(r->mesh, primary_ids, secondary_ids, penalty);
    }
  } catch (std::exception const& e) {
    r->err = e.what();
  }
  return r;
}

// Route every snapshot through ModelData::WriteExodusOutput -> nimble::ExodusOutput as well (without
// NIMBLE_HAVE_EXODUS the reference writes a text file, src/nimble_exodus_output.cc:86-88, 326-549).  Call before begin.
int
nsmref_enable_output(void* h, const char* filename)
{
  auto& r = *static_cast<RefRun*>(h);
  try {
    r.dm->InitializeOutput(filename);
    r.write_output = true;
    return 0;
  } catch (std::exception const& e) {
    r.err = e.what();
    return 1;
  }
}

// kernels launched by the device library behind the model data (0 for the serial reference): proof of which path ran
long
nsmref_device_launches(void* h)
{
#ifdef NSM_REF_BINDING
  auto* md = dynamic_cast<nsm_binding::B200ModelData*>(static_cast<RefRun*>(h)->md);
  return md ? md->DeviceLaunches() : -1;
#else
  (void)h;
  return 0;
#endif
}

void
nsmref_close(void* h)
{
  delete static_cast<RefRun*>(h);
}

// Pre-loop part of ExplicitTimeIntegrator::Integrate (explicit_time_integrator.cc:123-147).
double
nsmref_begin(void* h)
{
  auto& r = *static_cast<RefRun*>(h);
  r.md->ComputeLumpedMass(*r.dm);
  double initial_time = r.parser.InitialTime();
  double final_time   = r.parser.FinalTime();
  r.time_current = r.time_previous = initial_time;
  r.num_load_steps                 = r.parser.NumLoadSteps();
  r.output_frequency               = r.parser.OutputFrequency();
  r.dt_user                        = (final_time - initial_time) / r.num_load_steps;
  r.step                           = 0;
  r.md->ApplyInitialConditions(*r.dm);
  r.md->ApplyKinematicConditions(*r.dm, 0.0, 0.0);
  if (r.keep_snapshots) take_snapshot(r);  // data_manager.WriteOutput(time_current), :149
  return r.md->GetCriticalTimeStep();
}

// n passes of the loop body (explicit_time_integrator.cc:177-278); the contact branch (:232-249) runs when the deck has
// a `contact:` line, through RefContact.
double
nsmref_advance(void* h, int n)
{
  auto& r              = *static_cast<RefRun*>(h);
  auto& model_data     = *r.md;
  auto& data_manager   = *r.dm;
  auto  displacement   = model_data.GetVectorNodeData("displacement");
  auto  velocity       = model_data.GetVectorNodeData("velocity");
  auto  acceleration   = model_data.GetVectorNodeData("acceleration");
  auto  internal_force = model_data.GetVectorNodeData("internal_force");
  auto  external_force = model_data.GetVectorNodeData("external_force");
  auto const lumped_mass = model_data.GetScalarNodeData("lumped_mass");
  const int  num_nodes   = static_cast<int>(r.mesh.GetNumNodes());

  for (int i = 0; i < n; ++i, ++r.step) {
    const int step           = r.step;
    bool      is_output_step = false;
    if (r.output_frequency != 0) {
      if (step % r.output_frequency == 0 || step == r.num_load_steps - 1) is_output_step = true;
    }
    r.time_previous = r.time_current;
    r.time_current += r.dt_user;
    const double delta_time      = r.time_current - r.time_previous;
    const double half_delta_time = 0.5 * delta_time;

    velocity += half_delta_time * acceleration;
    model_data.UpdateWithNewVelocity(data_manager, half_delta_time);
    model_data.ApplyKinematicConditions(data_manager, r.time_current, r.time_previous);
    displacement += delta_time * velocity;
    model_data.UpdateWithNewDisplacement(data_manager, delta_time);
    model_data.ApplyKinematicConditions(data_manager, r.time_current, r.time_previous);
    model_data.ComputeExternalForce(data_manager, r.time_previous, r.time_current, is_output_step);
    model_data.ComputeInternalForce(
        data_manager, r.time_previous, r.time_current, is_output_step, displacement, internal_force);
    if (r.contact) {
      auto contact_force   = model_data.GetVectorNodeData("contact_force");
      r.contact_pairs_last = r.contact->Compute(displacement.data(), contact_force.data(), num_nodes);
      for (int k = 0; k < num_nodes; ++k) {
        const double oneOverM = 1.0 / lumped_mass(k);
        acceleration(k, 0)    = oneOverM * (internal_force(k, 0) + external_force(k, 0) + contact_force(k, 0));
        acceleration(k, 1)    = oneOverM * (internal_force(k, 1) + external_force(k, 1) + contact_force(k, 1));
        acceleration(k, 2)    = oneOverM * (internal_force(k, 2) + external_force(k, 2) + contact_force(k, 2));
      }
    } else {
      for (int k = 0; k < num_nodes; ++k) {
        const double oneOverM = 1.0 / lumped_mass(k);
        acceleration(k, 0)    = oneOverM * (internal_force(k, 0) + external_force(k, 0));
        acceleration(k, 1)    = oneOverM * (internal_force(k, 1) + external_force(k, 1));
        acceleration(k, 2)    = oneOverM * (internal_force(k, 2) + external_force(k, 2));
      }
    }
    velocity += half_delta_time * acceleration;
    model_data.UpdateWithNewVelocity(data_manager, half_delta_time);
    if (is_output_step) {
      model_data.ApplyKinematicConditions(data_manager, r.time_current, r.time_previous);
      if (r.keep_snapshots) take_snapshot(r);
    }
    model_data.UpdateStates(data_manager);
  }
  return r.time_current;
}

// ---- contact: what RefContact built and one evaluation on a given displacement -------------------------------
// which: 0 primary quads (4 ints each), 1 secondary quads, 2 contact node ids; returns the count (quads / nodes) and
// fills out when it is non-null
long
nsmref_contact_ints(void* h, int which, int* out)
{
  auto& r = *static_cast<RefRun*>(h);
  if (!r.contact) return -1;
  const std::vector<int>& v = which == 0 ? r.contact->primary_quads : which == 1 ? r.contact->secondary_quads : r.contact->contact_node_ids;
  if (out) std::copy(v.begin(), v.end(), out);
  return (long)v.size() / (which == 2 ? 1 : 4);
}

// which: 0 characteristic lengths of the primary quads, 1 of the contact nodes
long
nsmref_contact_doubles(void* h, int which, double* out)
{
  auto& r = *static_cast<RefRun*>(h);
  if (!r.contact) return -1;
  const std::vector<double>& v = which == 0 ? r.contact->primary_char_len : r.contact->contact_node_char_len;
  if (out) std::copy(v.begin(), v.end(), out);
  return (long)v.size();
}

// contact force of a displacement field ([n][3] in, [n][3] out); returns the number of enforced pairs
long
nsmref_contact_force(void* h, const double* displacement, double* contact_force)
{
  auto& r = *static_cast<RefRun*>(h);
  if (!r.contact) return -1;
  return r.contact->Compute(displacement, contact_force, static_cast<int>(r.mesh.GetNumNodes()));
}

long
nsmref_contact_pairs_last(void* h)
{
  return static_cast<RefRun*>(h)->contact_pairs_last;
}

// ModelData::ComputeInternalForce alone on the current displacement (no state roll).
void
nsmref_internal_force(void* h)
{
  auto& r     = *static_cast<RefRun*>(h);
  auto  disp  = r.md->GetVectorNodeData("displacement");
  auto  force = r.md->GetVectorNodeData("internal_force");
  r.md->ComputeInternalForce(*r.dm, 0.0, 0.0, false, disp, force);
}

int
nsmref_num_nodes(void* h)
{
  return static_cast<int>(static_cast<RefRun*>(h)->mesh.GetNumNodes());
}

// Pointer to the live AoS storage of a nodal field (writable: tests set displacement directly).
double*
nsmref_node_field(void* h, const char* label)
{
  auto& r  = *static_cast<RefRun*>(h);
  int   id = r.md->GetFieldId(label);
  if (id < 0) return nullptr;
  return r.md->GetNodeData(id);
}

// which = 0: element_data_np1 (as written by the last force call), 1: element_data_n
// (after UpdateStates this holds the newest state).  Layout [elem][ipt][F9, sigma6].
long
nsmref_elem_data(void* h, int block_id, int which, double* out)
{
  auto& r = *static_cast<RefRun*>(h);
  auto& v = which == 0 ? r.md->GetElementDataNew(block_id) : r.md->GetElementDataOld(block_id);
  if (out) std::memcpy(out, v.data(), v.size() * sizeof(double));
  return static_cast<long>(v.size());
}

// doubles per integration point of a block: 15 (F 9, sigma 6) + the material's state variables
int
nsmref_elem_stride(void* h, int block_id)
{
  auto& r = *static_cast<RefRun*>(h);
  return static_cast<int>(r.md->GetElementDataLabels().at(block_id).size()) / 8;
}

// The material seam alone, through the reference's Material virtual (src/nimble_material.h:241-256): n_points
// integration points, state arrays [n_points][n_state].  Returns the material's state-variable count, -1 on error.
int
nsmref_material_stress(const char* material_string, int n_points, const double* F_n, const double* F_np1, const double* s_n,
                       double* s_np1, const double* state_n, double* state_np1)
{
  try {
    nsm_oracle::StateMaterialFactory factory;
    factory.parse_and_create(material_string);
    auto mat = factory.get_material();
    alignas(16) static char dm_storage[sizeof(nimble::DataManager)];
    nimble::DataManager&    dm = *reinterpret_cast<nimble::DataManager*>(dm_storage);  // forwarded, never dereferenced
    mat->GetStress(0, n_points, 0.0, 0.0, F_n, F_np1, s_n, s_np1, state_n, state_np1, dm, false);
    return mat->NumStateVariables();
  } catch (...) {
    return -1;
  }
}

int
nsmref_num_snapshots(void* h)
{
  return static_cast<int>(static_cast<RefRun*>(h)->snaps.size());
}

double
nsmref_snapshot_time(void* h, int i)
{
  return static_cast<RefRun*>(h)->snaps.at(i).time;
}

long
nsmref_snapshot_node(void* h, int i, const char* label, double* out)
{
  auto& s  = static_cast<RefRun*>(h)->snaps.at(i);
  auto  it = s.node.find(label);
  if (it == s.node.end()) return -1;
  if (out) std::memcpy(out, it->second.data(), it->second.size() * sizeof(double));
  return static_cast<long>(it->second.size());
}

long
nsmref_snapshot_elem(void* h, int i, int block_id, double* out)
{
  auto& v = static_cast<RefRun*>(h)->snaps.at(i).elem.at(block_id);
  if (out) std::memcpy(out, v.data(), v.size() * sizeof(double));
  return static_cast<long>(v.size());
}

// Derived labels of a block joined by '\n' (e.g. "volume\nstress_xx\n...") in storage order.
long
nsmref_derived_labels(void* h, int block_id, char* out, long cap)
{
  auto&       r = *static_cast<RefRun*>(h);
  std::string s;
  for (auto const& l : r.md->GetDerivedElementDataLabelsForOutput().at(block_id)) s += l + "\n";
  if (out && cap > 0) {
    std::strncpy(out, s.c_str(), static_cast<size_t>(cap - 1));
    out[cap - 1] = 0;
  }
  return static_cast<long>(s.size());
}

long
nsmref_snapshot_derived(void* h, int i, int block_id, int label_index, double* out)
{
  auto& v = static_cast<RefRun*>(h)->snaps.at(i).derived.at(block_id).at(label_index);
  if (out) std::memcpy(out, v.data(), v.size() * sizeof(double));
  return static_cast<long>(v.size());
}

// ---------------------------------------------------------------------------------------------
// CPU baseline: the unmodified reference per-element code (nimble::Block::ComputeInternalForce)
// + the node-wise loop of explicit_time_integrator.cc, on T contiguous element chunks, one
// Block + element-data vector + private force buffer per chunk (rank-style decomposition without
// MPI; the reference itself has no threaded serial path).  Returns seconds for `steps` steps.
// material_string e.g. "neohookean density 7.8 bulk_modulus 1.6e12 shear_modulus 0.8e12".
// Node arrays are AoS [n][3]; u, v, a are advanced in place; f receives the last internal force.
// ---------------------------------------------------------------------------------------------
double
nsmref_bench_steps(
    const char*   material_string,
    int           n_nodes,
    const double* ref_coord,
    int           n_elem,
    const int*    conn,
    const double* lumped_mass,
    double*       u,
    double*       v,
    double*       a,
    double*       f,
    double        dt,
    int           steps,
    int           threads)
{
  if (threads < 1) threads = 1;
  struct Chunk
  {
    nimble::Block       block;
    int                 e0, e1;
    int                 n0 = 0, n1 = 0;  // node range touched by this chunk (private force buffer covers it)
    std::vector<double> data_n, data_np1, force;
    std::vector<int>    gids;
  };
  nsm_oracle::StateMaterialFactory    factory;
  std::vector<std::unique_ptr<Chunk>> chunks;
  std::vector<std::string>            labels(120, "x");
  nimble::Parser                      parser;
  ArrayMesh                           mesh;
  // Block::ComputeInternalForce only forwards the DataManager reference to Material::GetStress,
  // which ignores it (src/nimble_material.cc:68-93, 226-250); a never-dereferenced reference suffices.
  alignas(16) static char dm_storage[sizeof(nimble::DataManager)];
  nimble::DataManager&    dm = *reinterpret_cast<nimble::DataManager*>(dm_storage);
  for (int t = 0; t < threads; ++t) {
    auto c = std::unique_ptr<Chunk>(new Chunk);
    c->e0  = static_cast<int>(static_cast<long>(n_elem) * t / threads);
    c->e1  = static_cast<int>(static_cast<long>(n_elem) * (t + 1) / threads);
    c->block.Initialize(material_string, factory);
    int ne = c->e1 - c->e0;
    c->n0 = n_nodes;
    c->n1 = 0;
    for (long i = 8L * c->e0; i < 8L * c->e1; ++i) {
      c->n0 = std::min(c->n0, conn[i]);
      c->n1 = std::max(c->n1, conn[i] + 1);
    }
    if (ne == 0) c->n0 = c->n1 = 0;
    if (threads > 1) c->force.assign(static_cast<size_t>(c->n1 - c->n0) * 3, 0.0);
    c->gids.resize(ne);
    for (int e = 0; e < ne; ++e) c->gids[e] = c->e0 + e;
    // offsets tables: labels in the order of Block::GetDataLabelsAndLengths (src/nimble_block.cc:84-108)
    std::vector<std::pair<std::string, nimble::Length>> ll;
    c->block.GetDataLabelsAndLengths(ll);
    std::vector<std::string> comp;
    for (auto& p : ll) {
      auto cl = nimble::GetComponentLabels(p.first, p.second, 3);
      comp.insert(comp.end(), cl.begin(), cl.end());
    }
    labels = comp;
    c->data_n.assign(static_cast<size_t>(ne) * comp.size(), 0.0);
    c->data_np1.assign(static_cast<size_t>(ne) * comp.size(), 0.0);
    std::vector<std::string> none;
    c->block.InitializeElementData(ne, c->gids, comp, none, c->data_n, c->data_np1, factory, dm);
    chunks.push_back(std::move(c));
  }
  const long   ndof = static_cast<long>(n_nodes) * 3;
  const double hdt  = 0.5 * dt;
  auto         t0   = std::chrono::steady_clock::now();
  for (int s = 0; s < steps; ++s) {
    for (long i = 0; i < ndof; ++i) v[i] += hdt * a[i];
    for (long i = 0; i < ndof; ++i) u[i] += dt * v[i];
    auto work = [&](int t) {
      Chunk& c = *chunks[t];
      // private buffer is indexed by global node id through an offset base pointer
      double* dst = threads > 1 ? c.force.data() - 3L * c.n0 : f;
      if (threads > 1)
        std::fill(c.force.begin(), c.force.end(), 0.0);
      else
        std::fill(f, f + ndof, 0.0);
      c.block.ComputeInternalForce(
          ref_coord, u, v, dst, 0.0, dt, c.e1 - c.e0, conn + 8L * c.e0, c.gids.data(), labels, c.data_n, c.data_np1,
          dm, false);
      c.data_n.swap(c.data_np1);
    };
    if (threads == 1) {
      work(0);
    } else {
      std::vector<std::thread> pool;
      for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
      for (auto& th : pool) th.join();
      std::fill(f, f + ndof, 0.0);
      for (int t = 0; t < threads; ++t) {
        const double* src = chunks[t]->force.data() - 3L * chunks[t]->n0;
        for (long i = 3L * chunks[t]->n0; i < 3L * chunks[t]->n1; ++i) f[i] += src[i];
      }
    }
    for (int k = 0; k < n_nodes; ++k) {
      const double oneOverM = 1.0 / lumped_mass[k];
      a[3 * k + 0]          = oneOverM * (f[3 * k + 0] + 0.0);
      a[3 * k + 1]          = oneOverM * (f[3 * k + 1] + 0.0);
      a[3 * k + 2]          = oneOverM * (f[3 * k + 2] + 0.0);
    }
    for (long i = 0; i < ndof; ++i) v[i] += hdt * a[i];
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}


// The reference's own expression parser on one point (src/nimble_expression_parser.h:694-760): pins
// nimblesm_b200/host/expression.cc.  Returns 0 and the value, or 1 when the reference throws.
int
nsmref_expression_eval(const char* text, double x, double y, double z, double t, double* out)
{
  try {
    ExpressionParsing::BoundaryConditionFunctor f{std::string(text)};
    f.x = x, f.y = y, f.z = z, f.t = t;
    *out = f.eval();
    return 0;
  } catch (...) {
    return 1;
  }
}

}  // extern "C"
