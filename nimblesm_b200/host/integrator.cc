// nimblesm_b200/host/integrator.cc — see integrator.h.
#include "integrator.h"

#include "contact_manager.h"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <thread>

namespace nimble_b200 {

namespace {

thread_local int t_rank = 0;

// Viewify::operator+=(alpha * w) of the reference (src/nimble_view.h:181-214): prod = alpha * 1.0; data += prod * rhs
void
axpy(Viewify<2> dest, double alpha, const Viewify<2>& rhs)
{
  const double prod = alpha * 1.0;
  const long   n    = (long)dest.size()[0] * dest.stride()[0];
  double*      d    = dest.data();
  const double* r   = rhs.data();
  for (long i = 0; i < n; ++i) d[i] += prod * r[i];
}

// BlockMaterialInterfaceFactory of the B200 build: stress through the device seam on host arrays
class B200BlockMaterialInterfaceFactory : public BlockMaterialInterfaceFactoryBase
{
 public:
  std::shared_ptr<BlockMaterialInterfaceBase>
  create(double time_n, double time_np1, const FieldIds&, const std::vector<BlockData>& blocks, ModelDataBase* model_data_ptr) const override
  {
    auto* md = dynamic_cast<ModelData*>(model_data_ptr);
    if (!md) throw std::invalid_argument("BlockMaterialInterfaceFactory::create needs a nimble_b200::ModelData");
    return std::make_shared<BlockMaterialInterface>(time_n, time_np1, blocks, std::map<int, BlockMaterialInterface::Arrays>(), md->Device());
  }
};

}  // namespace

int
ExplicitTimeIntegrator::Integrate()
{
  DataManager&  data_manager = GetDataManager();
  const Parser& parser       = data_manager.GetParser();
  const int     my_rank      = App().Rank();
  const bool    talk         = my_rank == 0 && !App().Options().quiet;
  auto          model_base   = data_manager.GetModelData();
  auto*         model_data   = dynamic_cast<ModelData*>(model_base.get());
  if (!model_data) throw std::runtime_error("ExplicitTimeIntegrator needs a nimble_b200::ModelData");
  const int num_nodes = (int)Mesh().GetNumNodes();

  auto displacement   = model_data->GetVectorNodeData("displacement");
  auto velocity       = model_data->GetVectorNodeData("velocity");
  auto acceleration   = model_data->GetVectorNodeData("acceleration");
  auto internal_force = model_data->GetVectorNodeData("internal_force");
  auto external_force = model_data->GetVectorNodeData("external_force");

  // contact entities (explicit_time_integrator.cc:66-99): block names of the `contact:` line -> ids on this rank ->
  // skin faces / contact nodes, sent to the device once
  auto       contact_manager = GetContactManager(data_manager);
  const bool contact_enabled = parser.HasContact();
  Viewify<2> contact_force;
  if (contact_enabled) {
    std::vector<std::string> contact_primary_block_names, contact_secondary_block_names;
    double                   penalty_parameter = 0.0;
    ParseContactCommand(parser.ContactString(), contact_primary_block_names, contact_secondary_block_names, penalty_parameter);
    std::vector<int> contact_primary_block_ids, contact_secondary_block_ids;
    Mesh().BlockNamesToOnProcessorBlockIds(contact_primary_block_names, contact_primary_block_ids);
    Mesh().BlockNamesToOnProcessorBlockIds(contact_secondary_block_names, contact_secondary_block_ids);
    contact_manager->SetPenaltyParameter(penalty_parameter);
    contact_manager->CreateContactEntities(Mesh(), *data_manager.GetVectorCommunicator(), contact_primary_block_ids, contact_secondary_block_ids);
    contact_force = model_data->GetVectorNodeData("contact_force");
    if (parser.ContactVisualization())  // explicit_time_integrator.cc:93-98
      contact_manager->InitializeContactVisualization(IOFileName(parser.ContactVisualizationFileName(), "e", "out", my_rank, App().Options().num_ranks));
  }
  const bool contact_visualization = contact_enabled && parser.ContactVisualization();

  model_data->ComputeLumpedMass(data_manager);
  double critical_time_step = model_data->GetCriticalTimeStep();
  auto   group              = data_manager.GetVectorCommunicator()->Group();
  if (group && group->NumRanks() > 1) critical_time_step = group->MinAll(my_rank, critical_time_step);
  auto lumped_mass = model_data->GetScalarNodeData("lumped_mass");

  const double initial_time = parser.InitialTime(), final_time = parser.FinalTime();
  double       time_current = initial_time, time_previous = initial_time;
  const int    num_load_steps = parser.NumLoadSteps(), output_frequency = parser.OutputFrequency();
  const double user_specified_time_step = (final_time - initial_time) / num_load_steps;
  if (final_time < initial_time)
    throw std::invalid_argument("Final time: " + std::to_string(final_time) + " is less than initial time: " + std::to_string(initial_time) + "\n");

  model_data->ApplyInitialConditions(data_manager);
  model_data->ApplyKinematicConditions(data_manager, 0.0, 0.0);
  model_data->PushNodalFields();
  data_manager.WriteOutput(time_current);
  if (contact_visualization) contact_manager->ContactVisualizationWriteStep(time_current, false);  // (:152: entities at model coordinates)

  if (talk) {
    std::cout << "\nUser specified time step:              " << user_specified_time_step << std::endl;
    std::cout << "Approximate maximum stable time step:  " << critical_time_step << "\n" << std::endl;
    if (user_specified_time_step > critical_time_step)
      std::cout << "**** WARNING:  The user specified time step exceeds the computed maximum stable time step.\n" << std::endl;
    std::cout << "Explicit time integration:\n    0% complete" << std::endl;
  }
  auto is_output = [&](int step) { return output_frequency != 0 && (step % output_frequency == 0 || step == num_load_steps - 1); };
  auto progress  = [&](int step) {
    if (!talk) return;
    if (10 * (step + 1) % num_load_steps == 0 && step != num_load_steps - 1)
      std::cout << "   " << (int)(100.0 * (double)(step + 1) / num_load_steps) << "% complete" << std::endl;
    else if (step == num_load_steps - 1)
      std::cout << "  100% complete\n" << std::endl;
  };

  // the reference's timer regions (explicit_time_integrator.cc:172-279): host wall clock per region in the
  // call-by-call sequence; in the fused path the device reports the split of a run of steps (CUDA events)
  double     total_dynamics_time = 0.0, total_force_time = 0.0, total_exodus_write_time = 0.0, total_vector_reduction_time = 0.0, total_contact_time = 0.0;
  auto       seconds_since       = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); };
  const auto t0 = std::chrono::steady_clock::now();
  // contact across partitions pools the surface displacements on the host every step (ContactManager, replicated
  // sub-model): those runs are sequenced call by call
  const bool call_by_call = reference_sequence_ || (contact_enabled && contact_manager->Replicated());
  if (talk && call_by_call && !reference_sequence_)
    std::cout << "(contact across " << (group ? group->NumRanks() : 1) << " mesh partitions: steps are sequenced call by call)" << std::endl;
  if (!call_by_call) {
    // ---- fused: one device call per run of steps that ends at an output step (or at a 10 % progress mark)
    int step = 0;
    while (step < num_load_steps) {
      int  run = 0;
      bool out = false;
      while (step + run < num_load_steps) {
        out = is_output(step + run);
        const bool mark = 10 * (step + run + 1) % num_load_steps == 0;
        ++run;
        if (out || mark || run >= 4096) break;
      }
      model_data->AdvanceOnDevice(data_manager, run, time_current, user_specified_time_step, out);
      for (int s = 0; s < run; ++s) progress(step + s);
      step += run;
      if (out) {
        const auto t_out = std::chrono::steady_clock::now();
        model_data->PullNodalFields();
        data_manager.WriteOutput(time_current);
        if (contact_visualization) contact_manager->ContactVisualizationWriteStep(time_current);
        total_exodus_write_time += seconds_since(t_out);
      }
    }
    model_data->PullNodalFields();
    total_force_time    = model_data->DeviceForceSeconds();
    total_dynamics_time = model_data->DeviceUpdateSeconds();
    total_contact_time  = model_data->DeviceContactSeconds();
  } else {
    // ---- reference sequence (explicit_time_integrator.cc:177-278), host-sequenced through the ModelData virtuals
    for (int step = 0; step < num_load_steps; ++step) {
      progress(step);
      const bool is_output_step = is_output(step);
      time_previous             = time_current;
      time_current += user_specified_time_step;
      const double delta_time = time_current - time_previous, half_delta_time = 0.5 * delta_time;
      auto t_region = std::chrono::steady_clock::now();
      axpy(velocity, half_delta_time, acceleration);
      model_data->UpdateWithNewVelocity(data_manager, half_delta_time);
      model_data->ApplyKinematicConditions(data_manager, time_current, time_previous);
      axpy(displacement, delta_time, velocity);
      model_data->UpdateWithNewDisplacement(data_manager, delta_time);
      model_data->ApplyKinematicConditions(data_manager, time_current, time_previous);
      total_dynamics_time += seconds_since(t_region);
      t_region = std::chrono::steady_clock::now();
      model_data->ComputeExternalForce(data_manager, time_previous, time_current, is_output_step);
      model_data->ComputeInternalForce(data_manager, time_previous, time_current, is_output_step, displacement, internal_force);
      total_force_time += seconds_since(t_region);
      t_region = std::chrono::steady_clock::now();
      if (contact_enabled) {  // :232-249
        contact_manager->ComputeContactForce(step + 1, false, contact_force);
        total_contact_time += seconds_since(t_region);
        t_region = std::chrono::steady_clock::now();
        for (int i = 0; i < num_nodes; ++i) {
          const double one_over_m = 1.0 / lumped_mass(i);
          for (int c = 0; c < 3; ++c) acceleration(i, c) = one_over_m * (internal_force(i, c) + external_force(i, c) + contact_force(i, c));
        }
      } else {
        for (int i = 0; i < num_nodes; ++i) {
          const double one_over_m = 1.0 / lumped_mass(i);
          for (int c = 0; c < 3; ++c) acceleration(i, c) = one_over_m * (internal_force(i, c) + external_force(i, c));
        }
      }
      axpy(velocity, half_delta_time, acceleration);
      model_data->UpdateWithNewVelocity(data_manager, half_delta_time);
      total_dynamics_time += seconds_since(t_region);
      if (is_output_step) {
        const auto t_out = std::chrono::steady_clock::now();
        model_data->ApplyKinematicConditions(data_manager, time_current, time_previous);
        data_manager.WriteOutput(time_current);
        if (contact_visualization) contact_manager->ContactVisualizationWriteStep(time_current);  // :273
        total_exodus_write_time += seconds_since(t_out);
      }
      model_data->UpdateStates(data_manager);
    }
  }
  step_loop_seconds_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (group) group->Barrier();
  const int num_ranks = group ? group->NumRanks() : 1;
  if (my_rank == 0 && parser.WriteTimingDataFile()) {
    // TimingInfo::BinaryWrite (src/nimble_timing_utils.cc:70-94): nimble_timing_data_n<ranks>_<time stamp>.log, one
    // tab-separated line: ranks, simulation, internal force, contact, exodus write, vector reduction
    char        stamp[64];
    const auto  now = std::chrono::system_clock::now();
    std::time_t tt  = std::chrono::system_clock::to_time_t(now);
    const long  us  = (long)(std::chrono::duration_cast<std::chrono::microseconds>(now.time_since_epoch()).count() % 1000000);
    size_t      end = std::strftime(stamp, sizeof stamp, "%Y.%m.%d.%H.%M.%S.", std::localtime(&tt));
    std::snprintf(stamp + end, sizeof stamp - end, "%06ld", us);
    std::ofstream fs("nimble_timing_data_n" + std::to_string(num_ranks) + "_" + stamp + ".log", std::ios::binary | std::ios::trunc);
    if (!fs.is_open())
      std::cerr << "Failed to open timing data file" << std::endl;
    else
      fs << num_ranks << "\t" << step_loop_seconds_ << "\t" << total_force_time << "\t" << total_contact_time << "\t" << total_exodus_write_time << "\t"
         << total_vector_reduction_time << "\n";
  }
  if (talk) {
    // the reference's closing summary (explicit_time_integrator.cc:307-322)
    const double upd = (double)Mesh().GetNumElements() * num_load_steps / (step_loop_seconds_ > 0 ? step_loop_seconds_ : 1.0);
    std::cout << "======== Timing data: ========\n";
    std::cout << "Total step time = " << step_loop_seconds_ << " s (" << upd << " element-updates/s on rank 0, output included)\n";
    std::cout << " --- Update A, V, U: " << total_dynamics_time << (call_by_call ? "" : "  (device time; includes the shared-node exchange)") << '\n';
    std::cout << " --- Force: " << total_force_time << (call_by_call ? "" : "  (device time of the element kernels)") << "\n";
    if (contact_enabled)
      std::cout << " --- Contact time: " << total_contact_time << (call_by_call ? "" : "  (device time of the contact kernels)") << '\n';
    if (num_ranks > 1) std::cout << " --- Vector Reduction = " << total_vector_reduction_time << "  (on the device, inside the update time)\n";
    std::cout << " --- Exodus Write = " << total_exodus_write_time << "\n";
  }
  return 0;
}

std::shared_ptr<BlockMaterialInterfaceFactoryBase>
NimbleApplication::CreateBlockMaterialInterfaceFactory()
{
  return std::make_shared<B200BlockMaterialInterfaceFactory>();
}

std::unique_ptr<IntegratorBase>
NimbleApplication::CreateIntegrator(GenesisMesh& mesh, DataManager& data_manager)
{
  return std::unique_ptr<IntegratorBase>(new ExplicitTimeIntegrator(*this, mesh, data_manager, options_.reference_sequence));
}

int
NimbleApplication::Rank() const
{
  return t_rank;
}

int
NimbleApplication::Run(int argc, char** argv)
{
  RunOptions o;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--gpus" && i + 1 < argc)
      o.num_ranks = std::atoi(argv[++i]);
    else if (a == "--devices" && i + 1 < argc) {
      // CUDA device per rank; naming a device twice runs those ranks on one GPU in lockstep (RankGroup::SetLockstep)
      o.devices.clear();
      std::string list = argv[++i];
      for (size_t p = 0; p <= list.size();) {
        const size_t q = std::min(list.find(',', p), list.size());
        o.devices.push_back(std::atoi(list.substr(p, q - p).c_str()));
        p = q + 1;
      }
    } else if (a == "--assembly" && i + 1 < argc)
      o.assembly = std::string(argv[++i]) == "atomic" ? NSM_ASSEMBLY_ATOMIC : NSM_ASSEMBLY_ORDERED;
    else if (a == "--flags" && i + 1 < argc)
      o.flags = (unsigned)std::atoi(argv[++i]);
    else if (a == "--reference_sequence")
      o.reference_sequence = true;
    else if (a == "--quiet")
      o.quiet = true;
    else if (a == "--use_tpetra" || a == "--use_vt" || a == "--use_uq") {
      std::cerr << "NimbleSM_b200: option " << a << " belongs to reference subsystems outside the B200 hex8 explicit path\n";
      return 1;
    } else if (!a.empty() && a[0] == '-') {
      std::cerr << "NimbleSM_b200: unknown option " << a << "\n";
      return 1;
    } else
      o.input_file = a;
  }
  if (o.input_file.empty()) {
    std::cerr << "Usage: NimbleSM_b200 [--gpus N] [--devices a,b,...] [--assembly atomic|ordered] [--flags F] [--reference_sequence] [--quiet] <input deck>\n";
    return 1;
  }
  return Run(o);
}

int
NimbleApplication::Run(const RunOptions& options)
{
  options_ = options;
  if (options_.num_ranks < 1) options_.num_ranks = 1;
  if (options_.devices.empty())
    for (int r = 0; r < options_.num_ranks; ++r) options_.devices.push_back(r);
  if ((int)options_.devices.size() != options_.num_ranks) {
    std::cerr << "NimbleSM_b200: --devices names " << options_.devices.size() << " devices for " << options_.num_ranks << " ranks\n";
    return 1;
  }
  auto             group = std::make_shared<RankGroup>(options_.num_ranks);
  for (int r = 0; r < options_.num_ranks; ++r)
    for (int q = 0; q < r; ++q)
      if (options_.devices[q] == options_.devices[r]) group->SetLockstep(true);
  // Ranks on one GPU wait for each other INSIDE kernels; a kernel's code must therefore never be loaded lazily at
  // its first launch (module loading synchronises the device and would wait for the peer's spinning kernel).
  if (group->Lockstep()) setenv("CUDA_MODULE_LOADING", "EAGER", 1);
  std::vector<int> status(options_.num_ranks, 0);
  if (options_.num_ranks == 1) {
    status[0] = ExecRank(0, group);
  } else {
    std::vector<std::thread> threads;
    for (int r = 0; r < options_.num_ranks; ++r) threads.emplace_back([&, r] { status[r] = ExecRank(r, group); });
    for (auto& t : threads) t.join();
  }
  for (int s : status)
    if (s) return s;
  return 0;
}

int
NimbleApplication::ExecRank(int rank, std::shared_ptr<RankGroup> group)
{
  t_rank = rank;
  try {
    // InitializeSubsystems / MainLoop (src/nimble.cc:176-237, 323-372)
    auto parser = CreateParser();
    parser->SetInputFilename(options_.input_file);
    parser->SetRankID(rank);
    parser->SetNumRanks(options_.num_ranks);
    parser->Initialize();
    if (rank == 0 && !options_.quiet)
      std::cout << "\n-- NimbleSM_b200 (" << nsm_b200_version() << ")\n-- input deck " << options_.input_file << ", " << options_.num_ranks
                << " rank(s), one B200 each\n";
    GenesisMesh mesh;
    const std::string piece = IOFileName(parser->GenesisFileName(), "g", "", rank, options_.num_ranks);
    if (options_.num_ranks > 1 && !std::ifstream(piece).good() && std::ifstream(parser->GenesisFileName()).good()) {
      // no Nemesis pieces on disk: decompose the serial mesh here (the reference needs an offline SEACAS decomp run)
      mesh.ReadFile(parser->GenesisFileName());
      mesh.KeepPart(mesh.RcbElementPartition(options_.num_ranks), rank);
      if (rank == 0 && !options_.quiet)
        std::cout << "-- " << parser->GenesisFileName() << " decomposed in the driver: recursive coordinate bisection of elements into "
                  << options_.num_ranks << " parts\n";
    } else {
      mesh.ReadFile(piece);
    }
    // one GPU per rank by default.  Ranks may share a device only in lockstep (--devices 0,0; RankGroup::SetLockstep):
    // a rank's in-kernel wait for its peer must not sit in front of a device-synchronising call of that peer's thread.
    DataManager data_manager(*parser, mesh, options_.devices[rank], options_.assembly, options_.flags, group);
    data_manager.SetBlockMaterialInterfaceFactory(CreateBlockMaterialInterfaceFactory());
    data_manager.GetModelData()->InitializeBlocks(data_manager, CreateMaterialFactory());
    data_manager.InitializeOutput(IOFileName(parser->ExodusFileName(), "e", "out", rank, options_.num_ranks));
    auto integrator = CreateIntegrator(mesh, data_manager);
    const int status = integrator->Integrate();
    data_manager.GetExodusOutput()->Close();
    return status;
  } catch (std::exception const& e) {
    std::cerr << "Standard exception: " << e.what() << std::endl;
    // a rank that dies leaves the others waiting in a collective: this build runs all ranks in one process
    if (options_.num_ranks > 1) std::exit(1);
    return 1;
  }
}

}  // namespace nimble_b200
