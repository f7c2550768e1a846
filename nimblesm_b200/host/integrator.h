// nimblesm_b200/host/integrator.h — integrator drivers: IntegratorBase (src/integrators/integrator_base.h:54-70),
// ExplicitTimeIntegrator (src/integrators/explicit_time_integrator.{h,cc}) and the application shell
// NimbleApplication with the reference's factory customisation points (src/nimble.h:85-91, src/nimble.cc:261-372).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "data_manager.h"

namespace nimble_b200 {

class NimbleApplication;

class IntegratorBase
{
 public:
  IntegratorBase(NimbleApplication& app, GenesisMesh& mesh, DataManager& data_manager)
      : nimble_app_(&app), mesh_(&mesh), data_manager_(&data_manager)
  {
  }
  virtual ~IntegratorBase() = default;
  virtual int
  Integrate() = 0;
  NimbleApplication&
  App() noexcept
  {
    return *nimble_app_;
  }
  GenesisMesh&
  Mesh() noexcept
  {
    return *mesh_;
  }
  DataManager&
  GetDataManager() noexcept
  {
    return *data_manager_;
  }

 private:
  NimbleApplication* nimble_app_;
  GenesisMesh*       mesh_;
  DataManager*       data_manager_;
};

// Central-difference / velocity-Verlet loop of the reference (explicit_time_integrator.cc:60-330).
//   fused (default)     runs of steps between output steps execute as ONE device call (ModelData::AdvanceOnDevice):
//                       v, u, BC, f_int, a never leave the GPU; the host sees fields on output steps only
//   reference sequence  the loop body is issued call by call through the ModelDataBase virtuals on host views,
//                       exactly as ExplicitTimeIntegrator::Integrate does (axpy on the host, UpdateWithNewVelocity /
//                       UpdateWithNewDisplacement / ComputeInternalForce crossing to the device) -- what a drop-in
//                       ModelData sees inside the unmodified reference integrator
class ExplicitTimeIntegrator : public IntegratorBase
{
 public:
  ExplicitTimeIntegrator(NimbleApplication& app, GenesisMesh& mesh, DataManager& data_manager, bool reference_sequence = false)
      : IntegratorBase(app, mesh, data_manager), reference_sequence_(reference_sequence)
  {
  }
  int
  Integrate() override;
  double
  StepLoopSeconds() const
  {
    return step_loop_seconds_;
  }

 private:
  bool   reference_sequence_;
  double step_loop_seconds_ = 0.0;
};

struct RunOptions
{
  std::string input_file;
  int         num_ranks          = 1;     // one thread + one GPU per rank
  std::vector<int> devices;               // CUDA device of every rank (--devices a,b,...); default: rank r -> device r
  int         assembly           = NSM_ASSEMBLY_ORDERED;
  unsigned    flags              = NSM_FLAG_CACHE_REF_JACOBIAN;
  bool        reference_sequence = false;
  bool        quiet              = false;
};

class NimbleApplication
{
 public:
  NimbleApplication()          = default;
  virtual ~NimbleApplication() = default;
  // CLI: NimbleSM_b200 [--gpus N] [--assembly atomic|ordered] [--reference_sequence] [--quiet] <input deck>
  // Returns the process exit code; parse / setup errors are reported and give 1 (src/nimble.cc:129-137).
  int
  Run(int argc, char** argv);
  int
  Run(const RunOptions& options);
  int
  Rank() const;
  int
  NumRanks() const
  {
    return options_.num_ranks;
  }
  const RunOptions&
  Options() const
  {
    return options_;
  }

 protected:
  // customisation points of the reference application (src/nimble.h:85-91)
  virtual std::unique_ptr<Parser>
  CreateParser()
  {
    return std::unique_ptr<Parser>(new Parser());
  }
  virtual std::shared_ptr<MaterialFactoryBase>
  CreateMaterialFactory()
  {
    return std::make_shared<MaterialFactory>();
  }
  virtual std::shared_ptr<BlockMaterialInterfaceFactoryBase>
  CreateBlockMaterialInterfaceFactory();
  virtual std::unique_ptr<IntegratorBase>
  CreateIntegrator(GenesisMesh& mesh, DataManager& data_manager);

 private:
  int
  ExecRank(int rank, std::shared_ptr<RankGroup> group);
  RunOptions options_;
};

}  // namespace nimble_b200
