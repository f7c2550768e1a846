// tests/ref_binding/b200_model_data.h — THE BINDING OF INTEGRATION.md §2, COMPILED AGAINST THE REFERENCE'S HEADERS.
//
// Test infrastructure: this translation unit includes /root/reference/src/*.h, so it is built only where the
// reference tree exists (tests/ref_binding/Makefile, output tests/ref_binding/_build/libref_binding.so, which travels
// to the GPU box like the other prebuilt checkers).  It is what a NimbleSM maintainer adds to the reference tree to
// put the B200 library behind the reference's own model-data interface:
//
//   class B200ModelData : public nimble::ModelData          (a nimble::ModelDataBase, src/nimble_model_data_base.h:86-258)
//
// The serial nimble::ModelData keeps everything that is bookkeeping -- nodal storage and Viewify views, element-data
// labels and containers, output-field selection, the Exodus writer calls -- and the four virtuals that COMPUTE are
// overridden to call the C ABI of include/nsm_b200.h: InitializeBlocks (adds the device model), ComputeLumpedMass,
// ComputeInternalForce, UpdateStates.  With it the reference's unmodified BoundaryConditionManager, DataManager glue
// and explicit loop (oracle/ref_glue.cc restates only src/nimble_data_manager.cc:70-198 and
// src/integrators/explicit_time_integrator.cc:123-278, which cannot be compiled here) drive the GPU; the tests
// require the results to equal the serial reference's bit for bit (ORDERED assembly).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "nimble_model_data.h"

struct nsm_b200_ctx;

namespace nsm_binding {

class B200ModelData : public nimble::ModelData
{
 public:
  explicit B200ModelData(int device = 0, int assembly = 1 /* NSM_ASSEMBLY_ORDERED */, unsigned flags = 0x2 /* cache b^-1 */);
  ~B200ModelData() override;

  void
  InitializeBlocks(nimble::DataManager& data_manager, const std::shared_ptr<nimble::MaterialFactoryBase>& material_factory_base) override;
  void
  ComputeLumpedMass(nimble::DataManager& data_manager) override;
  void
  ComputeInternalForce(nimble::DataManager& data_manager, double time_previous, double time_current, bool is_output_step,
                       const nimble::Viewify<2>& displacement, nimble::Viewify<2>& force) override;
  void
  UpdateStates(const nimble::DataManager& data_manager) override;

  // element data of the reference's containers <- device records (what WriteExodusOutput / derived data read)
  void
  PullElementData();
  long
  DeviceLaunches() const;

 private:
  void
  check(int status, const char* what) const;
  nsm_b200_ctx* ctx_ = nullptr;
  int           assembly_;
  unsigned      flags_;
};

}  // namespace nsm_binding
