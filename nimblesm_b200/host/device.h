// nimblesm_b200/host/device.h — RAII owner of one nsm_b200_ctx (include/nsm_b200.h) and the error convention
// of the host layer: a non-zero ABI status becomes std::runtime_error carrying nsm_b200_last_error(), the
// analogue of NIMBLE_ABORT's message (src/nimble_macros.h:49-72).  No CPU fallback: when the library reports
// "no CUDA device" every host entry point that needs arithmetic fails with that message.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>

#include "../../include/nsm_b200.h"

namespace nimble_b200 {

class DeviceContext
{
 public:
  explicit DeviceContext(int device = 0)
  {
    const int rc = nsm_b200_create(device, &ctx_);
    if (rc != NSM_OK) throw std::runtime_error(std::string("nsm_b200_create: ") + nsm_b200_last_error(nullptr));
  }
  ~DeviceContext()
  {
    nsm_b200_destroy(ctx_);
  }
  DeviceContext(const DeviceContext&) = delete;
  DeviceContext&
  operator=(const DeviceContext&) = delete;
  nsm_b200_ctx*
  get() const
  {
    return ctx_;
  }
  // throws when an ABI call failed
  void
  check(int rc, const char* what) const
  {
    if (rc != NSM_OK) throw std::runtime_error(std::string(what) + ": " + nsm_b200_last_error(ctx_));
  }

 private:
  nsm_b200_ctx* ctx_ = nullptr;
};

}  // namespace nimble_b200
