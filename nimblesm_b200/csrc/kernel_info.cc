// nimblesm_b200/csrc/kernel_info.cc — nsm_b200_kernel_info(): what scripts/sass_hot_loop.py read off the SASS of
// the very object this library is linked from (see the Makefile).  The text is JSON:
//   {"source_sha": <sha256/16 of hex8_kernels.cuh + hex8_math.cuh>, "kernels": {"mat<M>_ordered<O>_mode<F>":
//    {"dp": DP warp-instructions per warp pass over 4 elements, "other": ..., "dp_lane_instr_per_element": dp * 8,
//     "hot_instructions", "cold_instructions", "reg", "stack", "mix": {...}}}}
#include "../../include/nsm_b200.h"

extern "C" const char*
nsm_b200_kernel_info(void)
{
  static const char text[] =
#include "kernel_info.inc"
      ;
  return text;
}
