// scripts/micro/issue_mix.cu — does a DP warp-instruction hold a B200 scheduler's ISSUE PORT for 2 cycles, or
// only the FP64 PIPE?  (DESIGN.md §10's "2 cycles per DP + 1 per other instruction" model of the element kernel.)
//
// One loop body of NDP independent DADD/DMUL (8 chains) interleaved evenly with NOTHER independent non-DP
// instructions (8 chains), fully unrolled like the element kernel's pass (1121 DP + 736 other per warp-pass,
// profiles/r01A_ncu_elem_neo_f2_summary.txt), run at the element kernel's occupancy (2 CTAs x 8 warps per SM =
// 4 warps per scheduler).  Reported: scheduler cycles per loop body per warp = elapsed SM cycles / iterations /
// (warps per scheduler).
// (clock64 from the block's start to the finish of its LAST warp; the CUDA-event time of the launch is printed beside it.)
//   issue-port model  : 2*NDP + NOTHER          (DP blocks the port for both cycles)
//   pipe-only model   : max(2*NDP, NDP + NOTHER) (the second cycle of a DP issue is free for another warp's non-DP)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue_mix issue_mix.cu ; run: ./issue_mix
#include <cstdio>
#include <cuda_runtime.h>

enum { kInt = 0, kSel = 1, kLds = 2 };

template <int I>
struct DpOp
{
  static __device__ __forceinline__ void
  run(double (&x)[8], double m, double c)
  {
    if (I & 1)
      asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[I & 7]) : "d"(c));
    else
      asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[I & 7]) : "d"(m));
  }
};

template <int KIND, int I>
struct OtherOp
{
  static __device__ __forceinline__ void
  run(unsigned (&y)[8], float (&z)[8], unsigned k, const unsigned* sm)
  {
    if (KIND == kInt) {
      if (I & 1)
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[I & 7]) : "r"(k), "r"(y[(I + 3) & 7]));
      else
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[I & 7]) : "r"(k), "r"(y[(I + 5) & 7]));
    } else if (KIND == kSel) {  // FSETP + FSEL pairs (the eigen-solver's selects) and LOP3
      if ((I % 3) == 0)
        asm volatile("{ .reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %1, %0, p; }" : "+f"(z[I & 7]) : "f"(z[(I + 3) & 7]));
      else
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[I & 7]) : "r"(k), "r"(y[(I + 3) & 7]));
    } else {  // every 8th non-DP instruction a shared-memory load (the kernel's LDS share: 53 + 37 of 736)
      if ((I & 7) == 0)
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(y[I & 7]) : "r"((unsigned)__cvta_generic_to_shared(sm + ((y[(I + 1) & 7] + threadIdx.x) & 255))));
      else
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[I & 7]) : "r"(k), "r"(y[(I + 3) & 7]));
    }
  }
};

// Bresenham interleave of NDP and NOTHER instructions over positions [LO, HI); split in halves so that the
// template recursion stays logarithmic
template <int NDP, int NOTHER, int KIND, int LO, int HI>
struct Body
{
  static __device__ __forceinline__ void
  run(double (&x)[8], unsigned (&y)[8], float (&z)[8], double m, double c, unsigned k, const unsigned* sm)
  {
    if (HI - LO == 1) {
      constexpr int       N = NDP + NOTHER;
      constexpr long long a = (long long)LO * NDP / N, b = (long long)(LO + 1) * NDP / N;
      if (b > a)
        DpOp<(int)a>::run(x, m, c);
      else
        OtherOp<KIND, LO - (int)a>::run(y, z, k, sm);
    } else {
      constexpr int MID = HI - LO > 1 ? (LO + HI) / 2 : HI;
      Body<NDP, NOTHER, KIND, LO, MID>::run(x, y, z, m, c, k, sm);
      Body<NDP, NOTHER, KIND, MID, HI>::run(x, y, z, m, c, k, sm);
    }
  }
};
template <int NDP, int NOTHER, int KIND, int LO>
struct Body<NDP, NOTHER, KIND, LO, LO>
{
  static __device__ __forceinline__ void
  run(double (&)[8], unsigned (&)[8], float (&)[8], double, double, unsigned, const unsigned*)
  {
  }
};

template <int NDP, int NOTHER, int KIND>
__global__ void __launch_bounds__(256, 2)
mix_kernel(double* out, long long* cycles, int iters, double m, double c, unsigned k)
{
  __shared__ unsigned           sm[256];
  __shared__ unsigned long long t_end;  // the LAST warp's finish time: warps of one scheduler do not advance in step
  sm[threadIdx.x] = threadIdx.x * k;
  if (threadIdx.x == 0) t_end = 0ull;
  __syncthreads();
  double   x[8];
  unsigned y[8];
  float    z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 1.0 + i + threadIdx.x, y[i] = threadIdx.x * 7 + i, z[i] = (float)(i + threadIdx.x);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) Body<NDP, NOTHER, KIND, 0, NDP + NOTHER>::run(x, y, z, m, c, k, sm);
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) atomicMax(&t_end, (unsigned long long)t1);
  __syncthreads();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + (double)y[i] + (double)z[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = (long long)t_end - t0;
}

template <int NDP, int NOTHER, int KIND>
void
run(const char* what, int sms)
{
  const int  blocks = sms * 2, iters = 200;
  double*    out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * blocks * 256);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  float best = 1e30f;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    mix_kernel<NDP, NOTHER, KIND><<<blocks, 256>>>(out, cyc, iters, 0.9999999, 1e-9, 0x9e3779b9u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  long long h[1024];
  cudaMemcpy(h, cyc, sizeof(long long) * (blocks < 1024 ? blocks : 1024), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks && i < 1024; ++i) avg += (double)h[i];
  avg /= blocks < 1024 ? blocks : 1024;
  const double per_body = avg / iters / 4.0;  // 16 warps per SM = 4 per scheduler
  const double port = 2.0 * NDP + NOTHER, pipe = (2.0 * NDP > NDP + NOTHER) ? 2.0 * NDP : NDP + NOTHER;
  printf("%-28s DP %4d other %4d : %8.1f scheduler cycles per body per warp | issue-port model %6.0f (x%.3f)  pipe-only model %6.0f (x%.3f) | %.3f ms\n",
         what, NDP, NOTHER, per_body, port, per_body / port, pipe, per_body / pipe, best);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
  cudaFree(out), cudaFree(cyc);
}

int
main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs, 2 CTAs x 8 warps per SM\n", p.name, p.multiProcessorCount);
  const int s = p.multiProcessorCount;
  run<1121, 0, kInt>("DP only", s);
  run<0, 736, kInt>("INT only (LOP3/IMAD)", s);
  run<1121, 736, kInt>("neohookean mix, LOP3/IMAD", s);
  run<1121, 736, kSel>("neohookean mix, FSETP/SEL", s);
  run<1121, 736, kLds>("neohookean mix, +LDS", s);
  run<1121, 368, kInt>("half the non-DP", s);
  run<1121, 1121, kInt>("1:1", s);
  run<539, 355, kInt>("elastic mix (539:355)", s);
  return 0;
}
