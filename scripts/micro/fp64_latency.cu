// scripts/micro/fp64_latency.cu — dependent-issue latency and per-SMSP issue interval of the B200 FP64 pipe
// (DADD / DMUL / DFMA), measured with clock64() around fully unrolled chains.  Build:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int ILP>
__global__ void chain(double* out, long long* cycles, double a, double b)
{
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (OP == 0) x[i] = x[i] + b;
        if (OP == 1) x[i] = x[i] * b;
        if (OP == 2) x[i] = fma(x[i], b, a);
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP, int ILP>
void run(const char* name, int threads)
{
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * 4096);
  cudaMalloc(&cyc, sizeof(long long));
  chain<OP, ILP><<<1, threads>>>(out, cyc, 1.0, 1.0000001);
  chain<OP, ILP><<<1, threads>>>(out, cyc, 1.0, 1.0000001);
  long long h = 0;
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  const double n = 64.0 * 32.0 * ILP;
  printf("%-5s ILP=%d warps/SM=%2d (per SMSP %d): %.2f cycles per instruction per warp, %.2f cycles per warp-instruction per SMSP\n",
         name, ILP, threads / 32, threads / 128 ? threads / 128 : 1, h / n, h / n / (threads >= 128 ? threads / 128 : 1));
  cudaFree(out);
  cudaFree(cyc);
}

int main()
{
  run<0, 1>("DADD", 32);
  run<1, 1>("DMUL", 32);
  run<2, 1>("DFMA", 32);
  run<0, 2>("DADD", 32);
  run<0, 4>("DADD", 32);
  run<0, 8>("DADD", 32);
  run<2, 8>("DFMA", 32);
  run<0, 1>("DADD", 128);
  run<0, 1>("DADD", 256);
  run<0, 1>("DADD", 512);
  run<0, 1>("DADD", 1024);
  run<0, 2>("DADD", 512);
  run<0, 4>("DADD", 512);
  run<1, 4>("DMUL", 512);
  run<2, 4>("DFMA", 512);
  return 0;
}
