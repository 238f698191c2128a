// Instruction-fetch microbenchmark (sm_100a): IPC of straight-line code of a given footprint, as a function of the
// number of warps per SM sub-partition and of whether those warps run the code in step or spread over it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ifetch ifetch.cu && ./ifetch
#include <cstdio>
#include <cuda_runtime.h>

#define R8(x) x x x x x x x x
#define R64(x) R8(R8(x))
// 8 independent FFMA chains: a single warp could issue one per cycle if instructions were always available
#define STEP "fma.rn.f32 %0, %0, %8, %9; fma.rn.f32 %1, %1, %8, %9; fma.rn.f32 %2, %2, %8, %9; fma.rn.f32 %3, %3, %8, %9;" \
             "fma.rn.f32 %4, %4, %8, %9; fma.rn.f32 %5, %5, %8, %9; fma.rn.f32 %6, %6, %8, %9; fma.rn.f32 %7, %7, %8, %9;\n"

template <int BLOCKS64>   // body = BLOCKS64 * 64 * 8 FFMA = BLOCKS64 * 8 KB of SASS
__global__ void body_kernel(float* out, int iters, int spread_cycles, float b, float c) {
  float a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  int warp = threadIdx.x >> 5;
  if (spread_cycles > 0) {   // desynchronise the warps of a sub-partition (warps w, w+4, w+8 ... share one)
    long long t0 = clock64();
    while (clock64() - t0 < (long long)(warp >> 2) * spread_cycles) {}
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < BLOCKS64; ++k)
      asm volatile(R64(STEP) : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(b), "f"(c));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int B>
void run(float* out, int warps_per_smsp, int spread) {
  int block = warps_per_smsp * 4 * 32;
  long long inst_per_warp_iter = (long long)B * 64 * 8;
  int iters = (int)(40000000LL / inst_per_warp_iter / warps_per_smsp) + 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  body_kernel<B><<<148, block>>>(out, 2, spread, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  body_kernel<B><<<148, block>>>(out, iters, spread, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  double inst_per_smsp = (double)inst_per_warp_iter * iters * warps_per_smsp;
  double cycles = ms * 1e-3 * 1.965e9;   // assumes the max SM clock
  printf("body %4d KB  warps/SMSP %2d  spread %6d  ms %8.3f  IPC/SMSP %.3f  (cycles per 128B line per SMSP %.1f)\n", B * 8,
         warps_per_smsp, spread, ms, inst_per_smsp / cycles, cycles / (inst_per_smsp / 8.0));
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  int wl[] = {1, 2, 4, 8};
  for (int w : wl) run<1>(out, w, 0);         // 8 KB?  (1 * 64 * 8 * 16 B = 8 KB)
  for (int w : wl) run<2>(out, w, 0);         // 16 KB
  for (int w : wl) run<3>(out, w, 0);         // 24 KB
  for (int w : wl) run<8>(out, w, 0);         // 64 KB
  for (int w : wl) run<16>(out, w, 0);        // 128 KB
  // same footprints, warps of a sub-partition started 3000 cycles apart (spread over the body)
  for (int w : {2, 4, 8}) run<2>(out, w, 3000);
  for (int w : {2, 4, 8}) run<3>(out, w, 3000);
  for (int w : {2, 4, 8}) run<8>(out, w, 3000);
  return 0;
}
