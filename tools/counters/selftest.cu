// Self-test of libom_counters.so with a trivial kernel: prints smsp__inst_executed.sum of 1 CTA x 32 threads x a known loop.
//   nvcc -arch=sm_100a selftest.cu -o selftest -L. -lom_counters -Wl,-rpath,'$ORIGIN' && ./selftest
#include <cstdio>
#include <cuda_runtime.h>
extern "C" {
int omc_setup(int, const char* const*, int);
int omc_num_passes();
int omc_pass_begin();
int omc_pass_end(int*);
int omc_finish(double*, int*);
const char* omc_last_error();
}
__global__ void spin(float* out, int n) {
  float x = threadIdx.x;
  for (int i = 0; i < n; ++i) x = x * 1.0001f + 0.5f;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
int main() {
  float* d;
  cudaMalloc(&d, 148 * 128 * sizeof(float));
  spin<<<148, 128>>>(d, 1000);
  cudaDeviceSynchronize();
  const char* metrics[] = {"smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"};
  if (omc_setup(0, metrics, 3)) { printf("setup failed: %s\n", omc_last_error()); return 1; }
  printf("passes: %d\n", omc_num_passes());
  int done = 0, guard = 0;
  while (!done && guard++ < 16) {
    if (omc_pass_begin()) { printf("begin failed: %s\n", omc_last_error()); return 1; }
    spin<<<148, 128>>>(d, 1000);
    cudaDeviceSynchronize();
    if (omc_pass_end(&done)) { printf("end failed: %s\n", omc_last_error()); return 1; }
  }
  double v[3] = {0, 0, 0};
  int nr = 0;
  if (omc_finish(v, &nr)) { printf("finish failed: %s\n", omc_last_error()); return 1; }
  printf("ranges %d  inst_executed %.0f (expect ~ 148*4 warps * ~3000+)  dram read %.0f write %.0f\n", nr, v[0], v[1], v[2]);
  return 0;
}
