// launch_cost.cu -- host-side cost of one kernel launch on this box: <<<>>> vs cudaLaunchKernelEx without / with the programmatic
// stream-serialisation attribute, for a tiny kernel that waits at entry (griddepcontrol).
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -cudart shared -o bin/launch_cost launch_cost.cu
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p, int v) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0) *p = v;
}
int main() {
  int* d; cudaMalloc(&d, 4);
  const int n = 20000;
  for (int mode = 0; mode < 3; ++mode) {
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; ++i) {
      if (mode == 0) k<<<1, 64>>>(d, i);
      else {
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(1); cfg.blockDim = dim3(64); cfg.stream = 0;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = mode == 2 ? 1 : 0;
        cudaLaunchKernelEx(&cfg, k, d, i);
      }
    }
    auto t1 = std::chrono::steady_clock::now();
    cudaDeviceSynchronize();
    auto t2 = std::chrono::steady_clock::now();
    printf("%-44s host enqueue %.2f us/launch, incl. drain %.2f us/launch\n",
           mode == 0 ? "<<<>>>" : (mode == 1 ? "cudaLaunchKernelEx, no attribute" : "cudaLaunchKernelEx, programmatic serialisation"),
           std::chrono::duration<double, std::micro>(t1 - t0).count() / n, std::chrono::duration<double, std::micro>(t2 - t0).count() / n);
  }
  return 0;
}
