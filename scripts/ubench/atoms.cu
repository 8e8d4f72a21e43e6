// Micro-benchmark: shared-memory atomic throughput on sm_100a, the way the step kernel uses it (one atomicAdd with
// return value per particle to a pseudo-random counter out of 2048..8192; 2 CTAs x 512 threads per SM).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o atoms atoms.cu ; prints cycles per warp instruction.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512, 2) k(unsigned *out, int iters, int nctr, long long *cyc) {
  extern __shared__ unsigned s[];
  for (int i = threadIdx.x; i < nctr; i += blockDim.x) s[i] = 0;
  __syncthreads();
  unsigned h = threadIdx.x * 2654435761u + blockIdx.x * 40503u, acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    h = h * 1664525u + 1013904223u;
    unsigned a = (h >> 8) % (unsigned)nctr;
    if (MODE == 0) acc += atomicAdd(&s[a], 1u);              // ATOMS with return, spread addresses
    else if (MODE == 1) atomicAdd(&s[a], 1u);                // RED-like (no return)
    else if (MODE == 2) acc += s[a];                         // plain LDS
    else if (MODE == 3) { s[a] = h; }                        // plain STS
    else if (MODE == 4) acc += atomicAdd(&s[(threadIdx.x & 31) + 32 * (a & 63)], 1u);  // conflict-free banks
    else if (MODE == 5) acc += __match_any_sync(0xffffffffu, a & 1023u);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  if (acc == 0x12345678u) out[0] = acc + s[0];
}
int main() {
  unsigned *out; long long *cyc, h;
  cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  const char *names[] = {"ATOMS ret spread", "atomic noret spread", "LDS spread", "STS spread", "ATOMS ret bank=lane", "match_any"};
  for (int nctr : {2048, 8192}) {
    for (int mode = 0; mode < 6; mode++) {
      size_t sm = 48 * 1024;
      auto launch = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        kern<<<296, 512, sm>>>(out, 100, nctr, cyc);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        kern<<<296, 512, sm>>>(out, iters, nctr, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        // 32 warps per SM each issue `iters` warp instructions
        printf("ctr=%d %-22s %.2f ms  %.1f cycles per warp-instruction per SM (%.2f per lane)\n", nctr, names[mode], ms,
               (double)h / (32.0 * iters), (double)h / (32.0 * iters) / 32.0);
      };
      switch (mode) {
        case 0: launch(k<0>); break; case 1: launch(k<1>); break; case 2: launch(k<2>); break;
        case 3: launch(k<3>); break; case 4: launch(k<4>); break; case 5: launch(k<5>); break;
      }
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
