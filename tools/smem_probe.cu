// smem_probe.cu — how many cycles does a warp-wide shared-memory access with RANDOM per-lane addresses cost on
// this part, per access width?  Answers the layout question of the hot table (SoA 4/8-byte words vs one 16-byte
// AoS cell): run with 1 warp (latency) and 16 warps per SM (throughput), all SMs.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smem_probe tools/smem_probe.cu && ./smem_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int CELLS = 1024;     // random cell index range (16-byte cells -> 16 KB per warp region)
constexpr int ITERS = 4096;

template <int W, bool STORE, bool SPREAD>
__global__ void probe(unsigned long long* out, int warps_per_cta) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* base = smem + (size_t)warp * CELLS * 16;
  for (int i = lane; i < CELLS * 4; i += 32) ((uint32_t*)base)[i] = i;
  __syncwarp();
  uint32_t s = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x * 7919u;
  uint32_t acc = 0;
  const long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < ITERS; ++it) {
    s = s * 1664525u + 1013904223u;
    // SPREAD: conflict-free by construction (lane-owned column), else a random cell
    const uint32_t cell = SPREAD ? (((s >> 10) & (CELLS / 32 - 1)) * 32 + lane) : ((s >> 10) & (CELLS - 1));
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(base) + cell * W;  // an array of W-byte elements
    if (STORE) {
      if (W == 4) asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(s));
      else if (W == 8) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(addr), "r"(s), "r"(acc));
      else asm volatile("st.shared.v4.u32 [%0], {%1,%2,%1,%2};" ::"r"(addr), "r"(s), "r"(acc));
    } else {
      uint32_t a, b, c, d;
      if (W == 4) { asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(addr)); acc += a; }
      else if (W == 8) { asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr)); acc += a ^ b; }
      else { asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr)); acc += a ^ b ^ c ^ d; }
    }
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) out[1] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
}

template <int W, bool STORE, bool SPREAD>
void run(const char* name, int warps) {
  unsigned long long* d;
  cudaMalloc(&d, 16);
  const size_t smem = (size_t)warps * CELLS * 16;
  cudaFuncSetAttribute(probe<W, STORE, SPREAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<W, STORE, SPREAD><<<148, warps * 32, smem>>>(d, warps);
  probe<W, STORE, SPREAD><<<148, warps * 32, smem>>>(d, warps);
  unsigned long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double per_instr = (double)h[0] / ITERS;
  // with `warps` warps sharing the SM's pipe, cycles per warp instruction of pipe time = per_instr / warps
  printf("%-28s warps/SM %2d  cycles per iteration (one warp's view) %7.2f  -> pipe cycles per warp access %6.2f\n", name, warps, per_instr,
         per_instr / warps);
  cudaFree(d);
}

int main() {
  for (int warps : {1, 12}) {
    run<4, false, false>("LDS.32  random cells", warps);
    run<8, false, false>("LDS.64  random cells", warps);
    run<16, false, false>("LDS.128 random cells", warps);
    run<4, true, false>("STS.32  random cells", warps);
    run<8, true, false>("STS.64  random cells", warps);
    run<16, true, false>("STS.128 random cells", warps);
    run<4, false, true>("LDS.32  lane-owned banks", warps);
    run<8, false, true>("LDS.64  lane-owned banks", warps);
    run<16, false, true>("LDS.128 lane-owned banks", warps);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
