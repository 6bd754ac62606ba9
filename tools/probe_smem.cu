// Shared-memory pipe micro-benchmarks (B200): cycles per warp-instruction at saturation (24 warps per SM) for the access shapes
// the elimination uses.  Results quoted in profiles/probes_r2.txt.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

enum { STS128_1LANE, STS128_4Q, STS128_8SAMEQ, STS64_4Q, LDS128_BCAST, LDS64_BCAST, LDS128_8ADDR, LDS128_16ADDR_2H, SHFL_IDX, LDS128_2ADDR, STS128_2H, NMODES };
const char* names[NMODES] = {"STS.128 one lane", "STS.128 8 lanes in 4 quarter-warps", "STS.128 8 lanes in ONE quarter-warp", "STS.64 8 lanes in 4 quarters",
  "LDS.128 broadcast (1 address)", "LDS.64 broadcast (1 address)", "LDS.128 8 addresses (6-lane segments)", "LDS.128 16 addresses (3-lane groups, 2 halves)",
  "SHFL.IDX (segment source)", "LDS.128 2 addresses (one per half-warp)", "STS.128 2 lanes, one per half-warp"};

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters)
{
  extern __shared__ float4 sm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float4* base = sm + wid * 256;      // 4 KB per warp
  for (int i = lane; i < 256; i += 32) base[i] = make_float4(i, 1, 2, 3);
  __syncwarp();
  float4 acc = make_float4(0, 0, 0, 0);
  uint32_t a = (uint32_t)__cvta_generic_to_shared(base);
  bool act = false; uint32_t off = 0;
  if (MODE == STS128_1LANE) { act = lane == 5; }
  if (MODE == STS128_4Q || MODE == STS64_4Q) { act = (lane % 4) == 1; off = (lane / 4) * 144; }       // lanes 1,5,9,...,29: two per quarter
  if (MODE == STS128_8SAMEQ) { act = lane < 8; off = lane * 144; }
  if (MODE == STS128_2H) { act = (lane == 3) || (lane == 19); off = (lane / 16) * 144; }
  if (MODE == LDS128_8ADDR) off = (lane / 4) * 144;
  if (MODE == LDS128_16ADDR_2H) off = (lane / 2) * 144;
  if (MODE == LDS128_2ADDR) off = (lane / 16) * 144;
  const int src = (lane / 6) * 6 + 2;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const uint32_t ad = a + off + r * 16;
      if (MODE == STS128_1LANE || MODE == STS128_4Q || MODE == STS128_8SAMEQ || MODE == STS128_2H)
        asm volatile("{ .reg .pred q; setp.ne.u32 q, %5, 0; @q st.shared.v4.f32 [%0], {%1,%2,%3,%4}; }" :: "r"(ad), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w), "r"((uint32_t)act) : "memory");
      else if (MODE == STS64_4Q)
        asm volatile("{ .reg .pred q; setp.ne.u32 q, %3, 0; @q st.shared.v2.f32 [%0], {%1,%2}; }" :: "r"(ad), "f"(acc.x), "f"(acc.y), "r"((uint32_t)act) : "memory");
      else if (MODE == LDS64_BCAST) {
        float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(ad) : "memory"); acc.x += v.x; acc.y += v.y;
      } else if (MODE == SHFL_IDX) {
        acc.x += __shfl_sync(0xffffffffu, acc.y + r, src);
      } else {
        float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ad) : "memory");
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

template <int MODE> float run(float* d, int iters)
{
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4096);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0); probe<MODE><<<148 * 3, 256, 8 * 4096>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main()
{
  float* d; cudaMalloc(&d, 1024);
  const int iters = 20000;
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float ms[NMODES];
  ms[0] = run<0>(d, iters); ms[1] = run<1>(d, iters); ms[2] = run<2>(d, iters); ms[3] = run<3>(d, iters); ms[4] = run<4>(d, iters); ms[5] = run<5>(d, iters);
  ms[6] = run<6>(d, iters); ms[7] = run<7>(d, iters); ms[8] = run<8>(d, iters); ms[9] = run<9>(d, iters); ms[10] = run<10>(d, iters);
  for (int m = 0; m < NMODES; m++) {
    // per SM: 24 warps x iters x 8 instructions; cycles = ms * clk(kHz)
    const double cyc = (double)ms[m] * clk / (24.0 * iters * 8.0);
    printf("%-48s %8.3f ms  %.2f SM-cycles per warp-instruction\n", names[m], ms[m], cyc);
  }
  printf("last error: %s (clock %d kHz)\n", cudaGetErrorString(cudaGetLastError()), clk);
  return 0;
}
