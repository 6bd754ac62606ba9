// Hardware probes used to pin two facts the design relies on (run on the GPU box; results are quoted in profiles/probes_r2.txt):
//  1. what `__shfl_down_sync(__activemask(), v, off)` returns for lanes whose source lane does not exist in a 30-thread block — the
//     reference sums its convergence norms that way (kernel_GPUHC_..._TrunPaths.cu:236-239);
//  2. the issue rate of packed FP32 (FFMA2) against scalar FFMA on this GPU.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o probes tools/probes.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__global__ void dirty_regs(float* out)      // leaves non-zero values in the registers of all 32 lanes of many warps
{
  float a[24];
  for (int i = 0; i < 24; i++) a[i] = 1000.0f + threadIdx.x * 3.0f + i;
  for (int k = 0; k < 64; k++)
    for (int i = 0; i < 24; i++) a[i] = a[i] * 1.0001f + a[(i + 1) % 24];
  float s = 0; for (int i = 0; i < 24; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void shfl_probe(unsigned long long* tally, int* src, float* fsum)
{
  const int tx = threadIdx.x;
  // (a) which lane does each shuffle read?  value = lane id + 100
  for (int k = 0, off = 16; off > 0; off >>= 1, k++)
    src[(blockIdx.x * 5 + k) * 32 + tx] = __shfl_down_sync(__activemask(), tx + 100, off);
  // (b) the reference's reduction with a 2-bit tally per contributing lane
  unsigned long long v = 1ull << (2 * tx);
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(__activemask(), v, off);
  tally[blockIdx.x * 32 + tx] = v;
  // (c) the same in float with the value pattern of the reference (r_sqrt_sols += shfl_down)
  float f = 1.0f + tx;
  for (int off = 16; off > 0; off >>= 1) f += __shfl_down_sync(__activemask(), f, off);
  fsum[blockIdx.x * 32 + tx] = f;
}

template <int PACKED>
__global__ void __launch_bounds__(256) fma_rate(float* out, int iters)
{
  unsigned long long a[8];
  float s[16];
  for (int i = 0; i < 16; i++) s[i] = (threadIdx.x + i) * 1e-3f;
  for (int i = 0; i < 8; i++) asm("mov.b64 %0, {%1, %2};" : "=l"(a[i]) : "f"(s[2 * i]), "f"(s[2 * i + 1]));
  unsigned long long b, c;
  asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(0.999f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(1e-3f));
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if (PACKED) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
      } else {
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(0.999f), "f"(1e-3f));
      }
    }
  }
  float t = 0;
  for (int i = 0; i < 8; i++) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); t += x + y; }
  for (int i = 0; i < 16; i++) t += s[i];
  if (t == 123456.789f) out[0] = t;
}

int main()
{
  float* d; cudaMalloc(&d, 1 << 24);
  unsigned long long* dt; int* ds; float* df;
  const int NB = 64;
  cudaMalloc(&dt, NB * 32 * 8); cudaMalloc(&ds, NB * 5 * 32 * 4); cudaMalloc(&df, NB * 32 * 4);
  for (int rep = 0; rep < 3; rep++) {
    dirty_regs<<<148 * 8, 256>>>(d);
    cudaMemset(dt, 0, NB * 32 * 8); cudaMemset(ds, 0xff, NB * 5 * 32 * 4);
    shfl_probe<<<NB, 30>>>(dt, ds, df);
    cudaDeviceSynchronize();
    static unsigned long long ht[NB * 32]; static int hs[NB * 5 * 32]; static float hf[NB * 32];
    cudaMemcpy(ht, dt, sizeof ht, cudaMemcpyDeviceToHost); cudaMemcpy(hs, ds, sizeof hs, cudaMemcpyDeviceToHost);
    cudaMemcpy(hf, df, sizeof hf, cudaMemcpyDeviceToHost);
    // summarise over blocks: are all blocks identical?
    int same = 1;
    for (int b = 1; b < NB; b++) {
      if (ht[b * 32] != ht[0] || hf[b * 32] != hf[0]) same = 0;
      for (int k = 0; k < 5 * 32; k++) if (hs[b * 5 * 32 + k] != hs[k]) same = 0;
    }
    printf("rep %d: all %d blocks identical: %s\n", rep, NB, same ? "yes" : "NO");
    for (int k = 0, off = 16; k < 5; k++, off >>= 1) {
      printf("  shfl_down off %2d, value read by lanes 0..29 (100 + source lane):", off);
      for (int l = 0; l < 30; l++) printf(" %d", hs[k * 32 + l]);
      printf("\n");
    }
    printf("  lane-0 tally (how often each lane's value is in the sum): ");
    for (int l = 0; l < 32; l++) printf("%llu", (ht[0] >> (2 * l)) & 3ull);
    printf("\n  lane-0 float sum of (1 + lane): %.1f   (sum over 30 lanes = 465)\n", hf[0]);
    if (!same) {
      for (int b = 0; b < 8; b++) { printf("   block %d tally ", b); for (int l = 0; l < 32; l++) printf("%llu", (ht[b * 32] >> (2 * l)) & 3ull); printf(" fsum %.1f\n", hf[b * 32]); }
    }
  }
  // FMA rates
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, grid = 148 * 8, block = 256;
  for (int packed = 0; packed < 2; packed++) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0);
      if (packed) fma_rate<1><<<grid, block>>>(d, iters); else fma_rate<0><<<grid, block>>>(d, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double flops = 2.0 * 64.0 * iters * (double)grid * block;
    const double instr = (packed ? 32.0 : 64.0) * iters * (double)grid * block / 32.0;
    printf("%s: %.3f ms  %.1f TFLOP/s  %.3f warp-instructions/ns\n", packed ? "FFMA2 (fma.rn.f32x2)" : "FFMA  (fma.rn.f32)  ", best, flops / best / 1e9, instr / best / 1e6);
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
