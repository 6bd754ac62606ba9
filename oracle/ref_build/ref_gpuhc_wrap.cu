// C-ABI driver for the UNMODIFIED reference GPU-HC++ launch wrappers
//   kernel_GPUHC_trifocal_2op1p_30x30_PH_CodeOpt_TrunPaths            (/root/reference/magmaHC/gpu-kernels/…_TrunPaths.cu:292-386)
//   kernel_GPUHC_trifocal_2op1p_30x30_PH_CodeOpt_TrunPaths_TrunRANSAC (…_TrunRANSAC.cu:329-453)
// compiled for sm_100a by oracle/Makefile into oracle/_ref/libref_gpuhc.so.  It reproduces what GPU_HC_Solver does
// around them: pointer arrays (GPU_HC_Solver.cpp:352-353), the L2 persisting window over the index table
// (:117,364-378) and pre-loading the tracks with the start solutions (:208,342).
// BASELINE / second-oracle infrastructure for the GPU box only; the product never loads it.
#include <cstdio>
#include <cuda_runtime.h>
#include "magma_v2.h"
#include "magmaHC-kernels.hpp"

namespace {
__global__ void set_pointers(magmaFloatComplex** arr, magmaFloatComplex* base, int stride, int n)
{ int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) arr[i] = base + (size_t)i * stride; }
__global__ void preload_tracks(magmaFloatComplex* tracks, const magmaFloatComplex* start, int n_paths)
{
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)n_paths * 31) tracks[i] = start[i % (312 * 31)];
}
magmaFloatComplex** g_start_arr = nullptr;
magmaFloatComplex** g_track_arr = nullptr;
int g_track_cap = 0;
}

extern "C" {

// Prepares pointer arrays + tracks (untimed part of the reference flow).  All pointers are device pointers.
int ref_gpuhc_prepare(void* stream, int n_hyp, float* d_start_sols, float* d_tracks, int* d_unified_index, int index_ints)
{
  cudaStream_t s = (cudaStream_t)stream;
  const int n_paths = n_hyp * 312;
  if (!g_start_arr && cudaMalloc(&g_start_arr, 312 * sizeof(void*)) != cudaSuccess) return 1;
  if (n_paths > g_track_cap) {
    if (g_track_arr) cudaFree(g_track_arr);
    if (cudaMalloc(&g_track_arr, (size_t)(n_paths + n_hyp) * sizeof(void*)) != cudaSuccess) return 2;
    g_track_cap = n_paths;
  }
  set_pointers<<<(312 + 127) / 128, 128, 0, s>>>(g_start_arr, (magmaFloatComplex*)d_start_sols, 31, 312);
  set_pointers<<<(n_paths + 127) / 128, 128, 0, s>>>(g_track_arr, (magmaFloatComplex*)d_tracks, 31, n_paths);
  preload_tracks<<<(int)(((size_t)n_paths * 31 + 255) / 256), 256, 0, s>>>((magmaFloatComplex*)d_tracks, (const magmaFloatComplex*)d_start_sols, n_paths);
  // L2 persisting window over the index table, as GPU_HC_Solver.cpp:117,364-378
  const size_t bytes = (size_t)index_ints * sizeof(int);
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes + (1 << 20));
  cudaStreamAttrValue attr;
  attr.accessPolicyWindow.base_ptr = (void*)d_unified_index;
  attr.accessPolicyWindow.num_bytes = bytes;
  attr.accessPolicyWindow.hitRatio = 1.0f;
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr);
  return (int)cudaGetLastError();
}

// Re-load the tracks with the start solutions (needed before every launch: the kernel tracks in place).
int ref_gpuhc_reload_tracks(void* stream, int n_hyp, float* d_start_sols, float* d_tracks)
{
  const int n_paths = n_hyp * 312;
  preload_tracks<<<(int)(((size_t)n_paths * 31 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((magmaFloatComplex*)d_tracks, (const magmaFloatComplex*)d_start_sols, n_paths);
  return (int)cudaGetLastError();
}

int ref_gpuhc_track(void* stream, int n_hyp, int max_steps, int max_corr, int dt_inc,
                    float* d_start_params, float* d_target, float* d_diff, int* d_unified_index,
                    bool* d_conv, bool* d_inf, float* d_debug)
{
  magma_queue q; q.s = (cudaStream_t)stream;
  kernel_GPUHC_trifocal_2op1p_30x30_PH_CodeOpt_TrunPaths(&q, n_hyp, max_steps, max_corr, dt_inc, g_start_arr, g_track_arr,
      (magmaFloatComplex*)d_start_params, (magmaFloatComplex*)d_target, (magmaFloatComplex*)d_diff, d_unified_index,
      d_conv, d_inf, (magmaFloatComplex*)d_debug);
  return (int)cudaGetLastError();
}

int ref_gpuhc_track_abort(void* stream, int n_hyp, int n_edgels, int max_steps, int max_corr, int dt_inc,
                          float* d_start_params, float* d_target, float* d_diff, int* d_unified_index,
                          float* d_edgels, float* d_K, bool* d_conv, bool* d_inf, float* d_debug,
                          bool* d_found, int* d_found_index)
{
  magma_queue q; q.s = (cudaStream_t)stream;
  kernel_GPUHC_trifocal_2op1p_30x30_PH_CodeOpt_TrunPaths_TrunRANSAC(&q, n_hyp, n_edgels, max_steps, max_corr, dt_inc,
      g_start_arr, g_track_arr, (magmaFloatComplex*)d_start_params, (magmaFloatComplex*)d_target, (magmaFloatComplex*)d_diff,
      d_unified_index, d_edgels, d_K, d_conv, d_inf, (magmaFloatComplex*)d_debug, d_found, d_found_index);
  return (int)cudaGetLastError();
}

}  // extern "C"
