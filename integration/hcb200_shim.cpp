// hcb200_shim.cpp — the reference's kernel entry points on top of include/hcb200.h.
//
// Add this ONE file to the reference tree (e.g. magmaHC/gpu-kernels/), remove the four kernel_GPUHC_*.cu files from
// magmaHC/CMakeLists.txt and link -lhcb200: the reference's GPU_HC_Solver.cpp (magmaHC/GPU_HC_Solver.cpp:395-433) then runs on
// this library without a source change.  It is compiled against the reference's own MAGMA headers (magma_queue_t is a MAGMA
// type), which is why it is shipped as source and is not part of libhcb200.so.  Signatures: magmaHC/gpu-kernels/magmaHC-kernels.hpp:24-105.
//
// oracle/Makefile (target ref_dropin) builds exactly that configuration in the build container — the UNMODIFIED reference
// GPU_HC_Solver.cpp / Data_Reader.cpp / Evaluations.cpp + this file + libhcb200.so — and tests/test_gpu_full.py runs it on the B200.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "magma_v2.h"
#include "magmaHC-kernels.hpp"
#include "hcb200.h"

namespace {

// one workspace per device (the launches zero it themselves), grown when a larger round arrives; sized for HCB200_FLAG_SPLIT_LONG_PATHS
void* workspace(int n_hyp = 0)
{
  static void* ws[64] = {nullptr};
  static size_t bytes[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return nullptr;
  const size_t need = n_hyp <= HCB200_SPLIT_MAX_HYPOTHESES ? hcb200_workspace_bytes_for(n_hyp) : hcb200_workspace_bytes();
  if (!ws[dev] || bytes[dev] < need) {
    if (ws[dev]) { cudaDeviceSynchronize(); cudaFree(ws[dev]); }
    bytes[dev] = need;
    if (cudaMalloc(&ws[dev], need) != cudaSuccess) { ws[dev] = nullptr; bytes[dev] = 0; }
  }
  return ws[dev];
}

// d_startSols_array[0] / d_Track_array[0] are the bases of the contiguous buffers the pointer arrays were built from
// (magma_cset_pointer, GPU_HC_Solver.cpp:352-353).  The base of a pointer array is read back ONCE — the reference builds each array a
// single time from a buffer it never re-allocates (GPU_HC_Solver.cpp:137-184, 352-353) — and cached, so that later launches only
// enqueue work and return, as the reference's wrappers do.  A host layer that re-points an array must call hcb200_shim_forget().
struct BaseCache { const void* array; void* base; };
BaseCache g_bases[32] = {};
template <class T>
T* first_entry(T** d_pointer_array, cudaStream_t s)
{
  for (const BaseCache& c : g_bases)
    if (c.array == (const void*)d_pointer_array && c.base) return (T*)c.base;
  T* p = nullptr;
  cudaMemcpyAsync(&p, d_pointer_array, sizeof p, cudaMemcpyDeviceToHost, s);      // on the launch stream: ordered after the kernel that fills the array
  cudaStreamSynchronize(s);
  for (BaseCache& c : g_bases)
    if (!c.array) { c.array = (const void*)d_pointer_array; c.base = (void*)p; break; }
  return p;
}

void report(const char* who, int rc)
{
  if (rc != 0) printf("%s: %s\n", who, hcb200_error_string(rc));      // the reference only prints launch failures (…TrunPaths.cu:383)
}

real_Double_t track(magma_queue_t q, int n_hyp, int max_steps, int max_corr, int dt_inc, magmaFloatComplex** d_startSols_array,
                    magmaFloatComplex** d_Track_array, magmaFloatComplex* d_startParams, magmaFloatComplex* d_targetParams,
                    magmaFloatComplex* d_diffParams, bool* d_conv, bool* d_inf)
{
  cudaStream_t s = q->cuda_stream();
  report("hcb200_track",
         hcb200_track(s, n_hyp, max_steps, max_corr, dt_inc,
                      HCB200_FLAG_PRUNE_PATHS | (n_hyp <= HCB200_SPLIT_MAX_HYPOTHESES ? HCB200_FLAG_SPLIT_LONG_PATHS : 0u),
                      (const float*)first_entry(d_startSols_array, s), (const float*)d_startParams, (const float*)d_targetParams,
                      (const float*)d_diffParams, (float*)first_entry(d_Track_array, s), (uint8_t*)d_conv, (uint8_t*)d_inf,
                      nullptr, workspace(n_hyp)));
  return 0.0;
}

real_Double_t track_abort(magma_queue_t q, int n_hyp, int n_edgels, int max_steps, int max_corr, int dt_inc,
                          magmaFloatComplex** d_startSols_array, magmaFloatComplex** d_Track_array, magmaFloatComplex* d_startParams,
                          magmaFloatComplex* d_targetParams, magmaFloatComplex* d_diffParams, float* d_edgels, float* d_K,
                          bool* d_conv, bool* d_inf, bool* d_found, int* d_found_index)
{
  cudaStream_t s = q->cuda_stream();
  report("hcb200_track_abort",
         hcb200_track_abort(s, n_hyp, n_edgels, max_steps, max_corr, dt_inc, HCB200_FLAG_PRUNE_PATHS,
                            (const float*)first_entry(d_startSols_array, s), (const float*)d_startParams,
                            (const float*)d_targetParams, (const float*)d_diffParams, d_edgels, d_K,
                            (float*)first_entry(d_Track_array, s), (uint8_t*)d_conv, (uint8_t*)d_inf, (uint8_t*)d_found,
                            d_found_index, nullptr, nullptr, workspace()));
  return 0.0;
}

}  // namespace

// `bool` is one byte on every CUDA host ABI, so the flag arrays are handed over as uint8_t*.  The index-table arguments are
// accepted and ignored: the polynomial system is compiled into the library.

real_Double_t kernel_GPUHC_trifocal_2op1p_30x30_PH_CodeOpt_TrunPaths(
    magma_queue_t my_queue, int sub_RANSAC_iters, int HC_max_steps, int HC_max_correction_steps, int HC_delta_t_incremental_steps,
    magmaFloatComplex** d_startSols_array, magmaFloatComplex** d_Track_array, magmaFloatComplex* d_startParams,
    magmaFloatComplex* d_targetParams, magmaFloatComplex* d_diffParams, int* /*d_unified_dHdx_dHdt_Index*/,
    bool* d_is_GPU_HC_Sol_Converge, bool* d_is_GPU_HC_Sol_Infinity, magmaFloatComplex* /*d_Debug_Purpose*/)
{
  return track(my_queue, sub_RANSAC_iters, HC_max_steps, HC_max_correction_steps, HC_delta_t_incremental_steps, d_startSols_array,
               d_Track_array, d_startParams, d_targetParams, d_diffParams, d_is_GPU_HC_Sol_Converge, d_is_GPU_HC_Sol_Infinity);
}

real_Double_t kernel_GPUHC_trifocal_2op1p_30x30_PH_CodeOpt_TrunPaths_TrunRANSAC(
    magma_queue_t my_queue, int sub_RANSAC_iters, int Num_Of_Triplet_Edgels, int HC_max_steps, int HC_max_correction_steps,
    int HC_delta_t_incremental_steps, magmaFloatComplex** d_startSols_array, magmaFloatComplex** d_Track_array,
    magmaFloatComplex* d_startParams, magmaFloatComplex* d_targetParams, magmaFloatComplex* d_diffParams,
    int* /*d_unified_dHdx_dHdt_Index*/, float* d_Triplet_Edge_Locations, float* d_Intrinsic_Matrix,
    bool* d_is_GPU_HC_Sol_Converge, bool* d_is_GPU_HC_Sol_Infinity, magmaFloatComplex* /*d_Debug_Purpose*/,
    bool* d_Found_Trifocal_Sols, int* d_Trifocal_Sols_Batch_Index)
{
  return track_abort(my_queue, sub_RANSAC_iters, Num_Of_Triplet_Edgels, HC_max_steps, HC_max_correction_steps,
                     HC_delta_t_incremental_steps, d_startSols_array, d_Track_array, d_startParams, d_targetParams, d_diffParams,
                     d_Triplet_Edge_Locations, d_Intrinsic_Matrix, d_is_GPU_HC_Sol_Converge, d_is_GPU_HC_Sol_Infinity,
                     d_Found_Trifocal_Sols, d_Trifocal_Sols_Batch_Index);
}

// The pre-Ampere twins (separate dHdx / dHdt index tables) are referenced by GPU_HC_Solver.cpp:417-433 and therefore have to
// link; on this library there is one code path, so they forward to the same launches.
real_Double_t kernel_GPUHC_trifocal_2op1p_30x30_PH_CodeOpt_TrunPaths_Volta(
    magma_queue_t my_queue, int sub_RANSAC_iters, int HC_max_steps, int HC_max_correction_steps, int HC_delta_t_incremental_steps,
    magmaFloatComplex** d_startSols_array, magmaFloatComplex** d_Track_array, magmaFloatComplex* d_startParams,
    magmaFloatComplex* d_targetParams, magmaFloatComplex* d_diffParams, int* /*d_dHdx_Index*/, int* /*d_dHdt_Index*/,
    bool* d_is_GPU_HC_Sol_Converge, bool* d_is_GPU_HC_Sol_Infinity, magmaFloatComplex* /*d_Debug_Purpose*/)
{
  return track(my_queue, sub_RANSAC_iters, HC_max_steps, HC_max_correction_steps, HC_delta_t_incremental_steps, d_startSols_array,
               d_Track_array, d_startParams, d_targetParams, d_diffParams, d_is_GPU_HC_Sol_Converge, d_is_GPU_HC_Sol_Infinity);
}

real_Double_t kernel_GPUHC_trifocal_2op1p_30x30_PH_CodeOpt_TrunPaths_TrunRANSAC_Volta(
    magma_queue_t my_queue, int sub_RANSAC_iters, int Num_Of_Triplet_Edgels, int HC_max_steps, int HC_max_correction_steps,
    int HC_delta_t_incremental_steps, magmaFloatComplex** d_startSols_array, magmaFloatComplex** d_Track_array,
    magmaFloatComplex* d_startParams, magmaFloatComplex* d_targetParams, magmaFloatComplex* d_diffParams,
    int* /*d_dHdx_Index*/, int* /*d_dHdt_Index*/, float* d_Triplet_Edge_Locations, float* d_Intrinsic_Matrix,
    bool* d_is_GPU_HC_Sol_Converge, bool* d_is_GPU_HC_Sol_Infinity, magmaFloatComplex* /*d_Debug_Purpose*/,
    bool* d_Found_Trifocal_Sols, int* d_Trifocal_Sols_Batch_Index)
{
  return track_abort(my_queue, sub_RANSAC_iters, Num_Of_Triplet_Edgels, HC_max_steps, HC_max_correction_steps,
                     HC_delta_t_incremental_steps, d_startSols_array, d_Track_array, d_startParams, d_targetParams, d_diffParams,
                     d_Triplet_Edge_Locations, d_Intrinsic_Matrix, d_is_GPU_HC_Sol_Converge, d_is_GPU_HC_Sol_Infinity,
                     d_Found_Trifocal_Sols, d_Trifocal_Sols_Batch_Index);
}

// forget the cached pointer-array bases (only needed by a host layer that re-points d_startSols_array / d_Track_array)
extern "C" void hcb200_shim_forget() { for (auto& c : g_bases) c = BaseCache{}; }
