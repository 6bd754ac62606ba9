// Host + device stand-in for MAGMA's public header on top of the CUDA runtime, written for this repo so that the UNMODIFIED
// reference host layer (/root/reference/magmaHC/GPU_HC_Solver.cpp, Data_Reader.cpp, Evaluations.cpp) compiles and RUNS without
// MAGMA: types / macros / operators come from ../gshim/magma_v2.h, this file adds the ~20 runtime calls GPU_HC_Solver.cpp makes
// (device selection, queues, allocation, set/get matrix, pointer arrays, wall clock) with MAGMA's documented semantics.
// Used only by the drop-in check (oracle/Makefile target ref_dropin).  Test infrastructure.
#ifndef HCB200_REFHSHIM_MAGMA_V2_H
#define HCB200_REFHSHIM_MAGMA_V2_H
#include "../gshim/magma_v2.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <memory>

static inline int magma_init() { return 0; }
static inline int magma_finalize() { return 0; }
static inline void magma_print_environment() {}
static inline void magma_setdevice(magma_device_t d) { cudaSetDevice(d); }
static inline void magma_getdevice(magma_device_t* d) { cudaGetDevice(d); }
static inline void magma_getdevices(magma_device_t* devs, magma_int_t size, magma_int_t* n)
{ int c = 0; cudaGetDeviceCount(&c); *n = std::min(c, size); for (int i = 0; i < *n; i++) devs[i] = i; }
static inline magma_int_t magma_getdevice_arch()           // MAGMA: major*100 + minor*10
{ int d = 0; cudaGetDevice(&d); cudaDeviceProp p; cudaGetDeviceProperties(&p, d); return p.major * 100 + p.minor * 10; }
static inline void magma_queue_create(magma_device_t dev, magma_queue_t* q)
{ cudaSetDevice(dev); *q = new magma_queue; cudaStreamCreate(&(*q)->s); }
static inline void magma_queue_destroy(magma_queue_t q) { if (q) { cudaStreamDestroy(q->s); delete q; } }
static inline void magma_queue_sync(magma_queue_t q) { cudaStreamSynchronize(q->s); }
static inline int magma_cmalloc_cpu(magmaFloatComplex** p, size_t n) { *p = (magmaFloatComplex*)malloc(std::max<size_t>(n, 1) * sizeof(magmaFloatComplex)); return *p ? 0 : 1; }
static inline int magma_free_cpu(void* p) { free(p); return 0; }
static inline int magma_malloc(void** p, size_t bytes) { return cudaMalloc(p, std::max<size_t>(bytes, 1)) == cudaSuccess ? 0 : 1; }
static inline int magma_cmalloc(magmaFloatComplex** p, size_t n) { return magma_malloc((void**)p, n * sizeof(magmaFloatComplex)); }
static inline int magma_free(void* p) { return cudaFree(p) == cudaSuccess ? 0 : 1; }
// column-major m x n copies, leading dimensions in elements; the non-async MAGMA forms return after the copy is done
static inline void magma_csetmatrix(magma_int_t m, magma_int_t n, const magmaFloatComplex* hA, magma_int_t lda, magmaFloatComplex* dB, magma_int_t ldb, magma_queue_t q)
{ cudaMemcpy2DAsync(dB, (size_t)ldb * sizeof(magmaFloatComplex), hA, (size_t)lda * sizeof(magmaFloatComplex), (size_t)m * sizeof(magmaFloatComplex), n, cudaMemcpyHostToDevice, q->s); cudaStreamSynchronize(q->s); }
static inline void magma_cgetmatrix(magma_int_t m, magma_int_t n, const magmaFloatComplex* dA, magma_int_t lda, magmaFloatComplex* hB, magma_int_t ldb, magma_queue_t q)
{ cudaMemcpy2DAsync(hB, (size_t)ldb * sizeof(magmaFloatComplex), dA, (size_t)lda * sizeof(magmaFloatComplex), (size_t)m * sizeof(magmaFloatComplex), n, cudaMemcpyDeviceToHost, q->s); cudaStreamSynchronize(q->s); }
// output_array[i] = input + i*batch_offset + row + column*lda  (magmablas set_pointer)
static inline void magma_cset_pointer(magmaFloatComplex** output_array, magmaFloatComplex* input, magma_int_t lda, magma_int_t row, magma_int_t column, magma_int_t batch_offset, magma_int_t batchCount, magma_queue_t q)
{
  std::unique_ptr<magmaFloatComplex*[]> h(new magmaFloatComplex*[std::max(batchCount, 1)]);
  for (magma_int_t i = 0; i < batchCount; i++) h[i] = input + (size_t)i * batch_offset + row + (size_t)column * lda;
  cudaMemcpyAsync(output_array, h.get(), (size_t)batchCount * sizeof(magmaFloatComplex*), cudaMemcpyHostToDevice, q->s);
  cudaStreamSynchronize(q->s);
}
static inline double magma_wtime() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#endif
