// Device-capable stand-in for MAGMA's public header, written for this repo so that the UNMODIFIED reference GPU-HC++
// kernels (/root/reference/magmaHC/gpu-kernels/*.cu, dev-*.cuh, gpu-idx-evals/*.cuh) compile for sm_100a without MAGMA.
// Only what those files touch: the complex type, MAGMA_C_* macros, the operators of magma_operators.h,
// magmablas_syncwarp and a queue that hands out a CUDA stream.  Test / baseline infrastructure only.
#ifndef HCB200_REFGSHIM_MAGMA_V2_H
#define HCB200_REFGSHIM_MAGMA_V2_H
#include <cuda_runtime.h>
#include <cuComplex.h>
#include <array>
#include <cassert>
#include <vector>
#include <string>

typedef cuFloatComplex magmaFloatComplex;
typedef magmaFloatComplex* magmaFloatComplex_ptr;
typedef int magma_int_t;
typedef int magma_device_t;
typedef double real_Double_t;
struct magma_queue { cudaStream_t s; cudaStream_t cuda_stream() { return s; } };
typedef magma_queue* magma_queue_t;

#define MAGMA_C_MAKE(r, i)   make_cuFloatComplex((float)(r), (float)(i))
#define MAGMA_C_REAL(a)      ((a).x)
#define MAGMA_C_IMAG(a)      ((a).y)
#define MAGMA_C_ZERO         make_cuFloatComplex(0.0f, 0.0f)
#define MAGMA_C_ONE          make_cuFloatComplex(1.0f, 0.0f)
#define MAGMA_C_NEG_ONE      make_cuFloatComplex(-1.0f, 0.0f)
#define MAGMA_C_DIV(a, b)    cuCdivf((a), (b))
#define MAGMA_D_ZERO         (0.0)
#define MAGMA_D_ONE          (1.0)

#define HD __host__ __device__ static inline
HD magmaFloatComplex operator+(const magmaFloatComplex a, const magmaFloatComplex b) { return make_cuFloatComplex(a.x + b.x, a.y + b.y); }
HD magmaFloatComplex operator-(const magmaFloatComplex a, const magmaFloatComplex b) { return make_cuFloatComplex(a.x - b.x, a.y - b.y); }
HD magmaFloatComplex operator-(const magmaFloatComplex a) { return make_cuFloatComplex(-a.x, -a.y); }
HD magmaFloatComplex operator*(const magmaFloatComplex a, const magmaFloatComplex b) { return make_cuFloatComplex(a.x * b.x - a.y * b.y, a.y * b.x + a.x * b.y); }
HD magmaFloatComplex operator*(const magmaFloatComplex a, const float s) { return make_cuFloatComplex(a.x * s, a.y * s); }
HD magmaFloatComplex operator*(const float s, const magmaFloatComplex a) { return make_cuFloatComplex(a.x * s, a.y * s); }
HD magmaFloatComplex operator/(const magmaFloatComplex a, const float s) { return make_cuFloatComplex(a.x / s, a.y / s); }
HD magmaFloatComplex operator/(const magmaFloatComplex a, const magmaFloatComplex b) { return cuCdivf(a, b); }
HD magmaFloatComplex& operator+=(magmaFloatComplex& a, const magmaFloatComplex b) { a.x += b.x; a.y += b.y; return a; }
HD magmaFloatComplex& operator-=(magmaFloatComplex& a, const magmaFloatComplex b) { a.x -= b.x; a.y -= b.y; return a; }
HD magmaFloatComplex& operator*=(magmaFloatComplex& a, const magmaFloatComplex b) { a = a * b; return a; }
HD magmaFloatComplex& operator*=(magmaFloatComplex& a, const float s) { a.x *= s; a.y *= s; return a; }
#undef HD
__device__ static inline void magmablas_syncwarp() { __syncwarp(); }
#endif
