// Minimal host-side stand-in for MAGMA's public header, written for this repo so that the UNMODIFIED
// reference CPU-HC sources (/root/reference/magmaHC/{CPU_HC_Solver,Data_Reader,Evaluations}.cpp,
// cpuhc-solvers/CPUHC_Generic_Solver_Eval_by_Indx.cpp) compile without MAGMA installed.
// It provides only the surface those files touch: the complex type, MAGMA_C_* macros, the arithmetic
// operators of MAGMA's magma_operators.h (documented semantics: SURVEY.md §8c item 4), host malloc/free,
// a wall clock and thread-count setters.  Test infrastructure only (see oracle/README.md).
#ifndef HCB200_REFSHIM_MAGMA_V2_H
#define HCB200_REFSHIM_MAGMA_V2_H
#include <cuComplex.h>
// headers the reference gets transitively from the real MAGMA / yaml-cpp headers
#include <array>
#include <cassert>
#include <vector>
#include <string>
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <omp.h>

typedef cuFloatComplex magmaFloatComplex;
typedef magmaFloatComplex* magmaFloatComplex_ptr;
typedef int magma_int_t;
typedef int magma_device_t;
typedef double real_Double_t;
struct magma_queue;
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

static inline magmaFloatComplex operator+(const magmaFloatComplex a, const magmaFloatComplex b) { return make_cuFloatComplex(a.x + b.x, a.y + b.y); }
static inline magmaFloatComplex operator-(const magmaFloatComplex a, const magmaFloatComplex b) { return make_cuFloatComplex(a.x - b.x, a.y - b.y); }
static inline magmaFloatComplex operator-(const magmaFloatComplex a) { return make_cuFloatComplex(-a.x, -a.y); }
static inline magmaFloatComplex operator*(const magmaFloatComplex a, const magmaFloatComplex b) { return make_cuFloatComplex(a.x * b.x - a.y * b.y, a.y * b.x + a.x * b.y); }
static inline magmaFloatComplex operator*(const magmaFloatComplex a, const float s) { return make_cuFloatComplex(a.x * s, a.y * s); }
static inline magmaFloatComplex operator*(const float s, const magmaFloatComplex a) { return make_cuFloatComplex(a.x * s, a.y * s); }
static inline magmaFloatComplex operator/(const magmaFloatComplex a, const float s) { return make_cuFloatComplex(a.x / s, a.y / s); }
static inline magmaFloatComplex operator/(const magmaFloatComplex a, const magmaFloatComplex b) { return cuCdivf(a, b); }
static inline magmaFloatComplex& operator+=(magmaFloatComplex& a, const magmaFloatComplex b) { a.x += b.x; a.y += b.y; return a; }
static inline magmaFloatComplex& operator-=(magmaFloatComplex& a, const magmaFloatComplex b) { a.x -= b.x; a.y -= b.y; return a; }
static inline magmaFloatComplex& operator*=(magmaFloatComplex& a, const magmaFloatComplex b) { a = a * b; return a; }
static inline magmaFloatComplex& operator*=(magmaFloatComplex& a, const float s) { a.x *= s; a.y *= s; return a; }

static inline int magma_init() { return 0; }
static inline int magma_finalize() { return 0; }
static inline int magma_cmalloc_cpu(magmaFloatComplex** p, size_t n) { *p = (magmaFloatComplex*)malloc(n * sizeof(magmaFloatComplex)); return *p ? 0 : 1; }
static inline int magma_imalloc_cpu(magma_int_t** p, size_t n) { *p = (magma_int_t*)malloc(n * sizeof(magma_int_t)); return *p ? 0 : 1; }
static inline int magma_free_cpu(void* p) { free(p); return 0; }
static inline double magma_wtime() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static inline void magma_set_lapack_numthreads(int) {}
static inline void magma_set_omp_numthreads(int n) { omp_set_num_threads(n); }
#endif
