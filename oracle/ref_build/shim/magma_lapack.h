// LAPACK binding for the reference CPU-HC build: the image has no system LAPACK, but the OpenBLAS bundled with
// the scipy wheel exports the Fortran-ABI routine as `scipy_cgesv_` (LP64).  SURVEY.md App. D.1.
#ifndef HCB200_REFSHIM_MAGMA_LAPACK_H
#define HCB200_REFSHIM_MAGMA_LAPACK_H
#include "magma_v2.h"
extern "C" void scipy_cgesv_(const int* n, const int* nrhs, magmaFloatComplex* A, const int* lda, int* ipiv,
                             magmaFloatComplex* B, const int* ldb, int* info);
#define lapackf77_cgesv scipy_cgesv_
#endif
