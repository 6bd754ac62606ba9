// C-ABI wrapper that drives the UNMODIFIED reference CPU-HC solver (class CPU_HC_Solver,
// /root/reference/magmaHC/CPU_HC_Solver.{hpp,cpp} + cpuhc-solvers/CPUHC_Generic_Solver_Eval_by_Indx.cpp)
// the way the reference's own driver does (cmd/magmaHC-main.cpp:124-158), and hands the per-path results
// back as plain arrays.  Built only by oracle/Makefile into oracle/_ref/libref_cpuhc.so.
// TEST INFRASTRUCTURE: used by tests/, tools/make_golden.py and bench.py's reference arm; never by the product.
#include <array>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>
#include <unistd.h>
#include <yaml-cpp/yaml.h>
#include "magma_v2.h"

int g_hcb200_ref_num_hyp = 100;   // read through the NUM_OF_RANSAC_ITERATIONS macro (see generated definitions.hpp)

// reach the per-path arrays, which the reference keeps private
#define private public
#define class struct
#include "CPU_HC_Solver.hpp"
#undef class
#undef private

extern "C" {

// Runs n_hyp hypotheses x 312 paths with the reference CPU-HC.  `bin_dir` must be <tree>/build/bin of a tree made by
// fixtures.materialize_tree (the reference opens "../../problems/..." relative to the cwd).
// in_target_params (optional, n_hyp*34 complex): overrides the sampler's target parameters.
// Returns 0 on success.
int ref_cpuhc_run(const char* bin_dir, int n_hyp, unsigned seed, int dataset_index, int n_cores,
                  const float* in_target_params,
                  float* out_tracks, unsigned char* out_conv, unsigned char* out_inf,
                  float* out_target_params, double* out_seconds)
{
  char cwd[4096];
  if (!getcwd(cwd, sizeof cwd)) return 1;
  if (chdir(bin_dir) != 0) return 2;
  g_hcb200_ref_num_hyp = n_hyp;
  int rc = 0;
  try {
    YAML::Node cfg = YAML::LoadFile("../../problems/trifocal_2op1p_30x30/gpuhc_settings.yaml");
    cfg.set("Num_Of_Cores", std::to_string(n_cores));
    CPU_HC_Solver s(cfg);
    s.Allocate_Arrays();
    if (!s.Read_Problem_Data()) rc = 3;
    if (!rc && !s.Read_RANSAC_Data(dataset_index)) rc = 4;
    if (!rc) {
      s.Prepare_Target_Params(seed);
      const int P1 = s.Num_Of_Params + 1;
      if (in_target_params) {
        for (int i = 0; i < n_hyp * P1; i++) {
          s.h_Target_Params[i] = MAGMA_C_MAKE(in_target_params[2 * i], in_target_params[2 * i + 1]);
          s.h_diff_params[i]   = s.h_Target_Params[i] - s.h_Start_Params[i % P1];
        }
      }
      s.Set_Initial_Array_Vals();
      double t = s.CPUHC_Generic_Solver_Eval_by_Indx(n_cores, cpu_eval_indx_dHdX_trifocal_2op1p_30,
                                                      cpu_eval_indx_dHdt_trifocal_2op1p_30,
                                                      cpu_eval_indx_H_trifocal_2op1p_30);
      if (out_seconds) *out_seconds = t;
      const int n_paths = n_hyp * s.Num_Of_Tracks, V1 = s.Num_Of_Vars + 1;
      if (out_tracks) std::memcpy(out_tracks, s.h_CPU_HC_Track_Sols, sizeof(magmaFloatComplex) * (size_t)n_paths * V1);
      for (int i = 0; i < n_paths; i++) {
        if (out_conv) out_conv[i] = s.h_is_Track_Converged[i] ? 1 : 0;
        if (out_inf)  out_inf[i]  = s.h_is_Track_Inf_Failed[i] ? 1 : 0;
      }
      if (out_target_params) std::memcpy(out_target_params, s.h_Target_Params, sizeof(magmaFloatComplex) * (size_t)n_hyp * P1);
      s.Free_Triplet_Edgels_Mem();
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_cpuhc_run: %s\n", e.what());
    rc = 5;
  }
  if (chdir(cwd) != 0 && !rc) rc = 6;
  return rc;
}

// The reference's three index-table evaluators (cpu-jacobian-evals/cpu-eval-indx_trifocal_2op1p_30x30.hpp:22-89),
// exposed so the oracle's evaluators can be pinned term by term.  A is column-major 30x30 (LAPACK layout).
void ref_eval_dHdX(const int* dHdx_index, const float* x31, const float* p34, float* A900)
{ cpu_eval_indx_dHdX_trifocal_2op1p_30(30, 8, 40, 5, dHdx_index, (magmaFloatComplex*)x31, (magmaFloatComplex*)p34, (magmaFloatComplex*)A900); }
void ref_eval_dHdt(const int* dHdt_index, const float* x31, const float* p34, const float* dp34, float* b30)
{ cpu_eval_indx_dHdt_trifocal_2op1p_30(30, 16, 6, dHdt_index, (magmaFloatComplex*)x31, (magmaFloatComplex*)p34, (magmaFloatComplex*)b30, (magmaFloatComplex*)dp34); }
void ref_eval_H(const int* dHdt_index, const float* x31, const float* p34, float* b30)
{ cpu_eval_indx_H_trifocal_2op1p_30(30, 16, 6, dHdt_index, (magmaFloatComplex*)x31, (magmaFloatComplex*)p34, (magmaFloatComplex*)b30); }

// LAPACK cgesv exactly as the reference calls it (CPUHC_Generic_Solver_Eval_by_Indx.cpp:93); A column-major, overwritten.
int ref_cgesv(float* A900, float* b30)
{ int n = 30, nrhs = 1, info = 0, ipiv[30]; lapackf77_cgesv(&n, &nrhs, (magmaFloatComplex*)A900, &n, ipiv, (magmaFloatComplex*)b30, &n, &info); return info; }

}  // extern "C"
