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
#include "util.hpp"
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

// The same solver on ANOTHER problem folder (<tree>/problems/<problem_name>/ in the reference's layout): the reference's CPU-HC takes every size
// from gpuhc_settings.yaml and its evaluators walk the index tables with run-time sizes, so it runs any problem given as data.  The RANSAC
// front end (edgels -> target parameters) is trifocal-specific and is not used: the caller supplies the target parameters
// (n_hyp * (Num_Of_Params + 1) complex, last entry of every hypothesis = 1).
int ref_cpuhc_run_problem(const char* bin_dir, const char* problem_name, int n_hyp, int n_cores, const float* in_target_params,
                          float* out_tracks, unsigned char* out_conv, unsigned char* out_inf, double* out_seconds)
{
  char cwd[4096];
  if (!in_target_params || !problem_name) return 7;
  if (!getcwd(cwd, sizeof cwd)) return 1;
  if (chdir(bin_dir) != 0) return 2;
  g_hcb200_ref_num_hyp = n_hyp;
  int rc = 0;
  try {
    YAML::Node cfg = YAML::LoadFile(std::string("../../problems/") + problem_name + "/gpuhc_settings.yaml");
    cfg.set("Num_Of_Cores", std::to_string(n_cores));
    CPU_HC_Solver s(cfg);
    s.Allocate_Arrays();
    if (!s.Read_Problem_Data()) rc = 3;
    if (!rc) {
      const int P1 = s.Num_Of_Params + 1;
      for (int i = 0; i < n_hyp * P1; i++) {
        s.h_Target_Params[i] = MAGMA_C_MAKE(in_target_params[2 * i], in_target_params[2 * i + 1]);
        s.h_diff_params[i]   = s.h_Target_Params[i] - s.h_Start_Params[i % P1];
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
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_cpuhc_run_problem: %s\n", e.what());
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

// Support counting of ONE end point with the reference's own MVG helpers (class util, magmaHC/util.hpp:29-209), called the way
// Evaluations::Transform_GPUHC_Sols_to_Trifocal_Relative_Pose + ::get_Solution_with_Maximal_Support (Evaluations.cpp:298-504) call them,
// but with the candidate's own end point (the reference indexes the first track for every candidate).  Returns 1
// if the path passes the reference's candidate gates (:317-335): |Im| of the six Cayley parameters < IMAG_PART_TOL, eight depths >= 0.
// pose24 = R21 (9, row-major), t21 (3), R31 (9), t31 (3) as the reference normalises them.  This pins the arithmetic of
// hcb200_score_tracks / host/mvg.hpp to util.hpp (tests/golden/ref_util_support.npz).
int ref_util_support(const float* x31_re_im, const float* locations, int n_edgels, const float* K9, int* n21, int* n31, float* pose24)
{
  *n21 = 0; *n31 = 0;
  const magmaFloatComplex* x = (const magmaFloatComplex*)x31_re_im;
  for (int vi = 0; vi < 6; vi++) if (!(fabs(MAGMA_C_IMAG(x[24 + vi])) < IMAG_PART_TOL)) return 0;
  for (int di = 0; di < 8; di++) if (!(MAGMA_C_REAL(x[di]) >= 0)) return 0;
  util u;
  float* t21 = new float[3]; float* t31 = new float[3]; float* R21 = new float[9]; float* R31 = new float[9];
  float r21[3], r31[3], K[9];
  for (int i = 0; i < 3; i++) { t21[i] = MAGMA_C_REAL(x[18 + i]); t31[i] = MAGMA_C_REAL(x[21 + i]); r21[i] = MAGMA_C_REAL(x[24 + i]); r31[i] = MAGMA_C_REAL(x[27 + i]); }
  for (int i = 0; i < 9; i++) K[i] = K9[i];
  u.Normalize_Translation_Vector(t21);
  u.Normalize_Translation_Vector(t31);
  u.Cayley_To_Rotation_Matrix(r21, R21);
  u.Cayley_To_Rotation_Matrix(r31, R31);
  for (int i = 0; i < 9; i++) { pose24[i] = R21[i]; pose24[12 + i] = R31[i]; }
  for (int i = 0; i < 3; i++) { pose24[9 + i] = t21[i]; pose24[21 + i] = t31[i]; }
  float g1[3], g2[3], g3[3];
  for (int ei = 0; ei < n_edgels; ei++) {
    const float* e = locations + (size_t)ei * 6;
    g1[0] = e[0]; g1[1] = e[1]; g1[2] = 1.0f; g2[0] = e[2]; g2[1] = e[3]; g2[2] = 1.0f; g3[0] = e[4]; g3[1] = e[5]; g3[2] = 1.0f;
    const float rho21 = u.get_depth_rho(g1, g2, R21, t21);
    const float err21 = u.get_Reprojection_Pixels_Error(g1, g2, R21, t21, K, rho21);
    const float rho31 = u.get_depth_rho(g1, g3, R31, t31);
    const float err31 = u.get_Reprojection_Pixels_Error(g1, g3, R31, t31, K, rho31);
    if (err21 < REPROJ_ERROR_INLIER_THRESH) (*n21)++;
    if (err31 < REPROJ_ERROR_INLIER_THRESH) (*n31)++;
  }
  delete[] t21; delete[] t31; delete[] R21; delete[] R31;
  return 1;
}

// the two helpers on their own (util.hpp:169-209), for value-level goldens: out = (rho, reprojection error in pixels)
void ref_util_pair(const float* gamma1, const float* gamma2, const float* R9, const float* T3, const float* K9, float* out2)
{
  util u;
  float g1[3] = {gamma1[0], gamma1[1], 1.0f}, g2[3] = {gamma2[0], gamma2[1], 1.0f}, R[9], T[3], K[9];
  for (int i = 0; i < 9; i++) { R[i] = R9[i]; K[i] = K9[i]; }
  for (int i = 0; i < 3; i++) T[i] = T3[i];
  out2[0] = u.get_depth_rho(g1, g2, R, T);
  out2[1] = u.get_Reprojection_Pixels_Error(g1, g2, R, T, K, out2[0]);
}

}  // extern "C"
