// See Evaluations.hpp.
#include "Evaluations.hpp"

#include <cmath>
#include <cstdio>
#include <iomanip>
#include <iostream>
#include <set>

using hcb200::complex32;
namespace mvg = hcb200::mvg;

Evaluations::Evaluations(std::string Output_Files_Path, std::string GPU_or_CPU, int tracks, int vars)
    : WRITE_FILES_PATH(std::move(Output_Files_Path)), evaluate_GPUHC_or_CPUHC(std::move(GPU_or_CPU)),
      num_of_tracks(tracks), num_of_variables(vars)
{
  std::string sols, steps;
  if (evaluate_GPUHC_or_CPUHC == "GPU-HC") { sols = "GPU_Converged_HC_tracks.txt"; steps = "GPUHC_Steps_of_Actual_Solutions.txt"; }
  else if (evaluate_GPUHC_or_CPUHC == "CPU-HC") { sols = "CPU_Converged_HC_tracks.txt"; steps = "CPUHC_Steps_of_Actual_Solutions.txt"; }
  else hcb200::log_error("Invalid GPU_or_CPU input parameter for the Evaluation constructor.");
  HC_Track_Sols_File.open(WRITE_FILES_PATH + sols);
  if (!HC_Track_Sols_File.is_open()) hcb200::log_file_error(WRITE_FILES_PATH + sols);
  HC_Actual_Sols_Steps_File.open(WRITE_FILES_PATH + steps);
  if (!HC_Actual_Sols_Steps_File.is_open()) hcb200::log_file_error(WRITE_FILES_PATH + steps);
}

Evaluations::~Evaluations() { HC_Track_Sols_File.close(); HC_Actual_Sols_Steps_File.close(); }

void Evaluations::Flush_Out_Data()
{
  Num_Of_Inf_Sols = Num_Of_Coverged_Sols = Num_Of_Real_Sols = Num_Of_Unique_Sols = 0;
  Percentage_Of_Convergence = Percentage_Of_Inf_Sols = Percentage_Of_Real_Sols = Percentage_Of_Unique_Sols = 0.0f;
  success_flag = false;
  Min_Residual_R21 = Min_Residual_R31 = Min_Residual_t21 = Min_Residual_t31 = 100.0f;
  real_track_indices.clear(); HC_steps_of_actual_solutions.clear(); Per_Hypothesis_Counts.clear();
  normalized_t21s.clear(); normalized_t31s.clear(); normalized_R21s.clear(); normalized_R31s.clear(); F21s.clear(); F31s.clear();
  Max_Reproj_Inliers_Support_Views21_Index.clear(); Max_Reproj_Inliers_Support_Views31_Index.clear();
  Best_Candidate_Path_Index = -1;
}

// Format of {GPU,CPU}_Converged_HC_tracks.txt (Evaluations.cpp:120-143 of the reference): per RANSAC iteration a header
// line, then for every converged track its running index and 30 lines "re<TAB>im" with 20 significant digits.
void Evaluations::Write_Converged_Sols(complex32* tracks, bool* conv)
{
  hcb200::log_info("Writing HC converged solutions to a file ...");
  int counter = 0;
  const int stride = num_of_variables + 1;
  for (int ri = 0; ri < num_of_ransac_iters; ri++) {
    HC_Track_Sols_File << "-------------------- RANSAC Iteration " << ri + 1 << " --------------------\n\n";
    for (int bs = 0; bs < num_of_tracks; bs++, counter++) {
      const size_t path = (size_t)ri * num_of_tracks + bs;
      if (!conv[path]) continue;
      HC_Track_Sols_File << counter << "\n";
      for (int v = 0; v < num_of_variables; v++)
        HC_Track_Sols_File << std::setprecision(20) << tracks[path * stride + v].x << "\t" << std::setprecision(20) << tracks[path * stride + v].y << "\n";
      HC_Track_Sols_File << "\n";
    }
    HC_Track_Sols_File << "\n";
  }
}

// Evaluations.cpp:145-167: converged, infinity-failed, and "real" = converged with all 30 |imag| <= 1e-4
void Evaluations::Evaluate_HC_Sols(complex32* tracks, bool* conv, bool* inf, int ri)
{
  const int stride = num_of_variables + 1;
  std::array<unsigned, 3> c = {0, 0, 0};
  for (int bs = 0; bs < num_of_tracks; bs++) {
    const size_t path = (size_t)ri * num_of_tracks + bs;
    if (conv[path]) c[0]++;
    if (inf[path]) c[1]++;
    if (conv[path]) {
      bool real = true;
      for (int v = 0; v < num_of_variables && real; v++) real = (std::fabs(tracks[path * stride + v].y) <= ZERO_IMAG_PART_TOL_FOR_SP);
      if (real) c[2]++;
    }
  }
  Num_Of_Coverged_Sols += c[0]; Num_Of_Inf_Sols += c[1]; Num_Of_Real_Sols += c[2];
  Per_Hypothesis_Counts.push_back(c);
}

void Evaluations::Evaluate_RANSAC_HC_Sols(complex32* tracks, bool* conv, bool* inf)
{
  for (int ri = 0; ri < num_of_ransac_iters; ri++) Evaluate_HC_Sols(tracks, conv, inf, ri);
  const float total = (float)(num_of_tracks * num_of_ransac_iters);
  Percentage_Of_Convergence = (float)Num_Of_Coverged_Sols / total;
  Percentage_Of_Inf_Sols = (float)Num_Of_Inf_Sols / total;
  Percentage_Of_Real_Sols = (float)Num_Of_Real_Sols / total;
}

void Evaluations::Set_RANSAC_HC_Sol_Counts(const unsigned* c, int n_hypotheses)
{
  for (int ri = 0; ri < n_hypotheses; ri++) {
    Num_Of_Coverged_Sols += c[3 * ri]; Num_Of_Inf_Sols += c[3 * ri + 1]; Num_Of_Real_Sols += c[3 * ri + 2];
    Per_Hypothesis_Counts.push_back({c[3 * ri], c[3 * ri + 1], c[3 * ri + 2]});
  }
  const float total = (float)(num_of_tracks * num_of_ransac_iters);
  Percentage_Of_Convergence = (float)Num_Of_Coverged_Sols / total;
  Percentage_Of_Inf_Sols = (float)Num_Of_Inf_Sols / total;
  Percentage_Of_Real_Sols = (float)Num_Of_Real_Sols / total;
}

// Evaluations.cpp:184-233: a converged track is unique if no later track agrees with it in every variable to 1e-4
void Evaluations::Find_Unique_Sols(complex32* tracks, bool* conv)
{
  const int stride = num_of_variables + 1;
  std::set<int> duplicates_of_current, skip;
  for (int bs = 0; bs < num_of_tracks; bs++) {
    if (!conv[bs]) continue;
    if (!skip.empty()) { if (skip.count(bs)) continue; duplicates_of_current.clear(); }
    for (int ds = bs + 1; ds < num_of_tracks; ds++) {
      bool same = true;
      for (int v = 0; v < num_of_variables && same; v++)
        same = std::fabs(tracks[bs * stride + v].x - tracks[ds * stride + v].x) < DUPLICATE_SOL_DIFF_TOL &&
               std::fabs(tracks[bs * stride + v].y - tracks[ds * stride + v].y) < DUPLICATE_SOL_DIFF_TOL;
      if (same) duplicates_of_current.insert(ds);
    }
    if (duplicates_of_current.empty()) { Num_Of_Unique_Sols++; Unique_Sols_Index.push_back(bs); }
    else skip = duplicates_of_current;
  }
}

void Evaluations::Convert_Trifocal_Translation(complex32* x)
{
  raw_t21 = {x[18].x, x[19].x, x[20].x};
  raw_t31 = {x[21].x, x[22].x, x[23].x};
  normalized_t21 = mvg::normalized(raw_t21);
  normalized_t31 = mvg::normalized(raw_t31);
}

void Evaluations::Convert_Trifocal_Rotation(complex32* x)
{
  normalized_R21 = mvg::cayley_to_rotation({x[24].x, x[25].x, x[26].x});
  normalized_R31 = mvg::cayley_to_rotation({x[27].x, x[28].x, x[29].x});
}

// Candidate gate of the reference (Evaluations.cpp:298-358): converged, |imag| of the six Cayley parameters < 1e-5 and all eight
// depths >= 0; each passing path contributes ITS OWN pose (the reference converts path 0 every time).
void Evaluations::Transform_GPUHC_Sols_to_Trifocal_Relative_Pose(complex32* tracks, bool* conv, float* IntrinsicMatrix)
{
  for (int i = 0; i < 9; i++) K[i] = IntrinsicMatrix[i];
  const int stride = num_of_variables + 1;
  const int n_paths = num_of_tracks * num_of_ransac_iters;
  for (int bs = 0; bs < n_paths; bs++) {
    if (!conv[bs]) continue;
    complex32* x = tracks + (size_t)bs * stride;
    bool ok = true;
    for (int vi = 24; vi < 30 && ok; vi++) ok = std::fabs(x[vi].y) < IMAG_PART_TOL;
    for (int di = 0; di < 8 && ok; di++) ok = x[di].x >= 0;
    if (!ok) continue;
    Convert_Trifocal_Translation(x);
    Convert_Trifocal_Rotation(x);
    normalized_t21s.push_back(normalized_t21); normalized_t31s.push_back(normalized_t31);
    normalized_R21s.push_back(normalized_R21); normalized_R31s.push_back(normalized_R31);
    F21s.push_back(mvg::fundamental_matrix(K, normalized_R21, normalized_t21));
    F31s.push_back(mvg::fundamental_matrix(K, normalized_R31, normalized_t31));
    real_track_indices.push_back(bs);
  }
}

float Evaluations::get_Rotation_Residual(float* GT_R, std::array<float, 9> Sol_R)
{
  mvg::Mat3 G; for (int i = 0; i < 9; i++) G[i] = GT_R[i];
  return mvg::rotation_residual(G, Sol_R);
}

float Evaluations::get_Translation_Residual(float* GT_Transl, std::array<float, 3> Sol_Transl)
{ return mvg::translation_residual({GT_Transl[0], GT_Transl[1], GT_Transl[2]}, Sol_Transl); }

// Evaluations.cpp:382-504.  Support of every candidate = number of edgel triplets whose reprojection error is < 2 px, for the
// view pairs (1,2) and (1,3).  The pose returned is the candidate maximising min(support21, support31) (first one on ties);
// the index lists hold every candidate that reaches the per-pair maximum.
bool Evaluations::get_Solution_with_Maximal_Support(unsigned n_edgels, float* loc, float* /*tangents*/, float* Kin)
{
  Max_Num_Of_Reproj_Inliers_Views21 = Max_Num_Of_Reproj_Inliers_Views31 = 0;
  Max_Reproj_Inliers_Support_Views21_Index.clear(); Max_Reproj_Inliers_Support_Views31_Index.clear();
  std::vector<unsigned> s21(normalized_t21s.size()), s31(normalized_t21s.size());
  for (size_t c = 0; c < normalized_t21s.size(); c++) {
    unsigned n21 = 0, n31 = 0;
    for (unsigned e = 0; e < n_edgels; e++) {
      const float* g = loc + (size_t)e * 6;
      const mvg::Vec3 g1 = {g[0], g[1], 1.0f}, g2 = {g[2], g[3], 1.0f}, g3 = {g[4], g[5], 1.0f};
      const float rho21 = mvg::depth_rho(g1, g2, normalized_R21s[c], normalized_t21s[c]);
      const float rho31 = mvg::depth_rho(g1, g3, normalized_R31s[c], normalized_t31s[c]);
      if (mvg::reprojection_error_pixels(g1, g2, normalized_R21s[c], normalized_t21s[c], Kin, rho21) < REPROJ_ERROR_INLIER_THRESH) n21++;
      if (mvg::reprojection_error_pixels(g1, g3, normalized_R31s[c], normalized_t31s[c], Kin, rho31) < REPROJ_ERROR_INLIER_THRESH) n31++;
    }
    s21[c] = n21; s31[c] = n31;
    if (n21 > Max_Num_Of_Reproj_Inliers_Views21) Max_Num_Of_Reproj_Inliers_Views21 = n21;
    if (n31 > Max_Num_Of_Reproj_Inliers_Views31) Max_Num_Of_Reproj_Inliers_Views31 = n31;
  }
  if (normalized_t21s.empty()) return false;
  size_t best = 0;
  for (size_t c = 0; c < s21.size(); c++) {
    if (s21[c] == Max_Num_Of_Reproj_Inliers_Views21) Max_Reproj_Inliers_Support_Views21_Index.push_back((int)c);
    if (s31[c] == Max_Num_Of_Reproj_Inliers_Views31) Max_Reproj_Inliers_Support_Views31_Index.push_back((int)c);
    if (std::min(s21[c], s31[c]) > std::min(s21[best], s31[best])) best = c;
  }
  R21_w_Max_Supports = normalized_R21s[best]; t21_w_Max_Supports = normalized_t21s[best];
  R31_w_Max_Supports = normalized_R31s[best]; t31_w_Max_Supports = normalized_t31s[best];
  Best_Candidate_Path_Index = real_track_indices[best];
  return true;
}

void Evaluations::Set_Selected_Solution(complex32* x, int path_index, unsigned support21, unsigned support31)
{
  Convert_Trifocal_Translation(x);
  Convert_Trifocal_Rotation(x);
  R21_w_Max_Supports = normalized_R21; t21_w_Max_Supports = normalized_t21;
  R31_w_Max_Supports = normalized_R31; t31_w_Max_Supports = normalized_t31;
  Max_Num_Of_Reproj_Inliers_Views21 = support21; Max_Num_Of_Reproj_Inliers_Views31 = support31;
  Best_Candidate_Path_Index = path_index;
}

static void split_gt(float GT_Pose[12], mvg::Mat3& R, mvg::Vec3& t)
{
  for (int i = 0; i < 9; i++) R[i] = GT_Pose[i];                 // rows 0-2 of the 4x3 file = R (row-major), row 3 = t
  t = mvg::normalized({GT_Pose[9], GT_Pose[10], GT_Pose[11]});
}

void Evaluations::Measure_Relative_Pose_Error(float GT_Pose21[12], float GT_Pose31[12])
{
  mvg::Mat3 R21, R31; mvg::Vec3 t21, t31;
  split_gt(GT_Pose21, R21, t21); split_gt(GT_Pose31, R31, t31);
  Min_Residual_R21 = mvg::rotation_residual(R21, R21_w_Max_Supports);
  Min_Residual_R31 = mvg::rotation_residual(R31, R31_w_Max_Supports);
  Min_Residual_t21 = mvg::translation_residual(t21, t21_w_Max_Supports);
  Min_Residual_t31 = mvg::translation_residual(t31, t31_w_Max_Supports);
  success_flag = Min_Residual_t21 < TRANSL_RESIDUAL_TOL && Min_Residual_t31 < TRANSL_RESIDUAL_TOL &&
                 Min_Residual_R21 < ROT_RESIDUAL_TOL && Min_Residual_R31 < ROT_RESIDUAL_TOL;
}

void Evaluations::Measure_Relative_Pose_Error_from_All_Real_Sols(float GT_Pose21[12], float GT_Pose31[12], complex32* /*debug*/)
{
  mvg::Mat3 R21, R31; mvg::Vec3 t21, t31;
  split_gt(GT_Pose21, R21, t21); split_gt(GT_Pose31, R31, t31);
  for (size_t si = 0; si < normalized_R21s.size(); si++) {
    const float rR21 = mvg::rotation_residual(R21, normalized_R21s[si]), rR31 = mvg::rotation_residual(R31, normalized_R31s[si]);
    const float rt21 = mvg::translation_residual(t21, normalized_t21s[si]), rt31 = mvg::translation_residual(t31, normalized_t31s[si]);
    if (rR21 < Min_Residual_R21) Min_Residual_R21 = rR21;
    if (rR31 < Min_Residual_R31) Min_Residual_R31 = rR31;
    if (rt21 < Min_Residual_t21) Min_Residual_t21 = rt21;
    if (rt31 < Min_Residual_t31) Min_Residual_t31 = rt31;
    if (rt21 < TRANSL_RESIDUAL_TOL && rt31 < TRANSL_RESIDUAL_TOL && rR21 < ROT_RESIDUAL_TOL && rR31 < ROT_RESIDUAL_TOL) success_flag = true;
  }
}

void Evaluations::Check_Deviations_of_Veridical_Sol_from_GT(complex32* x, float GT_Pose21[12], float GT_Pose31[12])
{
  Convert_Trifocal_Translation(x);
  Convert_Trifocal_Rotation(x);
  mvg::Mat3 R21, R31; mvg::Vec3 t21, t31;
  split_gt(GT_Pose21, R21, t21); split_gt(GT_Pose31, R31, t31);
  std::cout << "GT translation_21 = (" << t21[0] << ", " << t21[1] << ", " << t21[2] << ")\n"
            << "GT translation_31 = (" << t31[0] << ", " << t31[1] << ", " << t31[2] << ")\n"
            << "Sol translation_21 = (" << normalized_t21[0] << ", " << normalized_t21[1] << ", " << normalized_t21[2] << ")\n"
            << "Sol translation_31 = (" << normalized_t31[0] << ", " << normalized_t31[1] << ", " << normalized_t31[2] << ")\n";
  Min_Residual_R21 = mvg::rotation_residual(R21, normalized_R21); Min_Residual_R31 = mvg::rotation_residual(R31, normalized_R31);
  Min_Residual_t21 = mvg::translation_residual(t21, normalized_t21); Min_Residual_t31 = mvg::translation_residual(t31, normalized_t31);
  std::cout << "Residuals in Rotations:    (R21) " << Min_Residual_R21 << " (R31) " << Min_Residual_R31 << "\n"
            << "Residuals in Translations: (t21) " << Min_Residual_t21 << " (t31) " << Min_Residual_t31 << std::endl;
}
