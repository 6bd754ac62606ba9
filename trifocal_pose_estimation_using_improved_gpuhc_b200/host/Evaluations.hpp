// Evaluations — host-side statistics, pose extraction, inlier support and result files.
// Public interface of the reference class (magmaHC/Evaluations.hpp:32-100); file formats of SURVEY.md App. A.4.
// Where the reference has indexing bugs (every candidate is converted from path 0, flags are read
// with a doubled offset) this class implements the INTENDED behaviour — each candidate is converted from its own end point.
#ifndef HCB200_HOST_EVALUATIONS_HPP
#define HCB200_HOST_EVALUATIONS_HPP
#include <array>
#include <fstream>
#include <string>
#include <vector>

#include "definitions.hpp"
#include "mvg.hpp"

class Evaluations {
public:
  Evaluations(std::string Output_Files_Path, std::string GPU_or_CPU, int num_of_tracks, int num_of_vars);
  ~Evaluations();

  void Set_Num_Of_RANSAC_Iterations(int n) { num_of_ransac_iters = n; }      // macro NUM_OF_RANSAC_ITERATIONS in the reference

  void Write_Converged_Sols(hcb200::complex32* h_HC_Track_Sols, bool* h_is_HC_Sol_Converge);
  void Evaluate_HC_Sols(hcb200::complex32* h_HC_Track_Sols, bool* h_is_HC_Sol_Converge, bool* h_is_HC_Sol_Infinity, int ransac_sample_offset);
  void Evaluate_RANSAC_HC_Sols(hcb200::complex32* h_HC_Track_Sols, bool* h_is_HC_Sol_Converge, bool* h_is_HC_Sol_Infinity);
  // the same statistics from per-hypothesis (converged, infinity, real) counts that were reduced on the device (hcb200_count_solutions)
  void Set_RANSAC_HC_Sol_Counts(const unsigned* counts_conv_inf_real, int n_hypotheses);
  void Find_Unique_Sols(hcb200::complex32* h_GPU_HC_Track_Sols, bool* h_is_GPU_HC_Sol_Converge);

  void Convert_Trifocal_Translation(hcb200::complex32* h_GPU_HC_Track_Sols);
  void Convert_Trifocal_Rotation(hcb200::complex32* h_GPU_HC_Track_Sols);
  void Transform_GPUHC_Sols_to_Trifocal_Relative_Pose(hcb200::complex32* h_GPU_HC_Track_Sols, bool* h_is_GPU_HC_Sol_Converge, float* IntrinsicMatrix);
  float get_Rotation_Residual(float* GT_R, std::array<float, 9> Sol_R);
  float get_Translation_Residual(float* GT_Transl, std::array<float, 3> Sol_Transl);
  void Measure_Relative_Pose_Error_from_All_Real_Sols(float GT_Pose21[12], float GT_Pose31[12], hcb200::complex32* h_Debug_Purpose);
  void Measure_Relative_Pose_Error(float GT_Pose21[12], float GT_Pose31[12]);
  bool get_Solution_with_Maximal_Support(unsigned Num_Of_Triplet_Edgels, float* h_Triplet_Edge_Locations, float* h_Triplet_Edge_Tangents, float* K);
  // adopt ONE end point (selected on the device by hcb200_score_tracks) as the pose with maximal support
  void Set_Selected_Solution(hcb200::complex32* track, int path_index, unsigned support21, unsigned support31);
  void Check_Deviations_of_Veridical_Sol_from_GT(hcb200::complex32* h_GPU_HC_Track_Sols, float GT_Pose21[12], float GT_Pose31[12]);
  void Flush_Out_Data();
  void Write_HC_Steps_of_Actual_Solutions(std::vector<int> steps) { for (int s : steps) HC_Actual_Sols_Steps_File << s << "\n"; }

  // statistics of the last Evaluate_RANSAC_HC_Sols
  unsigned Num_Of_Coverged_Sols = 0, Num_Of_Inf_Sols = 0, Num_Of_Real_Sols = 0, Num_Of_Unique_Sols = 0;
  std::vector<std::array<unsigned, 3>> Per_Hypothesis_Counts;         // (converged, infinity, real) per RANSAC iteration
  float Percentage_Of_Convergence = 0, Percentage_Of_Inf_Sols = 0, Percentage_Of_Real_Sols = 0, Percentage_Of_Unique_Sols = 0;
  float Min_Residual_R21 = 100, Min_Residual_R31 = 100, Min_Residual_t21 = 100, Min_Residual_t31 = 100;
  bool success_flag = false;
  std::vector<int> HC_steps_of_actual_solutions;
  unsigned Max_Num_Of_Reproj_Inliers_Views21 = 0, Max_Num_Of_Reproj_Inliers_Views31 = 0;
  std::array<float, 9> R21_w_Max_Supports{}, R31_w_Max_Supports{};
  std::array<float, 3> t21_w_Max_Supports{}, t31_w_Max_Supports{};
  std::vector<int> Max_Reproj_Inliers_Support_Views21_Index, Max_Reproj_Inliers_Support_Views31_Index;
  int Best_Candidate_Path_Index = -1;                                  // hypothesis*312 + track of the selected pose

  // candidates collected by Transform_GPUHC_Sols_to_Trifocal_Relative_Pose
  std::vector<std::array<float, 3>> normalized_t21s, normalized_t31s;
  std::vector<std::array<float, 9>> normalized_R21s, normalized_R31s, F21s, F31s;
  std::vector<int> real_track_indices;

private:
  std::string WRITE_FILES_PATH, evaluate_GPUHC_or_CPUHC;
  std::ofstream HC_Track_Sols_File, HC_Actual_Sols_Steps_File;
  const int num_of_tracks, num_of_variables;
  int num_of_ransac_iters = NUM_OF_RANSAC_ITERATIONS;
  float K[9] = {0};
  std::array<float, 3> normalized_t21{}, normalized_t31{}, raw_t21{}, raw_t31{};
  std::array<float, 9> normalized_R21{}, normalized_R31{};
  std::vector<int> Unique_Sols_Index;
};
#endif
