// Data_Reader — problem-definition and RANSAC-dataset text parsers.
// Same public interface and file formats as the reference class (magmaHC/Data_Reader.hpp:32-56, .cpp:37-338;
// formats: SURVEY.md App. A.3), written without MAGMA: `hcb200::complex32` is layout-identical to magmaFloatComplex.
#ifndef HCB200_HOST_DATA_READER_HPP
#define HCB200_HOST_DATA_READER_HPP
#include <array>
#include <string>
#include <vector>

#include "definitions.hpp"

class Data_Reader {
public:
  Data_Reader(std::string problem_dir, std::string ransac_data_dir, int num_of_tracks, int num_of_vars, int num_of_params);

  // ---- minimal-problem definition ------------------------------------------------------------------------
  bool Read_Start_Params(hcb200::complex32*& h_Start_Params);          // 33 lines "re im"; entry 33 := 1
  bool Read_Target_Params(hcb200::complex32*& h_Target_Params);        // (the reference never calls it; kept for parity)
  bool Read_Start_Sols(hcb200::complex32*& h_Start_Sols);              // 312 x 30 lines -> [312][31], entry 30 := 1
  bool Feed_Start_Sols_for_Intermediate_Homotopy(hcb200::complex32*& h_Start_Sols, hcb200::complex32*& h_Homotopy_Sols,
                                                 int RANSAC_Iters_per_GPU);
  // sizes of the caller's index tables (dHdx_Max_Terms * dHdx_Max_Parts * vars^2, dHdt_Max_Terms * dHdt_Max_Parts * vars); once set,
  // the two readers below refuse a file of any other length instead of writing past the table (the reference does not check)
  void Set_Index_Table_Sizes(size_t dHdx_size, size_t dHdt_size) { dHdx_capacity_ = dHdx_size; dHdt_capacity_ = dHdt_size; }
  template <typename T> bool Read_dHdx_Indices(T*& h_dHdx_Index);
  template <typename T> bool Read_dHdt_Indices(T*& h_dHdt_Index);
  template <typename T> bool Read_unified_dHdx_dHdt_Indices(T*& h_unified, T* h_dHdx_Index, T* h_dHdt_Index, int dHdx_size, int dHdt_size);

  // ---- RANSAC data ---------------------------------------------------------------------------------------
  int get_Num_Of_Triplet_Edgels(int tp_index);                         // parses Triplet_Edgels_XXX.txt, returns #lines
  bool Read_Camera_Poses(float Pose21[12], float Pose31[12], int tp_index);
  bool Read_Intrinsic_Matrix(float* h_Intrinsic_Matrix);
  void Read_Triplet_Edgels(float*& Triplet_Edge_Locations, float*& Triplet_Edge_Tangents);   // [E][6] each

  void Print_Out_Target_Params_from_Triplet_Edgels(int sample_index, std::vector<std::array<int, 3>> target_params_match_indices,
                                                   hcb200::complex32* h_Target_Params);

  static std::string padded_index(int index);                          // 7 -> "007"

private:
  size_t dHdx_capacity_ = 0, dHdt_capacity_ = 0;
  std::string problem_dir_, ransac_dir_;
  const int num_of_tracks, num_of_variables, num_of_params;
  std::vector<float> edgel_rows_;                                      // 12 floats per triplet, file order
};
#endif
