// See Data_Reader.hpp.  Every file is read as a whitespace-separated token stream, exactly like the reference's
// `stream >> value` loops, so line breaks and tabs do not matter.
#include "Data_Reader.hpp"

#include <cstdio>
#include <fstream>
#include <iostream>

namespace hcb200 {
void log_info(const std::string& msg) { std::printf("\033[1;32m[INFO] %s\033[0m\n", msg.c_str()); }
void log_error(const std::string& msg) { std::printf("\033[1;31m[ERROR] %s\033[0m\n", msg.c_str()); }
void log_file_error(const std::string& path) { std::printf("\033[1;31m[ERROR] File %s not found!\033[0m\n", path.c_str()); }
}  // namespace hcb200

namespace {
// Reads every token of `path` as T; false when the file cannot be opened.
template <typename T>
bool slurp(const std::string& path, std::vector<T>& out)
{
  std::ifstream in(path);
  if (!in) { hcb200::log_file_error(path); return false; }
  T v;
  while (in >> v) out.push_back(v);
  return true;
}

bool read_complex_list(const std::string& path, hcb200::complex32* dst, int max_entries)
{
  std::vector<float> v;
  if (!slurp(path, v)) return false;
  const int n = std::min<int>((int)v.size() / 2, max_entries);
  for (int i = 0; i < n; i++) dst[i] = hcb200::make_c32(v[2 * i], v[2 * i + 1]);
  return true;
}
}  // namespace

Data_Reader::Data_Reader(std::string problem_dir, std::string ransac_data_dir, int tracks, int vars, int params)
    : problem_dir_(std::move(problem_dir)), ransac_dir_(std::move(ransac_data_dir)),
      num_of_tracks(tracks), num_of_variables(vars), num_of_params(params) {}

std::string Data_Reader::padded_index(int index)
{
  std::string s = std::to_string(index);
  return s.size() >= 3 ? s : std::string(3 - s.size(), '0') + s;
}

bool Data_Reader::Read_Start_Params(hcb200::complex32*& h_Start_Params)
{
  if (!read_complex_list(problem_dir_ + "/start_params.txt", h_Start_Params, num_of_params)) return false;
  h_Start_Params[num_of_params] = hcb200::make_c32(1.0f, 0.0f);
  return true;
}

bool Data_Reader::Read_Target_Params(hcb200::complex32*& h_Target_Params)
{
  if (!read_complex_list(problem_dir_ + "/target_params.txt", h_Target_Params, num_of_params)) return false;
  h_Target_Params[num_of_params] = hcb200::make_c32(1.0f, 0.0f);
  return true;
}

bool Data_Reader::Read_Start_Sols(hcb200::complex32*& h_Start_Sols)
{
  std::vector<float> v;
  if (!slurp(problem_dir_ + "/start_sols.txt", v)) return false;
  const int stride = num_of_variables + 1;
  const int n_sols = std::min<int>((int)v.size() / (2 * num_of_variables), num_of_tracks);
  for (int s = 0; s < n_sols; s++)
    for (int d = 0; d < num_of_variables; d++) {
      const int t = 2 * (s * num_of_variables + d);
      h_Start_Sols[s * stride + d] = hcb200::make_c32(v[t], v[t + 1]);
    }
  for (int s = 0; s < num_of_tracks; s++) h_Start_Sols[s * stride + num_of_variables] = hcb200::make_c32(1.0f, 0.0f);
  return n_sols == num_of_tracks;
}

bool Data_Reader::Feed_Start_Sols_for_Intermediate_Homotopy(hcb200::complex32*& h_Start_Sols, hcb200::complex32*& h_Homotopy_Sols,
                                                             int RANSAC_Iters_per_GPU)
{
  const size_t block = (size_t)num_of_tracks * (num_of_variables + 1);
  for (int ri = 0; ri < RANSAC_Iters_per_GPU; ri++)
    std::copy(h_Start_Sols, h_Start_Sols + block, h_Homotopy_Sols + (size_t)ri * block);
  return true;
}

template <typename T>
bool Data_Reader::Read_dHdx_Indices(T*& h_dHdx_Index)
{
  std::vector<int> v;
  if (!slurp(problem_dir_ + "/dHdx_indx.txt", v)) return false;
  if (dHdx_capacity_ && v.size() != dHdx_capacity_) {       // the caller sized its table from the YAML keys: a file of another length is an error
    hcb200::log_error("dHdx_indx.txt holds " + std::to_string(v.size()) + " indices, the settings file implies " + std::to_string(dHdx_capacity_));
    return false;
  }
  for (size_t i = 0; i < v.size(); i++) h_dHdx_Index[i] = (T)v[i];
  return true;
}

template <typename T>
bool Data_Reader::Read_dHdt_Indices(T*& h_dHdt_Index)
{
  std::vector<int> v;
  if (!slurp(problem_dir_ + "/dHdt_indx.txt", v)) return false;
  if (dHdt_capacity_ && v.size() != dHdt_capacity_) {
    hcb200::log_error("dHdt_indx.txt holds " + std::to_string(v.size()) + " indices, the settings file implies " + std::to_string(dHdt_capacity_));
    return false;
  }
  for (size_t i = 0; i < v.size(); i++) h_dHdt_Index[i] = (T)v[i];
  return true;
}

template <typename T>
bool Data_Reader::Read_unified_dHdx_dHdt_Indices(T*& h_unified, T* h_dHdx_Index, T* h_dHdt_Index, int dHdx_size, int dHdt_size)
{
  std::copy(h_dHdx_Index, h_dHdx_Index + dHdx_size, h_unified);
  std::copy(h_dHdt_Index, h_dHdt_Index + dHdt_size, h_unified + dHdx_size);
  return true;
}

int Data_Reader::get_Num_Of_Triplet_Edgels(int tp_index)
{
  edgel_rows_.clear();
  const std::string path = ransac_dir_ + "/Triplet_Edgels/Triplet_Edgels_" + padded_index(tp_index) + ".txt";
  if (!slurp(path, edgel_rows_)) return 0;
  edgel_rows_.resize(edgel_rows_.size() / 12 * 12);      // complete lines only, like the 12-value extraction loop
  return (int)(edgel_rows_.size() / 12);
}

void Data_Reader::Read_Triplet_Edgels(float*& Triplet_Edge_Locations, float*& Triplet_Edge_Tangents)
{
  // line = x1 y1 tx1 ty1 x2 y2 tx2 ty2 x3 y3 tx3 ty3  ->  locations [x1 y1 x2 y2 x3 y3], tangents likewise
  const size_t n = edgel_rows_.size() / 12;
  for (size_t e = 0; e < n; e++)
    for (int view = 0; view < 3; view++)
      for (int c = 0; c < 2; c++) {
        Triplet_Edge_Locations[e * 6 + 2 * view + c] = edgel_rows_[e * 12 + 4 * view + c];
        Triplet_Edge_Tangents[e * 6 + 2 * view + c] = edgel_rows_[e * 12 + 4 * view + 2 + c];
      }
}

bool Data_Reader::Read_Camera_Poses(float Pose21[12], float Pose31[12], int tp_index)
{
  const std::string idx = padded_index(tp_index);
  const std::string paths[2] = {ransac_dir_ + "/GT_Poses21/GT_Poses21_" + idx + ".txt", ransac_dir_ + "/GT_Poses31/GT_Poses31_" + idx + ".txt"};
  float* dst[2] = {Pose21, Pose31};
  for (int k = 0; k < 2; k++) {
    std::vector<float> v;
    if (!slurp(paths[k], v)) return false;
    for (size_t i = 0; i < v.size() && i < 12; i++) dst[k][i] = v[i];      // rows 0-2 = R (row-major), row 3 = t
  }
  return true;
}

bool Data_Reader::Read_Intrinsic_Matrix(float* h_Intrinsic_Matrix)
{
  std::vector<float> v;
  if (!slurp(ransac_dir_ + "/Intrinsic_Matrix.txt", v)) return false;
  for (size_t i = 0; i < v.size() && i < 9; i++) h_Intrinsic_Matrix[i] = v[i];
  return true;
}

void Data_Reader::Print_Out_Target_Params_from_Triplet_Edgels(int sample_index, std::vector<std::array<int, 3>> picks,
                                                              hcb200::complex32* h_Target_Params)
{
  const std::array<int, 3>& t = picks[sample_index];
  std::cout << "\nPrinting triplet edgel indices: " << t[0] << " " << t[1] << " " << t[2] << "\n"
            << "Converting from triplet edgels to target parameters:\n";
  const hcb200::complex32* p = h_Target_Params + (size_t)sample_index * (num_of_params + 1);
  for (int i = 0; i <= num_of_params; i++) std::printf("(%.10f, %.10f)\n", p[i].x, p[i].y);
}

template bool Data_Reader::Read_dHdx_Indices<int>(int*&);
template bool Data_Reader::Read_dHdt_Indices<int>(int*&);
template bool Data_Reader::Read_unified_dHdx_dHdt_Indices<int>(int*&, int*, int*, int, int);
template bool Data_Reader::Read_dHdx_Indices<char>(char*&);
template bool Data_Reader::Read_dHdt_Indices<char>(char*&);
template bool Data_Reader::Read_unified_dHdx_dHdt_Indices<char>(char*&, char*, char*, int, int);
