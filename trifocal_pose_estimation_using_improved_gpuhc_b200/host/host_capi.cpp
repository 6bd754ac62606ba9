// C handle API over GPU_HC_Solver (include/hcb200_host.h).
#include <cmath>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>

#include "GPU_HC_Solver.hpp"
#include "hcb200_host.h"

struct hcb200_solver { GPU_HC_Solver* impl; };

extern "C" {

hcb200_solver* hcb200_solver_create(const char* settings_yaml, const char* overrides)
{
  try {
    YAML::Node cfg = YAML::LoadFile(settings_yaml);
    if (overrides) {
      std::istringstream in(overrides);
      std::string kv;
      while (std::getline(in, kv, ';')) {
        const size_t eq = kv.find('=');
        if (eq != std::string::npos) cfg.set(YAML::strip(kv.substr(0, eq)), YAML::strip(kv.substr(eq + 1)));
      }
    }
    hcb200_solver* h = new hcb200_solver;
    h->impl = new GPU_HC_Solver(cfg);
    return h;
  } catch (const std::exception& e) {
    std::cerr << "hcb200_solver_create: " << e.what() << std::endl;
    return nullptr;
  }
}

void hcb200_solver_destroy(hcb200_solver* s) { if (s) { delete s->impl; delete s; } }
int hcb200_solver_allocate(hcb200_solver* s) { s->impl->Allocate_Arrays(); return 0; }
int hcb200_solver_read_problem(hcb200_solver* s) { return s->impl->Read_Problem_Data() ? 0 : 1; }
int hcb200_solver_read_ransac(hcb200_solver* s, int i) { return s->impl->Read_RANSAC_Data(i) ? 0 : 1; }
int hcb200_solver_prepare(hcb200_solver* s, unsigned seed) { s->impl->Prepare_Target_Params(seed); return 0; }
int hcb200_solver_set_abort_arrays(hcb200_solver* s) { s->impl->Set_RANSAC_Abort_Arrays(); return 0; }
int hcb200_solver_h2d(hcb200_solver* s) { s->impl->Data_Transfer_From_Host_To_Device(); s->impl->Set_CUDA_Stream_Attributes(); return 0; }
int hcb200_solver_solve(hcb200_solver* s) { s->impl->Solve_by_GPU_HC(); return 0; }
int hcb200_solver_free_round(hcb200_solver* s) { s->impl->Free_Triplet_Edgels_Mem(); s->impl->Free_Arrays_for_Aborting_RANSAC(); return 0; }
int hcb200_solver_set_pruning(hcb200_solver* s, int on) { s->impl->Set_Pruning(on != 0); return 0; }

int hcb200_solver_num_hypotheses(hcb200_solver* s) { return s->impl->Num_Of_RANSAC_Iterations(); }
double hcb200_solver_kernel_seconds(hcb200_solver* s) { return s->impl->multi_GPUs_time; }

int hcb200_solver_totals(hcb200_solver* s, unsigned out[3])
{
  if (s->impl->Collect_Num_Of_Coverged_Sols.empty()) return 1;
  out[0] = s->impl->Collect_Num_Of_Coverged_Sols.back();
  out[1] = s->impl->Collect_Num_Of_Real_Sols.back();
  out[2] = s->impl->Collect_Num_Of_Inf_Sols.back();
  return 0;
}

int hcb200_solver_per_hypothesis(hcb200_solver* s, unsigned* out)
{
  const auto& c = s->impl->Per_Hypothesis_Counts();
  for (size_t i = 0; i < c.size(); i++) { out[3 * i] = c[i][0]; out[3 * i + 1] = c[i][1]; out[3 * i + 2] = c[i][2]; }
  return (int)c.size();
}

int hcb200_solver_copy_results(hcb200_solver* s, float* tracks, uint8_t* conv, uint8_t* inf)
{
  const size_t n = (size_t)s->impl->Num_Of_Paths();
  if (tracks) std::memcpy(tracks, s->impl->Track_Sols(), n * 31 * sizeof(hcb200::complex32));
  if (conv) std::memcpy(conv, s->impl->Sol_Converge(), n);
  if (inf) std::memcpy(inf, s->impl->Sol_Infinity(), n);
  return 0;
}

int hcb200_solver_copy_target_params(hcb200_solver* s, float* out)
{
  size_t off = 0;
  for (int g = 0; g < MAX_NUM_OF_GPUS; g++) {
    const int H = s->impl->Sub_RANSAC_Iters(g);
    if (!H) continue;
    std::memcpy(out + off, s->impl->Target_Params(g), (size_t)H * 34 * sizeof(hcb200::complex32));
    off += (size_t)H * 34 * 2;
  }
  return 0;
}

int hcb200_solver_best(hcb200_solver* s, hcb200_best_record* rec, int* pose_found, float residuals[4])
{
  if (rec) *rec = s->impl->Best_Record();
  if (pose_found) *pose_found = s->impl->Found_Pose() ? 1 : 0;
  if (residuals) for (int i = 0; i < 4; i++) residuals[i] = s->impl->Pose_Residuals()[i];
  return 0;
}

int hcb200_solver_selected(hcb200_solver* s, int* path_id, unsigned support21_31[2])
{
  if (path_id) *path_id = s->impl->Selected_Path();
  if (support21_31) { support21_31[0] = s->impl->Selected_Support()[0]; support21_31[1] = s->impl->Selected_Support()[1]; }
  return s->impl->Found_Pose() ? 0 : 1;
}

int hcb200_solver_shard_size(hcb200_solver* s, int gpu_id) { return (gpu_id >= 0 && gpu_id < MAX_NUM_OF_GPUS) ? s->impl->Sub_RANSAC_Iters(gpu_id) : 0; }

// ---- file-format access without a device (Data_Reader + the settings reader), used by the CPU-only tests ----------
int hcb200_reader_load(const char* problem_dir, const char* ransac_dir, int dataset_index,
                       float* start_sols, float* start_params, int* dHdx, int* dHdt,
                       int* n_edgels, float* locations, float* tangents, int edgel_capacity,
                       float* pose21, float* pose31, float* K)
{
  Data_Reader rd(problem_dir, ransac_dir, HCB200_NUM_TRACKS, HCB200_NUM_VARS, HCB200_NUM_PARAMS);
  hcb200::complex32* ss = (hcb200::complex32*)start_sols;
  hcb200::complex32* sp = (hcb200::complex32*)start_params;
  if (!rd.Read_Start_Params(sp)) return 1;
  if (!rd.Read_Start_Sols(ss)) return 2;
  if (!rd.Read_dHdx_Indices<int>(dHdx)) return 3;
  if (!rd.Read_dHdt_Indices<int>(dHdt)) return 4;
  const int n = rd.get_Num_Of_Triplet_Edgels(dataset_index);
  if (n_edgels) *n_edgels = n;
  if (n == 0) return 5;
  if (n > edgel_capacity) return 6;
  if (!rd.Read_Camera_Poses(pose21, pose31, dataset_index)) return 7;
  if (!rd.Read_Intrinsic_Matrix(K)) return 8;
  rd.Read_Triplet_Edgels(locations, tangents);
  return 0;
}

// Host-side support of ONE end point with the Evaluations / mvg arithmetic (the reference's host scoring, per-solution):
// returns 1 if the path is a pose candidate and fills the two inlier counts, 0 otherwise.  No device involved.
int hcb200_host_score_track(const float* track31, const float* locations, int n_edgels, const float* K, int* n21, int* n31)
{
  namespace mvg = hcb200::mvg;
  const hcb200::complex32* x = (const hcb200::complex32*)track31;
  for (int vi = 24; vi < 30; vi++) if (!(std::fabs(x[vi].y) < IMAG_PART_TOL)) return 0;
  for (int di = 0; di < 8; di++) if (!(x[di].x >= 0)) return 0;
  const mvg::Vec3 t21 = mvg::normalized({x[18].x, x[19].x, x[20].x}), t31 = mvg::normalized({x[21].x, x[22].x, x[23].x});
  const mvg::Mat3 R21 = mvg::cayley_to_rotation({x[24].x, x[25].x, x[26].x}), R31 = mvg::cayley_to_rotation({x[27].x, x[28].x, x[29].x});
  int c21 = 0, c31 = 0;
  for (int e = 0; e < n_edgels; e++) {
    const float* g = locations + (size_t)e * 6;
    const mvg::Vec3 g1 = {g[0], g[1], 1.0f}, g2 = {g[2], g[3], 1.0f}, g3 = {g[4], g[5], 1.0f};
    if (mvg::reprojection_error_pixels(g1, g2, R21, t21, K, mvg::depth_rho(g1, g2, R21, t21)) < REPROJ_ERROR_INLIER_THRESH) c21++;
    if (mvg::reprojection_error_pixels(g1, g3, R31, t31, K, mvg::depth_rho(g1, g3, R31, t31)) < REPROJ_ERROR_INLIER_THRESH) c31++;
  }
  *n21 = c21; *n31 = c31;
  return 1;
}

int hcb200_settings_lookup(const char* settings_yaml, const char* key, char* out, int capacity)
{
  try {
    YAML::Node cfg = YAML::LoadFile(settings_yaml);
    if (!cfg.has(key)) return 1;
    const std::string v = cfg[key].as<std::string>();
    if ((int)v.size() + 1 > capacity) return 2;
    std::memcpy(out, v.c_str(), v.size() + 1);
    return 0;
  } catch (const std::exception&) { return 3; }
}

}  // extern "C"
