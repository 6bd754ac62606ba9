// GPU_HC_Solver — host driver of one RANSAC round on 1..8 B200s.
// Public interface of the reference class (magmaHC/GPU_HC_Solver.hpp:91-120) so the reference's driver code
// (cmd/magmaHC-main.cpp:24-66) works unchanged against it; implemented on the CUDA runtime and the C ABI of
// include/hcb200.h — no MAGMA, no index tables on the device, no pointer arrays.
//
// Multi-GPU = the reference's only parallelism (GPU_HC_Solver.cpp:85-88,390-506): hypotheses are split into contiguous
// shards sub_RANSAC_iters[g] = H/N + (g < H%N), one launch per GPU on its own stream from a single host thread; results
// land directly in the stacked host arrays, and the per-GPU best-pose records (64 B) are reduced on the host.
#ifndef HCB200_HOST_GPU_HC_SOLVER_HPP
#define HCB200_HOST_GPU_HC_SOLVER_HPP
#include <array>
#include <memory>
#include <string>
#include <vector>

#include "Data_Reader.hpp"
#include "Evaluations.hpp"
#include "definitions.hpp"
#include "hcb200.h"
#include "yaml_lite.hpp"

typedef double real_Double_t;

class GPU_HC_Solver {
public:
  //> timers (seconds).  multi_GPUs_time brackets launch -> sync exactly like the reference (GPU_HC_Solver.cpp:384-446);
  //> gpu_time[g] is the CUDA-event time of GPU g's launch (the reference leaves it at 0).
  real_Double_t gpu_time[MAX_NUM_OF_GPUS] = {0.0};
  real_Double_t transfer_h2d_time[MAX_NUM_OF_GPUS] = {0.0};
  real_Double_t transfer_d2h_time[MAX_NUM_OF_GPUS] = {0.0};
  double multi_GPUs_time = 0.0;

  GPU_HC_Solver() {}
  explicit GPU_HC_Solver(YAML::Node Problem_Settings_File);
  ~GPU_HC_Solver();

  bool Read_Problem_Data();
  bool Read_RANSAC_Data(int tp_index);
  void Allocate_Arrays();
  void Prepare_Target_Params(unsigned rand_seed_);
  void Data_Transfer_From_Host_To_Device();
  void Set_CUDA_Stream_Attributes();
  void Set_RANSAC_Abort_Arrays();
  void Solve_by_GPU_HC();
  void Export_Data();
  void Free_Triplet_Edgels_Mem();
  void Free_Arrays_for_Aborting_RANSAC();

  //> one entry per Solve_by_GPU_HC call.  (The reference stores the real count in ..._Inf_Sols and vice versa,
  //> GPU_HC_Solver.cpp:522-524; here every vector holds what its name says and the statistics FILE keeps the reference's
  //> column order  converged <TAB> real <TAB> infinity.)
  std::vector<unsigned> Collect_Num_Of_Coverged_Sols;
  std::vector<unsigned> Collect_Num_Of_Inf_Sols;
  std::vector<unsigned> Collect_Num_Of_Real_Sols;

  // ---- additions for tests / benchmarks / callers that want the raw results --------------------------------
  int Num_Of_RANSAC_Iterations() const { return num_ransac_iters; }
  int Num_Of_Paths() const { return num_ransac_iters * Num_Of_Tracks; }
  const hcb200::complex32* Track_Sols() { Fetch_Results_To_Host(); return h_GPU_HC_Track_Sols_Stack; }     // [H*312][31]
  const bool* Sol_Converge() { Fetch_Results_To_Host(); return h_is_GPU_HC_Sol_Converge_Stack; }
  const bool* Sol_Infinity() { Fetch_Results_To_Host(); return h_is_GPU_HC_Sol_Infinity_Stack; }
  void Fetch_Results_To_Host();      // Lazy_Results: copy every end point and flag back now (no-op when they already are on the host)
  const hcb200::complex32* Target_Params(int gpu_id) { Fetch_Target_Params_To_Host(); return h_Target_Params[gpu_id]; }
  void Fetch_Target_Params_To_Host();    // device-side Prepare_Target_Params: bring the parameters back when the host asks for them
  const int* Picked_Edgels(int gpu_id) const { return h_picked[gpu_id]; }     // [H_g][3] edgel indices drawn by the rand() stream
  const std::vector<std::array<unsigned, 3>>& Per_Hypothesis_Counts() const { return per_hypothesis_counts; }
  const hcb200_best_record& Best_Record() const { return best_record; }                 // path_id is GLOBAL (stacked) numbering
  const std::vector<int>& Found_Path_Ids() const { return found_path_ids; }
  bool Found_Pose() const { return found_pose; }
  const std::array<float, 4>& Pose_Residuals() const { return pose_residuals; }         // R21, R31, t21, t31 vs ground truth
  int Selected_Path() const { return selected_path; }                                   // global path id of the pose with maximal support
  const std::array<unsigned, 2>& Selected_Support() const { return selected_support; }
  int Sub_RANSAC_Iters(int gpu_id) const { return sub_RANSAC_iters[gpu_id]; }
  void Set_Verbose(bool v) { verbose = v; }
  void Set_Pruning(bool on) { prune_paths = on; }       // the reference GPU kernels always prune (…TrunPaths.cu:148-154)

private:
  struct DeviceShard {       // everything GPU g owns
    int device = 0;
    void* stream = nullptr;
    void *ev_start = nullptr, *ev_stop = nullptr;
    float *d_start_sols = nullptr, *d_start_params = nullptr, *d_target = nullptr, *d_diff = nullptr, *d_tracks = nullptr;
    unsigned char *d_conv = nullptr, *d_inf = nullptr, *d_found = nullptr;
    void* d_ws = nullptr;
    float *d_edgels = nullptr, *d_K = nullptr;
    float* d_tangents = nullptr;       // device-side Prepare_Target_Params: edgel tangents and the picked edgel triplets of this shard
    int* d_picked = nullptr;
    unsigned* d_counts = nullptr;      // per-hypothesis (converged, infinity, real)
    int* d_found_index = nullptr;
    hcb200_best_record* d_best = nullptr;
    int* d_support = nullptr;                 // [paths][2] inlier supports from hcb200_score_tracks
    hcb200_best_record* d_score_best = nullptr;
    float* d_refine_sums = nullptr;           // [paths][2] from hcb200_refine_tracks (only with Refine_Iterations > 0)
    int path_offset = 0;     // first path of this shard in the stacked arrays
  };
  DeviceShard shard[MAX_NUM_OF_GPUS];

  // host arrays (pinned where they are DMA targets)
  hcb200::complex32* h_Start_Sols = nullptr;
  hcb200::complex32* h_Start_Params = nullptr;
  hcb200::complex32* h_Target_Params[MAX_NUM_OF_GPUS] = {nullptr};
  hcb200::complex32* h_diffParams[MAX_NUM_OF_GPUS] = {nullptr};
  hcb200::complex32* h_GPU_HC_Track_Sols_Stack = nullptr;
  bool* h_is_GPU_HC_Sol_Converge_Stack = nullptr;
  bool* h_is_GPU_HC_Sol_Infinity_Stack = nullptr;
  bool* h_Found_Trifocal_Sols[MAX_NUM_OF_GPUS] = {nullptr};
  int* h_Trifocal_Sols_Batch_Index[MAX_NUM_OF_GPUS] = {nullptr};
  hcb200_best_record* h_best[MAX_NUM_OF_GPUS] = {nullptr};
  hcb200_best_record* h_score_best[MAX_NUM_OF_GPUS] = {nullptr};
  int* h_dHdx_Index = nullptr;          // parsed for format parity; the device code has the system compiled in
  int* h_dHdt_Index = nullptr;
  float* h_Camera_Intrinsic_Matrix = nullptr;
  float* h_Triplet_Edge_Locations = nullptr;
  float* h_Triplet_Edge_Tangents = nullptr;
  float h_Camera_Pose21[12] = {0}, h_Camera_Pose31[12] = {0};

  std::shared_ptr<Data_Reader> Load_Problem_Data;
  std::shared_ptr<Evaluations> Evaluate_GPUHC_Sols;
  YAML::Node Problem_Setting_YAML_File;
  std::string Problem_File_Path, RANSAC_Data_File_Path, Write_Files_Path;
  std::string HC_problem, HC_print_problem_name, RANSAC_Dataset_Name;
  int GPUHC_Max_Steps = 80, GPUHC_Max_Correction_Steps = 3, GPUHC_delta_t_incremental_steps = 4;
  int Num_Of_Vars = 30, Num_Of_Params = 33, Num_Of_Tracks = 312;
  int dHdx_Max_Terms = 8, dHdx_Max_Parts = 5, dHdt_Max_Terms = 16, dHdt_Max_Parts = 6;
  int dHdx_Index_Size = 0, dHdt_Index_Size = 0;
  bool Abort_RANSAC_by_Good_Sol = false;
  int Num_Of_GPUs = 1, device_count = 0;
  int num_ransac_iters = NUM_OF_RANSAC_ITERATIONS;
  int Num_Of_Triplet_Edgels = 0;
  int sub_RANSAC_iters[MAX_NUM_OF_GPUS] = {0};
  bool arrays_allocated = false, abort_arrays_allocated = false, edgels_allocated = false;
  bool verbose = true, prune_paths = true;
  bool device_scoring = true;       // final support counting + pose selection on the GPU (YAML key Device_Scoring)
  int refine_iterations = 0;        // Newton refinement of converged end points on the GPU before they are copied back
                                    // (YAML key Refine_Iterations; 0 = the reference's behaviour)
  bool device_edgels_allocated = false;
  // Prepare_Target_Params on the device (hcb200_build_target_params): the host keeps the glibc rand() stream and ships 12 bytes per
  // hypothesis instead of 544.  YAML key Device_Target_Params: auto (default: from 2048 hypotheses up) | true | false.
  bool device_target_params = false;
  bool target_params_on_host = true;     // h_Target_Params / h_diffParams hold this round's values
  int* h_picked[MAX_NUM_OF_GPUS] = {nullptr};
  // Round statistics on the device (hcb200_count_solutions; YAML key Device_Statistics, default true): 12 bytes per hypothesis come
  // back instead of the host walking every end point.  The stacked pinned result arrays are allocated lazily, AFTER the tracker
  // launches, so that pinning gigabytes of host memory overlaps the GPU work instead of preceding it.
  bool device_statistics = true;
  // Long paths are parked at 4/5 of the step cap and finished by idle warps (HCB200_FLAG_SPLIT_LONG_PATHS; YAML key Split_Long_Paths,
  // default true): same results, the default round ends 3.6 % sooner.  Costs 20 bytes of device workspace per path.
  bool split_long_paths = true;
  // Early abort across GPUs (YAML key Abort_Across_GPUs, default false = the reference's per-GPU flag, GPU_HC_Solver.cpp:329,402): the first GPU
  // that finds a pose raises the other GPUs' flags through NVLink peer stores (hcb200_track_abort_peers), so the round ends when ANY GPU has one.
  bool abort_across_gpus = false;
  bool result_stacks_allocated = false;
  bool lazy_results = false, results_on_host = false;
  hcb200::complex32 h_selected_track[32];
  unsigned* h_counts[MAX_NUM_OF_GPUS] = {nullptr};
  void Allocate_Result_Stacks();
public:
  double phase_seconds[8] = {0};     // allocate, read, prepare, h2d, solve, d2h, statistics, scoring (last round)
private:
  int device_edgel_capacity = 0;

  std::vector<std::array<unsigned, 3>> per_hypothesis_counts;
  hcb200_best_record best_record{};
  std::vector<int> found_path_ids;
  bool found_pose = false;
  std::array<float, 4> pose_residuals{{100.f, 100.f, 100.f, 100.f}};
  int selected_path = -1;
  std::array<unsigned, 2> selected_support{{0u, 0u}};

  void check_multiGPUs();
  int device_of(int gpu_id) const { return (Num_Of_GPUs == 1) ? SET_GPU_DEVICE_ID : gpu_id; }
};
#endif
