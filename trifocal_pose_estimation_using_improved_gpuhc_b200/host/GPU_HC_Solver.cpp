// See GPU_HC_Solver.hpp.  Flow of one round (same call order as the reference driver, cmd/magmaHC-main.cpp:32-66):
//   ctor -> Allocate_Arrays -> Read_Problem_Data -> Read_RANSAC_Data(i) -> Prepare_Target_Params(seed)
//        -> Set_RANSAC_Abort_Arrays -> Data_Transfer_From_Host_To_Device -> Set_CUDA_Stream_Attributes -> Solve_by_GPU_HC
//        -> Free_Triplet_Edgels_Mem -> Free_Arrays_for_Aborting_RANSAC
#include "GPU_HC_Solver.hpp"

#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

using hcb200::complex32;

#define HC_CUDA(call)                                                                                       \
  do {                                                                                                      \
    cudaError_t e_ = (call);                                                                                \
    if (e_ != cudaSuccess) {                                                                                \
      std::fprintf(stderr, "\033[1;31m[CUDA ERROR] %s:%d %s -> %s\033[0m\n", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      std::exit(2);          /* no CPU fallback: a failing device call is fatal (the reference only prints) */                     \
    }                                                                                                       \
  } while (0)

namespace {
// Every public method that walks over the shards leaves the caller's current device as it found it (the reference's
// magma_setdevice loops do not, GPU_HC_Solver.cpp:392; a library living inside someone else's process should).
struct DeviceGuard {
  int saved = -1;
  DeviceGuard() { if (cudaGetDevice(&saved) != cudaSuccess) saved = -1; }
  ~DeviceGuard() { if (saved >= 0) cudaSetDevice(saved); }
};
}  // namespace

static double wall_seconds()
{ return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

GPU_HC_Solver::GPU_HC_Solver(YAML::Node cfg) : Problem_Setting_YAML_File(cfg)
{
  DeviceGuard keep_callers_device;
  HC_problem                      = cfg["problem_name"].as<std::string>();
  HC_print_problem_name           = cfg["problem_print_out_name"].as<std::string>();
  GPUHC_Max_Steps                 = cfg["GPUHC_Max_Steps"].as<int>();
  GPUHC_Max_Correction_Steps      = cfg["GPUHC_Max_Correction_Steps"].as<int>();
  GPUHC_delta_t_incremental_steps = cfg["GPUHC_Num_Of_Steps_to_Increase_Delta_t"].as<int>();
  Num_Of_Vars                     = cfg["Num_Of_Vars"].as<int>();
  Num_Of_Params                   = cfg["Num_Of_Params"].as<int>();
  Num_Of_Tracks                   = cfg["Num_Of_Tracks"].as<int>();
  dHdx_Max_Terms                  = cfg["dHdx_Max_Terms"].as<int>();
  dHdx_Max_Parts                  = cfg["dHdx_Max_Parts"].as<int>();
  dHdt_Max_Terms                  = cfg["dHdt_Max_Terms"].as<int>();
  dHdt_Max_Parts                  = cfg["dHdt_Max_Parts"].as<int>();
  Abort_RANSAC_by_Good_Sol        = cfg["Abort_RANSAC_by_Good_Sol"].as<bool>();
  RANSAC_Dataset_Name             = cfg["RANSAC_Dataset"].as<std::string>();
  Num_Of_GPUs                     = cfg["Num_Of_GPUs"].as<int>();
  // optional keys (not in the reference's file): iteration count (a macro there) and the tree root
  num_ransac_iters                = cfg.as_or<int>("Num_Of_RANSAC_Iterations", NUM_OF_RANSAC_ITERATIONS);
  const std::string root          = cfg.as_or<std::string>("Repo_Root", "../../");
  verbose                         = cfg.as_or<bool>("Verbose", true);
  device_scoring                  = cfg.as_or<bool>("Device_Scoring", true);
  refine_iterations               = cfg.as_or<int>("Refine_Iterations", 0);
  device_statistics               = cfg.as_or<bool>("Device_Statistics", true);
  split_long_paths                = cfg.as_or<bool>("Split_Long_Paths", true);
  abort_across_gpus               = cfg.as_or<bool>("Abort_Across_GPUs", false);
  // Lazy_Results: auto (default) | true | false.  When statistics and scoring run on the GPU the host needs 12 bytes per hypothesis and one
  // 128-byte record per GPU; the 248 bytes per path of end points are then copied back only when somebody asks for them (Track_Sols(),
  // Sol_Converge(), Sol_Infinity(), hcb200_solver_copy_results).  auto = lazy from 2048 hypotheses up.
  const std::string lz            = cfg.as_or<std::string>("Lazy_Results", "auto");
  lazy_results                    = device_statistics && device_scoring && ((lz == "true") || (lz == "auto" && num_ransac_iters >= 2048));
  const std::string dtp           = cfg.as_or<std::string>("Device_Target_Params", "auto");
  device_target_params            = (dtp == "true") || (dtp == "auto" && num_ransac_iters >= 2048);

  if (HC_problem != "trifocal_2op1p_30x30" || Num_Of_Vars != HCB200_NUM_VARS || Num_Of_Params != HCB200_NUM_PARAMS ||
      Num_Of_Tracks != HCB200_NUM_TRACKS) {
    hcb200::log_error("this build has the trifocal_2op1p_30x30 system (30 vars, 33 params, 312 tracks) compiled in");
    std::exit(1);
  }

  HC_CUDA(cudaGetDeviceCount(&device_count));
  check_multiGPUs();
  for (int g = 0; g < Num_Of_GPUs; g++) {
    shard[g].device = device_of(g);
    HC_CUDA(cudaSetDevice(shard[g].device));
    cudaStream_t s; cudaEvent_t a, b;
    HC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    HC_CUDA(cudaEventCreate(&a)); HC_CUDA(cudaEventCreate(&b));
    shard[g].stream = s; shard[g].ev_start = a; shard[g].ev_stop = b;
  }
  int offset = 0;
  for (int g = 0; g < Num_Of_GPUs; g++) {        // GPU_HC_Solver.cpp:85-88
    sub_RANSAC_iters[g] = num_ransac_iters / Num_Of_GPUs + ((g < (num_ransac_iters % Num_Of_GPUs)) ? 1 : 0);
    shard[g].path_offset = offset * Num_Of_Tracks;
    offset += sub_RANSAC_iters[g];
    if (verbose) std::printf("GPU %2d computes %2d RANSAC iterations\n", g, sub_RANSAC_iters[g]);
  }
  dHdx_Index_Size = Num_Of_Vars * Num_Of_Vars * dHdx_Max_Terms * dHdx_Max_Parts;
  dHdt_Index_Size = Num_Of_Vars * dHdt_Max_Terms * dHdt_Max_Parts;

  Problem_File_Path     = root + "problems/" + HC_problem;
  RANSAC_Data_File_Path = root + "RANSAC_Data/" + HC_problem + "/" + RANSAC_Dataset_Name;
  Write_Files_Path      = root + WRITE_FILES_FOLDER;
  Evaluate_GPUHC_Sols = std::make_shared<Evaluations>(Write_Files_Path, "GPU-HC", Num_Of_Tracks, Num_Of_Vars);
  Evaluate_GPUHC_Sols->Set_Num_Of_RANSAC_Iterations(num_ransac_iters);
}

void GPU_HC_Solver::check_multiGPUs()
{
  if (Num_Of_GPUs == 1 && verbose) hcb200::log_info("Only 1 GPU is used. Device ID = " + std::to_string(SET_GPU_DEVICE_ID));
  if (Num_Of_GPUs < 1 || Num_Of_GPUs > MAX_NUM_OF_GPUS) {
    hcb200::log_error("Requested GPUs larger than MAX_NUM_OF_GPUS");
    std::printf("\033[1;31m[Requested GPUs] %d\t[Max GPUs] %d\033[0m\n", Num_Of_GPUs, MAX_NUM_OF_GPUS);
    std::exit(1);
  }
  if (Num_Of_GPUs > device_count) {
    hcb200::log_error("Not enough GPUs");
    std::printf("\033[1;31m[Requested GPUs] %d\t[Available GPUs] %d\033[0m\n", Num_Of_GPUs, device_count);
    std::exit(1);
  }
}

void GPU_HC_Solver::Allocate_Arrays()
{
  if (arrays_allocated) return;          // sizes are fixed by the settings file: a second call (one per round in some drivers) is a no-op
  DeviceGuard keep_callers_device;
  const size_t V1 = Num_Of_Vars + 1, P1 = Num_Of_Params + 1, n_paths = (size_t)Num_Of_Paths();
  h_Start_Sols   = (complex32*)std::malloc(sizeof(complex32) * Num_Of_Tracks * V1);
  h_Start_Params = (complex32*)std::malloc(sizeof(complex32) * P1);
  h_dHdx_Index = new int[dHdx_Index_Size];
  h_dHdt_Index = new int[dHdt_Index_Size];
  h_Camera_Intrinsic_Matrix = new float[9];
  const double t_alloc = wall_seconds();
  (void)n_paths;                       // the stacked result arrays are pinned lazily (Allocate_Result_Stacks), while the GPUs track
  for (int g = 0; g < Num_Of_GPUs; g++) {
    DeviceShard& d = shard[g];
    const size_t H = sub_RANSAC_iters[g], paths = H * Num_Of_Tracks;
    HC_CUDA(cudaSetDevice(d.device));
    HC_CUDA(cudaMallocHost((void**)&h_Target_Params[g], sizeof(complex32) * P1 * (H ? H : 1)));
    HC_CUDA(cudaMallocHost((void**)&h_diffParams[g], sizeof(complex32) * P1 * (H ? H : 1)));
    HC_CUDA(cudaMallocHost((void**)&h_picked[g], sizeof(int) * 3 * (H ? H : 1)));
    HC_CUDA(cudaMallocHost((void**)&h_counts[g], sizeof(unsigned) * 3 * (H ? H : 1)));
    HC_CUDA(cudaMalloc((void**)&d.d_counts, sizeof(unsigned) * 3 * (H ? H : 1)));
    if (device_target_params) HC_CUDA(cudaMalloc((void**)&d.d_picked, sizeof(int) * 3 * (H ? H : 1)));
    HC_CUDA(cudaMalloc((void**)&d.d_start_sols, sizeof(complex32) * Num_Of_Tracks * V1));
    HC_CUDA(cudaMalloc((void**)&d.d_start_params, sizeof(complex32) * P1));
    HC_CUDA(cudaMalloc((void**)&d.d_target, sizeof(complex32) * P1 * (H ? H : 1)));
    HC_CUDA(cudaMalloc((void**)&d.d_diff, sizeof(complex32) * P1 * (H ? H : 1)));
    HC_CUDA(cudaMalloc((void**)&d.d_tracks, sizeof(complex32) * V1 * (paths ? paths : 1)));
    HC_CUDA(cudaMalloc((void**)&d.d_conv, paths ? paths : 1));
    HC_CUDA(cudaMalloc((void**)&d.d_inf, paths ? paths : 1));
    HC_CUDA(cudaMalloc(&d.d_ws, (split_long_paths && H <= HCB200_SPLIT_MAX_HYPOTHESES) ? hcb200_workspace_bytes_for(H) : hcb200_workspace_bytes()));
    // load the tracker kernels on this device now (CUDA loads modules lazily): the reference's timed region — launch to
    // sync, GPU_HC_Solver.cpp:384-446 — would otherwise include a one-off module load as long as the round itself
    { int regs = 0; hcb200_kernel_info(0, &regs, nullptr, nullptr, nullptr, nullptr); hcb200_kernel_info(1, &regs, nullptr, nullptr, nullptr, nullptr); }
    HC_CUDA(cudaMalloc((void**)&d.d_support, 2 * sizeof(int) * (paths ? paths : 1)));
    HC_CUDA(cudaMalloc((void**)&d.d_score_best, sizeof(hcb200_best_record)));
    HC_CUDA(cudaMallocHost((void**)&h_score_best[g], sizeof(hcb200_best_record)));
  }
  arrays_allocated = true;
  phase_seconds[0] = wall_seconds() - t_alloc;
}

// Lazy results: bring every end point and flag to the stacked host arrays now (all GPUs copy at the same time).
void GPU_HC_Solver::Fetch_Results_To_Host()
{
  if (results_on_host) return;
  DeviceGuard keep_callers_device;
  Allocate_Result_Stacks();
  const size_t V1 = Num_Of_Vars + 1;
  for (int g = 0; g < Num_Of_GPUs; g++) {
    DeviceShard& d = shard[g];
    const size_t paths = (size_t)sub_RANSAC_iters[g] * Num_Of_Tracks;
    if (!paths) continue;
    cudaStream_t s = (cudaStream_t)d.stream;
    HC_CUDA(cudaSetDevice(d.device));
    HC_CUDA(cudaMemcpyAsync(h_GPU_HC_Track_Sols_Stack + (size_t)d.path_offset * V1, d.d_tracks, sizeof(complex32) * V1 * paths, cudaMemcpyDeviceToHost, s));
    HC_CUDA(cudaMemcpyAsync(h_is_GPU_HC_Sol_Converge_Stack + d.path_offset, d.d_conv, paths, cudaMemcpyDeviceToHost, s));
    HC_CUDA(cudaMemcpyAsync(h_is_GPU_HC_Sol_Infinity_Stack + d.path_offset, d.d_inf, paths, cudaMemcpyDeviceToHost, s));
  }
  for (int g = 0; g < Num_Of_GPUs; g++) {
    if (!sub_RANSAC_iters[g]) continue;
    HC_CUDA(cudaSetDevice(shard[g].device));
    HC_CUDA(cudaStreamSynchronize((cudaStream_t)shard[g].stream));
  }
  results_on_host = true;
}

void GPU_HC_Solver::Allocate_Result_Stacks()
{
  if (result_stacks_allocated) return;
  const size_t V1 = Num_Of_Vars + 1, n_paths = (size_t)Num_Of_Paths();
  HC_CUDA(cudaMallocHost((void**)&h_GPU_HC_Track_Sols_Stack, sizeof(complex32) * n_paths * V1));
  HC_CUDA(cudaMallocHost((void**)&h_is_GPU_HC_Sol_Converge_Stack, n_paths * sizeof(bool)));
  HC_CUDA(cudaMallocHost((void**)&h_is_GPU_HC_Sol_Infinity_Stack, n_paths * sizeof(bool)));
  result_stacks_allocated = true;
}

bool GPU_HC_Solver::Read_Problem_Data()
{
  Load_Problem_Data = std::make_shared<Data_Reader>(Problem_File_Path, RANSAC_Data_File_Path, Num_Of_Tracks, Num_Of_Vars, Num_Of_Params);
  Load_Problem_Data->Set_Index_Table_Sizes((size_t)dHdx_Index_Size, (size_t)dHdt_Index_Size);
  if (!Load_Problem_Data->Read_Start_Params(h_Start_Params)) { hcb200::log_error("Start Parameters not loaded successfully!"); return false; }
  if (!Load_Problem_Data->Read_Start_Sols(h_Start_Sols)) { hcb200::log_error("Start Solutions not loaded successfully!"); return false; }
  // The evaluation-index tables are part of the problem definition the reference ships; the device code was generated from
  // them (codegen/gen_eval.py).  They are parsed here so a missing/short file is reported exactly like in the reference.
  if (!Load_Problem_Data->Read_dHdx_Indices<int>(h_dHdx_Index)) { hcb200::log_error("dH/dx Evaluation Indices not loaded successfully!"); return false; }
  if (!Load_Problem_Data->Read_dHdt_Indices<int>(h_dHdt_Index)) { hcb200::log_error("dH/dt Evaluation Indices not loaded successfully!"); return false; }
  return true;
}

bool GPU_HC_Solver::Read_RANSAC_Data(int tp_index)
{
  Num_Of_Triplet_Edgels = Load_Problem_Data->get_Num_Of_Triplet_Edgels(tp_index);
  if (Num_Of_Triplet_Edgels == 0) return false;
  h_Triplet_Edge_Locations = new float[(size_t)Num_Of_Triplet_Edgels * 6];
  h_Triplet_Edge_Tangents  = new float[(size_t)Num_Of_Triplet_Edgels * 6];
  edgels_allocated = true;
  if (!Load_Problem_Data->Read_Camera_Poses(h_Camera_Pose21, h_Camera_Pose31, tp_index)) { hcb200::log_error("Camera Extrinsic Matrices not loaded successfully!"); return false; }
  if (!Load_Problem_Data->Read_Intrinsic_Matrix(h_Camera_Intrinsic_Matrix)) { hcb200::log_error("Camera Intrinsic Matrices not loaded successfully!"); return false; }
  Load_Problem_Data->Read_Triplet_Edgels(h_Triplet_Edge_Locations, h_Triplet_Edge_Tangents);
  return true;
}

// Hypothesis sampling + target parameters (reference GPU_HC_Solver.cpp:252-306): ONE rand() stream consumed in GPU-major
// order, so the hypothesis sequence does not depend on Num_Of_GPUs.  Parameter layout: SURVEY.md App. A.2.
void GPU_HC_Solver::Prepare_Target_Params(unsigned rand_seed_)
{
  std::srand(rand_seed_);
  const int P1 = Num_Of_Params + 1;
  for (int g = 0; g < Num_Of_GPUs; g++) {
    for (int ti = 0; ti < sub_RANSAC_iters[g]; ti++) {
      unsigned e[3];
      do { for (int r = 0; r < 3; r++) e[r] = std::rand() % Num_Of_Triplet_Edgels; }
      while (!(e[0] != e[1] && e[1] != e[2]));          // the reference never tests e0 != e2
      for (int r = 0; r < 3; r++) h_picked[g][(size_t)ti * 3 + r] = (int)e[r];
      if (device_target_params) continue;               // the parameters themselves are gathered on the device (Data_Transfer_…)
      complex32* T = h_Target_Params[g] + (size_t)ti * P1;
      complex32* D = h_diffParams[g] + (size_t)ti * P1;
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 6; j++) T[i * 6 + j] = hcb200::make_c32(h_Triplet_Edge_Locations[(size_t)e[i] * 6 + j], 0.0f);
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 6; j++) T[18 + i * 6 + j] = hcb200::make_c32(h_Triplet_Edge_Tangents[(size_t)e[i] * 6 + j], 0.0f);
      T[30] = hcb200::make_c32(1.0f, 0.0f);
      T[31] = hcb200::make_c32(0.5f, 0.0f);
      T[32] = hcb200::make_c32(1.0f, 0.0f);
      T[33] = hcb200::make_c32(1.0f, 0.0f);
      for (int i = 0; i < P1; i++) D[i] = hcb200::make_c32(T[i].x - h_Start_Params[i].x, T[i].y - h_Start_Params[i].y);
    }
  }
  target_params_on_host = !device_target_params;
}

// Target parameters of GPU g as the host sees them; in device mode they are fetched from the GPU on first use.
void GPU_HC_Solver::Fetch_Target_Params_To_Host()
{
  if (target_params_on_host) return;
  DeviceGuard keep_callers_device;
  const size_t P1 = Num_Of_Params + 1;
  for (int g = 0; g < Num_Of_GPUs; g++) {
    const size_t H = sub_RANSAC_iters[g];
    if (!H) continue;
    HC_CUDA(cudaSetDevice(shard[g].device));
    HC_CUDA(cudaMemcpyAsync(h_Target_Params[g], shard[g].d_target, sizeof(complex32) * P1 * H, cudaMemcpyDeviceToHost, (cudaStream_t)shard[g].stream));
    HC_CUDA(cudaMemcpyAsync(h_diffParams[g], shard[g].d_diff, sizeof(complex32) * P1 * H, cudaMemcpyDeviceToHost, (cudaStream_t)shard[g].stream));
  }
  for (int g = 0; g < Num_Of_GPUs; g++) {
    if (!sub_RANSAC_iters[g]) continue;
    HC_CUDA(cudaSetDevice(shard[g].device));
    HC_CUDA(cudaStreamSynchronize((cudaStream_t)shard[g].stream));
  }
  target_params_on_host = true;
}

void GPU_HC_Solver::Set_RANSAC_Abort_Arrays()
{
  DeviceGuard keep_callers_device;
  if (!Abort_RANSAC_by_Good_Sol) return;
  for (int g = 0; g < Num_Of_GPUs; g++) {
    DeviceShard& d = shard[g];
    const size_t paths = (size_t)sub_RANSAC_iters[g] * Num_Of_Tracks;
    HC_CUDA(cudaSetDevice(d.device));
    h_Found_Trifocal_Sols[g] = new bool[1];
    h_Found_Trifocal_Sols[g][0] = false;
    h_Trifocal_Sols_Batch_Index[g] = new int[paths ? paths : 1];
    for (size_t i = 0; i < paths; i++) h_Trifocal_Sols_Batch_Index[g][i] = -1;
    h_best[g] = new hcb200_best_record();
    HC_CUDA(cudaMalloc((void**)&d.d_found, sizeof(bool)));
    HC_CUDA(cudaMalloc((void**)&d.d_found_index, (paths ? paths : 1) * sizeof(int)));
    HC_CUDA(cudaMalloc((void**)&d.d_best, sizeof(hcb200_best_record)));
  }
  if (abort_across_gpus && Num_Of_GPUs > 1) {             // every GPU will store into every other GPU's flag (NVLink peer mappings)
    for (int g = 0; g < Num_Of_GPUs; g++)
      for (int q = 0; q < Num_Of_GPUs; q++) {
        const int rc = hcb200_enable_peer_access(shard[g].device, shard[q].device);
        if (rc != 0) { std::fprintf(stderr, "[ERROR] Abort_Across_GPUs: GPU %d cannot reach GPU %d: %s\n", g, q, hcb200_error_string(rc)); std::exit(2); }
      }
  }
  abort_arrays_allocated = true;
}

void GPU_HC_Solver::Data_Transfer_From_Host_To_Device()
{
  DeviceGuard keep_callers_device;
  const size_t V1 = Num_Of_Vars + 1, P1 = Num_Of_Params + 1;
  const double t0 = wall_seconds();
  for (int g = 0; g < Num_Of_GPUs; g++) {                  // enqueue every copy on every GPU first ...
    DeviceShard& d = shard[g];
    const size_t H = sub_RANSAC_iters[g], paths = H * Num_Of_Tracks;
    cudaStream_t s = (cudaStream_t)d.stream;
    HC_CUDA(cudaSetDevice(d.device));
    HC_CUDA(cudaMemcpyAsync(d.d_start_sols, h_Start_Sols, sizeof(complex32) * Num_Of_Tracks * V1, cudaMemcpyHostToDevice, s));
    HC_CUDA(cudaMemcpyAsync(d.d_start_params, h_Start_Params, sizeof(complex32) * P1, cudaMemcpyHostToDevice, s));
    if (Abort_RANSAC_by_Good_Sol || device_scoring || device_target_params) {   // the edgel triplets feed the abort test, the final scoring, the parameter gather
      if (d.d_edgels && device_edgel_capacity < Num_Of_Triplet_Edgels) {        // a larger dataset file than the buffers were sized for
        cudaFree(d.d_edgels); cudaFree(d.d_tangents); d.d_edgels = d.d_tangents = nullptr;
      }
      if (!d.d_edgels) {
        if (!d.d_K) HC_CUDA(cudaMalloc((void**)&d.d_K, 9 * sizeof(float)));
        HC_CUDA(cudaMalloc((void**)&d.d_edgels, (size_t)Num_Of_Triplet_Edgels * 6 * sizeof(float)));
        HC_CUDA(cudaMalloc((void**)&d.d_tangents, (size_t)Num_Of_Triplet_Edgels * 6 * sizeof(float)));
        device_edgels_allocated = true;
      }
      HC_CUDA(cudaMemcpyAsync(d.d_edgels, h_Triplet_Edge_Locations, (size_t)Num_Of_Triplet_Edgels * 6 * sizeof(float), cudaMemcpyHostToDevice, s));
      HC_CUDA(cudaMemcpyAsync(d.d_K, h_Camera_Intrinsic_Matrix, 9 * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    if (H && !device_target_params) {
      HC_CUDA(cudaMemcpyAsync(d.d_target, h_Target_Params[g], sizeof(complex32) * P1 * H, cudaMemcpyHostToDevice, s));
      HC_CUDA(cudaMemcpyAsync(d.d_diff, h_diffParams[g], sizeof(complex32) * P1 * H, cudaMemcpyHostToDevice, s));
    } else if (H) {       // 12 bytes per hypothesis go over PCIe; the 34 parameters and target - start are gathered on the GPU
      HC_CUDA(cudaMemcpyAsync(d.d_tangents, h_Triplet_Edge_Tangents, (size_t)Num_Of_Triplet_Edgels * 6 * sizeof(float), cudaMemcpyHostToDevice, s));
      HC_CUDA(cudaMemcpyAsync(d.d_picked, h_picked[g], sizeof(int) * 3 * H, cudaMemcpyHostToDevice, s));
      const int rc = hcb200_build_target_params(d.stream, (int)H, d.d_picked, Num_Of_Triplet_Edgels, d.d_edgels, d.d_tangents,
                                                (const float*)d.d_start_params, (float*)d.d_target, (float*)d.d_diff);
      if (rc != 0) { std::fprintf(stderr, "[ERROR] target-parameter launch failed on GPU %d: %s\n", g, hcb200_error_string(rc)); std::exit(2); }
    }
    if (Abort_RANSAC_by_Good_Sol) {
      if (paths) HC_CUDA(cudaMemcpyAsync(d.d_found_index, h_Trifocal_Sols_Batch_Index[g], paths * sizeof(int), cudaMemcpyHostToDevice, s));
      HC_CUDA(cudaMemcpyAsync(d.d_found, h_Found_Trifocal_Sols[g], sizeof(bool), cudaMemcpyHostToDevice, s));
    }
  }
  device_edgel_capacity = std::max(device_edgel_capacity, Num_Of_Triplet_Edgels);
  for (int g = 0; g < Num_Of_GPUs; g++) {                  // ... then wait for all of them
    HC_CUDA(cudaSetDevice(shard[g].device));
    HC_CUDA(cudaStreamSynchronize((cudaStream_t)shard[g].stream));
    transfer_h2d_time[g] = wall_seconds() - t0;
  }
  phase_seconds[3] = wall_seconds() - t0;
}

// The reference pins its 152 KB index table in L2 here (GPU_HC_Solver.cpp:364-378).  There is no table any more: the
// operand words live in shared memory and the system is compiled in, so nothing needs a persisting window.
void GPU_HC_Solver::Set_CUDA_Stream_Attributes() {}

void GPU_HC_Solver::Solve_by_GPU_HC()
{
  DeviceGuard keep_callers_device;
  if (verbose) std::cout << "GPU computing ..." << std::endl << std::endl;
  const unsigned prune_flag = prune_paths ? HCB200_FLAG_PRUNE_PATHS : 0u;
  const size_t V1 = Num_Of_Vars + 1;

  multi_GPUs_time = wall_seconds();
  for (int g = 0; g < Num_Of_GPUs; g++) {               // one asynchronous launch per GPU (GPU_HC_Solver.cpp:390-436)
    DeviceShard& d = shard[g];
    if (!sub_RANSAC_iters[g]) continue;
    HC_CUDA(cudaSetDevice(d.device));
    HC_CUDA(cudaEventRecord((cudaEvent_t)d.ev_start, (cudaStream_t)d.stream));
    int rc;
    const unsigned flags = prune_flag | ((split_long_paths && sub_RANSAC_iters[g] <= HCB200_SPLIT_MAX_HYPOTHESES) ? HCB200_FLAG_SPLIT_LONG_PATHS : 0u);
    if (Abort_RANSAC_by_Good_Sol && abort_across_gpus && Num_Of_GPUs > 1) {
      uint8_t* peers[MAX_NUM_OF_GPUS];
      int n_peers = 0;
      for (int q = 0; q < Num_Of_GPUs; q++)
        if (q != g && sub_RANSAC_iters[q]) peers[n_peers++] = shard[q].d_found;
      rc = hcb200_track_abort_peers(d.stream, sub_RANSAC_iters[g], Num_Of_Triplet_Edgels, GPUHC_Max_Steps, GPUHC_Max_Correction_Steps,
                                    GPUHC_delta_t_incremental_steps, flags, d.d_start_sols, d.d_start_params, d.d_target, d.d_diff,
                                    d.d_edgels, d.d_K, d.d_tracks, d.d_conv, d.d_inf, d.d_found, d.d_found_index, d.d_best, nullptr, d.d_ws,
                                    peers, n_peers);
    } else if (Abort_RANSAC_by_Good_Sol)
      rc = hcb200_track_abort(d.stream, sub_RANSAC_iters[g], Num_Of_Triplet_Edgels, GPUHC_Max_Steps, GPUHC_Max_Correction_Steps,
                              GPUHC_delta_t_incremental_steps, flags, d.d_start_sols, d.d_start_params, d.d_target, d.d_diff,
                              d.d_edgels, d.d_K, d.d_tracks, d.d_conv, d.d_inf, d.d_found, d.d_found_index, d.d_best, nullptr, d.d_ws);
    else
      rc = hcb200_track(d.stream, sub_RANSAC_iters[g], GPUHC_Max_Steps, GPUHC_Max_Correction_Steps, GPUHC_delta_t_incremental_steps,
                        flags, d.d_start_sols, d.d_start_params, d.d_target, d.d_diff, d.d_tracks, d.d_conv, d.d_inf, nullptr, d.d_ws);
    if (rc != 0) { std::fprintf(stderr, "[ERROR] tracker launch failed on GPU %d: %s\n", g, hcb200_error_string(rc)); std::exit(2); }
    HC_CUDA(cudaEventRecord((cudaEvent_t)d.ev_stop, (cudaStream_t)d.stream));
  }
  if (!lazy_results) Allocate_Result_Stacks();          // first round only: pin the stacked host arrays while the GPUs are tracking
  for (int g = 0; g < Num_Of_GPUs; g++) {               // GPU_HC_Solver.cpp:440-444
    HC_CUDA(cudaSetDevice(shard[g].device));
    HC_CUDA(cudaStreamSynchronize((cudaStream_t)shard[g].stream));
  }
  multi_GPUs_time = wall_seconds() - multi_GPUs_time;
  phase_seconds[4] = multi_GPUs_time;

  // optional polish of the converged end points (outside the reference's timed region; Evaluations::Find_Unique_Sols compares
  // end points at DUPLICATE_SOL_DIFF_TOL = 1e-4, which raw single-precision end points of the same root do not always meet)
  if (refine_iterations > 0) {
    for (int g = 0; g < Num_Of_GPUs; g++) {
      DeviceShard& d = shard[g];
      const size_t paths = (size_t)sub_RANSAC_iters[g] * Num_Of_Tracks;
      if (!paths) continue;
      HC_CUDA(cudaSetDevice(d.device));
      if (!d.d_refine_sums) HC_CUDA(cudaMalloc((void**)&d.d_refine_sums, paths * 2 * sizeof(float)));
      const int rc = hcb200_refine_tracks(d.stream, (int)paths, refine_iterations, d.d_target, d.d_conv, d.d_tracks, d.d_refine_sums, d.d_ws);
      if (rc != 0) { std::fprintf(stderr, "[ERROR] refinement launch failed on GPU %d: %s\n", g, hcb200_error_string(rc)); std::exit(2); }
    }
  }

  // results: every GPU copies straight into its slice of the stacked host arrays (reference: per-GPU copies + memcpy, :449-506)
  const double t0_d2h = wall_seconds();
  for (int g = 0; g < Num_Of_GPUs; g++) {
    DeviceShard& d = shard[g];
    const size_t paths = (size_t)sub_RANSAC_iters[g] * Num_Of_Tracks;
    if (!paths) continue;
    cudaStream_t s = (cudaStream_t)d.stream;
    HC_CUDA(cudaSetDevice(d.device));
    float ms = 0.f;
    HC_CUDA(cudaEventElapsedTime(&ms, (cudaEvent_t)d.ev_start, (cudaEvent_t)d.ev_stop));
    gpu_time[g] = ms * 1e-3;
    if (!lazy_results) {
      HC_CUDA(cudaMemcpyAsync(h_GPU_HC_Track_Sols_Stack + (size_t)d.path_offset * V1, d.d_tracks, sizeof(complex32) * V1 * paths, cudaMemcpyDeviceToHost, s));
      HC_CUDA(cudaMemcpyAsync(h_is_GPU_HC_Sol_Converge_Stack + d.path_offset, d.d_conv, paths, cudaMemcpyDeviceToHost, s));
      HC_CUDA(cudaMemcpyAsync(h_is_GPU_HC_Sol_Infinity_Stack + d.path_offset, d.d_inf, paths, cudaMemcpyDeviceToHost, s));
    }
    if (device_statistics) {                            // per-hypothesis counts are reduced on the GPU while the tracks are in flight
      const int rc = hcb200_count_solutions(d.stream, sub_RANSAC_iters[g], (const float*)d.d_tracks, (const uint8_t*)d.d_conv, (const uint8_t*)d.d_inf, d.d_counts);
      if (rc != 0) { std::fprintf(stderr, "[ERROR] statistics launch failed on GPU %d: %s\n", g, hcb200_error_string(rc)); std::exit(2); }
      HC_CUDA(cudaMemcpyAsync(h_counts[g], d.d_counts, sizeof(unsigned) * 3 * sub_RANSAC_iters[g], cudaMemcpyDeviceToHost, s));
    }
    if (Abort_RANSAC_by_Good_Sol) {
      HC_CUDA(cudaMemcpyAsync(h_Found_Trifocal_Sols[g], d.d_found, sizeof(bool), cudaMemcpyDeviceToHost, s));
      HC_CUDA(cudaMemcpyAsync(h_Trifocal_Sols_Batch_Index[g], d.d_found_index, paths * sizeof(int), cudaMemcpyDeviceToHost, s));
      HC_CUDA(cudaMemcpyAsync(h_best[g], d.d_best, sizeof(hcb200_best_record), cudaMemcpyDeviceToHost, s));
    }
  }
  for (int g = 0; g < Num_Of_GPUs; g++) {               // all GPUs copy at the same time; wait once for each
    if (!sub_RANSAC_iters[g]) continue;
    HC_CUDA(cudaSetDevice(shard[g].device));
    HC_CUDA(cudaStreamSynchronize((cudaStream_t)shard[g].stream));
    transfer_d2h_time[g] = wall_seconds() - t0_d2h;
  }
  phase_seconds[5] = wall_seconds() - t0_d2h;
  results_on_host = !lazy_results;

  // tiny gather: the best record of every GPU, reduced on the host (smallest global path id wins)
  found_path_ids.clear();
  best_record = hcb200_best_record();
  best_record.path_id = -1;
  if (Abort_RANSAC_by_Good_Sol) {
    for (int g = 0; g < Num_Of_GPUs; g++) {
      const size_t paths = (size_t)sub_RANSAC_iters[g] * Num_Of_Tracks;
      if (!paths) continue;
      for (size_t i = 0; i < paths; i++)
        if (h_Trifocal_Sols_Batch_Index[g][i] != -1) found_path_ids.push_back(shard[g].path_offset + h_Trifocal_Sols_Batch_Index[g][i]);
      if (h_best[g]->found && (!best_record.found || shard[g].path_offset + h_best[g]->path_id < best_record.path_id)) {
        best_record = *h_best[g];
        best_record.path_id += shard[g].path_offset;
      }
      best_record.n_passed = (int)found_path_ids.size();
    }
  }

  if (verbose) {
    std::cout << "---------------------------------------------------------------------------------" << std::endl;
    std::cout << "## Solving " << HC_print_problem_name << std::endl << std::endl;
    std::printf("## Timings:\n - GPU Computation Time = %7.2f (ms)\n", multi_GPUs_time * 1000);
  }

  const double t_stat = wall_seconds();
  if (device_statistics) {
    for (int g = 0; g < Num_Of_GPUs; g++)
      if (sub_RANSAC_iters[g]) Evaluate_GPUHC_Sols->Set_RANSAC_HC_Sol_Counts(h_counts[g], sub_RANSAC_iters[g]);
  } else {
    Fetch_Results_To_Host();
    Evaluate_GPUHC_Sols->Evaluate_RANSAC_HC_Sols(h_GPU_HC_Track_Sols_Stack, h_is_GPU_HC_Sol_Converge_Stack, h_is_GPU_HC_Sol_Infinity_Stack);
  }
  per_hypothesis_counts = Evaluate_GPUHC_Sols->Per_Hypothesis_Counts;
  phase_seconds[6] = wall_seconds() - t_stat;
  if (verbose) {
    std::cout << "\n## Evaluation of GPU-HC Solutions: " << std::endl;
    std::cout << " - Number of Converged Solutions:       " << Evaluate_GPUHC_Sols->Num_Of_Coverged_Sols << std::endl;
    std::cout << " - Number of Real Solutions:            " << Evaluate_GPUHC_Sols->Num_Of_Real_Sols << std::endl;
    std::cout << " - Number of Infinity Failed Solutions: " << Evaluate_GPUHC_Sols->Num_Of_Inf_Sols << std::endl;
  }
  Collect_Num_Of_Coverged_Sols.push_back(Evaluate_GPUHC_Sols->Num_Of_Coverged_Sols);
  Collect_Num_Of_Inf_Sols.push_back(Evaluate_GPUHC_Sols->Num_Of_Inf_Sols);
  Collect_Num_Of_Real_Sols.push_back(Evaluate_GPUHC_Sols->Num_Of_Real_Sols);

  selected_path = -1;
  selected_support = {0u, 0u};
  const double t_score = wall_seconds();
  if (device_scoring) {
    // support counting + selection on every GPU (hcb200_score_tracks), then the tiny gather: one 64-byte record per GPU,
    // reduced on the host — largest min(support21, support31), lowest global path id on ties
    for (int g = 0; g < Num_Of_GPUs; g++) {
      DeviceShard& d = shard[g];
      const int paths = sub_RANSAC_iters[g] * Num_Of_Tracks;
      if (!paths) continue;
      HC_CUDA(cudaSetDevice(d.device));
      const int rc = hcb200_score_tracks(d.stream, paths, d.d_tracks, d.d_conv, Num_Of_Triplet_Edgels, d.d_edgels, d.d_K, d.d_support,
                                         d.d_score_best, d.d_ws);
      if (rc != 0) { std::fprintf(stderr, "[ERROR] scoring launch failed on GPU %d: %s\n", g, hcb200_error_string(rc)); std::exit(2); }
      HC_CUDA(cudaMemcpyAsync(h_score_best[g], d.d_score_best, sizeof(hcb200_best_record), cudaMemcpyDeviceToHost, (cudaStream_t)d.stream));
    }
    unsigned best_min = 0;
    for (int g = 0; g < Num_Of_GPUs; g++) {
      if (!sub_RANSAC_iters[g]) continue;
      HC_CUDA(cudaSetDevice(shard[g].device));
      HC_CUDA(cudaStreamSynchronize((cudaStream_t)shard[g].stream));
      const hcb200_best_record& r = *h_score_best[g];
      if (!r.found) continue;
      const unsigned m = (unsigned)std::min(r.inliers21, r.inliers31);
      if (selected_path < 0 || m > best_min) {
        best_min = m;
        selected_path = shard[g].path_offset + r.path_id;
        selected_support = {(unsigned)r.inliers21, (unsigned)r.inliers31};
      }
    }
    found_pose = selected_path >= 0;
    if (found_pose) {
      // the selected end point alone (248 bytes) comes from the GPU that owns it; the stacked arrays may still be on the devices
      for (int g = 0; g < Num_Of_GPUs; g++) {
        const int first = shard[g].path_offset, paths = sub_RANSAC_iters[g] * Num_Of_Tracks;
        if (selected_path < first || selected_path >= first + paths) continue;
        HC_CUDA(cudaSetDevice(shard[g].device));
        HC_CUDA(cudaMemcpy(h_selected_track, shard[g].d_tracks + (size_t)(selected_path - first) * V1 * 2, sizeof(complex32) * V1, cudaMemcpyDeviceToHost));   // d_tracks counts floats
      }
      Evaluate_GPUHC_Sols->Set_Selected_Solution(h_selected_track, selected_path, selected_support[0], selected_support[1]);
    }
  } else {
    Fetch_Results_To_Host();
    Evaluate_GPUHC_Sols->Transform_GPUHC_Sols_to_Trifocal_Relative_Pose(h_GPU_HC_Track_Sols_Stack, h_is_GPU_HC_Sol_Converge_Stack, h_Camera_Intrinsic_Matrix);
    found_pose = Evaluate_GPUHC_Sols->get_Solution_with_Maximal_Support(Num_Of_Triplet_Edgels, h_Triplet_Edge_Locations, h_Triplet_Edge_Tangents, h_Camera_Intrinsic_Matrix);
    if (found_pose) {
      selected_path = Evaluate_GPUHC_Sols->Best_Candidate_Path_Index;
      selected_support = {Evaluate_GPUHC_Sols->Max_Num_Of_Reproj_Inliers_Views21, Evaluate_GPUHC_Sols->Max_Num_Of_Reproj_Inliers_Views31};
    }
  }
  phase_seconds[7] = wall_seconds() - t_score;
  if (verbose)
    std::printf("## Phases (s): allocate %.3f, host->device %.3f, GPU tracking %.3f, device->host %.3f, statistics %.3f, scoring %.3f\n",
                phase_seconds[0], phase_seconds[3], phase_seconds[4], phase_seconds[5], phase_seconds[6], phase_seconds[7]);
  pose_residuals = {100.f, 100.f, 100.f, 100.f};
  if (found_pose) {
    Evaluate_GPUHC_Sols->Measure_Relative_Pose_Error(h_Camera_Pose21, h_Camera_Pose31);
    pose_residuals = {Evaluate_GPUHC_Sols->Min_Residual_R21, Evaluate_GPUHC_Sols->Min_Residual_R31,
                      Evaluate_GPUHC_Sols->Min_Residual_t21, Evaluate_GPUHC_Sols->Min_Residual_t31};
    if (verbose) {
      std::printf("## Pose with maximal support: path %d, inliers (1,2) %u / (1,3) %u of %d\n", selected_path,
                  selected_support[0], selected_support[1], Num_Of_Triplet_Edgels);
      std::printf(" - residuals vs GT: R21 %.3g rad, R31 %.3g rad, t21 %.3g, t31 %.3g%s\n", pose_residuals[0], pose_residuals[1],
                  pose_residuals[2], pose_residuals[3], Evaluate_GPUHC_Sols->success_flag ? "   ### Found GT pose!" : "");
    }
  }
  Evaluate_GPUHC_Sols->Flush_Out_Data();
}

void GPU_HC_Solver::Export_Data() {}

void GPU_HC_Solver::Free_Arrays_for_Aborting_RANSAC()
{
  DeviceGuard keep_callers_device;
  if (!abort_arrays_allocated) return;
  for (int g = 0; g < Num_Of_GPUs; g++) {
    DeviceShard& d = shard[g];
    HC_CUDA(cudaSetDevice(d.device));
    delete[] h_Found_Trifocal_Sols[g]; h_Found_Trifocal_Sols[g] = nullptr;
    delete[] h_Trifocal_Sols_Batch_Index[g]; h_Trifocal_Sols_Batch_Index[g] = nullptr;
    delete h_best[g]; h_best[g] = nullptr;
    cudaFree(d.d_found); cudaFree(d.d_found_index); cudaFree(d.d_best);
    d.d_found = nullptr; d.d_found_index = nullptr; d.d_best = nullptr;
  }
  abort_arrays_allocated = false;
}

void GPU_HC_Solver::Free_Triplet_Edgels_Mem()
{
  DeviceGuard keep_callers_device;
  if (device_edgels_allocated) {
    for (int g = 0; g < Num_Of_GPUs; g++) {
      cudaSetDevice(shard[g].device);
      cudaFree(shard[g].d_edgels); cudaFree(shard[g].d_K); cudaFree(shard[g].d_tangents);
      shard[g].d_edgels = shard[g].d_K = shard[g].d_tangents = nullptr;
    }
    device_edgels_allocated = false;
    device_edgel_capacity = 0;
  }
  if (!edgels_allocated) return;
  delete[] h_Triplet_Edge_Locations; delete[] h_Triplet_Edge_Tangents;
  h_Triplet_Edge_Locations = h_Triplet_Edge_Tangents = nullptr;
  edgels_allocated = false;
}

GPU_HC_Solver::~GPU_HC_Solver()
{
  DeviceGuard keep_callers_device;
  Free_Arrays_for_Aborting_RANSAC();
  Free_Triplet_Edgels_Mem();
  for (int g = 0; g < Num_Of_GPUs && g < MAX_NUM_OF_GPUS; g++) {
    DeviceShard& d = shard[g];
    if (!d.stream) continue;
    cudaSetDevice(d.device);
    if (arrays_allocated) {
      cudaFree(d.d_start_sols); cudaFree(d.d_start_params); cudaFree(d.d_target); cudaFree(d.d_diff); cudaFree(d.d_tracks);
      cudaFree(d.d_conv); cudaFree(d.d_inf); cudaFree(d.d_ws); cudaFree(d.d_support); cudaFree(d.d_score_best); cudaFree(d.d_refine_sums);
      cudaFreeHost(h_Target_Params[g]); cudaFreeHost(h_diffParams[g]); cudaFreeHost(h_score_best[g]); cudaFreeHost(h_picked[g]);
      cudaFreeHost(h_counts[g]);
      cudaFree(d.d_picked); cudaFree(d.d_counts);
    }
    cudaEventDestroy((cudaEvent_t)d.ev_start); cudaEventDestroy((cudaEvent_t)d.ev_stop);
    cudaStreamDestroy((cudaStream_t)d.stream);
  }
  if (arrays_allocated) {
    std::free(h_Start_Sols); std::free(h_Start_Params);
    delete[] h_dHdx_Index; delete[] h_dHdt_Index; delete[] h_Camera_Intrinsic_Matrix;
    if (result_stacks_allocated) { cudaFreeHost(h_GPU_HC_Track_Sols_Stack); cudaFreeHost(h_is_GPU_HC_Sol_Converge_Stack); cudaFreeHost(h_is_GPU_HC_Sol_Infinity_Stack); }
  }
}
