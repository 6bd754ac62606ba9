// hc-main -p <another problem>: the original GPU-HC usage (reference README.md:25, cmd/magmaHC-main.cpp:204-236) for a problem folder that
// has been COMPILED into its own tracker library (make problem PROBLEM_DIR=problems/<name> -> lib/libhcb200_<name>.so).
// The folder's start system is tracked to the target parameters of problems/<name>/target_params.txt (Num_Of_RANSAC_Iterations copies of it
// when that key is given: a throughput run), the solution statistics go to the reference's files, the converged end points to
// Output_Write_Files/GPU_Converged_HC_Tracks.txt.  There is no CPU fallback: a missing library or device is fatal.
#include <dlfcn.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "Data_Reader.hpp"
#include "definitions.hpp"
#include "hcb200.h"
#include "yaml_lite.hpp"

namespace {

struct ProblemLibrary {          // the C ABI of include/hcb200.h, bound at run time because the library is per problem
  void* handle = nullptr;
  decltype(&hcb200_problem_info) problem_info = nullptr;
  decltype(&hcb200_workspace_bytes_for) workspace_bytes_for = nullptr;
  decltype(&hcb200_track) track = nullptr;
  decltype(&hcb200_count_solutions) count_solutions = nullptr;
  decltype(&hcb200_error_string) error_string = nullptr;
};

std::string executable_dir()
{
  char buf[4096];
  const ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
  if (n <= 0) return ".";
  std::string p(buf, (size_t)n);
  const size_t s = p.rfind('/');
  return s == std::string::npos ? "." : p.substr(0, s);
}

bool load(ProblemLibrary& L, const std::string& name)
{
  const std::string path = executable_dir() + "/libhcb200_" + name + ".so";
  L.handle = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!L.handle) {
    hcb200::log_error("no tracker library for problem '" + name + "' (" + path + "): compile the problem folder first — make problem PROBLEM_DIR=problems/" + name);
    return false;
  }
  L.problem_info = (decltype(L.problem_info))dlsym(L.handle, "hcb200_problem_info");
  L.workspace_bytes_for = (decltype(L.workspace_bytes_for))dlsym(L.handle, "hcb200_workspace_bytes_for");
  L.track = (decltype(L.track))dlsym(L.handle, "hcb200_track");
  L.count_solutions = (decltype(L.count_solutions))dlsym(L.handle, "hcb200_count_solutions");
  L.error_string = (decltype(L.error_string))dlsym(L.handle, "hcb200_error_string");
  return L.problem_info && L.workspace_bytes_for && L.track && L.count_solutions && L.error_string;
}

#define GP_CUDA(call)                                                                                                  \
  do {                                                                                                                 \
    cudaError_t e_ = (call);                                                                                           \
    if (e_ != cudaSuccess) { std::fprintf(stderr, "[ERROR] %s: %s\n", #call, cudaGetErrorString(e_)); std::exit(2); }  \
  } while (0)

}  // namespace

bool run_compiled_problem(YAML::Node cfg, const std::string& root)
{
  using hcb200::complex32;
  const std::string name = cfg["problem_name"].as<std::string>();
  const int V = cfg["Num_Of_Vars"].as<int>(), NP = cfg["Num_Of_Params"].as<int>(), T = cfg["Num_Of_Tracks"].as<int>();
  const int max_steps = cfg["GPUHC_Max_Steps"].as<int>(), max_corr = cfg["GPUHC_Max_Correction_Steps"].as<int>();
  const int dt_inc = cfg["GPUHC_Num_Of_Steps_to_Increase_Delta_t"].as<int>();
  const int H = cfg.as_or<int>("Num_Of_RANSAC_Iterations", 1);
  const bool prune = cfg.as_or<bool>("Prune_Paths", false);
  const bool split = cfg.as_or<bool>("Split_Long_Paths", true) && H <= HCB200_SPLIT_MAX_HYPOTHESES;
  if (H < 1) { hcb200::log_error("Num_Of_RANSAC_Iterations must be >= 1"); return false; }

  ProblemLibrary L;
  if (!load(L, name)) return false;
  int lv = 0, lp = 0, lt = 0, tri = 0;
  const char* lname = nullptr;
  L.problem_info(&lv, &lp, &lt, &tri, &lname);
  if (lv != V || lp != NP || lt != T || name != lname) {
    hcb200::log_error("libhcb200_" + name + ".so was compiled for " + std::string(lname ? lname : "?") + " (" + std::to_string(lv) + " unknowns, " +
                      std::to_string(lp) + " parameters, " + std::to_string(lt) + " paths); gpuhc_settings.yaml describes another problem — recompile it");
    return false;
  }

  const size_t V1 = (size_t)V + 1, P1 = (size_t)NP + 1, paths = (size_t)H * T;
  std::vector<complex32> start_sols((size_t)T * V1), start_params(P1), target(P1);
  complex32 *p_ss = start_sols.data(), *p_sp = start_params.data(), *p_tp = target.data();
  Data_Reader reader(root + "problems/" + name, root + "RANSAC_Data/" + name, T, V, NP);
  if (!reader.Read_Start_Params(p_sp)) { hcb200::log_error("Start Parameters"); return false; }
  if (!reader.Read_Target_Params(p_tp)) { hcb200::log_error("Target Parameters"); return false; }
  if (!reader.Read_Start_Sols(p_ss)) { hcb200::log_error("Start Solutions"); return false; }
  std::vector<complex32> targets(H * P1), diffs(H * P1);
  for (int h = 0; h < H; h++)
    for (size_t i = 0; i < P1; i++) {
      targets[h * P1 + i] = target[i];
      diffs[h * P1 + i] = hcb200::make_c32(target[i].x - start_params[i].x, target[i].y - start_params[i].y);      // GPU_HC_Solver.cpp:298-299
    }

  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) { hcb200::log_error("no CUDA device: hc-main has no CPU path"); std::exit(2); }
  float *d_ss, *d_sp, *d_tp, *d_df, *d_tr;
  unsigned char *d_cv, *d_inf;
  unsigned* d_counts;
  void* d_ws;
  cudaStream_t s;
  GP_CUDA(cudaStreamCreate(&s));
  GP_CUDA(cudaMalloc((void**)&d_ss, sizeof(complex32) * T * V1));
  GP_CUDA(cudaMalloc((void**)&d_sp, sizeof(complex32) * P1));
  GP_CUDA(cudaMalloc((void**)&d_tp, sizeof(complex32) * H * P1));
  GP_CUDA(cudaMalloc((void**)&d_df, sizeof(complex32) * H * P1));
  GP_CUDA(cudaMalloc((void**)&d_tr, sizeof(complex32) * paths * V1));
  GP_CUDA(cudaMalloc((void**)&d_cv, paths));
  GP_CUDA(cudaMalloc((void**)&d_inf, paths));
  GP_CUDA(cudaMalloc((void**)&d_counts, sizeof(unsigned) * 3 * H));
  GP_CUDA(cudaMalloc(&d_ws, L.workspace_bytes_for(split ? H : 0)));
  GP_CUDA(cudaMemcpyAsync(d_ss, start_sols.data(), sizeof(complex32) * T * V1, cudaMemcpyHostToDevice, s));
  GP_CUDA(cudaMemcpyAsync(d_sp, start_params.data(), sizeof(complex32) * P1, cudaMemcpyHostToDevice, s));
  GP_CUDA(cudaMemcpyAsync(d_tp, targets.data(), sizeof(complex32) * H * P1, cudaMemcpyHostToDevice, s));
  GP_CUDA(cudaMemcpyAsync(d_df, diffs.data(), sizeof(complex32) * H * P1, cudaMemcpyHostToDevice, s));

  const unsigned flags = (prune ? HCB200_FLAG_PRUNE_PATHS : 0u) | (split ? HCB200_FLAG_SPLIT_LONG_PATHS : 0u);
  cudaEvent_t e0, e1;
  GP_CUDA(cudaEventCreate(&e0));
  GP_CUDA(cudaEventCreate(&e1));
  float ms = 0.0f;
  for (int rep = 0; rep < 2; rep++) {        // the first launch loads the module; the second one is timed
    GP_CUDA(cudaEventRecord(e0, s));
    const int rc = L.track(s, H, max_steps, max_corr, dt_inc, flags, d_ss, d_sp, d_tp, d_df, d_tr, d_cv, d_inf, nullptr, d_ws);
    if (rc != 0) { std::fprintf(stderr, "[ERROR] tracker launch failed: %s\n", L.error_string(rc)); std::exit(2); }
    GP_CUDA(cudaEventRecord(e1, s));
    GP_CUDA(cudaStreamSynchronize(s));
    GP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  }
  const int rc = L.count_solutions(s, H, d_tr, d_cv, d_inf, d_counts);
  if (rc != 0) { std::fprintf(stderr, "[ERROR] statistics launch failed: %s\n", L.error_string(rc)); std::exit(2); }
  std::vector<unsigned> counts(3 * (size_t)H);
  std::vector<complex32> tracks((size_t)T * V1);
  std::vector<unsigned char> conv(T);
  GP_CUDA(cudaMemcpyAsync(counts.data(), d_counts, sizeof(unsigned) * 3 * H, cudaMemcpyDeviceToHost, s));
  GP_CUDA(cudaMemcpyAsync(tracks.data(), d_tr, sizeof(complex32) * T * V1, cudaMemcpyDeviceToHost, s));      // first hypothesis (they are all the same system)
  GP_CUDA(cudaMemcpyAsync(conv.data(), d_cv, T, cudaMemcpyDeviceToHost, s));
  GP_CUDA(cudaStreamSynchronize(s));

  unsigned long long n_conv = 0, n_inf = 0, n_real = 0;
  for (int h = 0; h < H; h++) { n_conv += counts[3 * h]; n_inf += counts[3 * h + 1]; n_real += counts[3 * h + 2]; }
  std::printf("\n## %s: %d x %d paths tracked on the GPU in %.3f ms\n", name.c_str(), H, T, ms);
  std::printf(" - [Number of converged solutions]  %llu\n - [Number of real solutions]       %llu\n - [Number of infinity failed paths] %llu\n",
              n_conv, n_real, n_inf);

  const std::string out_dir = root + WRITE_FILES_FOLDER;
  std::ofstream timings(out_dir + "GPU_Timings.txt");
  if (!timings.is_open()) hcb200::log_file_error(out_dir + "GPU_Timings.txt");
  timings << ms << "\n";
  std::ofstream stats(out_dir + "GPU_Sols_Statistics.txt");
  if (!stats.is_open()) hcb200::log_file_error(out_dir + "GPU_Sols_Statistics.txt");
  stats << n_conv << "\t" << n_real << "\t" << n_inf << "\n";
  std::ofstream sols(out_dir + "GPU_Converged_HC_Tracks.txt");
  if (!sols.is_open()) hcb200::log_file_error(out_dir + "GPU_Converged_HC_Tracks.txt");
  char line[64];
  for (int t = 0; t < T; t++) {
    if (!conv[t]) continue;
    sols << "track " << t << "\n";
    for (int v = 0; v < V; v++) {
      std::snprintf(line, sizeof line, "%.9g\t%.9g\n", tracks[t * V1 + v].x, tracks[t * V1 + v].y);
      sols << line;
    }
  }
  cudaFree(d_ss); cudaFree(d_sp); cudaFree(d_tp); cudaFree(d_df); cudaFree(d_tr); cudaFree(d_cv); cudaFree(d_inf); cudaFree(d_counts); cudaFree(d_ws);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(s);
  return true;
}
