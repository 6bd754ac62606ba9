// hc-main — command-line driver, same usage and output files as the reference executable
// (cmd/magmaHC-main.cpp:197-260):   ./hc-main -p trifocal_2op1p_30x30        (run from <root>/build/bin, or add -d <root>)
// It runs the GPU solver only: the reference's CPU-HC half is its baseline, not part of this product (no CPU fallback).
// Files written under <root>/Output_Write_Files/: GPU_Timings.txt (ms per round), GPU_Sols_Statistics.txt
// (converged <TAB> real <TAB> infinity per round) — SURVEY.md App. A.4.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "GPU_HC_Solver.hpp"

bool run_compiled_problem(YAML::Node settings, const std::string& root);      // generic_problem.cpp: -p <a problem other than the trifocal one>

static void print_help()
{
  std::printf("Usage: ./hc-main [options]\n\noptions:\n"
              "  -h, --help            show this help message and exit\n"
              "  -p, --problem NAME    problem folder name under <root>/problems (trifocal_2op1p_30x30: the RANSAC pipeline; any other\n"
              "                        folder: tracked to its target_params.txt with the library compiled by `make problem PROBLEM_DIR=...`)\n"
              "  -d, --directory ROOT  repository root holding problems/, RANSAC_Data/, Output_Write_Files/ (default ../../)\n"
              "  -s, --set KEY=VALUE   override one settings key (repeatable)\n");
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static bool run_GPU_HC_Solver(YAML::Node settings, const std::string& root)
{
  std::vector<double> all_ms;
  double t[8];
  t[0] = now_s();
  GPU_HC_Solver GPU_HC_(settings);
  t[1] = now_s();
  GPU_HC_.Allocate_Arrays();
  t[2] = now_s();
  for (int ti = 0; ti < TEST_RANSAC_TIMES; ti++) {
    if (!GPU_HC_.Read_Problem_Data()) return false;
    if (!GPU_HC_.Read_RANSAC_Data(ti)) return false;
    t[3] = now_s();
    GPU_HC_.Prepare_Target_Params(ti);
    GPU_HC_.Set_RANSAC_Abort_Arrays();
    t[4] = now_s();
    GPU_HC_.Data_Transfer_From_Host_To_Device();
    GPU_HC_.Set_CUDA_Stream_Attributes();
    t[5] = now_s();
    GPU_HC_.Solve_by_GPU_HC();
    t[6] = now_s();
    GPU_HC_.Free_Triplet_Edgels_Mem();
    GPU_HC_.Free_Arrays_for_Aborting_RANSAC();
    all_ms.push_back(GPU_HC_.multi_GPUs_time * 1000);
  }
  t[7] = now_s();
  std::printf("## Driver wall clock (s): solver object + CUDA context %.3f, allocate %.3f, read files %.3f, sample hypotheses %.3f, host->device %.3f, "
              "solve + evaluate %.3f, free round %.3f\n", t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5], t[7] - t[6]);
  double avg = 0, mx = 0, mn = 1e30, var = 0;
  for (double v : all_ms) { avg += v; mx = std::max(mx, v); mn = std::min(mn, v); }
  avg /= all_ms.size();
  for (double v : all_ms) var += (v - avg) * (v - avg);
  std::printf("\n## Running %d rounds of %d RANSAC iterations:\n", TEST_RANSAC_TIMES, GPU_HC_.Num_Of_RANSAC_Iterations());
  std::printf(" - [Average GPU Computation Time] %7.2f (ms)\n - [Maximal GPU Computation Time] %7.2f (ms)\n"
              " - [Minimal GPU Computation Time] %7.2f (ms)\n - [Std dev GPU Computation Time] %7.2f (ms)\n",
              avg, mx, mn, std::sqrt(var / all_ms.size()));

  const std::string out_dir = root + WRITE_FILES_FOLDER;
  std::ofstream timings(out_dir + "GPU_Timings.txt");
  if (!timings.is_open()) hcb200::log_file_error(out_dir + "GPU_Timings.txt");
  for (double v : all_ms) timings << v << "\n";
  std::ofstream stats(out_dir + "GPU_Sols_Statistics.txt");
  if (!stats.is_open()) hcb200::log_file_error(out_dir + "GPU_Sols_Statistics.txt");
  for (size_t i = 0; i < all_ms.size(); i++)
    stats << GPU_HC_.Collect_Num_Of_Coverged_Sols[i] << "\t" << GPU_HC_.Collect_Num_Of_Real_Sols[i] << "\t" << GPU_HC_.Collect_Num_Of_Inf_Sols[i] << "\n";
  return true;
}

int main(int argc, char** argv)
{
  std::string problem, root = "../../";
  std::vector<std::string> sets;
  if (argc < 2) { print_help(); return 0; }
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    if (a == "-h" || a == "--help") { print_help(); return 0; }
    else if ((a == "-p" || a == "--problem") && i + 1 < argc) problem = argv[++i];
    else if ((a == "-d" || a == "--directory") && i + 1 < argc) { root = argv[++i]; if (root.back() != '/') root += '/'; }
    else if ((a == "-s" || a == "--set") && i + 1 < argc) sets.push_back(argv[++i]);
    else { hcb200::log_error("Invalid input arguments!"); print_help(); return 0; }
  }
  if (problem.empty()) { hcb200::log_error("Invalid input arguments!"); print_help(); return 0; }
  YAML::Node settings;
  try {
    settings = YAML::LoadFile(root + "problems/" + problem + "/gpuhc_settings.yaml");
    settings.set("Repo_Root", root);
    for (const auto& kv : sets) {
      const size_t eq = kv.find('=');
      if (eq != std::string::npos) settings.set(YAML::strip(kv.substr(0, eq)), YAML::strip(kv.substr(eq + 1)));
    }
    std::cout << std::endl << settings << std::endl;
  } catch (const std::exception& e) {
    std::cerr << "Exception: " << e.what() << std::endl;
    return 0;
  }
  try {
    if (settings["problem_name"].as<std::string>() != "trifocal_2op1p_30x30") return run_compiled_problem(settings, root) ? 0 : 1;
    return run_GPU_HC_Solver(settings, root) ? 0 : 1;
  } catch (const std::exception& e) {          // a missing / malformed settings key
    std::cerr << "Exception: " << e.what() << std::endl;
    return 1;
  }
}
