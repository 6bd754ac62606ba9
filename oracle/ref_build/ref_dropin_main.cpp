// Drop-in check: the UNMODIFIED reference host layer (GPU_HC_Solver / Data_Reader / Evaluations, /root/reference/magmaHC)
// driven in the order of the reference's own driver (cmd/magmaHC-main.cpp:24-116, run_GPU_HC_Solver), with the reference's GPU
// kernels replaced by integration/hcb200_shim.cpp -> libhcb200.so.  Built by oracle/Makefile (target ref_dropin) into
// oracle/_ref/ref_gpuhc_on_hcb200; run from <tree>/build/bin like the reference.  Test infrastructure only.
//   usage: ref_gpuhc_on_hcb200 <problem_name> [n_hypotheses]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include "definitions.hpp"
#include "GPU_HC_Solver.hpp"
#include <yaml-cpp/yaml.h>

int g_hcb200_ref_num_hyp = 100;      // NUM_OF_RANSAC_ITERATIONS of the generated definitions.hpp (oracle/Makefile)

int main(int argc, char** argv)
{
  if (argc < 2) { printf("usage: %s <problem_name> [n_hypotheses]\n", argv[0]); return 2; }
  const std::string problem = argv[1];
  if (argc > 2) g_hcb200_ref_num_hyp = atoi(argv[2]);
  YAML::Node settings;
  try { settings = YAML::LoadFile("../../problems/" + problem + "/gpuhc_settings.yaml"); }
  catch (const std::exception& e) { std::cerr << "Exception: " << e.what() << std::endl; return 1; }

  GPU_HC_Solver GPU_HC_(settings);
  GPU_HC_.Allocate_Arrays();
  double all_gpu_runtime[TEST_RANSAC_TIMES];
  for (int ti = 0; ti < TEST_RANSAC_TIMES; ti++) {
    if (!GPU_HC_.Read_Problem_Data()) return 1;
    if (!GPU_HC_.Read_RANSAC_Data(ti)) return 1;
    GPU_HC_.Prepare_Target_Params(ti);
    GPU_HC_.Set_RANSAC_Abort_Arrays();
    GPU_HC_.Data_Transfer_From_Host_To_Device();
    GPU_HC_.Set_CUDA_Stream_Attributes();
    GPU_HC_.Solve_by_GPU_HC();
    GPU_HC_.Free_Triplet_Edgels_Mem();
    GPU_HC_.Free_Arrays_for_Aborting_RANSAC();
    all_gpu_runtime[ti] = GPU_HC_.multi_GPUs_time * 1000;
  }
  // the two files the reference driver writes (cmd/magmaHC-main.cpp:97-116)
  std::ofstream timings(std::string("../../") + WRITE_FILES_FOLDER + "GPU_Timings.txt");
  for (int i = 0; i < TEST_RANSAC_TIMES; i++) timings << all_gpu_runtime[i] << "\n";
  std::ofstream stats(std::string("../../") + WRITE_FILES_FOLDER + "GPU_Sols_Statistics.txt");
  for (int i = 0; i < TEST_RANSAC_TIMES; i++)
    stats << GPU_HC_.Collect_Num_Of_Coverged_Sols[i] << "\t" << GPU_HC_.Collect_Num_Of_Inf_Sols[i] << "\t" << GPU_HC_.Collect_Num_Of_Real_Sols[i] << "\n";
  return 0;
}
