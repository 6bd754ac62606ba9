/* hcb200_host.h — C handle API over the C++ host layer (class GPU_HC_Solver), for callers that cannot include C++
 * (the Python tests and bench.py use it through ctypes).  It mirrors, call for call, what the reference driver does with
 * its GPU_HC_Solver object (cmd/magmaHC-main.cpp:24-66 of the reference):
 *
 *   create(settings)  ->  allocate  ->  read_problem  ->  read_ransac(i)  ->  prepare(seed)  ->  set_abort_arrays
 *        ->  h2d  ->  solve  ->  (results)  ->  free_round  ->  destroy
 *
 * All functions return 0 on success unless stated otherwise.  Device work goes through include/hcb200.h. */
#ifndef HCB200_HOST_H
#define HCB200_HOST_H
#include <stdint.h>
#include "hcb200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hcb200_solver hcb200_solver;

/* settings_yaml: path of a gpuhc_settings.yaml (reference format).  overrides: optional "key=value;key=value" list applied
 * on top (e.g. "Num_Of_GPUs=2;Abort_RANSAC_by_Good_Sol=true;Num_Of_RANSAC_Iterations=8;Repo_Root=/tmp/tree/;Verbose=false"). */
hcb200_solver* hcb200_solver_create(const char* settings_yaml, const char* overrides);
void hcb200_solver_destroy(hcb200_solver* s);

int hcb200_solver_allocate(hcb200_solver* s);                       /* GPU_HC_Solver::Allocate_Arrays                    */
int hcb200_solver_read_problem(hcb200_solver* s);                   /* ::Read_Problem_Data                               */
int hcb200_solver_read_ransac(hcb200_solver* s, int dataset_index); /* ::Read_RANSAC_Data                                */
int hcb200_solver_prepare(hcb200_solver* s, unsigned seed);         /* ::Prepare_Target_Params                           */
int hcb200_solver_set_abort_arrays(hcb200_solver* s);               /* ::Set_RANSAC_Abort_Arrays                         */
int hcb200_solver_h2d(hcb200_solver* s);                            /* ::Data_Transfer_From_Host_To_Device + stream attrs */
int hcb200_solver_solve(hcb200_solver* s);                          /* ::Solve_by_GPU_HC                                 */
int hcb200_solver_free_round(hcb200_solver* s);                     /* ::Free_Triplet_Edgels_Mem + ::Free_Arrays_for_Aborting_RANSAC */
int hcb200_solver_set_pruning(hcb200_solver* s, int on);

/* results of the last solve */
int    hcb200_solver_num_hypotheses(hcb200_solver* s);
double hcb200_solver_kernel_seconds(hcb200_solver* s);              /* multi_GPUs_time                                   */
int    hcb200_solver_totals(hcb200_solver* s, unsigned out_conv_real_inf[3]);       /* file column order                 */
int    hcb200_solver_per_hypothesis(hcb200_solver* s, unsigned* out /*[H][3] conv, inf, real*/);
int    hcb200_solver_copy_results(hcb200_solver* s, float* tracks /*[H*312][31][2]*/, uint8_t* conv, uint8_t* inf);
int    hcb200_solver_copy_target_params(hcb200_solver* s, float* out /*[H][34][2], stacked over GPUs*/);
int    hcb200_solver_best(hcb200_solver* s, hcb200_best_record* rec, int* pose_found, float residuals_R21_R31_t21_t31[4]);
int    hcb200_solver_selected(hcb200_solver* s, int* path_id, unsigned support21_31[2]);   /* pose with maximal support; 0 if one exists */
int    hcb200_solver_shard_size(hcb200_solver* s, int gpu_id);

/* Reference text formats without a device: Data_Reader over <problem_dir> (start_sols.txt, start_params.txt, dHdx_indx.txt,
 * dHdt_indx.txt) and <ransac_dir> (Triplet_Edgels/, GT_Poses21/, GT_Poses31/, Intrinsic_Matrix.txt).  Buffers: start_sols
 * [312][31][2], start_params [34][2], dHdx [36000], dHdt [2880], locations/tangents [edgel_capacity][6], poses [12], K [9].
 * Returns 0, or the 1-based index of the first file that failed. */
int hcb200_reader_load(const char* problem_dir, const char* ransac_dir, int dataset_index,
                       float* start_sols, float* start_params, int* dHdx, int* dHdt,
                       int* n_edgels, float* locations, float* tangents, int edgel_capacity,
                       float* pose21, float* pose31, float* K);
/* Host scoring of one end point [31][2] with the arithmetic of Evaluations / mvg.hpp (reference Evaluations.cpp:298-504, util.hpp:169-209):
 * returns 1 and the two inlier counts if the path is a pose candidate, 0 otherwise.  The device kernel behind
 * hcb200_score_tracks must agree with it count for count. */
int hcb200_host_score_track(const float* track31, const float* locations, int n_edgels, const float* K, int* n21, int* n31);
/* value of one key of a gpuhc_settings.yaml as text; 0 found, 1 missing key, 2 buffer too small, 3 unreadable file */
int hcb200_settings_lookup(const char* settings_yaml, const char* key, char* out, int capacity);

#ifdef __cplusplus
}
#endif
#endif
