// Compile-time constants of the host layer.  Values are part of the parity contract with the reference
// (magmaHC/definitions.hpp:4-44); names are kept so reference users find them.
#ifndef HCB200_HOST_DEFINITIONS_HPP
#define HCB200_HOST_DEFINITIONS_HPP
#include <string>

#define WRITE_FILES_FOLDER                std::string("Output_Write_Files/")
#define MAX_NUM_OF_GPUS                   (8)
#define SET_GPU_DEVICE_ID                 (0)
// The reference fixes the RANSAC iteration count at compile time (definitions.hpp:12).  Here it is the default of the
// optional YAML key `Num_Of_RANSAC_Iterations`, so large synthetic sweeps need no rebuild.
#define NUM_OF_RANSAC_ITERATIONS          (100)
#define IMAG_PART_TOL                     (1e-5)
#define ROT_RESIDUAL_TOL                  (1e-1)
#define TRANSL_RESIDUAL_TOL               (1e-1)
#define TEST_RANSAC_TIMES                 (1)
#define REPROJ_ERROR_INLIER_THRESH        (2)
#define PASS_RANSAC_INLIER_SUPPORT_RATIO  (0.90)
#define DUPLICATE_SOL_DIFF_TOL            (1e-4)
#define ZERO_IMAG_PART_TOL_FOR_SP         (1e-4)

namespace hcb200 {
struct complex32 { float x, y; };            // layout of cuFloatComplex / magmaFloatComplex / float2
inline complex32 make_c32(float re, float im) { complex32 z; z.x = re; z.y = im; return z; }
void log_info(const std::string& msg);
void log_error(const std::string& msg);
void log_file_error(const std::string& path);
}
#endif
