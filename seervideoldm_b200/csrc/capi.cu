#include "common.cuh"
#include "seer_b200.h"
extern "C" const char* seer_b200_version(void) { return "seer_b200 0.2 (sm_100a)"; }
extern "C" int seer_b200_gemm_desc_size(void) { return (int)sizeof(SeerGemmDesc); }
