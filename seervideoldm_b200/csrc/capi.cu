#include "common.cuh"
#include "seer_b200.h"
extern "C" const char* seer_b200_version(void) { return "seer_b200 0.1 (sm_100a)"; }
