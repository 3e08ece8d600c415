#include "common.cuh"
#include "seer_b200.h"

#include <mutex>
#include <string.h>

namespace seer {

int env_cached(const char* name, int dflt) {
  struct Entry { char name[48]; int val; };
  static Entry tab[64];
  static int n = 0;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < n; ++i)
    if (strcmp(tab[i].name, name) == 0) return tab[i].val;
  const char* v = getenv(name);
  const int val = v && *v ? atoi(v) : dflt;
  if (n < 64 && strlen(name) < sizeof(tab[0].name)) {
    strcpy(tab[n].name, name);
    tab[n].val = val;
    ++n;
  }
  return val;
}

static thread_local char g_last_attention[96] = "none";
static thread_local char g_last_gemm[160] = "none";

void debug_note_attention(const char* what) { snprintf(g_last_attention, sizeof(g_last_attention), "%s", what); }
void debug_note_gemm(int bn, int cg, int stages, int nepi, int ring, int bstat, int epi_spec, int mode) {
  snprintf(g_last_gemm, sizeof(g_last_gemm), "gemm_tc_kernel<%d,%d> tcgen05 stages=%d nepi=%d ring=%d bstat=%d spec=%d mode=%d", bn, cg, stages,
           nepi, ring, bstat, epi_spec, mode);
}

}  // namespace seer

extern "C" const char* seer_b200_version(void) { return "seer_b200 0.3 (sm_100a)"; }
extern "C" int seer_b200_gemm_desc_size(void) { return (int)sizeof(SeerGemmDesc); }
extern "C" const char* seer_b200_debug_last_attention(void) { return seer::g_last_attention; }
extern "C" const char* seer_b200_debug_last_gemm(void) { return seer::g_last_gemm; }
