#include "common.cuh"
#include "seer_b200.h"

#include <mutex>
#include <string.h>

namespace seer {

struct EnvEntry { char name[48]; int val; int overridden; };
static EnvEntry g_env[64];
static int g_env_n = 0;
static std::mutex g_env_mu;

int env_cached(const char* name, int dflt) {
  std::lock_guard<std::mutex> lock(g_env_mu);
  for (int i = 0; i < g_env_n; ++i)
    if (strcmp(g_env[i].name, name) == 0) return g_env[i].overridden == 2 ? dflt : g_env[i].val;
  const char* v = getenv(name);
  const bool set = v && *v;
  if (g_env_n < 64 && strlen(name) < sizeof(g_env[0].name)) {
    strcpy(g_env[g_env_n].name, name);
    g_env[g_env_n].val = set ? atoi(v) : 0;
    g_env[g_env_n].overridden = set ? 1 : 2;      // 2: not set -> every caller's own default applies
    ++g_env_n;
  }
  return set ? atoi(v) : dflt;
}

// tuning tools (tools/gemm_bench.py --sweep): override / clear a cached switch inside the running process
static void env_override(const char* name, int value, int clear) {
  std::lock_guard<std::mutex> lock(g_env_mu);
  for (int i = 0; i < g_env_n; ++i)
    if (strcmp(g_env[i].name, name) == 0) {
      g_env[i].val = value;
      g_env[i].overridden = clear ? 2 : 1;
      return;
    }
  if (g_env_n < 64 && strlen(name) < sizeof(g_env[0].name)) {
    strcpy(g_env[g_env_n].name, name);
    g_env[g_env_n].val = value;
    g_env[g_env_n].overridden = clear ? 2 : 1;
    ++g_env_n;
  }
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
extern "C" void seer_b200_debug_setenv(const char* name, int value, int clear) { seer::env_override(name, value, clear); }
