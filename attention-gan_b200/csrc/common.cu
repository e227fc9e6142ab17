// Error plumbing and library identity for libattngan_b200 (C ABI in include/attngan_b200.h).
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "agb_common.cuh"

namespace agb {

static thread_local char g_err[512] = "";

static void vset(const char* fmt, va_list ap) { vsnprintf(g_err, sizeof(g_err), fmt, ap); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vset(fmt, ap);
  va_end(ap);
}
int fail_arg(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vset(fmt, ap);
  va_end(ap);
  return AGB_E_BADARG;
}
int fail_unsupported(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vset(fmt, ap);
  va_end(ap);
  return AGB_E_UNSUPPORTED;
}
static std::atomic<long long> g_launches{0};

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}

// ---- options ----------------------------------------------------------------------------------
static Options g_opt;
static std::once_flag g_opt_once;
static long long env_ll(const char* name, long long dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoll(e) : dflt;
}
static void load_options() {
  g_opt.damsm_chunk_bytes = env_ll("AGB_DAMSM_CHUNK_MB", 16384) << 20;
  g_opt.damsm_save_bytes = env_ll("AGB_DAMSM_SAVE_MB", 65536) << 20;
  g_opt.damsm_bwd = (int)env_ll("AGB_DAMSM_BWD", 0);
  g_opt.damsm_uniform_split = (int)env_ll("AGB_DAMSM_UNIFORM_SPLIT", 0);
  g_opt.damsm_img_block = (int)env_ll("AGB_DAMSM_IMG_BLOCK", 0);
  g_opt.damsm_dw_splits = (int)env_ll("AGB_DAMSM_DW_SPLITS", 16);
  g_opt.attn_fwd_stages = (int)env_ll("AGB_ATTN_FWD_STAGES", 0);
  g_opt.attn_fwd_ctas = (int)env_ll("AGB_ATTN_FWD_CTAS", 0);
  g_opt.attn_bwd_stages = (int)env_ll("AGB_ATTN_BWD_STAGES", 0);
  g_opt.attn_bwd_ctas = (int)env_ll("AGB_ATTN_BWD_CTAS", 0);
}
const Options& options() {
  std::call_once(g_opt_once, load_options);
  return g_opt;
}
int device_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// ---- optional per-kernel timing (CUDA events on the launching stream) -------------------------
struct ProfRec {
  cudaEvent_t a, b;
  int tag;
};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;       // pool, reused across enable() calls
static size_t g_prof_used = 0;
static bool g_prof_on = false;
static long long g_prof_dropped = 0;

int prof_begin(int tag, cudaStream_t st) {
  if (!g_prof_on) return -1;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_prof_used == g_prof.size()) {
    if (g_prof.size() >= 16384) {
      ++g_prof_dropped;
      return -1;
    }
    ProfRec r;
    r.tag = tag;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return -1;
    g_prof.push_back(r);
  }
  const int slot = (int)g_prof_used++;
  g_prof[slot].tag = tag;
  cudaEventRecord(g_prof[slot].a, st);
  return slot;
}
void prof_end(int slot, cudaStream_t st) {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[slot].b, st);
}

}  // namespace agb

extern "C" void agb_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(agb::g_prof_mu);
  agb::g_prof_on = on != 0;
  agb::g_prof_used = 0;
  agb::g_prof_dropped = 0;
}

extern "C" int agb_prof_read(int tag, double* total_ms, long long* launches) {
  std::lock_guard<std::mutex> lk(agb::g_prof_mu);
  double tot = 0.0;
  long long n = 0;
  for (size_t i = 0; i < agb::g_prof_used; ++i) {
    if (agb::g_prof[i].tag != tag) continue;
    cudaError_t e = cudaEventSynchronize(agb::g_prof[i].b);
    if (e != cudaSuccess) return (int)e;
    float ms = 0.f;
    e = cudaEventElapsedTime(&ms, agb::g_prof[i].a, agb::g_prof[i].b);
    if (e != cudaSuccess) return (int)e;
    tot += ms;
    ++n;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = n;
  return agb::g_prof_dropped ? -3 : 0;
}

extern "C" int agb_set_option(const char* name, long long value) {
  if (!name) return agb::fail_arg("null option name");
  agb::options();   // environment first, then the override
  const std::string n(name);
  agb::Options& o = agb::g_opt;
  if (n == "damsm_chunk_mb") o.damsm_chunk_bytes = value << 20;
  else if (n == "damsm_save_mb") o.damsm_save_bytes = value << 20;
  else if (n == "damsm_bwd") o.damsm_bwd = (int)value;
  else if (n == "damsm_uniform_split") o.damsm_uniform_split = (int)value;
  else if (n == "damsm_img_block") o.damsm_img_block = (int)value;
  else if (n == "damsm_dw_splits") o.damsm_dw_splits = (int)value;
  else if (n == "attn_fwd_stages") o.attn_fwd_stages = (int)value;
  else if (n == "attn_fwd_ctas") o.attn_fwd_ctas = (int)value;
  else if (n == "attn_bwd_stages") o.attn_bwd_stages = (int)value;
  else if (n == "attn_bwd_ctas") o.attn_bwd_ctas = (int)value;
  else return agb::fail_arg("unknown option '%s'", name);
  return 0;
}

extern "C" long long agb_launch_count(void) { return agb::g_launches.load(); }

extern "C" int agb_version(void) { return AGB_VERSION; }
extern "C" const char* agb_last_error(void) { return agb::g_err; }
