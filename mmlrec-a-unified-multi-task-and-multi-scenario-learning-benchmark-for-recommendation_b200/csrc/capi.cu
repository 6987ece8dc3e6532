// Process-wide bookkeeping of the C ABI: last-error text, launch counter, ABI version.
#include <stdlib.h>
#include <stdarg.h>
#include <atomic>

#include "common.cuh"

namespace mmlrec {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace mmlrec

extern "C" int mmlrec_abi_version(void) { return MMLREC_ABI_VERSION; }
extern "C" const char* mmlrec_last_error(void) { return mmlrec::g_err; }
namespace mmlrec {
int pdl_mode() {   // MMLREC_PDL: 0 off, 1 every patched kernel, 2 all but the GEMM, 3 the GEMM only (default)
  static int v = -1;
  if (v < 0) { const char* e = getenv("MMLREC_PDL"); v = e ? atoi(e) : 3; }
  return v;
}
bool pdl_enabled() { return pdl_mode() == 1 || pdl_mode() == 2; }
bool pdl_enabled_gemm() { return pdl_mode() == 1 || pdl_mode() == 3; }
}  // namespace mmlrec

extern "C" int64_t mmlrec_launch_count(void) { return (int64_t)mmlrec::g_launches.load(); }

// sizeof() of every ABI structure, so the ctypes mirror can be checked without a GPU.
extern "C" int64_t mmlrec_struct_size(int32_t which) {
  switch (which) {
    case 0: return sizeof(MmlrecHyper);
    case 1: return sizeof(MmlrecGemmF32);
    case 2: return sizeof(MmlrecGemmTcDesc);
    case 3: return sizeof(MmlrecGate);
    case 4: return sizeof(MmlrecExpertGrad);
    case 5: return sizeof(MmlrecHead);
    case 6: return sizeof(MmlrecGateLevel);
    default: return -1;
  }
}
