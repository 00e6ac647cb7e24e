// ABI bookkeeping: version, thread-local error string, launch accounting.
#include "common.cuh"
#include <atomic>
#include <string.h>

namespace petr {
char* last_error_buf() {
    static thread_local char buf[512] = "";
    return buf;
}
int64_t& launch_counter() {
    static thread_local int64_t n = 0;
    return n;
}
namespace tuning {
static std::atomic<int> g_flash_exact_only{0};
static std::atomic<int> g_flash_poly{2};
static std::atomic<int> g_gemm_staged{1};
static std::atomic<int> g_gemm_coalesce{1};
static std::atomic<int> g_gemm_skinny{1};
static std::atomic<int> g_gemm_band{0};
static std::atomic<int> g_gemm_debug{0};
static std::atomic<int> g_pdl{0};
bool flash_exact_only() { return g_flash_exact_only.load(std::memory_order_relaxed) != 0; }
int flash_poly() { return g_flash_poly.load(std::memory_order_relaxed); }
static std::atomic<void*> g_debug_trace{nullptr};
void* debug_trace() { return g_debug_trace.load(std::memory_order_relaxed); }
bool gemm_band() { return g_gemm_band.load(std::memory_order_relaxed) != 0; }
int gemm_debug() { return g_gemm_debug.load(std::memory_order_relaxed); }
bool pdl() { return g_pdl.load(std::memory_order_relaxed) != 0; }
bool gemm_skinny() { return g_gemm_skinny.load(std::memory_order_relaxed) != 0; }
bool gemm_coalesce() { return g_gemm_coalesce.load(std::memory_order_relaxed) != 0; }
bool gemm_staged() { return g_gemm_staged.load(std::memory_order_relaxed) != 0; }
struct Entry { const char* key; std::atomic<int>* v; int lo, hi; };
static const Entry kEntries[] = {
    {"flash_exact_only", &g_flash_exact_only, 0, 1},
    {"flash_poly", &g_flash_poly, 0, 4},
    {"gemm_staged", &g_gemm_staged, 0, 1},
    {"gemm_coalesce", &g_gemm_coalesce, 0, 1},
    {"gemm_skinny", &g_gemm_skinny, 0, 1},
    {"gemm_band", &g_gemm_band, 0, 1},
    {"gemm_debug", &g_gemm_debug, 0, 3},
    {"pdl", &g_pdl, 0, 1},
};
}  // namespace tuning
}  // namespace petr

extern "C" int petr_set_tuning(const char* key, int value) {
    if (key)
        for (const auto& e : petr::tuning::kEntries)
            if (!strcmp(key, e.key)) {
                if (value < e.lo || value > e.hi)
                    return petr::fail(PETR_E_BADARG, "petr_set_tuning: %s must be in [%d, %d]", key, e.lo, e.hi);
                e.v->store(value);
                return 0;
            }
    return petr::fail(PETR_E_BADARG, "petr_set_tuning: unknown key %s", key ? key : "(null)");
}
extern "C" void petr_debug_set_trace(void* device_buffer) { petr::tuning::g_debug_trace.store(device_buffer); }
extern "C" int petr_get_tuning(const char* key) {
    if (key)
        for (const auto& e : petr::tuning::kEntries)
            if (!strcmp(key, e.key)) return e.v->load();
    return -1;
}

extern "C" int petr_abi_version(void) { return PETR_B200_ABI_VERSION; }
extern "C" const char* petr_last_error_string(void) { return petr::last_error_buf(); }
extern "C" int64_t petr_launch_count(void) { return petr::launch_counter(); }
extern "C" void petr_launch_count_reset(void) { petr::launch_counter() = 0; }
