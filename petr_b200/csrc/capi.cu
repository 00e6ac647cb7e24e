// ABI bookkeeping: version, thread-local error string, launch accounting.
#include "common.cuh"

namespace petr {
char* last_error_buf() {
    static thread_local char buf[512] = "";
    return buf;
}
int64_t& launch_counter() {
    static thread_local int64_t n = 0;
    return n;
}
}  // namespace petr

extern "C" int petr_abi_version(void) { return PETR_B200_ABI_VERSION; }
extern "C" const char* petr_last_error_string(void) { return petr::last_error_buf(); }
extern "C" int64_t petr_launch_count(void) { return petr::launch_counter(); }
extern "C" void petr_launch_count_reset(void) { petr::launch_counter() = 0; }
