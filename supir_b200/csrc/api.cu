// supir_b200 — library-level C ABI: error string, version, launch counter.
#include "common.cuh"
#include "supir_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace supir {

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace supir

extern "C" const char* supir_last_error(void) { return supir::g_err; }
extern "C" int supir_version(void) { return 100; }
extern "C" long long supir_launch_count(void) { return supir::g_launches.load(); }
extern "C" void supir_reset_launch_count(void) { supir::g_launches.store(0); }
