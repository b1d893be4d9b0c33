// Library-level plumbing of libhicgat_sm100.so: version, thread-local error text, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace hicgat {
namespace {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
}  // namespace hicgat

extern "C" int hicgat_version(void) { return 100; }
extern "C" const char* hicgat_last_error(void) { return hicgat::g_err; }
extern "C" uint64_t hicgat_launch_count(void) { return hicgat::g_launches.load(std::memory_order_relaxed); }
