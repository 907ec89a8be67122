// cabi.cu -- error plumbing and the small host-only entry points of the C ABI (include/dme_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "type_quantize.cuh"

namespace dme {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return DME_ECUDA;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace dme

using namespace dme;

extern "C" const char *dme_last_error(void) { return g_err; }
extern "C" int dme_version(void) { return 100; }
extern "C" int64_t dme_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" float dme_uniform_x(uint64_t seed, uint64_t client) { return philox_client_uniform(seed, client); }

extern "C" int64_t dme_workspace_bytes(int64_t n, int64_t d) {
    if (n < 1 || d < 1) return 0;
    return ws_layout(n, d).total;
}
extern "C" int64_t dme_dir_entries(int64_t n, int64_t d) {
    if (n < 1 || d < 1) return 0;
    return n * ((d + kTile - 1) / kTile);
}
extern "C" int64_t dme_codes_bytes(int64_t n, int64_t d, int64_t m, int expect) {
    if (n < 1 || d < 1 || m < 1) return 0;
    const int64_t T = (d + kTile - 1) / kTile;
    int w = 32;
    if (expect) {
        // expected largest magnitude in a 4096-tile ~ (m/d) * (max|x| / mean|x|); light tails: factor ~6
        const double ell = (double)m / (double)d;
        w = 2;
        while (w < 32 && 6.0 * ell + 1.0 >= (double)(1u << (w - 1))) w <<= 1;
    }
    int64_t bytes = n * T * 512 * w;
    if (expect && w < 32) bytes += bytes / 2;
    return bytes + 4096;
}
