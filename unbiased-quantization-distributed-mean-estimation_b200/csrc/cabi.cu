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

// Optional per-kernel timing (bench.py's roofline leg): while enabled, a CUDA event is recorded on the profiled stream after
// every kernel launch of the library (DME_LAUNCH_CHECK); interval i is kernel i of the calls made since dme_profile_enable.
// Off by default; never enabled inside a timed region; one caller thread.
constexpr int kProfMax = 96;
struct Profile { bool on = false; bool have = false; cudaStream_t st = nullptr; int marks = 0; cudaEvent_t ev[kProfMax + 1]; const char *name[kProfMax]; };
static Profile g_prof;
void prof_launch(const char *name) {
    if (!g_prof.on || g_prof.marks >= kProfMax) return;
    g_prof.name[g_prof.marks] = name;
    cudaEventRecord(g_prof.ev[++g_prof.marks], g_prof.st);
}

}  // namespace dme

using namespace dme;

extern "C" const char *dme_last_error(void) { return g_err; }
extern "C" int dme_version(void) { return 100; }
extern "C" int64_t dme_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" float dme_uniform_x(uint64_t seed, uint64_t client) { return philox_client_uniform(seed, client); }
extern "C" void dme_add_launches(int64_t n) { count_launch((int)n); }

namespace dme {
__global__ void fill_uniforms_kernel(float *xu, int64_t n, const uint64_t *seedp, uint64_t client0) {
    const uint64_t seed = *seedp;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x)
        xu[c] = philox_client_uniform(seed, client0 + (uint64_t)c);
}
__global__ void bump_seed_kernel(uint64_t *seedp) { *seedp += 1; }
}  // namespace dme
extern "C" int dme_fill_uniforms(float *xu, int64_t n, uint64_t *seed_dev, uint64_t client0, int bump, dme_stream_t stream) {
    DME_REQUIRE(xu != nullptr && seed_dev != nullptr && n >= 1, "bad argument");
    const int64_t blocks = (n + 255) / 256;
    fill_uniforms_kernel<<<(unsigned)(blocks < 1024 ? blocks : 1024), 256, 0, (cudaStream_t)stream>>>(xu, n, seed_dev, client0);
    DME_LAUNCH_CHECK("fill_uniforms_kernel");
    if (bump) {
        bump_seed_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(seed_dev);
        DME_LAUNCH_CHECK("bump_seed_kernel");
    }
    return DME_OK;
}

extern "C" int64_t dme_workspace_bytes(int64_t n, int64_t d) {
    if (n < 1 || d < 1) return 0;
    return ws_layout(n, d).total;
}
extern "C" int64_t dme_dir_entries(int64_t n, int64_t d) {
    if (n < 1 || d < 1) return 0;
    return n * ((d + kCodeTile - 1) / kCodeTile);
}
extern "C" int64_t dme_codes_bytes(int64_t n, int64_t d, int64_t m, int expect) {
    if (n < 1 || d < 1 || m < 1) return 0;
    const int64_t T = (d + kCodeTile - 1) / kCodeTile;
    // primary slots (expected width, 128 * w0 bytes per code tile) + overflow space for wider tiles
    const int w0 = expected_width(m, d);
    const int64_t primary = n * T * 128 * w0;
    const int64_t overflow = expect ? (n * T * 128 * (w0 < 32 ? 2 * w0 : 0)) / 4 + 65536 : n * T * 128 * 32;
    return primary + overflow + 4096;
}

extern "C" int dme_profile_enable(int on, dme_stream_t stream) {
    if (on && !g_prof.have) {
        for (auto &e : g_prof.ev) DME_CUDA(cudaEventCreate(&e));
        g_prof.have = true;
    }
    g_prof.on = on != 0; g_prof.st = (cudaStream_t)stream; g_prof.marks = 0;
    if (on) DME_CUDA(cudaEventRecord(g_prof.ev[0], g_prof.st));
    return DME_OK;
}
extern "C" int dme_profile_read(float *ms, int cap) {
    DME_REQUIRE(ms != nullptr && cap >= 1, "bad argument");
    int k = 0;
    for (; k < g_prof.marks && k < cap; ++k) {
        DME_CUDA(cudaEventSynchronize(g_prof.ev[k + 1]));
        DME_CUDA(cudaEventElapsedTime(&ms[k], g_prof.ev[k], g_prof.ev[k + 1]));
    }
    return k;
}
extern "C" const char *dme_profile_name(int i) { return (i >= 0 && i < g_prof.marks) ? g_prof.name[i] : ""; }
