// common.cuh -- shared device/host helpers for libdme_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dme_b200.h"

namespace dme {

constexpr int kTile = 4096;           // coordinates per CTA tile of the row kernels (l1, Reznik passes, literal rows, FWHT blocks)
constexpr int kCodeTile = DME_TILE;   // coordinates per tile of the packed code: one directory entry, one field width
constexpr int kCodeChunks = kCodeTile / 16;   // 16-coordinate chunks per code tile (word q of chunk j at [q * kCodeChunks + j])
constexpr int kThreads = 256;        // threads per tile CTA
constexpr int kEpt = 16;             // coordinates per thread (blocked: thread t owns [16t, 16t+16))
constexpr int kWarps = kThreads / 32;

// ------------------------------------------------------------------ error plumbing (host)
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
void count_launch(int n = 1);
void prof_launch(const char *name);      // cabi.cu: per-kernel CUDA events while dme_profile_enable is on
#define DME_CUDA(expr)                                             \
    do {                                                           \
        cudaError_t _e = (expr);                                   \
        if (_e != cudaSuccess) return ::dme::cuda_fail(_e, #expr); \
    } while (0)
#define DME_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            ::dme::set_error(__VA_ARGS__); \
            return DME_EINVAL;            \
        }                                 \
    } while (0)
#define DME_LAUNCH_CHECK(name)                                          \
    do {                                                                \
        ::dme::count_launch();                                          \
        ::dme::prof_launch(name);                                       \
        cudaError_t _e = cudaGetLastError();                            \
        if (_e != cudaSuccess) return ::dme::cuda_fail(_e, "launch " name); \
    } while (0)

// ------------------------------------------------------------------ Philox4x32-10 (counter-based RNG)
// key = (seed lo, seed hi); counter = (c0, c1, c2, c3).  Same code on host and device.
struct Philox4 { uint32_t x, y, z, w; };
__host__ __device__ inline uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
__host__ __device__ inline Philox4 philox4x32_10(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}
constexpr uint32_t kStreamX = 0x584D44u;     // "XMD": per-client uniform X_c (AS:634)
constexpr uint32_t kStreamDiag = 0x444947u;  // "DIG": Rademacher diagonal (AS:117-120)
constexpr uint32_t kStreamDrive = 0x445256u; // "DRV": DRIVE's D (AS:735)
constexpr uint32_t kStreamScalar = 0x534353u;// "SCS": Scalar SQ uniforms (AS:783)
__host__ __device__ inline float u24_to_unit(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }
__host__ __device__ inline float philox_client_uniform(uint64_t seed, uint64_t client) {
    Philox4 p = philox4x32_10(seed, (uint32_t)client, (uint32_t)(client >> 32), 0u, kStreamX);
    return u24_to_unit(p.x);
}
// +-1 for coordinate i of stream `stream` (4 coordinates share one Philox block).
__host__ __device__ inline float philox_sign(uint64_t seed, uint64_t i, uint32_t stream, uint32_t aux = 0) {
    uint64_t blk = i >> 7;                       // 128 sign bits per Philox call
    Philox4 p = philox4x32_10(seed, (uint32_t)blk, (uint32_t)(blk >> 32), aux, stream);
    uint32_t word = ((i >> 5) & 3) == 0 ? p.x : ((i >> 5) & 3) == 1 ? p.y : ((i >> 5) & 3) == 2 ? p.z : p.w;
    return ((word >> (i & 31)) & 1u) ? 1.0f : -1.0f;
}

#ifdef __CUDACC__
// ------------------------------------------------------------------ memory-order helpers
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_f64(const double *p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed_s32(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// streaming 128-bit load that does not pollute L1
__device__ __forceinline__ float4 ldg_stream_f4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// ------------------------------------------------------------------ block-level helpers (kThreads threads)
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Fixed-association block sum (xor-butterfly inside a warp, then warps in index order). All threads get it.
__device__ __forceinline__ double block_sum_f64(double v, double *smem /* kWarps */) {
    v = warp_sum_f64(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    double t = smem[0];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) t += smem[w];
    return t;
}
#endif  // __CUDACC__

}  // namespace dme
