// tools/ubench3.cu -- bandwidth ceiling of the fused pass-A / pass-B access pattern (round 2).
// Persistent CTAs draw tickets; ticket i copies tile i of the matrix ("pass B": row c = i / T was streamed T + lead tickets
// earlier, so it should come from L2) and tile i + T + lead ("pass A": first touch, from HBM, L2 evict_last) with
// cp.async.bulk into a ring of shared-memory buffers.  No compute: one thread per CTA issues and waits.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench3 tools/ubench3.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n@p bra LD;\nbra LW;\nLD:\n}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ uint64_t policy(int kind) {
    uint64_t p = 0;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}

constexpr int kTileBytes = 16384;
// mode 0: fused (B tile + A tile per ticket), 1: B only (single stream from HBM), NB ring buffers
__global__ void __launch_bounds__(32) pattern_kernel(const char *base, long long T, long long rows, long long lead, int mode, int nbuf, int polA, int polB,
                                                       unsigned *ticket, float *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)nbuf * kTileBytes);
    const uint64_t pA = policy(polA), pB = policy(polB);
    if (threadIdx.x != 0) return;
    for (int i = 0; i < nbuf; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long nT = rows * T;
    const long long total = (mode & 1) == 0 ? nT + T + lead : nT;      // tickets
    int head = 0, tail = 0, inflight = 0;
    uint32_t phase_bits = 0;
    float acc = 0.f;
    bool done = false;
    long long next_static = blockIdx.x;
    while (!done || inflight > 0) {
        // fill the ring
        while (!done && inflight + ((mode & 1) == 0 ? 2 : 1) <= nbuf) {
            const long long i = (mode & 2) ? next_static : (long long)atomicAdd(ticket, 1u);
            next_static += gridDim.x;
            if (i >= total) { done = true; break; }
            const long long b = (mode & 1) == 0 ? i - T - lead : i;      // pass-B tile index (global tile order)
            if (b >= 0 && b < nT) {
                mbar_expect_tx(&bars[head], kTileBytes);
                bulk_g2s(smem_u32(smem + (size_t)head * kTileBytes), base + b * kTileBytes, kTileBytes, &bars[head], pB);
                head = head + 1 == nbuf ? 0 : head + 1; ++inflight;
            }
            if ((mode & 1) == 0 && i < nT) {
                mbar_expect_tx(&bars[head], kTileBytes);
                bulk_g2s(smem_u32(smem + (size_t)head * kTileBytes), base + i * kTileBytes, kTileBytes, &bars[head], pA);
                head = head + 1 == nbuf ? 0 : head + 1; ++inflight;
            }
        }
        if (inflight > 0) {
            mbar_wait(&bars[tail], (phase_bits >> tail) & 1u);
            phase_bits ^= 1u << tail;
            acc += *reinterpret_cast<volatile float *>(smem + (size_t)tail * kTileBytes + 64);
            tail = tail + 1 == nbuf ? 0 : tail + 1; --inflight;
        }
    }
    if (acc == 123.456f) *sink = acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float *sink; CK(cudaMalloc(&sink, 4));
    const long long rows = 48; long long rowBytes = 64ll << 20, T = rowBytes / kTileBytes;
    char *buf; CK(cudaMalloc(&buf, rows * (64ll << 20) + (1ll << 30))); CK(cudaMemset(buf, 0x11, rows * (64ll << 20) + (1ll << 30)));
    unsigned *ticket; CK(cudaMalloc(&ticket, 4));
    CK(cudaFuncSetAttribute(pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    auto run = [&](int mode, int cps, int nbuf, long long lead, int polA, int polB) {
        const size_t dyn = (size_t)nbuf * kTileBytes + 64 * 8;
        float best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaMemsetAsync(buf + rows * (64ll << 20), 0x22, 1ll << 30));     // flush L2
            CK(cudaMemset(ticket, 0, 4));
            CK(cudaEventRecord(e0));
            pattern_kernel<<<sms * cps, 32, dyn>>>(buf, T, rows, lead, mode, nbuf, polA, polB, ticket, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            best = fminf(best, time_ms(e0, e1));
        }
        const double alg = (double)rows * rowBytes;
        printf("{\"probe\":\"ab_pattern\",\"mode\":\"%s\",\"ctas_per_sm\":%d,\"nbuf\":%d,\"lead\":%lld,\"polA\":%d,\"polB\":%d,\"ms\":%.4f,\"alg_gbs\":%.1f,\"ms_for_8GiB\":%.3f}\n",
               mode == 0 ? "fused" : mode == 1 ? "single" : mode == 2 ? "fused_static" : "single_static", cps, nbuf, lead, polA, polB, best, alg / best * 1e-6, 8589934592.0 / (alg / best * 1e-6) * 1e-6);
        fflush(stdout);
    };
    for (long long mib : {64ll, 32ll, 48ll}) {
        rowBytes = mib << 20; T = rowBytes / kTileBytes;
        printf("{\"probe\":\"row_mib\",\"mib\":%lld}\n", mib);
        for (int cps : {2, 3}) {
            const long long G = (long long)sms * cps;
            const long long lead_static = (G - (T % G)) % G;         // T + lead is a multiple of G: the same CTA does pass A and pass B of a tile
            run(0, cps, 12 / cps, lead_static, 1, 2);
            run(2, cps, 12 / cps, lead_static, 1, 2);
            run(2, cps, 12 / cps, lead_static + G, 1, 2);
            run(2, cps, 12 / cps, lead_static + 1, 1, 2);              // same order, different CTA
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
