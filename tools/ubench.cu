// tools/ubench.cu -- B200 microbenchmarks that decide the design of the fused quantize kernel:
//   hbm_read        : streaming read bandwidth (float4 loads), the real ceiling of a read-dominated path
//   l2_reread       : read a buffer of S MiB twice back to back; second-pass GB/s tells whether S stays in L2
//   l2_pipeline     : the access pattern of the planned persistent kernel: tiles of row c+1 streamed from HBM
//                     interleaved with a re-read of row c, with different L2 eviction hints
//   pipes           : issue rates of the conversions / fp64 adds the scan needs (ops per clock per SM)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

enum Hint { H_NONE = 0, H_EVICT_LAST = 1, H_EVICT_FIRST = 2, H_NO_ALLOC = 3 };

template <int HINT>
__device__ __forceinline__ float4 ld4(const float *p, uint64_t pol) {
    float4 r;
    if (HINT == H_NONE) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
template <int HINT>
__device__ __forceinline__ uint64_t make_policy() {
    uint64_t pol = 0;
    if (HINT == H_EVICT_LAST) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    if (HINT == H_EVICT_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (HINT == H_NO_ALLOC) asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// Each CTA reads tiles of 4096 floats (256 threads x 4 float4, coalesced) in a grid-stride loop.
template <int HINT>
__global__ void __launch_bounds__(256) read_kernel(const float *__restrict__ p, int64_t ntiles, float *sink) {
    const uint64_t pol = make_policy<HINT>();
    float acc = 0.f;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const float *base = p + t * 4096 + threadIdx.x * 4;
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = ld4<HINT>(base + q * 1024, pol);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc += fabsf(v[q].x) + fabsf(v[q].y) + fabsf(v[q].z) + fabsf(v[q].w);
    }
    if (acc == 123.456f) *sink = acc;
}

// Pipeline pattern: work item i in [0, rows * 2 * T): phase = i / (2T); inside a phase even items read tile j of
// row (phase) "cold" (pass A) and odd items re-read tile j of row (phase - 1) (pass B).  Row r lives at p + r*rowElems.
template <int HA, int HB>
__global__ void __launch_bounds__(256) pipeline_kernel(const float *__restrict__ p, int64_t rowElems, int64_t T, int rows, unsigned *ticket, float *sink) {
    const uint64_t polA = make_policy<HA>(), polB = make_policy<HB>();
    __shared__ unsigned s_t;
    float acc = 0.f;
    const int64_t total = (int64_t)(rows + 1) * 2 * T;
    while (true) {
        if (threadIdx.x == 0) s_t = atomicAdd(ticket, 1u);
        __syncthreads();
        const int64_t i = s_t;
        __syncthreads();
        if (i >= total) break;
        const int64_t phase = i / (2 * T), r = i - phase * 2 * T, j = r >> 1;
        const bool isB = r & 1;
        const int64_t row = isB ? phase - 1 : phase;
        if (row < 0 || row >= rows) continue;
        const float *base = p + row * rowElems + j * 4096 + threadIdx.x * 4;
        float4 v[4];
        if (isB) {
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = ld4<HB>(base + q * 1024, polB);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = ld4<HA>(base + q * 1024, polA);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) acc += fabsf(v[q].x) + fabsf(v[q].y) + fabsf(v[q].z) + fabsf(v[q].w);
    }
    if (acc == 123.456f) *sink = acc;
}

// ---- pipe issue-rate probes: each thread runs ITER dependent-free ops on 8 independent registers
template <int OP>
__global__ void __launch_bounds__(256) pipe_kernel(float *out, int iters, float seed) {
    float f[8]; double d[8]; int k[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { f[i] = seed + i + threadIdx.x * 1e-3f; d[i] = (double)f[i]; k[i] = i + threadIdx.x; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) { f[i] = fmaf(f[i], 1.0000001f, 0.5f); }                                  // FFMA
            if (OP == 1) { d[i] = d[i] + 1.25; }                                                    // DADD
            if (OP == 2) { d[i] = (double)f[i]; f[i] = f[i] + __double2float_rn(d[i] * 0.0 + 1.0) * 0.f + 1.0f; }  // placeholder (see OP 5,6)
            if (OP == 3) { k[i] = __float2int_rd(f[i]); f[i] = f[i] + 0.75f; asm volatile("" : "+r"(k[i])); }      // F2I + FADD
            if (OP == 4) { k[i] = __funnelshift_l(k[i], k[(i + 1) & 7], 1); }                       // SHF
            if (OP == 5) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(f[i])); f[i] = f[i] + 0.75f; asm volatile("" : "+d"(d[i])); }   // F2F.F64.F32 + FADD
            if (OP == 6) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(d[i])); d[i] = d[i] + 1.25; asm volatile("" : "+f"(f[i])); } // F2F.F32.F64 + DADD
            if (OP == 7) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(f[i])); d[(i + 1) & 7] += d[i]; f[i] = f[i] + 0.75f; }        // cvt + DADD + FADD
            if (OP == 8) { f[i] = __fdiv_rn(f[i], seed); }                                          // IEEE fp32 division
            if (OP == 9) { f[i] = floorf(f[i]) + 0.3f; }                                            // FRND + FADD
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i] + (float)d[i] + (float)k[i];
    if (s == 123.456f) out[0] = s;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main(int argc, char **argv) {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int l2 = prop.l2CacheSize, persist = prop.persistingL2CacheMaxSize;
    int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("{\"probe\":\"device\",\"name\":\"%s\",\"sms\":%d,\"l2_bytes\":%d,\"persisting_l2_max\":%d,\"clock_khz\":%d}\n", prop.name, prop.multiProcessorCount, l2, persist, clk);
    const int sms = prop.multiProcessorCount;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float *sink; CK(cudaMalloc(&sink, 4));
    const int64_t big = (int64_t)4 << 30;   // 4 GiB
    float *buf; CK(cudaMalloc(&buf, big)); CK(cudaMemset(buf, 0x11, big));
    unsigned *ticket; CK(cudaMalloc(&ticket, 4));

    // ---- hbm_read
    for (int ctas_per_sm : {2, 4, 8}) {
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            read_kernel<H_NONE><<<sms * ctas_per_sm, 256>>>(buf, big / 4 / 4096, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            best = fminf(best, time_ms(e0, e1));
        }
        printf("{\"probe\":\"hbm_read\",\"ctas_per_sm\":%d,\"gib\":4,\"ms\":%.4f,\"gbs\":%.1f}\n", ctas_per_sm, best, big / best * 1e-6);
    }
    // ---- l2_reread: pass 1 (cold after flushing with a 1 GiB read elsewhere), pass 2 immediately after
    for (int mib : {8, 16, 32, 48, 64, 80, 96, 112, 128, 160}) {
        const int64_t bytes = (int64_t)mib << 20;
        float best1 = 1e9, best2 = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
            read_kernel<H_NONE><<<sms * 4, 256>>>(buf + (big / 4 / 2), ((int64_t)1 << 30) / 4 / 4096, sink);   // flush
            CK(cudaEventRecord(e0));
            read_kernel<H_NONE><<<sms * 4, 256>>>(buf, bytes / 4 / 4096, sink);
            CK(cudaEventRecord(e1));
            read_kernel<H_NONE><<<sms * 4, 256>>>(buf, bytes / 4 / 4096, sink);
            cudaEvent_t e2; CK(cudaEventCreate(&e2)); CK(cudaEventRecord(e2)); CK(cudaEventSynchronize(e2));
            best1 = fminf(best1, time_ms(e0, e1)); best2 = fminf(best2, time_ms(e1, e2));
            CK(cudaEventDestroy(e2));
        }
        printf("{\"probe\":\"l2_reread\",\"mib\":%d,\"pass1_ms\":%.4f,\"pass1_gbs\":%.1f,\"pass2_ms\":%.4f,\"pass2_gbs\":%.1f}\n", mib, best1, bytes / best1 * 1e-6, best2, bytes / best2 * 1e-6);
    }
    // ---- l2_pipeline
    auto run_pipe = [&](const char *name, auto kern, int mib, int rows, int ctas_per_sm) {
        const int64_t rowElems = ((int64_t)mib << 20) / 4, T = rowElems / 4096;
        float best = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
            read_kernel<H_NONE><<<sms * 4, 256>>>(buf + (big / 4 / 2) + (big / 4 / 4), ((int64_t)1 << 29) / 4 / 4096, sink);   // flush
            CK(cudaMemset(ticket, 0, 4));
            CK(cudaEventRecord(e0));
            kern<<<sms * ctas_per_sm, 256>>>(buf, rowElems, T, rows, ticket, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            best = fminf(best, time_ms(e0, e1));
        }
        const double alg = (double)rows * mib * 1048576.0;
        printf("{\"probe\":\"l2_pipeline\",\"hints\":\"%s\",\"row_mib\":%d,\"rows\":%d,\"ctas_per_sm\":%d,\"ms\":%.4f,\"alg_gbs\":%.1f}\n", name, mib, rows, ctas_per_sm, best, alg / best * 1e-6);
    };
    for (int mib : {4, 16, 32, 64}) {
        const int rows = (mib >= 64) ? 24 : 2048 / mib > 64 ? 64 : 2048 / mib;
        for (int cps : {4, 8}) {
            run_pipe("none/none", pipeline_kernel<H_NONE, H_NONE>, mib, rows, cps);
            run_pipe("last/first", pipeline_kernel<H_EVICT_LAST, H_EVICT_FIRST>, mib, rows, cps);
            run_pipe("none/first", pipeline_kernel<H_NONE, H_EVICT_FIRST>, mib, rows, cps);
            run_pipe("last/none", pipeline_kernel<H_EVICT_LAST, H_NONE>, mib, rows, cps);
        }
    }
    // ---- pipes
    const char *names[10] = {"ffma", "dadd", "skip", "f2i+fadd", "shf", "f2f.f64.f32+fadd", "f2f.f32.f64+dadd", "cvt64+dadd+fadd", "fdiv_rn", "frnd+fadd"};
    for (int op : {0, 1, 3, 4, 5, 6, 7, 8, 9}) {
        const int iters = 4096;
        float best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            switch (op) {
                case 0: pipe_kernel<0><<<sms * 8, 256>>>(sink, iters, 1.5f); break;
                case 1: pipe_kernel<1><<<sms * 8, 256>>>(sink, iters, 1.5f); break;
                case 3: pipe_kernel<3><<<sms * 8, 256>>>(sink, iters, 1.5f); break;
                case 4: pipe_kernel<4><<<sms * 8, 256>>>(sink, iters, 1.5f); break;
                case 5: pipe_kernel<5><<<sms * 8, 256>>>(sink, iters, 1.5f); break;
                case 6: pipe_kernel<6><<<sms * 8, 256>>>(sink, iters, 1.5f); break;
                case 7: pipe_kernel<7><<<sms * 8, 256>>>(sink, iters, 1.5f); break;
                case 8: pipe_kernel<8><<<sms * 8, 256>>>(sink, iters, 1.5f); break;
                case 9: pipe_kernel<9><<<sms * 8, 256>>>(sink, iters, 1.5f); break;
            }
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            best = fminf(best, time_ms(e0, e1));
        }
        const double ops = (double)sms * 8 * 256 * iters * 8;
        printf("{\"probe\":\"pipes\",\"op\":\"%s\",\"ms\":%.4f,\"gops\":%.1f,\"ops_per_sm_per_ns\":%.2f}\n", names[op], best, ops / best * 1e-6, ops / best * 1e-6 / sms);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
