// tools/ubench2.cu -- issue-rate probes for the fixed-point scan (round 2).  Every op is inline PTX inside an unrolled
// loop; check the SASS (cuobjdump -sass tools/ubench2 | grep -c F2I ...) before trusting a number.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench2 tools/ubench2.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

typedef unsigned long long u64;

template <int OP>
__global__ void __launch_bounds__(256) pipe_kernel(u64 *out, int iters, float seed) {
    float f[8]; double d[8]; uint32_t k[8]; u64 q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { f[i] = seed + i + threadIdx.x * 1e-3f; d[i] = (double)f[i]; k[i] = i + threadIdx.x; q[i] = k[i] * 77ull; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(1.0000001f), "f"(0.5f));
            if (OP == 1) asm volatile("add.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(1.25));
            if (OP == 2) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(f[i])); f[i] = __uint_as_float(((uint32_t)__double2hiint(d[i]) & 0x3fffffu) | 0x3f800000u); }
            if (OP == 3) { asm volatile("cvt.rni.u32.f32 %0, %1;" : "=r"(k[i]) : "f"(f[i])); f[i] = __uint_as_float((k[i] & 0x3fffffu) | 0x4b000000u); }
            if (OP == 4) { asm volatile("cvt.rni.u64.f32 %0, %1;" : "=l"(q[i]) : "f"(f[i])); f[i] = __uint_as_float(((uint32_t)q[i] & 0x3fffffu) | 0x4b000000u); }
            if (OP == 5) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(q[i]) : "r"(k[i]), "r"(k[(i + 1) & 7]));
            if (OP == 6) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %1;" : "+r"(k[i]), "+r"(k[(i + 4) & 7]) : "r"(k[(i + 1) & 7]));
            if (OP == 7) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[i]) : "l"(0x3f8000003f800000ull), "l"(0x3f0000003f000000ull));
            if (OP == 8) asm volatile("shf.l.wrap.b32 %0, %1, %0, 2;" : "+r"(k[i]) : "r"(k[(i + 1) & 7]));
            if (OP == 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0xfe;" : "+r"(k[i]) : "r"(k[(i + 1) & 7]), "r"(k[(i + 2) & 7]));
            if (OP == 10) { asm volatile("cvt.rni.u64.f32 %0, %1;" : "=l"(q[i]) : "f"(fabsf(f[i]))); f[i] = __uint_as_float(((uint32_t)(q[i] >> 32) & 0x3fffffu) | 0xcb000000u); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(q[(i + 4) & 7]) : "r"((uint32_t)q[i]), "r"(k[0])); }
            if (OP == 11) { double t; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(fabsf(f[i]))); asm volatile("add.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(t)); f[i] = __uint_as_float(((uint32_t)__double2loint(d[i]) & 0x3fffffu) | 0xbf800000u); }
            if (OP == 12) asm volatile("add.rz.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(0x4b0000004b000000ull));
            if (OP == 13) asm volatile("mul.lo.u32 %0, %0, %1;" : "+r"(k[i]) : "r"(k[(i + 1) & 7]));
            if (OP == 14) asm volatile("add.u64 %0, %0, %1;" : "+l"(q[i]) : "l"(q[(i + 1) & 7]));
            if (OP == 15) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]), "f"(f[(i + 2) & 7]));
            if (OP == 16) { asm volatile("cvt.rzi.u32.f32 %0, %1;" : "=r"(k[i]) : "f"(f[i])); f[i] = __uint_as_float((k[i] & 0x3fffffu) | 0x4b000000u); }
            if (OP == 17) { asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[i]) : "r"(k[i])); k[i] = __float_as_uint(f[i]) ^ 0x5555u; }
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (u64)f[i] + (u64)d[i] + k[i] + q[i];
    if (s == 123456789ull) out[0] = s;
}

// ---- composite: the planned B-phase + C-phase on a shared-memory resident tile (no global traffic): instruction-side bound.
// 128 threads, thread t owns 32 consecutive floats of a 4096-float tile (linear layout here; bank conflicts are not the point,
// so the tile is stored "thread-interleaved": element j of thread t at [(j/4)*128 + t]*4 + j%4 -> conflict-free 128-bit access).
__device__ __forceinline__ float4 lds128(uint32_t a) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts128u(uint32_t a, uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ uint4 lds128u(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ u64 f2_pack(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 f2_mul(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 f2_fma(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// VARIANT 0: F2I.U64 of |mp * 2^32| (hi = floor, lo = fraction);  1: magic floor + F2I.U32 of the scaled fraction
template <int VARIANT, int CPHASE>
__global__ void __launch_bounds__(128) composite_kernel(u64 *out, int iters, float rcp, float D, float M, uint32_t one) {
    __shared__ __align__(16) float tile[4096];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 4096; i += 128) tile[i] = (float)((i * 2654435761u) >> 8) * (1.0f / 16777216.0f) * 3.0f - 1.5f;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tile) + tid * 16u;
    u64 total = 0;
    const u64 R2 = f2_pack(rcp, rcp), ND = f2_pack(-D, -D), M2 = f2_pack(M, M);
    for (int it = 0; it < iters; ++it) {
        // ---------------- B-phase
        uint32_t lo[32];
        uint32_t sg0 = 0, sg1 = 0, hior = 0;
        u64 sum = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 v = lds128(base + q * 2048u);
            const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float xa = x[2 * h], xb = x[2 * h + 1];
                if (q < 4) { sg0 = __funnelshift_l(__float_as_uint(xa), sg0, 2); sg0 = __funnelshift_l(__float_as_uint(xb), sg0, 2); }
                else { sg1 = __funnelshift_l(__float_as_uint(xa), sg1, 2); sg1 = __funnelshift_l(__float_as_uint(xb), sg1, 2); }
                const u64 xx = f2_pack(xa, xb);
                const u64 q0 = f2_mul(xx, R2);
                const u64 rem = f2_fma(q0, ND, xx);
                const u64 pq = f2_fma(rem, R2, q0);
                const u64 mp = f2_mul(M2, pq);
                float ma, mb; f2_unpack(mp, ma, mb);
                if (VARIANT == 0) {
                    u64 fa, fb;
                    asm("cvt.rni.u64.f32 %0, %1;" : "=l"(fa) : "f"(fabsf(ma)));
                    asm("cvt.rni.u64.f32 %0, %1;" : "=l"(fb) : "f"(fabsf(mb)));
                    lo[4 * q + 2 * h] = (uint32_t)fa; lo[4 * q + 2 * h + 1] = (uint32_t)fb;
                    asm("lop3.b32 %0, %0, %1, %2, 0xfe;" : "+r"(hior) : "r"((uint32_t)(fa >> 32)), "r"((uint32_t)(fb >> 32)));
                } else {
                    // magic floor in the scaled domain: C = 2^55 (values < 2^55), fraction scaled by 2^32 -> F2I.U32
                    const float ta = __fadd_rz(fabsf(ma), 36028797018963968.0f), tb = __fadd_rz(fabsf(mb), 36028797018963968.0f);
                    const float fla = ta - 36028797018963968.0f, flb = tb - 36028797018963968.0f;
                    const float fra = fabsf(ma) - fla, frb = fabsf(mb) - flb;
                    uint32_t ia, ib;
                    asm("cvt.rni.u32.f32 %0, %1;" : "=r"(ia) : "f"(fra));
                    asm("cvt.rni.u32.f32 %0, %1;" : "=r"(ib) : "f"(frb));
                    lo[4 * q + 2 * h] = ia; lo[4 * q + 2 * h + 1] = ib;
                    asm("lop3.b32 %0, %0, %1, %2, 0xfe;" : "+r"(hior) : "r"(__float_as_uint(fla)), "r"(__float_as_uint(flb)));
                }
                asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(sum) : "r"(lo[4 * q + 2 * h]), "r"(one));
                asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(sum) : "r"(lo[4 * q + 2 * h + 1]), "r"(one));
            }
        }
        // warp scan of the thread sums (64-bit)
        u64 incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u64 up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (CPHASE) {
            // ---------------- C-phase: carry walk from the thread's start offset
            uint32_t acc = (uint32_t)(incl - sum) + (uint32_t)it * 0x9e3779b9u; total += (incl >> 32);
            uint32_t rb0 = 0, rb1 = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %1;" : "+r"(acc), "+r"(rb0) : "r"(lo[j]));
#pragma unroll
            for (int j = 16; j < 32; ++j) asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %1;" : "+r"(acc), "+r"(rb1) : "r"(lo[j]));
            // 16 compact bits (first coordinate on top) -> 2-bit fields, first coordinate at the bottom
            uint32_t w0 = __brev(rb0) >> 16, w1 = __brev(rb1) >> 16;
            w0 = (w0 | (w0 << 8)) & 0x00ff00ffu; w0 = (w0 | (w0 << 4)) & 0x0f0f0f0fu; w0 = (w0 | (w0 << 2)) & 0x33333333u; w0 = (w0 | (w0 << 1)) & 0x55555555u;
            w1 = (w1 | (w1 << 8)) & 0x00ff00ffu; w1 = (w1 | (w1 << 4)) & 0x0f0f0f0fu; w1 = (w1 | (w1 << 2)) & 0x33333333u; w1 = (w1 | (w1 << 1)) & 0x55555555u;
            total += (w0 | (sg0 & 0xaaaaaaaau)) + (w1 | (sg1 & 0xaaaaaaaau)) + hior;
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(base), "r"(((w0 ^ w1) & 0x007fffffu) | 0x3f000000u) : "memory");
        } else {
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) x ^= lo[j];
            total += incl + sg0 + sg1 + hior;
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(base), "r"((x & 0x007fffffu) | 0x3f000000u) : "memory");
        }
    }
    if (total == 123456789ull) out[0] = total;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    u64 *sink; CK(cudaMalloc(&sink, 8));
    const char *names[18] = {"ffma", "dadd", "f2f.f64.f32", "f2i.u32", "f2i.u64", "imad.wide", "addcc+addc(2)", "ffma2", "shf", "lop3", "f2i.u64+imad.wide(2)",
                             "f2f+dadd(2)", "fadd2.rz", "imul", "iadd64", "fmnmx3", "f2i.rz.u32", "i2f"};
    for (int op = 0; op < 18; ++op) {
        const int iters = 2048;
        float best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            switch (op) {
#define C(n) case n: pipe_kernel<n><<<sms * 8, 256>>>(sink, iters, 1.5f); break;
                C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15) C(16) C(17)
#undef C
            }
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            best = fminf(best, time_ms(e0, e1));
        }
        const double ops = (double)sms * 8 * 256 * iters * 8;
        printf("{\"probe\":\"pipes2\",\"op\":\"%s\",\"ms\":%.4f,\"asm_stmts_per_sm_per_clk\":%.2f}\n", names[op], best, ops / best * 1e-6 / sms / 1.965);
    }
    for (int v = 0; v < 4; ++v) {
        for (int cps : {4, 8}) {
            const int iters = 512;
            float best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaEventRecord(e0));
                if (v == 0) composite_kernel<0, 1><<<sms * cps, 128>>>(sink, iters, 1.0f / 1337.0f, 1337.0f, 358.0f * 4294967296.0f, 1u);
                if (v == 1) composite_kernel<1, 1><<<sms * cps, 128>>>(sink, iters, 1.0f / 1337.0f, 1337.0f, 358.0f * 4294967296.0f, 1u);
                if (v == 2) composite_kernel<0, 0><<<sms * cps, 128>>>(sink, iters, 1.0f / 1337.0f, 1337.0f, 358.0f * 4294967296.0f, 1u);
                if (v == 3) composite_kernel<1, 0><<<sms * cps, 128>>>(sink, iters, 1.0f / 1337.0f, 1337.0f, 358.0f * 4294967296.0f, 1u);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                best = fminf(best, time_ms(e0, e1));
            }
            const double coords = (double)sms * cps * iters * 4096.0;
            printf("{\"probe\":\"composite\",\"variant\":\"%s\",\"ctas_per_sm\":%d,\"ms\":%.4f,\"gcoords_per_s\":%.1f,\"ms_for_2^31\":%.3f}\n",
                   v == 0 ? "u64+C" : v == 1 ? "magic+u32+C" : v == 2 ? "u64 B only" : "magic+u32 B only", cps, best, coords / best * 1e-6, 2147483648.0 / (coords / best * 1e-6) * 1e-6);
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
