// type_quantize.cu -- unbiased type-based L1-ball quantizer (AS:609-641), packed-code emit, decode + mean
// (AS:640 + ND:133-147).  Hand-written for sm_100a; no CPU fallback.
//
// Numerics contract (oracle/dme_oracle.c header): fp64-accumulated L1 rounded to fp32; the elementwise chain
// of AS:625-631 in fp32 with IEEE division; prefix of fractional parts accumulated in fp64 and rounded to
// fp32 (what torch.cumsum does on CPU); r_i = [floor(c_i - X) - floor(c_{i-1} - X) == 1] in fp32.
//
// Kernel structure (one row = one client vector, tiles of 4096 coordinates, thread t owns 16 consecutive
// coordinates so that the in-thread part of the prefix is a plain sequential sum):
//   l1_kernel      : per-tile fp64 partial sums; the LAST tile of a row to finish reduces the partials in
//                    index order and publishes the row constants (deterministic, no float atomics).
//   scan_kernel<E> : single pass over the row in ticket order with a decoupled look-back across tiles.
//                    The exclusive prefix of a tile is the canonical left-to-right sum of tile aggregates
//                    (start at the nearest published inclusive prefix, add the aggregates after it in
//                    order), so the result does not depend on timing.  floor(c - X) of a tile's last
//                    coordinate is handed to the next tile through the descriptor (each a_i is computed once).
//   decode_mean_kernel : tile-major over d, clients in order in registers, one write of the mean.
#include <cstdlib>

#include "type_quantize.cuh"

namespace dme {

// ------------------------------------------------------------------ tile loads
// blocked: thread t owns coordinates [16t, 16t+16) of the tile (the scan needs consecutive coordinates per thread)
__device__ __forceinline__ void load_tile(const float *__restrict__ row, int64_t d, int64_t tile0, float (&x)[kEpt]) {
    const int64_t i0 = tile0 + (int64_t)threadIdx.x * kEpt;
    if (i0 + kEpt <= d) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 v = ldg_stream_f4(row + i0 + 4 * q);
            x[4 * q + 0] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kEpt; ++j) x[j] = (i0 + j < d) ? row[i0 + j] : 0.0f;
    }
}
// striped: fully coalesced 128-bit loads; used where the order inside the tile does not matter (L1 norm)
__device__ __forceinline__ void load_tile_striped(const float *__restrict__ row, int64_t d, int64_t tile0, float (&x)[kEpt]) {
    if (tile0 + kTile <= d) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 v = ldg_stream_f4(row + tile0 + q * 1024 + 4 * threadIdx.x);
            x[4 * q + 0] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int64_t i = tile0 + q * 1024 + 4 * threadIdx.x + e;
                x[4 * q + e] = (i < d) ? row[i] : 0.0f;
            }
    }
}

// ------------------------------------------------------------------ K1: L1 norms
// grid (T, n).  partial[c*T + t] = sum over the tile of |x| in fp64 (fixed association).
__global__ void __launch_bounds__(kThreads)
l1_kernel(const float *__restrict__ X, int64_t d, int64_t ld, int64_t m, int64_t T,
          double *__restrict__ partial, uint32_t *__restrict__ a_done, RowConst *__restrict__ consts,
          const float *__restrict__ x_inject, const float *__restrict__ l1_inject, uint64_t seed,
          uint64_t client0, float *__restrict__ l1_out) {
    __shared__ double s_red[kWarps];
    __shared__ uint32_t s_last;
    const int64_t c = blockIdx.y, t = blockIdx.x;
    const float *row = X + c * ld;
    float x[kEpt];
    load_tile_striped(row, d, t * kTile, x);
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < kEpt; ++j) s += (double)fabsf(x[j]);
    s = block_sum_f64(s, s_red);
    if (threadIdx.x == 0) {
        partial[c * T + t] = s;
        __threadfence();
        s_last = (atomicAdd(&a_done[c], 1u) == (uint32_t)(T - 1)) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // Last tile of the row: reduce the T partials in a fixed order (thread-strided, then block tree).
    const volatile double *pp = partial + c * T;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < T; i += kThreads) acc += pp[i];
    acc = block_sum_f64(acc, s_red);
    if (threadIdx.x == 0) {
        RowConst rc;
        rc.L1f = l1_inject ? l1_inject[c] : (float)acc;               // AS:624
        rc.D = __fadd_rn(rc.L1f, 1e-12f);                             // AS:625
        rc.mf = (float)m;
        rc.X = x_inject ? x_inject[c] : philox_client_uniform(seed, client0 + (uint64_t)c);   // AS:634
        rc.rcpD = __frcp_rn(rc.D);
        uint32_t fl = 0;
        // The fast chain (Markstein division, magic-number floor) is proven for these operand ranges only;
        // anything else takes the IEEE-div / floorf instantiation.  See DESIGN.md "Exactness of the fast chain".
        if (!(rc.D >= 9.5367431640625e-07f && rc.D <= 1.2676506e30f)) fl |= kRowExact;            // 2^-20 .. 2^100
        if ((__float_as_uint(rc.D) & 0x7fffffu) == 0x7fffffu) fl |= kRowExact;                      // 1/D rounding exception
        if (!(rc.X == 0.0f || (rc.X >= 5.9604644775390625e-08f && rc.X < 1.0f))) fl |= kRowExact;  // X on torch.rand's grid
        if (!(rc.mf <= 4194304.0f) || l1_inject) fl |= kRowGuardFloor;                              // m*p may reach 2^23
        rc.flags = fl;
        int lg = 0;
        while (((int64_t)1 << lg) < d) ++lg;
        rc.qshift = min(50, 62 - lg);
        rc.pad0 = 0;
        rc.q_up = scalbn(1.0, rc.qshift);
        rc.q_dn = scalbn(1.0, -rc.qshift);
        rc.pad1[0] = rc.pad1[1] = 0.0;
        consts[c] = rc;
        if (l1_out) l1_out[c] = rc.L1f;
    }
}

// ------------------------------------------------------------------ look-back (warp 0 of a tile CTA)
// Standard decoupled look-back over int64 fixed-point aggregates: integer addition is associative, so the
// exclusive prefix is the same whichever mixture of aggregates / inclusive prefixes happens to be visible.
__device__ __forceinline__ long long lookback_exclusive(TileDesc *rowdesc, int64_t t, int lane) {
    long long P = 0;
    int64_t top = t - 1;                       // nearest predecessor not yet accounted for
    while (top >= 0) {
        const int64_t idx = top - lane;
        uint32_t st;
        while (true) {
            st = idx >= 0 ? ld_acquire_u32(&rowdesc[idx].state) : 2u;          // virtual tiles < 0: inclusive 0
            const unsigned incl = __ballot_sync(0xffffffffu, st >= 2u);
            const unsigned zero = __ballot_sync(0xffffffffu, st == 0u);
            const int f = incl ? (__ffs(incl) - 1) : 32;                       // nearest inclusive in the window
            const unsigned need = (f >= 31) ? 0xffffffffu : ((1u << (f + 1)) - 1u);
            if ((zero & need) == 0u) {
                long long v = 0;
                if (lane < f) v = *reinterpret_cast<volatile long long *>(&rowdesc[idx].aggregate);
                else if (lane == f && idx >= 0) v = *reinterpret_cast<volatile long long *>(&rowdesc[idx].inclusive);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                P += v;
                if (f < 32) return P;
                break;
            }
            __nanosleep(20);
        }
        top -= 32;
    }
    return P;
}

enum Emit { kEmitArrays = 0, kEmitPacked = 1 };

// Arguments of the persistent quantize kernel.
struct StreamArgs {
    const float *X; int64_t d, ld, T, n, m;
    RowConst *consts; TileDesc *desc; WsHeader *hdr; double *partial; uint32_t *a_done; uint32_t *row_ready;
    const float *x_inject; const float *l1_inject; uint64_t seed, client0; float *l1_out;
    int64_t lag, total_items;          // the B stream trails the A stream by `lag` tiles
    // arrays
    int32_t *k_out; uint8_t *sgn_out; float *deq_out; int64_t ld_out;
    // packed
    uint32_t *codes; int64_t codes_bytes; uint64_t *dir; int W0;
    unsigned long long arena_base16;   // first 16-byte unit behind the primary slots
    int dbg;
};

// ---- async-copy / mbarrier primitives (TMA 1-D bulk copy, SASS UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}

// AS:625-631 for one coordinate.  EXACT: IEEE division + floorf.  Fast: x/D by Markstein's correction of x*rcp
// (correctly rounded for D in [2^-20, 2^100], 1/D correctly rounded, quotient normal) and floor by adding 2^23
// toward zero (exact for 0 <= mp < 2^23).
template <bool EXACT>
__device__ __forceinline__ void chain(float x, const RowConst &rc, float &flf, float &fr) {
    float mp;
    if (EXACT) {
        const float v = __fdiv_rn(x, rc.D);
        mp = __fmul_rn(rc.mf, fabsf(v));
        flf = floorf(mp);
    } else {
        const float ax = fabsf(x);
        const float q0 = __fmul_rn(ax, rc.rcpD);
        const float rem = __fmaf_rn(-q0, rc.D, ax);
        const float p = __fmaf_rn(rem, rc.rcpD, q0);
        mp = __fmul_rn(rc.mf, p);
        flf = __fsub_rn(__fadd_rz(mp, 8388608.0f), 8388608.0f);
    }
    fr = __fsub_rn(mp, flf);
}

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
struct TileScratch {
    unsigned long long tacc[8];
    double wtot[kWarps];
    double red[kWarps];
    double P;
    long long Pq;
    int alast[kWarps];
    PackScratch pack;
    uint32_t flag;
};

// Row constants from the finished L1 reduction (run by thread 0 of the CTA that completed the row).
__device__ __forceinline__ void make_row_const(const StreamArgs &a, int64_t c, double l1sum) {
    RowConst rc;
    rc.L1f = a.l1_inject ? a.l1_inject[c] : (float)l1sum;           // AS:624
    rc.D = __fadd_rn(rc.L1f, 1e-12f);                               // AS:625
    rc.mf = (float)a.m;
    rc.X = a.x_inject ? a.x_inject[c] : philox_client_uniform(a.seed, a.client0 + (uint64_t)c);   // AS:634
    rc.rcpD = __frcp_rn(rc.D);
    uint32_t fl = 0;
    // The fast chain (Markstein division, magic-number floor) is proven for these operand ranges only;
    // anything else takes the IEEE-div / floorf instantiation.  See DESIGN.md "Exactness of the fast chain".
    if (!(rc.D >= 9.5367431640625e-07f && rc.D <= 1.2676506e30f)) fl |= kRowExact;            // 2^-20 .. 2^100
    if ((__float_as_uint(rc.D) & 0x7fffffu) == 0x7fffffu) fl |= kRowExact;                      // 1/D rounding exception
    if (!(rc.X == 0.0f || (rc.X >= 5.9604644775390625e-08f && rc.X < 1.0f))) fl |= kRowExact;  // X on torch.rand's grid
    if (!(rc.mf <= 4194304.0f) || a.l1_inject) fl |= kRowGuardFloor;                            // m*p may reach 2^23
    rc.flags = fl;
    int lg = 0;
    while (((int64_t)1 << lg) < a.d) ++lg;
    rc.qshift = min(50, 62 - lg);
    rc.pad0 = 0;
    rc.q_up = scalbn(1.0, rc.qshift);
    rc.q_dn = scalbn(1.0, -rc.qshift);
    rc.pad1[0] = rc.pad1[1] = 0.0;
    a.consts[c] = rc;
    if (a.l1_out) a.l1_out[c] = rc.L1f;
}

// value of tile-local coordinate e (0..4095) from the staged tile; beyond `copied` floats fall back to global / zero
__device__ __forceinline__ float staged(const float *buf, int e, int copied, const float *row, int64_t tile0, int64_t d) {
    if (e < copied) return buf[e];
    const int64_t i = tile0 + e;
    return i < d ? row[i] : 0.0f;
}

// ---- pass A of one tile: |x| partial sum; the tile that completes a row reduces the partials and publishes the row
__device__ __forceinline__ void pass_a_tile(const StreamArgs &a, int64_t c, int64_t t, const float *buf, int copied, TileScratch &sc) {
    const float *row = a.X + c * a.ld;
    const int64_t tile0 = t * kTile;
    double s = 0.0;
    if (copied == kTile) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = *reinterpret_cast<const float4 *>(buf + q * 1024 + 4 * threadIdx.x);
            s += (double)fabsf(v.x); s += (double)fabsf(v.y); s += (double)fabsf(v.z); s += (double)fabsf(v.w);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) s += (double)fabsf(staged(buf, q * 1024 + 4 * threadIdx.x + e, copied, row, tile0, a.d));
    }
    s = block_sum_f64(s, sc.red);
    if (threadIdx.x == 0) {
        a.partial[c * a.T + t] = s;
        __threadfence();
        sc.flag = (atomicAdd(&a.a_done[c], 1u) == (uint32_t)(a.T - 1)) ? 1u : 0u;
    }
    __syncthreads();
    if (!sc.flag) return;
    __threadfence();
    const volatile double *pp = a.partial + c * a.T;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < a.T; i += kThreads) acc += pp[i];
    acc = block_sum_f64(acc, sc.red);
    if (threadIdx.x == 0) {
        make_row_const(a, c, acc);
        __threadfence();
        st_release_u32(&a.row_ready[c], 1u);
    }
}

// ---- pass B of one tile: AS:625-637 + emit
template <int EMIT, bool EXACT>
__device__ __forceinline__ void pass_b_tile(const StreamArgs &a, const RowConst &rc, int64_t c, int64_t t, const float *buf,
                                            int copied, TileScratch &sc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    TileDesc *rowdesc = a.desc + c * a.T;
    const float *row = a.X + c * a.ld;
    const int64_t tile0 = t * kTile;

    float x[kEpt];
    if (copied == kTile) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = *reinterpret_cast<const float4 *>(buf + kEpt * threadIdx.x + 4 * q);
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kEpt; ++j) x[j] = staged(buf, kEpt * threadIdx.x + j, copied, row, tile0, a.d);
    }

    float fl[kEpt];
    double frd[kEpt];
    double S = 0.0;
    if (!EXACT && (rc.flags & kRowGuardFloor)) {
        // m*p can reach 2^23 in this row: fall back to floorf for threads that actually see such a value
        float mx = 0.0f;
#pragma unroll
        for (int j = 0; j < kEpt; ++j) mx = fmaxf(mx, fabsf(x[j]));
        const bool big = !(__fmul_rn(rc.mf, __fmul_rn(mx, rc.rcpD)) < 4194304.0f);
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            float fr;
            if (big) chain<true>(x[j], rc, fl[j], fr); else chain<false>(x[j], rc, fl[j], fr);
            frd[j] = (double)fr;
            S += frd[j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            float fr;
            chain<EXACT>(x[j], rc, fl[j], fr);
            frd[j] = (double)fr;
            S += frd[j];
        }
    }
    // block scan of the thread sums (Kogge-Stone inside a warp, warps in order): fixed association
    double incl = S;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    double excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 0.0;
    if (lane == 31) sc.wtot[warp] = incl;
    __syncthreads();
    double wbase = 0.0, A = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        if (w == warp) wbase = A;
        A += sc.wtot[w];
    }
    const long long Aq = __double2ll_rn(A * rc.q_up);                // fixed point, 2^-qshift resolution
    if (warp == 0) {
        TileDesc *me = rowdesc + t;
        long long P = 0;
        unsigned long long t0 = 0, t1 = 0, t2 = 0;
        if ((a.dbg & 32) && lane == 0) t0 = gtime();
        if (t == 0) {
            if (lane == 0) { me->aggregate = Aq; me->inclusive = Aq; st_release_u32(&me->state, 2u); }
        } else {
            if (lane == 0) { me->aggregate = Aq; st_release_u32(&me->state, 1u); }
            if ((a.dbg & 32) && lane == 0) t1 = gtime();
            P = (a.dbg & 8) ? 0 : lookback_exclusive(rowdesc, t, lane);
            if ((a.dbg & 32) && lane == 0) t2 = gtime();
            if (lane == 0) { me->inclusive = P + Aq; st_release_u32(&me->state, 2u); }
        }
        if ((a.dbg & 32) && lane == 0 && t > 0) { sc.tacc[2] += t1 - t0; sc.tacc[3] += t2 - t1; sc.tacc[4] += gtime() - t2; }
        if (lane == 0) { sc.Pq = P; sc.P = __ll2double_rn(P) * rc.q_dn; }
    }
    __syncthreads();
    // prefix -> fp32 -> floor(c - X)   (AS:635-636)
    double C = sc.P + (wbase + excl);
    int av[kEpt];
#pragma unroll
    for (int j = 0; j < kEpt; ++j) {
        C += frd[j];
        // The prefix at a tile's last coordinate is DEFINED as the fixed-point inclusive prefix, so that the next
        // tile derives the same floor(c - X) from its own exclusive prefix (no hand-off between tiles).
        if (j == kEpt - 1 && threadIdx.x == kThreads - 1) C = __ll2double_rn(sc.Pq + Aq) * rc.q_dn;
        const float c32 = __double2float_rn(C);
        av[j] = __float2int_rd(__fsub_rn(c32, rc.X));
    }
    int aprev = __shfl_up_sync(0xffffffffu, av[kEpt - 1], 1);
    if (lane == 31) sc.alast[warp] = av[kEpt - 1];
    __syncthreads();
    if (lane == 0) {
        if (warp > 0) aprev = sc.alast[warp - 1];
        else aprev = __float2int_rd(__fsub_rn(__double2float_rn(sc.P), rc.X));     // t == 0: c_0 = 0 (AS:635)
    }
    float kf[kEpt];
#pragma unroll
    for (int j = 0; j < kEpt; ++j) {
        const int r = (av[j] - aprev == 1) ? 1 : 0;                                 // AS:636-637
        aprev = av[j];
        kf[j] = __fadd_rn(fl[j], (float)r);
    }

    const int64_t i0 = tile0 + (int64_t)threadIdx.x * kEpt;
    if (EMIT == kEmitArrays) {
        bool ovf = false;
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            const int64_t i = i0 + j;
            if (i >= a.d) break;
            if (a.deq_out) {
                const float v = __fdiv_rn(x[j], rc.D);
                const float sg = (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f);
                a.deq_out[c * a.ld_out + i] = __fdiv_rn(__fmul_rn(__fmul_rn(rc.L1f, sg), kf[j]), rc.mf);   // AS:640
            }
            if (a.k_out) {
                if (kf[j] >= 2147483648.0f) { ovf = true; a.k_out[c * a.ld_out + i] = 0x7fffffff; }
                else a.k_out[c * a.ld_out + i] = (int32_t)kf[j];
            }
            if (a.sgn_out) a.sgn_out[c * a.ld_out + i] = (uint8_t)(__float_as_uint(x[j]) >> 31);
        }
        if (ovf) atomicOr(&a.hdr->status, 1u);
    } else {
        uint32_t k[kEpt], sg[kEpt];
        bool ovf = false;
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            if (kf[j] >= 2147483648.0f) { ovf = true; k[j] = 0x7fffffffu; } else k[j] = (uint32_t)kf[j];
            sg[j] = __float_as_uint(x[j]) >> 31;
        }
        PackTarget pt{a.codes, a.codes_bytes, a.dir, a.hdr, a.W0, a.arena_base16};
        emit_packed_tile(pt, c * a.T + t, k, sg, ovf, sc.pack);
    }
}

// ------------------------------------------------------------------ the persistent quantize kernel
// Work items are handed out in ONE global order by an atomic ticket: even tickets are pass-A tiles (stream the row
// from HBM, L2 evict_last, partial L1 norm), odd tickets are pass-B tiles of the row `lag` tiles behind (re-read
// from L2, evict_first, scan + quantize + pack).  Every wait (row constants, look-back) is on a smaller ticket,
// and a ticket is only taken by a running CTA, so the schedule cannot deadlock.  Each CTA prefetches the tile of
// its NEXT ticket with a 1-D bulk async copy (TMA) while it works on the current one.
struct Item { int valid; int is_b; int64_t c, t; int copied; const float *src; };
__device__ __forceinline__ Item decode_item(const StreamArgs &a, int64_t i) {
    Item it; it.valid = 0; it.is_b = (int)(i & 1); it.c = 0; it.t = 0; it.copied = 0; it.src = nullptr;
    if (i >= a.total_items) return it;
    const int64_t s = it.is_b ? ((i >> 1) - a.lag) : (i >> 1);
    if (s < 0 || s >= a.n * a.T) return it;
    it.valid = 1;
    it.c = s / a.T; it.t = s - it.c * a.T;
    const int64_t rem = a.d - it.t * kTile;
    it.copied = rem >= kTile ? kTile : (int)(rem & ~(int64_t)3);
    it.src = a.X + it.c * a.ld + it.t * kTile;
    return it;
}

template <int EMIT>
__global__ void __launch_bounds__(kThreads, 2)
quantize_stream_kernel(StreamArgs a) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    float *buf0 = reinterpret_cast<float *>(dyn_smem);
    float *buf1 = reinterpret_cast<float *>(dyn_smem + kTile * 4);
    __shared__ uint64_t mbar[2];
    __shared__ long long s_ticket[2];
    __shared__ TileScratch sc;

    if (threadIdx.x == 0) {
        for (int q = 0; q < 8; ++q) sc.tacc[q] = 0;
        mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_ticket[0] = (long long)atomicAdd(&a.hdr->ticket, 1u);
    }
    __syncthreads();
    uint64_t pol_a = policy_evict_last(), pol_b = policy_evict_first();
    if (a.dbg & 1) { pol_b = pol_a; }                     // dbg: same policy for both passes
    if (a.dbg & 2) { pol_a = pol_b; }
    int64_t cur = s_ticket[0];
    Item it = decode_item(a, cur);
    if (threadIdx.x == 0 && it.valid && it.copied > 0) {
        mbar_expect_tx(&mbar[0], (uint32_t)it.copied * 4u);
        bulk_g2s(buf0, it.src, (uint32_t)it.copied * 4u, &mbar[0], it.is_b ? pol_b : pol_a);
    }
    uint32_t parity[2] = {0u, 0u};
    int stage = 0;
    while (cur < a.total_items) {
        if (it.valid) {
            const float *buf = stage ? buf1 : buf0;
            unsigned long long tq0 = 0, tq1 = 0;
            if ((a.dbg & 32) && threadIdx.x == 0) tq0 = gtime();
            if (it.is_b) {
                // row constants: published by the CTA that finished the row's pass A (a smaller ticket)
                if (threadIdx.x == 0) while (ld_acquire_u32(&a.row_ready[it.c]) == 0u) __nanosleep(50);
                __syncthreads();
            }
            if ((a.dbg & 32) && threadIdx.x == 0) tq1 = gtime();
            if (it.copied > 0) { mbar_wait(&mbar[stage], parity[stage]); parity[stage] ^= 1u; }
            if ((a.dbg & 32) && threadIdx.x == 0) { sc.tacc[0] += tq1 - tq0; sc.tacc[1] += gtime() - tq1; sc.tacc[it.is_b ? 6 : 5] -= gtime(); }
            if (!it.is_b) pass_a_tile(a, it.c, it.t, buf, it.copied, sc);
            else {
                RowConst rc;                  // L2 load: a neighbouring row's constants may sit stale in L1
                {
                    const uint4 *src = reinterpret_cast<const uint4 *>(&a.consts[it.c]);
                    uint4 *dst = reinterpret_cast<uint4 *>(&rc);
#pragma unroll
                    for (int q = 0; q < (int)(sizeof(RowConst) / 16); ++q) dst[q] = __ldcg(src + q);
                }
                if (rc.flags & kRowExact) pass_b_tile<EMIT, true>(a, rc, it.c, it.t, buf, it.copied, sc);
                else pass_b_tile<EMIT, false>(a, rc, it.c, it.t, buf, it.copied, sc);
            }
            if ((a.dbg & 32) && threadIdx.x == 0) sc.tacc[it.is_b ? 6 : 5] += gtime();
        }
        // Take the next ticket only now: a ticket that is held but not yet started would stall every later tile
        // of its row in the look-back (they wait for its aggregate).
        if (threadIdx.x == 0) s_ticket[stage ^ 1] = (long long)atomicAdd(&a.hdr->ticket, 1u);
        __syncthreads();
        cur = s_ticket[stage ^ 1];
        it = decode_item(a, cur);
        stage ^= 1;
        if (threadIdx.x == 0 && it.valid && it.copied > 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&mbar[stage], (uint32_t)it.copied * 4u);
            bulk_g2s(stage ? buf1 : buf0, it.src, (uint32_t)it.copied * 4u, &mbar[stage], it.is_b ? pol_b : pol_a);
        }
    }
    if ((a.dbg & 32) && threadIdx.x == 0)
        for (int q = 0; q < 8; ++q) atomicAdd(reinterpret_cast<unsigned long long *>(a.hdr->pad + 1) + q, sc.tacc[q]);
}

// ------------------------------------------------------------------ K7: decode + mean (tile-major)
template <int W>
__device__ __forceinline__ void decode_accumulate(const uint32_t *__restrict__ tw, float L1f, float mf, float nf, int biased,
                                                  float (&acc)[kEpt]) {
    constexpr int kPerWord = 32 / W;
#pragma unroll
    for (int q = 0; q < W / 2; ++q) {
        const uint32_t word = __ldg(tw + q * kThreads + threadIdx.x);
#pragma unroll
        for (int e = 0; e < kPerWord; ++e) {
            const int j = q * kPerWord + e;
            const uint32_t field = (W == 32) ? word : ((word >> (W * e)) & ((1u << W) - 1u));
            const uint32_t mag = (W == 32) ? (field & 0x7fffffffu) : (field & ((1u << (W - 1)) - 1u));
            if (mag != 0u) {
                const float sg = (field >> (W - 1)) ? -1.0f : 1.0f;
                const float kf = (float)mag;
                const float q32 = biased ? __fmul_rn(__fmul_rn(L1f, sg), __fdiv_rn(kf, mf))          // AS:687
                                         : __fdiv_rn(__fmul_rn(__fmul_rn(L1f, sg), kf), mf);        // AS:640
                acc[j] = __fadd_rn(acc[j], __fdiv_rn(q32, nf));                                       // ND:137
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads)
decode_mean_kernel(const uint32_t *__restrict__ codes, const uint64_t *__restrict__ dir, const float *__restrict__ l1,
                   int64_t n, int64_t d, int64_t T, float mf, float nf, int biased, float *__restrict__ mean, int accumulate) {
    const int64_t t = blockIdx.x;
    const int64_t i0 = t * kTile + (int64_t)threadIdx.x * kEpt;
    float acc[kEpt];
#pragma unroll
    for (int j = 0; j < kEpt; ++j) acc[j] = (accumulate && i0 + j < d) ? mean[i0 + j] : 0.0f;
    for (int64_t c = 0; c < n; ++c) {
        const uint64_t e = __ldg(dir + c * T + t);
        const int W = (int)(e & 0xffu);
        const uint32_t *tw = codes + (e >> 8) * 4ull;
        const float L1f = __ldg(l1 + c);
        switch (W) {
            case 2: decode_accumulate<2>(tw, L1f, mf, nf, biased, acc); break;
            case 4: decode_accumulate<4>(tw, L1f, mf, nf, biased, acc); break;
            case 8: decode_accumulate<8>(tw, L1f, mf, nf, biased, acc); break;
            case 16: decode_accumulate<16>(tw, L1f, mf, nf, biased, acc); break;
            case 32: decode_accumulate<32>(tw, L1f, mf, nf, biased, acc); break;
            default: break;   // width 0: tile was dropped (arena exhausted; status bit 2 is set)
        }
    }
    if (i0 + kEpt <= d) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4 *>(mean + i0 + 4 * q) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < kEpt; ++j)
            if (i0 + j < d) mean[i0 + j] = acc[j];
    }
}

// mean (+)= sum_c Q[c] / n for dequantised rows (ND:133-147), coalesced, clients in order.
__global__ void mean_accumulate_kernel(const float *__restrict__ Q, int64_t n, int64_t d, int64_t ld, float nf,
                                       float *__restrict__ mean, int accumulate) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    float acc = accumulate ? mean[i] : 0.0f;
    for (int64_t c = 0; c < n; ++c) acc = __fadd_rn(acc, __fdiv_rn(Q[c * ld + i], nf));
    mean[i] = acc;
}

// ------------------------------------------------------------------ host side
// Optional per-kernel timing (bench.py's roofline leg): CUDA events recorded on the caller's stream around each
// kernel of the type-quantizer path.  Off by default; never enabled inside a timed region.
struct Profile { bool on = false; cudaEvent_t ev[8]; bool have = false; int marks = 0; };
static Profile g_prof;
static void prof_mark(cudaStream_t st) {
    if (!g_prof.on) return;
    if (!g_prof.have) { for (auto &e : g_prof.ev) cudaEventCreate(&e); g_prof.have = true; }
    if (g_prof.marks < 8) cudaEventRecord(g_prof.ev[g_prof.marks++], st);
}
static void prof_reset() { g_prof.marks = 0; }

static int check_rows(const void *X, int64_t n, int64_t d, int64_t ld) {
    DME_REQUIRE(X != nullptr, "X is null");
    DME_REQUIRE(n >= 1 && n <= 65535, "n=%lld out of range [1, 65535]", (long long)n);
    DME_REQUIRE(d >= 1 && d <= ((int64_t)1 << 31), "d=%lld out of range [1, 2^31]", (long long)d);
    DME_REQUIRE(ld >= d && ld % 4 == 0, "ld=%lld must be >= d and a multiple of 4", (long long)ld);
    DME_REQUIRE(((uintptr_t)X & 15u) == 0, "X must be 16-byte aligned");
    return DME_OK;
}

int ws_prepare(void *ws, int64_t ws_bytes, int64_t n, int64_t d, cudaStream_t st, WsLayout *out, bool need_desc, bool need_sel) {
    const WsLayout L = ws_layout(n, d);
    DME_REQUIRE(ws != nullptr && ((uintptr_t)ws & 255u) == 0, "workspace must be non-null and 256-byte aligned");
    if (ws_bytes < L.total) {
        set_error("workspace too small: %lld < %lld bytes", (long long)ws_bytes, (long long)L.total);
        return DME_EWORKSPACE;
    }
    char *base = (char *)ws;
    DME_CUDA(cudaMemsetAsync(base, 0, (size_t)L.zero_bytes, st));
    if (need_desc) DME_CUDA(cudaMemsetAsync(base + L.off_desc, 0, sizeof(TileDesc) * (size_t)(n * L.T), st));
    if (need_sel) DME_CUDA(cudaMemsetAsync(base + L.off_sel, 0, sizeof(RowSelect) * (size_t)n, st));
    *out = L;
    return DME_OK;
}

int launch_l1(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, const WsLayout &L, void *ws,
              const float *x_inject, const float *l1_inject, uint64_t seed, uint64_t client0, float *l1_out, cudaStream_t st) {
    char *base = (char *)ws;
    dim3 grid((unsigned)L.T, (unsigned)n);
    l1_kernel<<<grid, kThreads, 0, st>>>(X, d, ld, m, L.T, (double *)(base + L.off_partial), (uint32_t *)(base + L.off_done),
                                         (RowConst *)(base + L.off_consts), x_inject, l1_inject, seed, client0, l1_out);
    DME_LAUNCH_CHECK("l1_kernel");
    return DME_OK;
}

int biased_quantize(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, const WsLayout &L, void *ws,
                    int32_t *k_out, uint8_t *sgn_out, float *deq_out, int64_t ld_out,
                    uint32_t *codes, int64_t codes_bytes, uint64_t *dir, cudaStream_t st);   // reznik.cu

}  // namespace dme

using namespace dme;

extern "C" int dme_l1_norms(const float *X, int64_t n, int64_t d, int64_t ld, float *l1_out, void *ws, int64_t ws_bytes,
                            dme_stream_t stream) {
    int rc = check_rows(X, n, d, ld);
    if (rc) return rc;
    DME_REQUIRE(l1_out != nullptr, "l1_out is null");
    cudaStream_t st = (cudaStream_t)stream;
    WsLayout L;
    rc = ws_prepare(ws, ws_bytes, n, d, st, &L, false, false);
    if (rc) return rc;
    return launch_l1(X, n, d, ld, 1, L, ws, nullptr, nullptr, 0, 0, l1_out, st);
}

static int quantize_common(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, int mode, const float *x_inject,
                           const float *l1_inject, uint64_t seed, uint64_t client0, int32_t *k_out, uint8_t *sgn_out,
                           float *deq_out, int64_t ld_out, uint32_t *codes, int64_t codes_bytes, uint64_t *dir,
                           float *l1_out, void *ws, int64_t ws_bytes, cudaStream_t st, bool packed) {
    int rc = check_rows(X, n, d, ld);
    if (rc) return rc;
    DME_REQUIRE(m >= (packed ? 1 : 0) && m < ((int64_t)1 << 40), "m=%lld out of range [%d, 2^40)", (long long)m, packed ? 1 : 0);
    DME_REQUIRE(mode == DME_MODE_UNBIASED || mode == DME_MODE_BIASED, "mode=%d unknown", mode);
    DME_REQUIRE(n * ((d + kTile - 1) / kTile) < ((int64_t)1 << 30), "n * tiles must be < 2^30");
    if (packed) {
        DME_REQUIRE(codes != nullptr && dir != nullptr && ((uintptr_t)codes & 15u) == 0, "codes/dir null or codes not 16-byte aligned");
        DME_REQUIRE(codes_bytes >= 1024, "codes_bytes too small");
    } else {
        DME_REQUIRE(k_out || sgn_out || deq_out, "no output requested");
        DME_REQUIRE(ld_out >= d, "ld_out=%lld < d", (long long)ld_out);
    }
    WsLayout L;
    rc = ws_prepare(ws, ws_bytes, n, d, st, &L, true, mode == DME_MODE_BIASED);
    if (rc) return rc;
    prof_reset();
    prof_mark(st);
    char *base = (char *)ws;
    if (mode == DME_MODE_BIASED) {
        rc = launch_l1(X, n, d, ld, m, L, ws, x_inject, l1_inject, seed, client0, l1_out, st);
        if (rc) return rc;
    }
    if (mode == DME_MODE_BIASED)
        return biased_quantize(X, n, d, ld, m, L, ws, k_out, sgn_out, deq_out, ld_out, codes, codes_bytes, dir, st);
    StreamArgs a;
    a.X = X; a.d = d; a.ld = ld; a.T = L.T; a.n = n; a.m = m;
    a.consts = (RowConst *)(base + L.off_consts);
    a.desc = (TileDesc *)(base + L.off_desc);
    a.hdr = (WsHeader *)base;
    a.partial = (double *)(base + L.off_partial);
    a.a_done = (uint32_t *)(base + L.off_done);
    a.row_ready = (uint32_t *)(base + L.off_ready);
    a.x_inject = x_inject; a.l1_inject = l1_inject; a.seed = seed; a.client0 = client0; a.l1_out = l1_out;
    a.k_out = k_out; a.sgn_out = sgn_out; a.deq_out = deq_out; a.ld_out = ld_out;
    a.codes = codes; a.codes_bytes = codes_bytes; a.dir = dir; a.W0 = expected_width(m > 0 ? m : 1, d);
    a.arena_base16 = 0;
    static int s_sms = 0, s_occ[2] = {0, 0};
    const size_t dyn = 2 * kTile * sizeof(float);
    if (s_sms == 0) {
        int dev = 0;
        DME_CUDA(cudaGetDevice(&dev));
        DME_CUDA(cudaDeviceGetAttribute(&s_sms, cudaDevAttrMultiProcessorCount, dev));
        DME_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s_occ[0], quantize_stream_kernel<kEmitArrays>, kThreads, dyn));
        DME_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s_occ[1], quantize_stream_kernel<kEmitPacked>, kThreads, dyn));
    }
    int64_t G = (int64_t)s_sms * (s_occ[packed ? 1 : 0] > 0 ? s_occ[packed ? 1 : 0] : 1);
    const int64_t nT = n * L.T;
    a.lag = L.T + G;
    a.dbg = 0;
    if (const char *e = getenv("DME_DBG")) a.dbg = atoi(e);            // development knobs (bit field)
    if (const char *e = getenv("DME_DBG_LAG")) a.lag = L.T + atoll(e);
    if (const char *e = getenv("DME_DBG_G")) G = atoll(e);
    a.total_items = 2 * (nT + a.lag);
    if (G > a.total_items) G = a.total_items;
    if (packed) {
        // primary slots occupy the front of the arena; the bump allocator hands out the space behind them
        const unsigned long long primary16 = (unsigned long long)nT * 32ull * (unsigned long long)a.W0;
        if ((long long)(primary16 * 16ull) > codes_bytes) {
            set_error("code arena too small for the primary slots: %lld < %llu bytes", (long long)codes_bytes, primary16 * 16ull);
            return DME_EWORKSPACE;
        }
        a.arena_base16 = primary16;
        quantize_stream_kernel<kEmitPacked><<<(unsigned)G, kThreads, dyn, st>>>(a);
    } else {
        quantize_stream_kernel<kEmitArrays><<<(unsigned)G, kThreads, dyn, st>>>(a);
    }
    DME_LAUNCH_CHECK("quantize_stream_kernel");
    prof_mark(st);
    return DME_OK;
}

extern "C" int dme_type_quantize(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, int mode, const float *x_inject,
                                 const float *l1_inject, uint64_t seed, uint64_t client0, int32_t *k_out, uint8_t *sgn_out,
                                 float *deq_out, int64_t ld_out, float *l1_out, void *ws, int64_t ws_bytes, dme_stream_t stream) {
    return quantize_common(X, n, d, ld, m, mode, x_inject, l1_inject, seed, client0, k_out, sgn_out, deq_out, ld_out, nullptr, 0,
                           nullptr, l1_out, ws, ws_bytes, (cudaStream_t)stream, false);
}

extern "C" int dme_type_encode(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, int mode, const float *x_inject,
                               const float *l1_inject, uint64_t seed, uint64_t client0, void *codes, int64_t codes_bytes,
                               uint64_t *dir, float *l1_out, void *ws, int64_t ws_bytes, dme_stream_t stream) {
    DME_REQUIRE(l1_out != nullptr, "l1_out is required by dme_type_encode (the decoder needs the norms)");
    return quantize_common(X, n, d, ld, m, mode, x_inject, l1_inject, seed, client0, nullptr, nullptr, nullptr, 0,
                           (uint32_t *)codes, codes_bytes, dir, l1_out, ws, ws_bytes, (cudaStream_t)stream, true);
}

extern "C" int dme_decode_mean(const void *codes, const uint64_t *dir, const float *l1, int64_t n, int64_t d, int64_t m, int mode,
                               int64_t n_total, float *mean, int accumulate, dme_stream_t stream) {
    DME_REQUIRE(codes && dir && l1 && mean, "null pointer argument");
    DME_REQUIRE(n >= 1 && d >= 1 && m >= 1 && n_total >= 1, "n, d, m, n_total must be >= 1");
    DME_REQUIRE(((uintptr_t)mean & 15u) == 0 && ((uintptr_t)codes & 15u) == 0, "mean and codes must be 16-byte aligned");
    const int64_t T = (d + kTile - 1) / kTile;
    decode_mean_kernel<<<(unsigned)T, kThreads, 0, (cudaStream_t)stream>>>((const uint32_t *)codes, dir, l1, n, d, T, (float)m,
                                                                           (float)n_total, mode == DME_MODE_BIASED, mean, accumulate);
    DME_LAUNCH_CHECK("decode_mean_kernel");
    prof_mark((cudaStream_t)stream);
    return DME_OK;
}

extern "C" int dme_profile_enable(int on) { g_prof.on = on != 0; prof_reset(); return DME_OK; }
// ms[i] = time between mark i and mark i+1 of the last profiled call (l1, scan, decode for dme_quantize_mean).
extern "C" int dme_profile_read(float *ms, int cap) {
    DME_REQUIRE(ms != nullptr && cap >= 1, "bad argument");
    int k = 0;
    for (; k + 1 < g_prof.marks && k < cap; ++k) {
        DME_CUDA(cudaEventSynchronize(g_prof.ev[k + 1]));
        DME_CUDA(cudaEventElapsedTime(&ms[k], g_prof.ev[k], g_prof.ev[k + 1]));
    }
    return k;
}

extern "C" int dme_quantize_mean(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, int mode, const float *x_inject,
                                 uint64_t seed, uint64_t client0, int64_t n_total, float *mean, int accumulate, void *codes,
                                 int64_t codes_bytes, uint64_t *dir, float *l1_out, void *ws, int64_t ws_bytes, dme_stream_t stream) {
    int rc = dme_type_encode(X, n, d, ld, m, mode, x_inject, nullptr, seed, client0, codes, codes_bytes, dir, l1_out, ws, ws_bytes, stream);
    if (rc) return rc;
    return dme_decode_mean(codes, dir, l1_out, n, d, m, mode, n_total, mean, accumulate, stream);
}

extern "C" int dme_mean_accumulate(const float *Q, int64_t n, int64_t d, int64_t ld, int64_t n_total, float *mean, int accumulate,
                                   dme_stream_t stream) {
    DME_REQUIRE(Q && mean && n >= 1 && d >= 1 && ld >= d && n_total >= 1, "bad argument");
    const int64_t blocks = (d + 255) / 256;
    mean_accumulate_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(Q, n, d, ld, (float)n_total, mean, accumulate);
    DME_LAUNCH_CHECK("mean_accumulate_kernel");
    return DME_OK;
}

extern "C" int dme_status(const void *ws, dme_stream_t stream) {
    DME_REQUIRE(ws != nullptr, "ws is null");
    uint32_t st = 0;
    DME_CUDA(cudaMemcpyAsync(&st, &((const WsHeader *)ws)->status, sizeof(st), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    DME_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (st & 2u) { set_error("code arena exhausted: enlarge codes_bytes (dme_codes_bytes(..., expect=0) is always enough)"); return DME_EWORKSPACE; }
    if (st & 1u) { set_error("a magnitude does not fit the requested integer output"); return DME_EOVERFLOW; }
    return DME_OK;
}
