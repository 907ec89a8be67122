// stream.cu -- the persistent quantize kernel of the unbiased type quantizer (AS:609-641):
// per-client L1 norm, scale to m, floor + systematic-sampling allocation of the fractional mass, sign/magnitude
// packing (or the dequantised output of the drop-in API), for all clients of one GPU in ONE launch.
//
// Schedule.  Work items are enumerated in one global order and dealt round-robin to G co-resident CTAs
// (cooperative launch, G odd).  Even items are pass-A tiles (stream a 16 KB tile of a row from HBM with an L2
// evict_last hint, add |x| in fp64), odd items are pass-B tiles of the row `lag` tiles behind (re-read the tile
// from L2 with evict_first, run AS:625-637, emit).  The A item at position (p+1)T + goff also reduces row p's tile
// sums and publishes the row constants, so that row p's pass B, which starts `lag - T - goff` tiles later, never
// waits for them.  Every wait is on an item with a smaller index, every CTA is resident: no deadlock.
//
// Inside a CTA: 8 compute warps + 1 service warp.
//   service warp : decodes the CTA's items, waits for the row constants, issues the TMA tensor copies (3-D map
//                  {32 floats, rows of 128 B, client}, SWIZZLE_128B: the blocked read "thread t owns coordinates
//                  [16t, 16t+16)" is free of bank conflicts, rows past the end of a client vector arrive as zeros),
//                  scans the 256 thread sums of a pass-B tile, publishes the tile aggregate, finishes the pass-A
//                  tile sums and resolves the decoupled look-back -- all the serial work, off the compute warps.
//   compute warps: pass A; pass B in two stages that are one tile apart (software pipeline of depth 2):
//                  stage 1 (tile i+1): division, floor, fractional parts (kept in registers as floats), thread sum;
//                  stage 2 (tile i)  : running fp64 prefix -> floor(c_j - X) for every coordinate, emit.
//                  The look-back of tile i has a whole tile pair of time to resolve.
//
// Stage 2 in closed form.  AS:636 evaluates t = floor(RN32(RN32(c) - X)) with c the fp64 prefix.  While c32 stays
// inside one binade [2^e + 1, 2^(e+1)), 2 <= e <= 22 (fp32 grid g = 2^(e-23)), this equals floor(c - Xp) when
// a = ceil(X/g - 1/2) is even and ceil(c - Xp) - 1 when a is odd, with Xp = g (a - 1/2) (proof in DESIGN.md
// section 3.1; ties of both roundings included).  So per coordinate: one DFMA (running sigma (c - Xp)) and one
// DADD.RM with the magic constant 1.5 * 2^52 whose low word is the floor -- no conversions; the 0/1 differences of
// consecutive floors telescope into one IMAD per coordinate that builds the 2-bit fields directly.  Threads whose
// prefixes cross a binade (or sit below 5.5) evaluate AS:636 literally.
//
// Look-back.  Every tile publishes its aggregate (int64 fixed point) as a 16-byte record and adds it, split in two
// 31-bit halves that each carry a contribution count, to the record of its block (32 tiles) and of its super-block
// (1024 tiles) with fire-and-forget 64-bit reductions.  A record is complete when both counts are full, so the
// exclusive prefix of a tile needs nothing but the stage 1 of the earlier tiles of its row: one round of independent
// 16-byte loads (tiles of its block, blocks of its super-block, earlier super-blocks).  Integer addition is
// associative: the result does not depend on timing.
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "type_quantize.cuh"

namespace dme {

constexpr int kRing = 2;                 // TMA ring depth (item j uses slot j & 1; refilled after the barrier of item j)
constexpr int kBlock = kThreads + 32;    // 8 compute warps + 1 service warp

struct __align__(16) Rec { unsigned long long v; uint32_t flag; uint32_t pad; };
struct __align__(16) Rec2 { unsigned long long lo, hi; };
typedef Rec TileRec;      // per tile: flag 1 = v is the tile aggregate (fixed point)
constexpr int kCntShift = 44;
constexpr unsigned long long kSumMask = (1ull << kCntShift) - 1ull;

__device__ __forceinline__ void rec_store(Rec *p, unsigned long long v, uint32_t flag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((uint32_t)v), "r"((uint32_t)(v >> 32)), "r"(flag), "r"(0u)
                 : "memory");
}
__device__ __forceinline__ uint32_t rec_load(const Rec *p, unsigned long long &v) {
    uint32_t a, b, f, z;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(f), "=r"(z) : "l"(p) : "memory");
    v = ((unsigned long long)b << 32) | a;
    return f;
}
__device__ __forceinline__ void rec2_load(const Rec2 *p, unsigned long long &lo, unsigned long long &hi) {
    uint32_t a, b, c, d;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
    lo = ((unsigned long long)b << 32) | a;
    hi = ((unsigned long long)d << 32) | c;
}
__device__ __forceinline__ void red_add_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- async-copy / mbarrier / named-barrier primitives
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ bool bar_or(int id, int n, bool pred) {
    uint32_t r;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.u32 q, %3, 0;\n"
        "bar.red.or.pred p, %1, %2, q;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(r) : "r"(id), "r"(n), "r"((uint32_t)pred) : "memory");
    return r != 0;
}
// named barriers: 1,2 = stage-1 data of a pass-B tile (by tile parity), 3,4 = pass-A thread sums (by parity),
// 5,6 = prefix of a pass-B tile resolved (by tile parity), 7 = compute warps only, 8 = compute warps, OR-reduce
constexpr int kBarB1 = 1, kBarPA = 3, kBarReady = 5, kBarCompute = 7, kBarOr = 8;

// one 16 KB box {32 floats, 128 rows, 1 client} at (0, row0, client) of the 3-D tensor map
__device__ __forceinline__ void tma_tile_g2s(uint32_t dst, const CUtensorMap *map, int row0, int client, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(dst), "l"(map), "r"(0), "r"(row0), "r"(client), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ double lds64(uint32_t addr) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#ifdef DME_TIMERS
#define TIC(k) do { if ((a.dbg & 32) && threadIdx.x == (k >= 4 ? kThreads : 0)) sc.tacc[k] -= gtime(); } while (0)
#define TOC(k) do { if ((a.dbg & 32) && threadIdx.x == (k >= 4 ? kThreads : 0)) sc.tacc[k] += gtime(); } while (0)
#else
#define TIC(k) do { } while (0)
#define TOC(k) do { } while (0)
#endif

struct StreamArgs {
    const float *X; int64_t d, ld, T, n, m;
    int64_t rows32;                        // full 128-byte rows per client vector (the part the tensor map covers)
    RowConst *consts; BinadeEntry *tabs; TileRec *desc; Rec2 *blocks; Rec2 *supers; int64_t TB, TS; WsHeader *hdr; Rec *partial; uint32_t *row_ready;
    const float *x_inject; const float *l1_inject; uint64_t seed, client0; float *l1_out;
    int64_t lag, goff, total_items, G;
    int step_c, step_t;                    // G = step_c * T + step_t: per-item advance of (client, tile) without a division
    int32_t *k_out; uint8_t *sgn_out; float *deq_out; int64_t ld_out;     // array outputs
    PackTarget pack; int packed;                                        // packed output
    int dbg;                                                            // development: bit 5 = phase timers
};

// One work item, decoded by the service warp when it issues the tile's copy and shared through a ring.
struct __align__(16) Item { int c, t; uint32_t flags; int fin_row; };
constexpr uint32_t kItValid = 1u, kItB = 2u, kItTma = 4u, kItTail = 8u, kItEnd = 16u;

__device__ __forceinline__ void floor_divmod(long long s, long long T, int &c, int &t) {
    long long q = s >= 0 ? s / T : -((-s + T - 1) / T);
    c = (int)q; t = (int)(s - q * T);
}
// Item decoder (service warp, lane-uniform): the CTA's items advance by G in the global order, i.e. by G positions in
// the A stream and in the B stream alternately, so (client, tile) pairs are updated without divisions.
struct Decoder {
    long long i; int cA, tA, cB, tB, cF, tF;
    __device__ __forceinline__ void init(const StreamArgs &a, long long g) {
        i = g;
        const long long firstA = (g & 1) ? g + a.G : g, firstB = (g & 1) ? g : g + a.G;
        floor_divmod(firstA >> 1, a.T, cA, tA);
        floor_divmod((firstA >> 1) - a.goff, a.T, cF, tF);
        floor_divmod((firstB >> 1) - a.lag, a.T, cB, tB);
    }
    __device__ __forceinline__ static void advance(int &c, int &t, const StreamArgs &a) {
        c += a.step_c; t += a.step_t;
        if (t >= (int)a.T) { t -= (int)a.T; ++c; }
    }
    __device__ __forceinline__ Item next(const StreamArgs &a) {
        Item it; it.c = 0; it.t = 0; it.fin_row = -1;
        const bool is_b = (i & 1) != 0;
        it.flags = is_b ? kItB : 0u;
        if (i >= a.total_items) { it.flags |= kItEnd; return it; }
        i += a.G;
        int c, t;
        if (!is_b) {
            c = cA; t = tA;
            if (tF == 0 && cF >= 1 && cF - 1 < a.n) it.fin_row = cF - 1;     // finaliser duty of this position
            advance(cA, tA, a);
            advance(cF, tF, a);
        } else {
            c = cB; t = tB;
            advance(cB, tB, a);
        }
        if (c < 0 || c >= a.n) return it;
        it.c = c; it.t = t;
        it.flags |= kItValid;
        if ((int64_t)t * (kTile / 32) < a.rows32) it.flags |= kItTma;             // at least one full row in this tile
        if (t == a.T - 1 && (a.d & 31)) it.flags |= kItTail;                      // d % 32 coordinates come straight from global
        return it;
    }
};

constexpr int kScanPad = kThreads + kThreads / 8 + 8;     // index t + (t >> 3): conflict-free for lane-strided-by-8 access
__device__ __forceinline__ int spad(int t) { return t + (t >> 3); }
struct TileInfo {            // what stage 2 needs about a pass-B tile (written by the service warp), by tile parity
    double P;                // exclusive prefix of the tile
    double EnLast;           // inclusive prefix at the last coordinate of the tile (from the fixed-point values)
    float flmax;             // largest floor in the tile
    int rcslot;
};
// Thread sums (stage 1 of a pass-B tile, pass A) travel to the service warp through the tile's own ring slot: every
// thread overwrites 8 of the 64 bytes only it has read, and the slot is refilled by the service warp after it has
// consumed them.
struct Scratch {
    double scanE[2][kScanPad];     // in-tile exclusive prefixes (entry kThreads = tile total), by tile parity
    double red[kWarps];            // finalize_row
    uint32_t flmaxw[2][kWarps];    // per-warp max floor (float bits), by tile parity
    TileInfo info[2];
    RowConst rc[4];                // row constants of the pass-B tiles in flight (slot = tile count & 3)
    BinadeEntry tab[4][kBinades];
    int rc_row[4];
    unsigned long long off16;      // arena offset of a wide tile (compute warps)
    Item items[kRing];
    uint64_t mbar[kRing];
    unsigned long long tacc[8];    // phase timers (dbg)
};

// Row constants + the binade table of AS:636's closed form (cold: once per client row).
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
    if (!(rc.X == 0.0f || (rc.X >= 5.9604644775390625e-08f && rc.X < 1.0f))) fl |= kRowExact;  // X on torch.rand's range
    if (!(rc.mf <= 4194304.0f) || a.l1_inject) fl |= kRowGuardFloor;                            // m*p may reach 2^23
    rc.flags = fl;
    int lg = 0;
    while (((int64_t)1 << lg) < a.d) ++lg;
    rc.qshift = min(50, 62 - lg);
    rc.pad0 = 0;
    rc.q_up = __longlong_as_double((long long)(1023 + rc.qshift) << 52);
    rc.q_dn = __longlong_as_double((long long)(1023 - rc.qshift) << 52);
    rc.pad1[0] = rc.pad1[1] = 0.0;
    a.consts[c] = rc;
    BinadeEntry *tab = a.tabs + c * kBinades;
    const double Xd = (double)rc.X;
#pragma unroll 1
    for (int e = 0; e < kBinades; ++e) {
        BinadeEntry b; b.Xp = 0.0; b.sigma = 0.0;
        if (!(fl & kRowExact) && e >= 2 && e <= 22) {
            const double g = __longlong_as_double((long long)(1023 + e - 23) << 52), ginv = __longlong_as_double((long long)(1023 + 23 - e) << 52);
            const double av = ceil(Xd * ginv - 0.5);                 // exact: X has 24 bits, X >= 2^-24 or X == 0
            b.sigma = (((long long)av) & 1) ? -1.0 : 1.0;
            b.Xp = -b.sigma * (g * (av - 0.5));                      // stored as -sigma * Xp: sigma (c - Xp) = fma(c, sigma, b.Xp)
        }
        tab[e] = b;
    }
    if (a.l1_out) a.l1_out[c] = rc.L1f;
}

// Cold path (compute warps): reduce row `row`'s tile sums in a fixed order (thread-strided, then warp trees in
// index order), publish the row.
__device__ __noinline__ void finalize_row(const StreamArgs &a, int row, Scratch &sc) {
    const Rec *pp = a.partial + (int64_t)row * a.T;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < a.T; i += kThreads) {
        unsigned long long v;
        while (rec_load(pp + i, v) == 0u) __nanosleep(64);
        acc += __longlong_as_double((long long)v);
    }
    acc = warp_sum_f64(acc);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sc.red[warp] = acc;
    bar_sync(kBarCompute, kThreads);
    if (threadIdx.x == 0) {
        double t = sc.red[0];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) t += sc.red[w];
        make_row_const(a, row, t);
        __threadfence();
        st_release_u32(&a.row_ready[row], 1u);
    }
    bar_sync(kBarCompute, kThreads);
}

// AS:625-631 for one coordinate, literal: IEEE division + floorf.
__device__ __forceinline__ void chain_exact(float x, const RowConst &rc, float &flf, float &fr) {
    const float v = __fdiv_rn(x, rc.D);
    const float mp = __fmul_rn(rc.mf, fabsf(v));
    flf = floorf(mp);
    fr = __fsub_rn(mp, flf);
}

// packed f32x2 arithmetic (sm_100: FMUL2 / FFMA2 / FADD2, one issue slot for two coordinates)
typedef unsigned long long f2;
__device__ __forceinline__ f2 f2_pack(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(f2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 f2_add_rz(f2 a, f2 b) { f2 r; asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 f2_sub(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

#ifndef DME_X2
#define DME_X2 1
#endif

// State of a pass-B tile between its two stages (registers of the compute warps).
struct BState {
    float fr[kEpt];         // fractional parts m p - floor(m p)
    uint32_t sgw;           // bits 2j+1 = IEEE sign of coordinate j (even bits: junk)
    uint32_t zmask;         // array output only: bit j = m * |v_j| is exactly zero (sign(v) = 0 in AS:640)
    float flmax_t;          // the thread's largest floor
    int c, t;
    uint32_t flags;         // item flags (0 = nothing parked)
};

// swizzled shared-memory offset of the 16-byte chunk q (0..3) of thread tid's 16 coordinates inside a 16 KB tile
// (SWIZZLE_128B: chunk index within the 128-byte row is XORed with row & 7); q enters as an XOR of (q << 4)
__device__ __forceinline__ uint32_t blocked_off_of(uint32_t tid) {
    const uint32_t row = tid >> 1;
    return row * 128u + ((((tid & 1u) << 2) ^ (row & 7u)) << 4);
}
__device__ __forceinline__ uint32_t blocked_off() { return blocked_off_of(threadIdx.x); }

// ---- pass A of one tile: the thread sums go to shared memory, the service warp finishes them
__device__ __forceinline__ void pass_a(const StreamArgs &a, const Item &it, uint32_t buf, Scratch &sc, int parity) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (it.flags & kItTma) {
        // any order will do: physical 16-byte chunks q*256 + tid (conflict-free), fixed association
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = lds128(buf + (uint32_t)(q * kThreads + threadIdx.x) * 16u);
            s0 += (double)fabsf(v.x); s1 += (double)fabsf(v.y); s2 += (double)fabsf(v.z); s3 += (double)fabsf(v.w);
        }
    }
    if ((it.flags & kItTail) && threadIdx.x < 32) {
        const int64_t i = a.rows32 * 32 + threadIdx.x;
        if (i < a.d) s0 += (double)fabsf(a.X[(int64_t)it.c * a.ld + i]);
    }
    sts64(buf + threadIdx.x * 16u, (s0 + s1) + (s2 + s3));      // over the thread's own first chunk
    bar_arrive(kBarPA + parity, kBlock);
    if (it.fin_row >= 0) finalize_row(a, it.fin_row, sc);
}

// ---- stage 1: everything that does not need the prefix of earlier tiles
template <int EMIT>
__device__ __forceinline__ void stage1(const StreamArgs &a, const Item &it, uint32_t buf, uint32_t park, const RowConst &rc, Scratch &sc, BState &st,
                                       int parity) {
    st.flags = it.flags; st.c = it.c; st.t = it.t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t off = blocked_off();
    float x[kEpt];
    if (it.flags & kItTma) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = lds128((buf + off) ^ (uint32_t)(q << 4));
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kEpt; ++j) x[j] = 0.0f;
    }
    if (it.flags & kItTail) {          // the last d % 32 coordinates of the row are not covered by the tensor map
        const int64_t i0 = (int64_t)it.t * kTile + (int64_t)threadIdx.x * kEpt, lo = a.rows32 * 32;
        const float *row = a.X + (int64_t)it.c * a.ld;
        if (i0 + kEpt > lo && i0 < a.d) {
#pragma unroll
            for (int j = 0; j < kEpt; ++j)
                if (i0 + j >= lo && i0 + j < a.d) x[j] = row[i0 + j];
        }
    }
    const bool exact = rc.flags & kRowExact;
    bool big = false;
    if (!exact && (rc.flags & kRowGuardFloor)) {
        // m*p can reach 2^23 in this row: threads that actually see such a value use floorf
        float mx = 0.0f;
#pragma unroll
        for (int j = 0; j < kEpt; ++j) mx = fmaxf(mx, fabsf(x[j]));
        big = !(__fmul_rn(rc.mf, __fmul_rn(mx, rc.rcpD)) < 4194304.0f);
    }
    uint32_t sg = 0;
#pragma unroll
    for (int j = kEpt - 1; j >= 0; --j) sg = __funnelshift_l(__float_as_uint(x[j]), sg, 2);     // bit 2j+1 = sign of x[j]
    st.sgw = sg;
    float flf[kEpt];
    if (exact || big) {
#pragma unroll
        for (int j = 0; j < kEpt; ++j) chain_exact(x[j], rc, flf[j], st.fr[j]);
    } else {
#if DME_X2
        // fast chain on pairs of |x|: x/D by Markstein's correction of x * rcp, floor by adding 2^23 toward zero
        const f2 R2 = f2_pack(rc.rcpD, rc.rcpD), ND = f2_pack(-rc.D, -rc.D), M2 = f2_pack(rc.mf, rc.mf), C2 = f2_pack(8388608.0f, 8388608.0f);
#pragma unroll
        for (int j = 0; j < kEpt; j += 2) {
            const f2 xx = f2_pack(__uint_as_float(__float_as_uint(x[j]) & 0x7fffffffu), __uint_as_float(__float_as_uint(x[j + 1]) & 0x7fffffffu));
            const f2 q0 = f2_mul(xx, R2);
            const f2 rem = f2_fma(q0, ND, xx);
            const f2 pq = f2_fma(rem, R2, q0);
            const f2 mp = f2_mul(M2, pq);
            const f2 tt = f2_add_rz(mp, C2);
            const f2 fl2 = f2_sub(tt, C2);
            const f2 fr2 = f2_sub(mp, fl2);
            f2_unpack(fl2, flf[j], flf[j + 1]);
            f2_unpack(fr2, st.fr[j], st.fr[j + 1]);
        }
#else
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            const float ax = fabsf(x[j]);
            const float q0 = __fmul_rn(ax, rc.rcpD);
            const float rem = __fmaf_rn(-q0, rc.D, ax);
            const float pq = __fmaf_rn(rem, rc.rcpD, q0);
            const float mp = __fmul_rn(rc.mf, pq);
            const float tt = __fadd_rz(mp, 8388608.0f);
            flf[j] = __fsub_rn(tt, 8388608.0f);
            st.fr[j] = __fsub_rn(mp, flf[j]);
        }
#endif
    }
    // park the floors (thread-private 64 bytes of the park buffer)
#pragma unroll
    for (int q = 0; q < 4; ++q) sts128((park + off) ^ (uint32_t)(q << 4), make_float4(flf[4 * q], flf[4 * q + 1], flf[4 * q + 2], flf[4 * q + 3]));
    if (EMIT == 0) {
        uint32_t z = 0;
#pragma unroll
        for (int j = 0; j < kEpt; ++j) z |= ((flf[j] == 0.0f && st.fr[j] == 0.0f) ? 1u : 0u) << j;
        st.zmask = z;
    }
    float mxf = flf[0];
#pragma unroll
    for (int j = 1; j < kEpt; ++j) mxf = fmaxf(mxf, flf[j]);
    st.flmax_t = mxf;
    // thread sum, left to right
    double run = (double)st.fr[0];
#pragma unroll
    for (int j = 1; j < kEpt; ++j) run += (double)st.fr[j];
    sts64((buf + off) ^ (((threadIdx.x >> 4) & 3u) << 4), run);                      // over one of the thread's own chunks
    const uint32_t wmx = __reduce_max_sync(0xffffffffu, __float_as_uint(mxf));     // floors are >= 0: bit order = value order
    if (lane == 0) sc.flmaxw[parity][warp] = wmx;
    bar_arrive(kBarB1 + parity, kBlock);
}

// AS:636 literally, for one prefix value
__device__ __forceinline__ int floor_ref(double c, float X) { return __float2int_rd(__fsub_rn(__double2float_rn(c), X)); }
constexpr double kMagic = 6755399441055744.0;     // 1.5 * 2^52: (u + kMagic) rounded down has floor(u) in its low word
__device__ __forceinline__ int floor_lo(double u) { return __double2loint(__dadd_rd(u, kMagic)); }
// double of a non-negative fp32 fraction by integer moves (no conversion unit).  0 maps to 2^-127, denormals to
// values below 2^-126: both vanish in every sum they enter.
__device__ __forceinline__ double frac_to_double(float f) {
    const uint32_t b = __float_as_uint(f);
    return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
}

// Prefix geometry of one thread in stage 2.
struct Geo {
    double E, En;        // prefix before the thread's first coordinate / at its last coordinate
    double sE, sig;      // closed form: sigma * (E - Xp), sigma
    double sEn;          // sigma * (En - Xp)
    bool fast;
};
__device__ __forceinline__ Geo make_geo(const BinadeEntry *tab, double E, double En) {
    Geo g;
    g.E = E; g.En = En;
    const int e = (__double2hiint(E) >> 20) - 1023;
    bool fast = e >= 2 && e <= 22;
    BinadeEntry b; b.Xp = 0.0; b.sigma = 0.0;
    if (fast) {
        b = tab[e];
        // every prefix of the thread, and their fp32 roundings, stay inside [2^e + 1, 2^(e+1))
        const int e_lo = (__double2hiint(E - 1.5) >> 20) - 1023, e_hi = (__double2hiint(En + 1.0) >> 20) - 1023;
        fast = (e_lo == e) && (e_hi == e) && (b.sigma != 0.0);
    }
    g.fast = fast;
    g.sig = b.sigma;
    g.sE = fma(E, b.sigma, b.Xp);         // exact: Xp is a multiple of 2^(e-24), |sigma| = 1 (b.Xp holds -sigma Xp)
    g.sEn = fma(En, b.sigma, b.Xp);
    return g;
}

// r_j = [floor(c_j - X) - floor(c_{j-1} - X) == 1] (AS:636-637) for the thread's 16 coordinates, as a bit mask
__device__ __forceinline__ uint32_t rbits_generic(const Geo &g, const BState &st, float X) {
    uint32_t rb = 0;
    if (g.fast) {
        const int sgi = g.sig > 0.0 ? 1 : -1;
        double u = g.sE;
        int Lp = floor_lo(u);
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            u = (j < kEpt - 1) ? fma(frac_to_double(st.fr[j]), g.sig, u) : g.sEn;
            const int L = floor_lo(u);
            rb |= (((L - Lp) * sgi == 1) ? 1u : 0u) << j;
            Lp = L;
        }
    } else {
        double c = g.E;
        int tp = floor_ref(c, X);
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            c = (j < kEpt - 1) ? c + (double)st.fr[j] : g.En;
            const int t = floor_ref(c, X);
            rb |= ((t - tp == 1) ? 1u : 0u) << j;
            tp = t;
        }
    }
    return rb;
}
__device__ __forceinline__ uint32_t spread16(uint32_t v) {       // bit j -> bit 2j
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}
// the same as 2-bit interleaved fields (bit 2j = r_j): the differences telescope into one IMAD per coordinate
__device__ __forceinline__ uint32_t rbits_interleaved(const Geo &g, const BState &st, float X) {
    if (g.fast) {
        const int sgi = g.sig > 0.0 ? 1 : -1;
        double u = g.sE;
        uint32_t acc = 0u - (uint32_t)floor_lo(u);
        int L14 = 0;
#pragma unroll
        for (int j = 0; j < kEpt - 1; ++j) {
            u = fma(frac_to_double(st.fr[j]), g.sig, u);
            const int L = floor_lo(u);
            // sum_j (L_j - L_{j-1}) 4^j  =  -L_{-1} - sum_{j<14} 3 * 4^j L_j + 4^14 L_14
            if (j < kEpt - 2) acc += (uint32_t)L * (0u - (3u << (2 * j)));
            else { acc += (uint32_t)L << (2 * j); L14 = L; }
        }
        acc *= (uint32_t)sgi;
        // the thread's last prefix is defined from the scan (the next thread starts from the same value), so it is
        // the one place where fp64 association could make a difference non-monotone: evaluated on its own
        const int L15 = floor_lo(g.sEn);
        if ((L15 - L14) * sgi == 1) acc |= 1u << 30;
        return acc;
    }
    return spread16(rbits_generic(g, st, X));
}

// smallest field width whose magnitude part holds k
__device__ __forceinline__ int width_for(float kmax) {
    int W = 2;
    while (W < 32 && kmax >= (float)(1u << (W - 1))) W <<= 1;
    return W;
}

// ---- stage 2: prefix -> floor(c - X) (AS:635-637), type vector, emit
template <int EMIT>
__device__ __forceinline__ void stage2(const StreamArgs &a, Scratch &sc, uint32_t park, BState &st, int parity) {
    bar_sync(kBarReady + parity, kBlock);                 // the service warp has resolved this tile's prefix
    if (!(st.flags & kItValid)) return;
    st.flags = 0;
    const TileInfo &ti = sc.info[parity];
    const RowConst &rc = sc.rc[ti.rcslot];
    const double Pd = ti.P;
    // The prefix at a thread's LAST coordinate is defined from the scan values (at the last coordinate of the tile
    // from the fixed-point inclusive prefix), so the next thread / tile starts from exactly the same value and
    // derives the same floor(c - X) for its predecessor: no hand-off is needed.
    const double E = Pd + sc.scanE[parity][spad(threadIdx.x)];
    double En = Pd + sc.scanE[parity][spad(threadIdx.x + 1)];
    if (threadIdx.x == kThreads - 1) En = ti.EnLast;
    const Geo g = make_geo(sc.tab[ti.rcslot], E, En);
    const float fm = ti.flmax;
    const uint32_t off = blocked_off();
    const int64_t i0 = (int64_t)st.t * kTile + (int64_t)threadIdx.x * kEpt;
    if (EMIT == 0) {
        const uint32_t rb = rbits_generic(g, st, rc.X);
        float fl[kEpt];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = lds128((park + off) ^ (uint32_t)(q << 4));
            fl[4 * q] = v.x; fl[4 * q + 1] = v.y; fl[4 * q + 2] = v.z; fl[4 * q + 3] = v.w;
        }
        bool ovf = false;
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            const float kf = __fadd_rn(fl[j], (float)((rb >> j) & 1u));
            const int64_t i = i0 + j;
            if (i >= a.d) continue;
            const uint32_t sbit = (st.sgw >> (2 * j + 1)) & 1u;
            if (a.deq_out) {
                // sign(v) of AS:640: v = x / D is zero exactly when m * |v| is (floor and fraction both zero, m > 0)
                const float sgf = ((st.zmask >> j) & 1u) ? 0.0f : (sbit ? -1.0f : 1.0f);
                a.deq_out[(int64_t)st.c * a.ld_out + i] = __fdiv_rn(__fmul_rn(__fmul_rn(rc.L1f, sgf), kf), rc.mf);
            }
            if (a.k_out) {
                if (kf >= 2147483648.0f) { ovf = true; a.k_out[(int64_t)st.c * a.ld_out + i] = 0x7fffffff; }
                else a.k_out[(int64_t)st.c * a.ld_out + i] = (int32_t)kf;
            }
            if (a.sgn_out) a.sgn_out[(int64_t)st.c * a.ld_out + i] = (uint8_t)sbit;
        }
        if (ovf) atomicOr(&a.hdr->status, 1u);
    } else {
        // tile-wide minimal field width: from the largest floor; only when the largest floor sits right below a
        // width boundary does it matter whether one of those coordinates also received a unit
        const int Wlo = width_for(fm), Whi = width_for(__fadd_rn(fm, 1.0f));
        int W = Wlo;
        uint32_t rb16 = 0;
        const bool certain2 = (Whi == 2);
        if (!certain2) {
            rb16 = rbits_generic(g, st, rc.X);
            if (Wlo != Whi) {
                bool hit = false;
                if (st.flmax_t == fm) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 v = lds128((park + off) ^ (uint32_t)(q << 4));
                        hit |= (v.x == fm && ((rb16 >> (4 * q)) & 1u)) | (v.y == fm && ((rb16 >> (4 * q + 1)) & 1u)) |
                               (v.z == fm && ((rb16 >> (4 * q + 2)) & 1u)) | (v.w == fm && ((rb16 >> (4 * q + 3)) & 1u));
                    }
                }
                W = bar_or(kBarOr, kThreads, hit) ? Whi : Wlo;
            }
        }
        if (fm >= 2147483520.0f && threadIdx.x == 0) atomicOr(&a.hdr->status, 1u);
        const int64_t slot_id = (int64_t)st.c * a.T + st.t;
        unsigned long long off16;
        if (W <= a.pack.W0) {
            off16 = (unsigned long long)slot_id * (32ull * a.pack.W0);
            if (threadIdx.x == 0) a.pack.dir[slot_id] = (off16 << 8) | (unsigned long long)W;
        } else {
            if (threadIdx.x == 0) {
                const unsigned long long units = 32ull * W;
                unsigned long long o = a.pack.arena_base16 + atomicAdd(&a.hdr->arena_top, units);
                if ((long long)((o + units) * 16ull) > a.pack.codes_bytes) { atomicOr(&a.hdr->status, 2u); o = ~0ull; }
                sc.off16 = o;
                a.pack.dir[slot_id] = (o == ~0ull) ? 0ull : ((o << 8) | (unsigned long long)W);
            }
            bar_sync(kBarCompute, kThreads);
            off16 = sc.off16;
            bar_sync(kBarCompute, kThreads);        // off16 may be rewritten by the next wide tile
        }
        if (off16 != ~0ull) {
            uint32_t *tw = a.pack.codes + off16 * 4ull;
            if (W == 2) {
                // fields [sign | magnitude bit]
                uint32_t kw;
                if (certain2) kw = rbits_interleaved(g, st, rc.X);
                else {
                    // largest floor is 1 and no such coordinate received a unit: k = floor + r is still <= 1
                    uint32_t kb = rb16;
                    if (st.flmax_t != 0.0f) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 v = lds128((park + off) ^ (uint32_t)(q << 4));
                            kb |= ((v.x != 0.0f) ? 1u : 0u) << (4 * q) | ((v.y != 0.0f) ? 1u : 0u) << (4 * q + 1) |
                                  ((v.z != 0.0f) ? 1u : 0u) << (4 * q + 2) | ((v.w != 0.0f) ? 1u : 0u) << (4 * q + 3);
                        }
                    }
                    kw = spread16(kb);
                }
                tw[threadIdx.x] = kw | (st.sgw & 0xaaaaaaaau);
            } else {
                uint32_t k[kEpt], sg[kEpt];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = lds128((park + off) ^ (uint32_t)(q << 4));
                    const float f4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int j = 4 * q + e;
                        const float ff = fminf(f4[e], 2147483520.0f);       // overflow already reported
                        k[j] = (uint32_t)ff + ((rb16 >> j) & 1u);
                        sg[j] = (st.sgw >> (2 * j + 1)) & 1u;
                    }
                }
                switch (W) {
                    case 4: pack_store<4>(k, sg, tw); break;
                    case 8: pack_store<8>(k, sg, tw); break;
                    case 16: pack_store<16>(k, sg, tw); break;
                    default: pack_store<32>(k, sg, tw); break;
                }
            }
        }
    }
}

// ------------------------------------------------------------------ service warp
#ifdef DME_TIMERS
__device__ unsigned long long g_polls;
#endif
// Exclusive fixed-point prefix of tile t of a row: earlier tiles of its block + earlier blocks of its super-block +
// earlier super-blocks (see the header).  All lanes return the result.
__device__ __forceinline__ long long lookback_resolve(const TileRec *tiles, const Rec2 *blocks, const Rec2 *supers, int t, int lane) {
    const int b = t >> 5, pos = t & 31, sb = b >> 5, bpos = b & 31;
    const bool has_t = lane < pos, has_b = lane < bpos;
    while (true) {
        bool ok = true;
        long long x = 0;
        if (has_t) {
            unsigned long long v;
            ok = rec_load(tiles + (t - 1 - lane), v) != 0u;
            x = (long long)v;
        }
        if (has_b) {
            unsigned long long lo, hi;
            rec2_load(blocks + (sb * 32 + lane), lo, hi);
            ok = ok && (lo >> kCntShift) == 32ull && (hi >> kCntShift) == 32ull;
            x += (long long)(((hi & kSumMask) << 31) + (lo & kSumMask));
        }
        for (int s = lane; s < sb; s += 32) {
            unsigned long long lo, hi;
            rec2_load(supers + s, lo, hi);
            ok = ok && (lo >> kCntShift) == 1024ull && (hi >> kCntShift) == 1024ull;
            x += (long long)(((hi & kSumMask) << 31) + (lo & kSumMask));
        }
        if (__all_sync(0xffffffffu, ok)) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            return x;
        }
#ifdef DME_TIMERS
        if (lane == 0) atomicAdd(&g_polls, 1ull);
#endif
        __nanosleep(64);
    }
}

struct Service {
    Decoder dec;
    int nb;              // pass-B items seen so far (tile parity = nb & 1, row-constant slot = nb & 3)
};

// decode the next item into ring slot `slot`, make its row constants available, start its copy
__device__ __forceinline__ void service_issue(const StreamArgs &a, const CUtensorMap *tmap, Scratch &sc, Service &sv, int slot, uint32_t ring0, int lane,
                                              uint64_t pol_a, uint64_t pol_b) {
    const Item it = sv.dec.next(a);
    if ((it.flags & kItB) && (it.flags & kItValid)) {
        const int rs = sv.nb & 3;
        if (sc.rc_row[rs] != it.c) {
            if (lane == 0) while (ld_acquire_u32(&a.row_ready[it.c]) == 0u) __nanosleep(64);
            __syncwarp();
            const uint4 *src = reinterpret_cast<const uint4 *>(&a.consts[it.c]);
            uint4 *dst = reinterpret_cast<uint4 *>(&sc.rc[rs]);
            if (lane < (int)(sizeof(RowConst) / 16)) dst[lane] = __ldcg(src + lane);
            const uint4 *ts = reinterpret_cast<const uint4 *>(a.tabs + (int64_t)it.c * kBinades);
            uint4 *td = reinterpret_cast<uint4 *>(sc.tab[rs]);
            if (lane < kBinades) td[lane] = __ldcg(ts + lane);
            if (lane == 0) sc.rc_row[rs] = it.c;
            __syncwarp();
        }
    }
    if (it.flags & kItB) ++sv.nb;
    __syncwarp();                                 // every lane has finished reading the slot
    if (lane == 0) {
        sc.items[slot] = it;
        if ((it.flags & kItValid) && (it.flags & kItTma)) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&sc.mbar[slot], (uint32_t)kTile * 4u);
            tma_tile_g2s(ring0 + (uint32_t)slot * kTile * 4u, tmap, it.t * (kTile / 32), it.c, &sc.mbar[slot], (it.flags & kItB) ? pol_b : pol_a);
        } else {
            mbar_arrive(&sc.mbar[slot]);          // nothing to copy: the phase completes at once (releases the item record)
        }
    }
    __syncwarp();
}

#ifndef DME_STREAM_CTAS
#define DME_STREAM_CTAS 3
#endif
template <int EMIT>
__global__ void __launch_bounds__(kBlock, DME_STREAM_CTAS)
quantize_stream_kernel(const __grid_constant__ StreamArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char dyn_smem[];      // ring slots, two park buffers, Scratch
    Scratch &sc = *reinterpret_cast<Scratch *>(dyn_smem + (size_t)(kRing + 2) * kTile * sizeof(float));

    if (threadIdx.x == 0) {
        for (int b = 0; b < kRing; ++b) mbar_init(&sc.mbar[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int q = 0; q < 4; ++q) sc.rc_row[q] = -1;
        for (int q = 0; q < 8; ++q) sc.tacc[q] = 0;
        if (a.dbg & 32) sc.tacc[6] -= gtime();
    }
    __syncthreads();
    const int64_t g = blockIdx.x;
    const uint32_t ring0 = smem_u32(dyn_smem);                         // SWIZZLE_128B wants 1024-byte aligned boxes
    const uint32_t park0 = ring0 + (uint32_t)kRing * kTile * 4u;       // two park buffers (floors of the tiles in flight)
    const uint32_t mbar0 = smem_u32(&sc.mbar[0]);

    if (threadIdx.x >= kThreads) {
        // ================================================================== service warp
        const int lane = threadIdx.x & 31;
        const uint64_t pol_a = policy_evict_last(), pol_b = policy_evict_first();
        Service sv;
        sv.dec.init(a, g);
        sv.nb = 0;
        for (int j = 0; j < kRing; ++j) service_issue(a, &tmap, sc, sv, j, ring0, lane, pol_a, pol_b);
        int nbd = 0, nad = 0;                   // pass-B / pass-A items completed by this warp
        for (int64_t j = 0;; ++j) {
            const int slot = (int)(j & 1);
            const Item it = sc.items[slot];     // written by this warp
            if (it.flags & kItEnd) break;
            if (it.flags & kItB) {
                const int par = nbd & 1;
                bar_sync(kBarB1 + par, kBlock);                                   // thread sums + floor maxima are in
                TIC(4);
                const int rs = nbd & 3;
                // scan of the 256 thread sums: lane l owns threads 8l .. 8l+7 (fixed association)
                double v[8];
                double tot = 0.0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t u = 8u * lane + i;
                    v[i] = lds64((ring0 + (uint32_t)slot * kTile * 4u + blocked_off_of(u)) ^ (((u >> 4) & 3u) << 4));
                    tot += v[i];
                }
                uint32_t fm = lane < kWarps ? sc.flmaxw[par][lane] : 0u;
                double incl = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                const double A = __shfl_sync(0xffffffffu, incl, 31);
                long long Aq = 0;
                if (it.flags & kItValid) {
                    const RowConst &rc = sc.rc[rs];
                    Aq = __double2ll_rn(A * rc.q_up);                             // fixed point, 2^-qshift resolution
                    if (lane == 0) {
                        rec_store(a.desc + (int64_t)it.c * a.T + it.t, (unsigned long long)Aq, 1u);
                        const unsigned long long lo = ((unsigned long long)Aq & 0x7fffffffull) + (1ull << kCntShift);
                        const unsigned long long hi = ((unsigned long long)Aq >> 31) + (1ull << kCntShift);
                        Rec2 *br = a.blocks + (int64_t)it.c * a.TB + (it.t >> 5), *sr = a.supers + (int64_t)it.c * a.TS + (it.t >> 10);
                        red_add_u64(&br->lo, lo); red_add_u64(&br->hi, hi);
                        red_add_u64(&sr->lo, lo); red_add_u64(&sr->hi, hi);
                    }
                }
                double ex = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0) ex = 0.0;
#pragma unroll
                for (int i = 0; i < 8; ++i) { sc.scanE[par][9 * lane + i] = ex; ex += v[i]; }
                if (lane == 31) sc.scanE[par][spad(kThreads)] = incl;
                fm = __reduce_max_sync(0xffffffffu, fm);
                TOC(4);
                service_issue(a, &tmap, sc, sv, slot, ring0, lane, pol_a, pol_b);  // the slot is free: next item of this parity
                // the A item that follows gives the predecessors time to publish; resolve after it
                const int slotA = slot ^ 1;
                const Item ia = sc.items[slotA];
                if (!(ia.flags & kItEnd) && !(ia.flags & kItB)) {
                    const int pa = nad & 1;
                    bar_sync(kBarPA + pa, kBlock);
                    const uint32_t bufA = ring0 + (uint32_t)slotA * kTile * 4u;
                    double tot2 = lds64(bufA + lane * 16u);
#pragma unroll
                    for (int i = 1; i < kThreads / 32; ++i) tot2 += lds64(bufA + (lane + 32 * i) * 16u);
                    tot2 = warp_sum_f64(tot2);
                    if (lane == 0 && (ia.flags & kItValid))
                        rec_store(&a.partial[(int64_t)ia.c * a.T + ia.t], (unsigned long long)__double_as_longlong(tot2), 1u);
                    service_issue(a, &tmap, sc, sv, slotA, ring0, lane, pol_a, pol_b);
                    ++nad; ++j;
                }
                if (it.flags & kItValid) {
                    TIC(5);
                    const long long P = lookback_resolve(a.desc + (int64_t)it.c * a.T, a.blocks + (int64_t)it.c * a.TB, a.supers + (int64_t)it.c * a.TS,
                                                         it.t, lane);
                    TOC(5);
                    if (lane == 0) {
                        const RowConst &rc = sc.rc[rs];
                        TileInfo ti;
                        ti.P = __ll2double_rn(P) * rc.q_dn;
                        ti.EnLast = __ll2double_rn(P + Aq) * rc.q_dn;
                        ti.flmax = __uint_as_float(fm);
                        ti.rcslot = rs;
                        sc.info[par] = ti;
                    }
                }
                __syncwarp();
                bar_arrive(kBarReady + par, kBlock);
                ++nbd;
            } else {
                const int pa = nad & 1;
                bar_sync(kBarPA + pa, kBlock);
                const uint32_t bufA = ring0 + (uint32_t)slot * kTile * 4u;
                double tot2 = lds64(bufA + lane * 16u);
#pragma unroll
                for (int i = 1; i < kThreads / 32; ++i) tot2 += lds64(bufA + (lane + 32 * i) * 16u);
                tot2 = warp_sum_f64(tot2);
                if (lane == 0 && (it.flags & kItValid))
                    rec_store(&a.partial[(int64_t)it.c * a.T + it.t], (unsigned long long)__double_as_longlong(tot2), 1u);
                service_issue(a, &tmap, sc, sv, slot, ring0, lane, pol_a, pol_b);
                ++nad;
            }
        }
    } else {
        // ================================================================== compute warps
        // software pipeline of depth 2 over the pass-B tiles: stage 1 of tile i+1, the pass-A item between, stage 2 of tile i
        BState s0, s1;
        s0.flags = 0; s1.flags = 0;
        int64_t j = 0;                      // local item index
        int nb = 0, na = 0;                 // pass-B / pass-A items started
        int pend0 = 0, pend1 = 0;           // a stage 2 is owed for state 0 / 1 (its READY barrier must be consumed)
        auto fetch = [&](Item &it, uint32_t &buf) -> bool {
            const int slot = (int)(j & 1);
            mbar_wait(mbar0 + 8u * slot, (uint32_t)((j >> 1) & 1));
            it = sc.items[slot];
            buf = ring0 + (uint32_t)slot * kTile * 4u;
            ++j;
            return !(it.flags & kItEnd);
        };
        auto run_a = [&](const Item &it, uint32_t buf) { TIC(1); pass_a(a, it, buf, sc, na & 1); TOC(1); ++na; };
        Item it; uint32_t buf = 0;
        bool more = fetch(it, buf);
        if (more && !(it.flags & kItB)) { run_a(it, buf); more = fetch(it, buf); }
        while (more) {
            // ---- B item -> state 0
            {
                const int par = nb & 1, rs = nb & 3;
                TIC(2);
                if (it.flags & kItValid) stage1<EMIT>(a, it, buf, park0 + (uint32_t)par * kTile * 4u, sc.rc[rs], sc, s0, par);
                else { s0.flags = 0; bar_arrive(kBarB1 + par, kBlock); }
                TOC(2);
                pend0 = 1; ++nb;
            }
            more = fetch(it, buf);
            if (more && !(it.flags & kItB)) { run_a(it, buf); more = fetch(it, buf); }
            if (pend1) { TIC(3); stage2<EMIT>(a, sc, park0 + (uint32_t)kTile * 4u, s1, 1); TOC(3); pend1 = 0; }
            if (!more) break;
            // ---- B item -> state 1
            {
                const int par = nb & 1, rs = nb & 3;
                TIC(2);
                if (it.flags & kItValid) stage1<EMIT>(a, it, buf, park0 + (uint32_t)par * kTile * 4u, sc.rc[rs], sc, s1, par);
                else { s1.flags = 0; bar_arrive(kBarB1 + par, kBlock); }
                TOC(2);
                pend1 = 1; ++nb;
            }
            more = fetch(it, buf);
            if (more && !(it.flags & kItB)) { run_a(it, buf); more = fetch(it, buf); }
            if (pend0) { TIC(3); stage2<EMIT>(a, sc, park0, s0, 0); TOC(3); pend0 = 0; }
        }
        // drain in tile order
        if (pend0 && pend1) {
            if (nb & 1) { stage2<EMIT>(a, sc, park0 + (uint32_t)kTile * 4u, s1, 1); stage2<EMIT>(a, sc, park0, s0, 0); }
            else { stage2<EMIT>(a, sc, park0, s0, 0); stage2<EMIT>(a, sc, park0 + (uint32_t)kTile * 4u, s1, 1); }
        } else if (pend0) stage2<EMIT>(a, sc, park0, s0, 0);
        else if (pend1) stage2<EMIT>(a, sc, park0 + (uint32_t)kTile * 4u, s1, 1);
    }
    if ((a.dbg & 32) && (threadIdx.x == 0 || threadIdx.x == kThreads)) {
        if (threadIdx.x == 0) sc.tacc[6] += gtime();
#ifdef DME_TIMERS
        if (blockIdx.x == 0 && threadIdx.x == 0) { sc.tacc[7] = g_polls; g_polls = 0; }
#endif
        for (int q = (threadIdx.x == 0 ? 0 : 4); q < (threadIdx.x == 0 ? 4 : 6); ++q)
            atomicAdd(reinterpret_cast<unsigned long long *>(a.hdr->pad + 1) + q, sc.tacc[q]);
        if (threadIdx.x == 0) for (int q = 6; q < 8; ++q) atomicAdd(reinterpret_cast<unsigned long long *>(a.hdr->pad + 1) + q, sc.tacc[q]);
    }
}

static int g_sms = 0, g_occ[2] = {0, 0};
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

int launch_stream(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, const WsLayout &L, void *ws,
                  const float *x_inject, const float *l1_inject, uint64_t seed, uint64_t client0,
                  int32_t *k_out, uint8_t *sgn_out, float *deq_out, int64_t ld_out,
                  uint32_t *codes, int64_t codes_bytes, uint64_t *dir, float *l1_out, cudaStream_t st, bool packed) {
    char *base = (char *)ws;
    StreamArgs a;
    a.X = X; a.d = d; a.ld = ld; a.T = L.T; a.n = n; a.m = m;
    a.rows32 = d / 32;
    a.consts = (RowConst *)(base + L.off_consts);
    a.tabs = (BinadeEntry *)(base + L.off_tab);
    a.TB = (L.T + 31) / 32;                                        // blocks of 32 tiles per row
    a.TS = (a.TB + 31) / 32;                                       // super-blocks of 32 blocks per row
    a.desc = (TileRec *)(base + L.off_desc);                       // tile records, then block records, then super-block records
    a.blocks = (Rec2 *)(base + L.off_desc + 16 * n * L.T);
    a.supers = (Rec2 *)(base + L.off_desc + 16 * n * (L.T + a.TB));
    a.hdr = (WsHeader *)base;
    a.partial = (Rec *)(base + L.off_partial);
    a.row_ready = (uint32_t *)(base + L.off_ready);
    a.x_inject = x_inject; a.l1_inject = l1_inject; a.seed = seed; a.client0 = client0; a.l1_out = l1_out;
    a.k_out = k_out; a.sgn_out = sgn_out; a.deq_out = deq_out; a.ld_out = ld_out;
    a.packed = packed ? 1 : 0;
    a.pack.codes = codes; a.pack.codes_bytes = codes_bytes; a.pack.dir = dir; a.pack.hdr = a.hdr;
    a.pack.W0 = expected_width(m > 0 ? m : 1, d);
    const int64_t nT = n * L.T;
    a.pack.arena_base16 = (unsigned long long)nT * 32ull * (unsigned long long)a.pack.W0;
    if (packed && (long long)(a.pack.arena_base16 * 16ull) > codes_bytes) {
        set_error("code arena too small for the primary slots: %lld < %llu bytes", (long long)codes_bytes, a.pack.arena_base16 * 16ull);
        return DME_EWORKSPACE;
    }
    const size_t dyn = (size_t)(kRing + 2) * kTile * sizeof(float) + sizeof(Scratch);
    if (g_sms == 0) {
        int dev = 0;
        DME_CUDA(cudaGetDevice(&dev));
        DME_CUDA(cudaFuncSetAttribute(quantize_stream_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        DME_CUDA(cudaFuncSetAttribute(quantize_stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        DME_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_occ[0], quantize_stream_kernel<0>, kBlock, dyn));
        DME_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_occ[1], quantize_stream_kernel<1>, kBlock, dyn));
        cudaDriverEntryPointQueryResult qres;
        void *fn = nullptr;
        DME_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (fn == nullptr || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return DME_ECUDA; }
        g_encode = (EncodeTiledFn)fn;
        DME_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    // 3-D view of the client rows: {32 floats, full 128-byte rows of a client, clients}; the last d % 32 coordinates
    // of every row are read directly by the kernel
    CUtensorMap tmap;
    {
        const cuuint64_t dims[3] = {32, (cuuint64_t)(a.rows32 > 0 ? a.rows32 : 1), (cuuint64_t)n};
        const cuuint64_t strides[2] = {128, (cuuint64_t)ld * 4};
        const cuuint32_t box[3] = {32, kTile / 32, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = g_encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)X, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for n=%lld d=%lld ld=%lld", (int)r, (long long)n, (long long)d, (long long)ld); return DME_ECUDA; }
    }
    const int occ = g_occ[packed ? 1 : 0];
    if (occ < 1) { set_error("quantize_stream_kernel does not fit on an SM"); return DME_ECUDA; }
    int64_t G = (int64_t)g_sms * occ;
    if (const char *e = getenv("DME_DBG_G")) G = atoll(e);
    if ((G & 1) == 0) --G;                       // odd: every CTA alternates pass-A and pass-B items
    if (G < 1) G = 1;
    a.goff = G / 2 + 16;
    a.lag = L.T + a.goff + G + 16;
    if (const char *e = getenv("DME_DBG_LAG")) a.lag = L.T + a.goff + atoll(e);
    const int64_t lenA = nT + a.goff + 1, lenB = nT + a.lag;
    a.total_items = 2 * (lenA > lenB ? lenA : lenB);
    a.G = G;
    a.step_c = (int)(G / L.T); a.step_t = (int)(G % L.T);
    a.dbg = 0;
    if (const char *e = getenv("DME_DBG")) a.dbg = atoi(e);
    void *args[] = {&a, &tmap};
    const void *fn = packed ? (const void *)quantize_stream_kernel<1> : (const void *)quantize_stream_kernel<0>;
    DME_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)G), dim3(kBlock), args, dyn, st));
    count_launch();
    return DME_OK;
}

}  // namespace dme
