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
// Inside a CTA.  Tiles arrive through a ring of 1-D bulk async copies (TMA, SASS UBLKCP) issued one item ahead.
// A pass-B tile is split in two stages: stage 1 (division, floor, fractional parts, block scan, publish the tile
// aggregate) and stage 2 (decoupled look-back, prefix -> floor(c - X), type vector, emit).  The pass-A item that
// follows every pass-B tile runs BETWEEN the two stages, which gives the look-back a head start while the tile's
// state (floor, fraction, signs: 33 registers) waits in registers; 64 registers per thread keep 4 CTAs (32 warps)
// resident per SM, which hides the rest.  (A fully pipelined variant with two tile states is kept behind
// DME_STREAM_PIPELINED=1; it needs 128 registers and measured slower.)
//
// Cross-CTA messages are 16-byte records {value, flag} written and read with single 128-bit accesses, so no
// fences are needed; tile aggregates travel as int64 fixed point, which makes the look-back result independent of
// timing (integer addition is associative).
#include <cooperative_groups.h>
#include <cstdlib>
#include <type_traits>

#include "type_quantize.cuh"

namespace dme {

#ifndef DME_STREAM_RING
#define DME_STREAM_RING 2
#endif
#ifndef DME_STREAM_AHEAD
#define DME_STREAM_AHEAD (DME_STREAM_RING - 1)
#endif
constexpr int kRing = DME_STREAM_RING;   // TMA ring depth
constexpr int kAhead = DME_STREAM_AHEAD; // items prefetched ahead (<= kRing - 1: a slot is refilled only after the
                                         // barrier inside the item that consumed it)
static_assert(kAhead >= 1 && kAhead <= kRing - 1, "ring too shallow");

struct __align__(16) Rec { unsigned long long v; uint32_t flag; uint32_t pad; };
typedef Rec TileRec;      // per tile: flag 1 = v is the tile aggregate, 2 = v is the inclusive prefix (fixed point)

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

// ---- async-copy / mbarrier primitives (TMA 1-D bulk copy)
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

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#ifdef DME_TIMERS
#define TIC(a, k) do { if (((a).dbg & 32) && threadIdx.x == 0) sc.tacc[k] -= gtime(); } while (0)
#define TOC(a, k) do { if (((a).dbg & 32) && threadIdx.x == 0) sc.tacc[k] += gtime(); } while (0)
#else
#define TIC(a, k) do { } while (0)
#define TOC(a, k) do { } while (0)
#endif

struct StreamArgs {
    const float *X; int64_t d, ld, T, n, m;
    RowConst *consts; TileRec *desc; Rec *blocks; int64_t TB; WsHeader *hdr; Rec *partial; uint32_t *row_ready;
    const float *x_inject; const float *l1_inject; uint64_t seed, client0; float *l1_out;
    int64_t lag, goff, total_items, G;
    int32_t *k_out; uint8_t *sgn_out; float *deq_out; int64_t ld_out;     // array outputs
    PackTarget pack; int packed;                                        // packed output
    int dbg;                                                            // development: bit 5 = phase timers
};

// One work item, decoded once by thread 0 when it issues the tile's bulk copy and shared through a ring.
struct __align__(16) Item { int c, t; int copied; int fin_row; int valid, is_b; int pad0, pad1; };
// Incremental item decoder (owned by the producer thread): the CTA's items advance by G in the global order, i.e. by
// G positions in the A stream and in the B stream alternately, so (client, tile) pairs are updated without divisions.
struct Decoder { long long i; int cA, tA, cB, tB, cF, tF; int pad; };
constexpr int kProducer = 32;     // warp 1 lane 0 decodes items and issues the bulk copies
constexpr int kSummer = 64;       // warp 2 lane 0 finishes the pass-A tile sum
__device__ __forceinline__ void floor_divmod(long long s, long long T, int &c, int &t) {
    long long q = s >= 0 ? s / T : -((-s + T - 1) / T);
    c = (int)q; t = (int)(s - q * T);
}
__device__ __forceinline__ void advance(int &c, int &t, int step, int T) {
    t += step;
    if (t >= T) { const int q = t / T; c += q; t -= q * T; }
}
__device__ __forceinline__ void decoder_init(Decoder &d, const StreamArgs &a, long long g) {
    d.i = g;
    const long long firstA = (g & 1) ? g + a.G : g, firstB = (g & 1) ? g : g + a.G;
    floor_divmod(firstA >> 1, a.T, d.cA, d.tA);
    floor_divmod((firstA >> 1) - a.goff, a.T, d.cF, d.tF);
    floor_divmod((firstB >> 1) - a.lag, a.T, d.cB, d.tB);
    d.pad = 0;
}
// Next item of the CTA (items alternate between the two streams because G is odd).
__device__ __forceinline__ Item decoder_next(Decoder &d, const StreamArgs &a) {
    Item it; it.valid = 0; it.is_b = (int)(d.i & 1); it.c = 0; it.t = 0; it.copied = 0; it.fin_row = -1; it.pad0 = it.pad1 = 0;
    const bool live = d.i < a.total_items;
    d.i += a.G;
    int c, t;
    if (!it.is_b) {
        c = d.cA; t = d.tA;
        if (live && d.tF == 0 && d.cF >= 1 && d.cF - 1 < a.n) it.fin_row = d.cF - 1;     // finaliser duty of this position
        advance(d.cA, d.tA, (int)a.G, (int)a.T);
        advance(d.cF, d.tF, (int)a.G, (int)a.T);
    } else {
        c = d.cB; t = d.tB;
        advance(d.cB, d.tB, (int)a.G, (int)a.T);
    }
    if (!live || c < 0 || c >= a.n) return it;
    it.valid = 1; it.c = c; it.t = t;
    const int64_t rem = a.d - (int64_t)t * kTile;
    it.copied = rem >= kTile ? kTile : (int)(rem & ~(int64_t)3);
    return it;
}

struct Scratch {
    double wtot[kWarps];
    double red[kWarps];
    double redA[kWarps];
    double P[2];
    long long Pq[2];
    RowConst rc[2];
    long long rc_row[2];
    PackScratch pack;
    Item items[kRing];
    Decoder dec;
    Rec lb[64];                // prefetched look-back windows: 32 tile records + 32 block records (warp 0)
    unsigned long long tacc[8];  // phase timers (dbg)
};

// value of tile-local coordinate e from the staged tile; beyond `copied` floats fall back to global / zero
__device__ __forceinline__ float staged(const float *buf, int e, int copied, const float *row, int64_t tile0, int64_t d) {
    if (e < copied) return buf[e];
    const int64_t i = tile0 + e;
    return i < d ? row[i] : 0.0f;
}

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

// Cold path: reduce row `row`'s tile sums in a fixed order (thread-strided, then the block tree), publish the row.
__device__ void finalize_row(const StreamArgs &a, int row, Scratch &sc) {
    const Rec *pp = a.partial + (int64_t)row * a.T;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < a.T; i += kThreads) {
        unsigned long long v;
        while (rec_load(pp + i, v) == 0u) __nanosleep(64);
        acc += __longlong_as_double((long long)v);
    }
    acc = block_sum_f64(acc, sc.red);
    if (threadIdx.x == 0) {
        make_row_const(a, row, acc);
        __threadfence();
        st_release_u32(&a.row_ready[row], 1u);
    }
}

// ---- pass A of one tile (+ the finaliser duty attached to this stream position)
__device__ __forceinline__ void pass_a(const StreamArgs &a, const Item &it, const float *buf, Scratch &sc) {
    if (it.valid) {
        const float *row = a.X + (int64_t)it.c * a.ld;
        const int64_t tile0 = (int64_t)it.t * kTile;
        double s = 0.0;
        if (it.copied == kTile) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v = *reinterpret_cast<const float4 *>(buf + q * 1024 + 4 * threadIdx.x);
                s += (double)fabsf(v.x); s += (double)fabsf(v.y); s += (double)fabsf(v.z); s += (double)fabsf(v.w);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int e = 0; e < 4; ++e) s += (double)fabsf(staged(buf, q * 1024 + 4 * threadIdx.x + e, it.copied, row, tile0, a.d));
        }
        // fixed association: xor butterfly inside a warp, warps in index order (same tree as block_sum_f64)
        s = warp_sum_f64(s);
        if ((threadIdx.x & 31) == 0) sc.redA[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == kSummer) {
            double tot = sc.redA[0];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) tot += sc.redA[w];
            rec_store(&a.partial[(int64_t)it.c * a.T + it.t], (unsigned long long)__double_as_longlong(tot), 1u);
        }
    }
    if (it.fin_row >= 0) finalize_row(a, it.fin_row, sc);
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

// State of a pass-B tile between its two stages (registers).
template <int EMIT>
struct BState {
    // floor(m p): integer for the packed output (magnitudes >= 2^31 raise DME_EOVERFLOW), float for the array output
    typename std::conditional<EMIT == 1, int, float>::type fl[kEpt];
    float fr[kEpt];
    uint32_t sign;          // bit j = IEEE sign of coordinate j
    double base;            // in-tile exclusive prefix of this thread's first coordinate
    double end;             // in-tile inclusive prefix of this thread's last coordinate (scan value)
    long long Aq;           // tile aggregate, fixed point
    int c, t;
    int live;
};

// ---- stage 1: everything that does not need the prefix of earlier tiles
template <int EMIT>
__device__ __forceinline__ void stage1(const StreamArgs &a, const Item &it, const float *buf, Scratch &sc, int slot, BState<EMIT> &st) {
    st.live = it.valid;
    if (!it.valid) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    st.c = it.c; st.t = it.t;
    // row constants: cached in shared memory; a new row waits for the finaliser (an earlier A item)
    if (sc.rc_row[slot] != it.c) {
        __syncthreads();
        if (threadIdx.x == 0) {
            while (ld_acquire_u32(&a.row_ready[it.c]) == 0u) __nanosleep(64);
            const uint4 *src = reinterpret_cast<const uint4 *>(&a.consts[it.c]);
            uint4 *dst = reinterpret_cast<uint4 *>(&sc.rc[slot]);
#pragma unroll
            for (int q = 0; q < (int)(sizeof(RowConst) / 16); ++q) dst[q] = __ldcg(src + q);
            sc.rc_row[slot] = it.c;
        }
        __syncthreads();
    }
    const RowConst &rc = sc.rc[slot];
    const float *row = a.X + (int64_t)it.c * a.ld;
    const int64_t tile0 = (int64_t)it.t * kTile;
    float x[kEpt];
    if (it.copied == kTile) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = *reinterpret_cast<const float4 *>(buf + kEpt * threadIdx.x + 4 * q);
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kEpt; ++j) x[j] = staged(buf, kEpt * threadIdx.x + j, it.copied, row, tile0, a.d);
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
    for (int j = kEpt - 1; j >= 0; --j) sg = __funnelshift_l(__float_as_uint(x[j]), sg, 1);     // bit j = sign of x[j]
    st.sign = sg;
    if (exact || big) {
        bool ovf = false;
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            float flf;
            chain<true>(x[j], rc, flf, st.fr[j]);
            if (EMIT == 1) {
                if (flf >= 2147483520.0f) { ovf = true; st.fl[j] = 0x7ffffffe; } else st.fl[j] = (int)flf;
            } else st.fl[j] = flf;
        }
        if (ovf) atomicOr(&a.hdr->status, 1u);
    } else {
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            // fast chain: x/D by Markstein's correction of x * rcp, floor by the 2^23 trick (see chain<false>)
            const float ax = fabsf(x[j]);
            const float q0 = __fmul_rn(ax, rc.rcpD);
            const float rem = __fmaf_rn(-q0, rc.D, ax);
            const float pq = __fmaf_rn(rem, rc.rcpD, q0);
            const float mp = __fmul_rn(rc.mf, pq);
            const float tt = __fadd_rz(mp, 8388608.0f);
            const float flf = __fsub_rn(tt, 8388608.0f);
            st.fr[j] = __fsub_rn(mp, flf);
            if (EMIT == 1) st.fl[j] = __float_as_int(tt) - 0x4b000000;       // the integer sits in the mantissa of tt
            else st.fl[j] = flf;
        }
    }
    // thread sum: four groups of four, each summed left to right, then combined left to right -- the same
    // association stage 2 uses for the running prefix, so the thread total equals its last prefix bit for bit
    double S;
    {
        double g4[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            g4[g] = (double)st.fr[4 * g];
#pragma unroll
            for (int e = 1; e < 4; ++e) g4[g] += (double)st.fr[4 * g + e];
        }
        S = ((g4[0] + g4[1]) + g4[2]) + g4[3];
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
    st.base = wbase + excl;
    st.end = wbase + incl;
    st.Aq = __double2ll_rn(A * rc.q_up);                 // fixed point, 2^-qshift resolution
    if (threadIdx.x == 0) rec_store(a.desc + (int64_t)it.c * a.T + it.t, (unsigned long long)st.Aq, 1u);
    // no trailing barrier: wtot and this ring slot are next written after the barriers of the following stage 2
}

// ------------------------------------------------------------------ two-level decoupled look-back (warp 0)
// With ~300 pass-B tiles in flight a flat look-back has to walk ~10 windows of 32 records, and that latency in turn
// keeps more tiles in flight.  Two levels bound it: tiles publish their aggregate {flag 1}; the LAST tile of every
// block of 32 tiles publishes the block aggregate {flag 1} as soon as its 31 predecessors' aggregates are in and the
// block's inclusive prefix {flag 2} once it knows its own.  A tile's exclusive prefix = aggregates of the earlier
// tiles of its block (one window) + block aggregates back to the nearest block-inclusive record (one window): two
// independent 16-byte loads per lane per poll.  All values are int64 fixed point, so any mixture gives the same sum.
#ifdef DME_TIMERS
__device__ unsigned long long g_polls;
#endif
struct LookArgs { const Rec *tiles; Rec *blocks; int t; };   // records of the tile's row
__device__ __forceinline__ void lookback_prefetch(const LookArgs &k, Rec *lb) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x, b = k.t >> 5, pos = k.t & 31;
        if (lane < pos)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(lb + lane)), "l"(k.tiles + (k.t - 1 - lane)) : "memory");
        if (b - 1 - lane >= 0)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(lb + 32 + lane)), "l"(k.blocks + (b - 1 - lane)) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
}
// Returns the exclusive fixed-point prefix P of tile k.t; Aq = the tile's own aggregate (for the block records).
__device__ __forceinline__ long long lookback_resolve(const LookArgs &k, long long Aq, const Rec *lb, int lane) {
    const int b = k.t >> 5, pos = k.t & 31;
    const bool has_t = lane < pos, has_b = (b - 1 - lane) >= 0;
    const bool block_last = pos == 31;
    bool ba_done = false;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    Rec rt = lb[lane], rb = lb[32 + lane];                         // first attempt: the prefetched copies
    while (true) {
        const uint32_t ft = has_t ? rt.flag : 1u;
        const uint32_t fb = has_b ? rb.flag : 2u;                  // blocks before the row start: inclusive prefix 0
        const unsigned t_none = __ballot_sync(0xffffffffu, ft == 0u);
        const unsigned b_incl = __ballot_sync(0xffffffffu, fb == 2u);
        const unsigned b_none = __ballot_sync(0xffffffffu, fb == 0u);
        const int f = b_incl ? (__ffs(b_incl) - 1) : 32;
        const unsigned need = (f >= 31) ? 0xffffffffu : ((1u << (f + 1)) - 1u);
        if (t_none == 0u) {
            long long xt = has_t ? (long long)rt.v : 0;
            if (block_last && !ba_done) {                           // block aggregate: unblocks later blocks early
                long long s = xt;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) rec_store(k.blocks + b, (unsigned long long)(s + Aq), 1u);
                ba_done = true;
            }
            if (f < 32 && (b_none & need) == 0u) {
                long long x = xt + ((has_b && lane <= f) ? (long long)rb.v : 0);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                if (block_last && lane == 0) rec_store(k.blocks + b, (unsigned long long)(x + Aq), 2u);
                return x;
            }
        }
#ifdef DME_TIMERS
        if (lane == 0) atomicAdd(&g_polls, 1ull);
#endif
        __nanosleep(32);
        if (has_t) rt.flag = rec_load(k.tiles + (k.t - 1 - lane), rt.v);
        if (has_b) rb.flag = rec_load(k.blocks + (b - 1 - lane), rb.v);
    }
}

// ---- stage 2: look-back, prefix -> floor(c - X) (AS:635-637), type vector, emit
template <int EMIT>
__device__ __forceinline__ void stage2(const StreamArgs &a, Scratch &sc, int slot, BState<EMIT> &st) {
    if (!st.live) return;
    st.live = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const RowConst &rc = sc.rc[slot];
    const TileRec *rowdesc = a.desc + (int64_t)st.c * a.T;
    if (warp == 0) {
        const LookArgs la{rowdesc, a.blocks + (int64_t)st.c * a.TB, st.t};
        TIC(a, 4);
        const long long P = lookback_resolve(la, st.Aq, sc.lb, lane);
        TOC(a, 4);
        if (lane == 0) { sc.Pq[slot] = P; sc.P[slot] = __ll2double_rn(P) * rc.q_dn; }
    }
    __syncthreads();
    const double Pd = sc.P[slot];
    // The prefix at the LAST coordinate of a warp is defined from the scan values (and at the last coordinate of the
    // tile from the fixed-point inclusive prefix), so the next warp / tile derives the same floor(c - X) for its
    // predecessor from its own exclusive prefix: no hand-off is needed.
    const double E = Pd + st.base;
    int aprev = __float2int_rd(__fsub_rn(__double2float_rn(E), rc.X));       // floor(c_{first-1} - X); c_0 = 0 (AS:635)
    int av[kEpt];
    {
        // running prefix, left to right inside the thread's four groups of four exactly as stage 1 summed them:
        // C_j = E + (group offset + in-group prefix), so the thread's last prefix equals E + S bit for bit
        double off = 0.0, grp = 0.0;
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            grp = (j & 3) ? grp + (double)st.fr[j] : (double)st.fr[j];
            double C = (j < 4) ? (E + grp) : (E + (off + grp));
            if ((j & 3) == 3) off = (j == 3) ? grp : off + grp;
            if (j == kEpt - 1 && lane == 31) C = Pd + st.end;
            if (j == kEpt - 1 && threadIdx.x == kThreads - 1) C = __ll2double_rn(sc.Pq[slot] + st.Aq) * rc.q_dn;
            av[j] = __float2int_rd(__fsub_rn(__double2float_rn(C), rc.X));   // AS:636
        }
    }
    // predecessor's floor: from the previous lane; lane 0 uses its own exclusive prefix (computed above)
    const int from_prev = __shfl_up_sync(0xffffffffu, av[kEpt - 1], 1);
    if (lane != 0) aprev = from_prev;
    const int64_t i0 = (int64_t)st.t * kTile + (int64_t)threadIdx.x * kEpt;
    if (EMIT == 0) {
        bool ovf = false;
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            const int r = (av[j] - aprev == 1) ? 1 : 0;                       // AS:636-637
            aprev = av[j];
            const float kf = __fadd_rn((float)st.fl[j], (float)r);
            const int64_t i = i0 + j;
            if (i >= a.d) continue;
            const uint32_t sbit = (st.sign >> j) & 1u;
            if (a.deq_out) {
                // sign(v) of AS:640: v = x / D is zero exactly when m * |v| is (floor and fraction both zero, m > 0)
                const float sgf = ((float)st.fl[j] == 0.0f && st.fr[j] == 0.0f) ? 0.0f : (sbit ? -1.0f : 1.0f);
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
        // type vector in place of av: k_j = floor + [floor(c_j - X) - floor(c_{j-1} - X) == 1]
        int kmax = 0;
        uint32_t kbits = 0;                 // bit j = k_j when every k_j <= 1 (the only case the 2-bit path is taken)
#pragma unroll
        for (int j = kEpt - 1; j >= 0; --j) {
            const int prev = (j == 0) ? aprev : av[j - 1];
            const int k = (int)st.fl[j] + ((av[j] - prev == 1) ? 1 : 0);
            av[j] = k;                      // av[j-1] is still the floor value when it is read in the next iteration
            kmax = max(kmax, k);
            kbits = kbits * 2u + (uint32_t)k;
        }
        // tile-wide minimal field width
        kmax = __reduce_max_sync(0xffffffffu, kmax);
        if (lane == 0) sc.pack.u32[warp] = (uint32_t)kmax;
        __syncthreads();
        uint32_t km = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) km = max(km, sc.pack.u32[w]);
        int W = 2;
        while (W < 32 && km >= (1u << (W - 1))) W <<= 1;
        const int64_t slot_id = (int64_t)st.c * a.T + st.t;
        unsigned long long off16;
        if (W <= a.pack.W0) {
            off16 = (unsigned long long)slot_id * (32ull * a.pack.W0);
            if (threadIdx.x == 0) a.pack.dir[slot_id] = (off16 << 8) | (unsigned long long)W;
        } else {
            if (threadIdx.x == 0) {
                const unsigned long long units = 32ull * W;
                unsigned long long off = a.pack.arena_base16 + atomicAdd(&a.hdr->arena_top, units);
                if ((long long)((off + units) * 16ull) > a.pack.codes_bytes) { atomicOr(&a.hdr->status, 2u); off = ~0ull; }
                sc.pack.off16 = off;
                a.pack.dir[slot_id] = (off == ~0ull) ? 0ull : ((off << 8) | (unsigned long long)W);
            }
            __syncthreads();
            off16 = sc.pack.off16;
        }
        if (off16 != ~0ull) {
            uint32_t *tw = a.pack.codes + off16 * 4ull;
            if (W == 2) {
                // fields [sign | magnitude bit]: interleave the two 16-bit masks
                uint32_t lo = kbits, hi = st.sign;
                lo = (lo | (lo << 8)) & 0x00ff00ffu; hi = (hi | (hi << 8)) & 0x00ff00ffu;
                lo = (lo | (lo << 4)) & 0x0f0f0f0fu; hi = (hi | (hi << 4)) & 0x0f0f0f0fu;
                lo = (lo | (lo << 2)) & 0x33333333u; hi = (hi | (hi << 2)) & 0x33333333u;
                lo = (lo | (lo << 1)) & 0x55555555u; hi = (hi | (hi << 1)) & 0x55555555u;
                tw[threadIdx.x] = lo | (hi << 1);
            } else {
                uint32_t k[kEpt], sg[kEpt];
#pragma unroll
                for (int j = 0; j < kEpt; ++j) { k[j] = (uint32_t)av[j]; sg[j] = (st.sign >> j) & 1u; }
                switch (W) {
                    case 4: pack_store<4>(k, sg, tw); break;
                    case 8: pack_store<8>(k, sg, tw); break;
                    case 16: pack_store<16>(k, sg, tw); break;
                    default: pack_store<32>(k, sg, tw); break;
                }
            }
        }
        // no trailing barrier: the scratch used here is next written after the first barrier of the next stage 2
    }
}

// Measured on B200 (d=2^24, n=128): fully pipelined (2 tile states, 128 regs, 2 CTAs/SM) 9.8 ms; no pipelining
// 3 CTAs 9.0 ms, 4 CTAs 8.0 ms; half-pipelined (below) 4 CTAs/SM, ring 2: 7.2 ms; ring 3: 7.6 ms.
#ifndef DME_STREAM_CTAS
#define DME_STREAM_CTAS 4
#endif
#ifndef DME_STREAM_PIPELINED
#define DME_STREAM_PIPELINED 0
#endif
template <int EMIT>
__global__ void __launch_bounds__(kThreads, DME_STREAM_CTAS)
quantize_stream_kernel(StreamArgs a) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ uint64_t mbar[kRing];
    __shared__ Scratch sc;

    if (threadIdx.x == 0) {
        for (int b = 0; b < kRing; ++b) mbar_init(&mbar[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sc.rc_row[0] = sc.rc_row[1] = -1;
        for (int q = 0; q < 8; ++q) sc.tacc[q] = 0;
        if (a.dbg & 32) sc.tacc[6] -= gtime();
    }
    __syncthreads();
    const uint64_t pol_a = policy_evict_last(), pol_b = policy_evict_first();
    const int64_t g = blockIdx.x;
    int64_t jn = 0;                      // local index of the next item to fetch
    uint32_t parity_bits = 0;            // phase parity of each ring slot

    auto issue = [&](int64_t j) {        // producer thread: decode local item j, start its bulk copy (if it has data)
        const Item it = decoder_next(sc.dec, a);
        const int b = (int)(j % kRing);
        sc.items[b] = it;
        if (it.valid && it.copied > 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&mbar[b], (uint32_t)it.copied * 4u);
            bulk_g2s(dyn_smem + (size_t)b * kTile * 4, a.X + (int64_t)it.c * a.ld + (int64_t)it.t * kTile, (uint32_t)it.copied * 4u, &mbar[b],
                     it.is_b ? pol_b : pol_a);
        }
    };
    if (threadIdx.x == kProducer) {
        decoder_init(sc.dec, a, g);
        for (int64_t j = 0; j < kAhead; ++j) issue(j);
    }
    __syncthreads();
    // fetch(): next item of this CTA, its staged tile ready in shared memory
    auto fetch = [&](Item &it, const float *&buf) -> bool {
        const int64_t j = jn++;
        if (g + j * a.G >= a.total_items) return false;
        if (threadIdx.x == kProducer) issue(j + kAhead);
        const int b = (int)(j % kRing);
        it = sc.items[b];
        buf = reinterpret_cast<const float *>(dyn_smem + (size_t)b * kTile * 4);
        if (it.valid && it.copied > 0) {
            TIC(a, 0);
            mbar_wait(&mbar[b], (parity_bits >> b) & 1u);
            parity_bits ^= 1u << b;
            TOC(a, 0);
        }
        // Every item must contain at least one CTA barrier: thread 0 runs the item ring and the TMA ring two items
        // ahead, and a slot may only be overwritten after all threads have passed a barrier since they read it.
        // Valid items have theirs inside pass A / stage 1; empty items (stream head and tail) get one here.
        if (!it.valid) __syncthreads();
        return true;
    };

#if DME_STREAM_PIPELINED
    BState<EMIT> s0, s1;
    s0.live = 0; s1.live = 0;
    Item it; const float *buf = nullptr;
    while (true) {
        // ---- pass-B tile into state 0 (an A item may come first)
        if (!fetch(it, buf)) break;
        if (!it.is_b) { TIC(a, 1); pass_a(a, it, buf, sc); TOC(a, 1); if (!fetch(it, buf)) break; }
        if (s1.live) lookback_prefetch(LookArgs{a.desc + (int64_t)s1.c * a.T, a.blocks + (int64_t)s1.c * a.TB, s1.t}, sc.lb);
        TIC(a, 2); stage1<EMIT>(a, it, buf, sc, 0, s0); TOC(a, 2);
        TIC(a, 3); stage2<EMIT>(a, sc, 1, s1); TOC(a, 3);
        // ---- pass-B tile into state 1
        if (!fetch(it, buf)) break;
        if (!it.is_b) { TIC(a, 1); pass_a(a, it, buf, sc); TOC(a, 1); if (!fetch(it, buf)) break; }
        if (s0.live) lookback_prefetch(LookArgs{a.desc + (int64_t)s0.c * a.T, a.blocks + (int64_t)s0.c * a.TB, s0.t}, sc.lb);
        TIC(a, 2); stage1<EMIT>(a, it, buf, sc, 1, s1); TOC(a, 2);
        TIC(a, 3); stage2<EMIT>(a, sc, 0, s0); TOC(a, 3);
    }
    // drain: the last parked tiles (their look-back falls back to polling when nothing was prefetched)
    if (s0.live) lookback_prefetch(LookArgs{a.desc + (int64_t)s0.c * a.T, a.blocks + (int64_t)s0.c * a.TB, s0.t}, sc.lb);
    stage2<EMIT>(a, sc, 0, s0);
    if (s1.live) lookback_prefetch(LookArgs{a.desc + (int64_t)s1.c * a.T, a.blocks + (int64_t)s1.c * a.TB, s1.t}, sc.lb);
    stage2<EMIT>(a, sc, 1, s1);
#else
    // half-pipelined variant: one tile state in registers; the pass-A item that follows a pass-B tile (items
    // alternate because G is odd) runs between the tile's two stages and gives its look-back a head start
    BState<EMIT> s0;
    s0.live = 0;
    Item it; const float *buf = nullptr;
    while (true) {
        if (!fetch(it, buf)) break;
        if (!it.is_b) { TIC(a, 1); pass_a(a, it, buf, sc); TOC(a, 1); continue; }
        TIC(a, 2); stage1<EMIT>(a, it, buf, sc, 0, s0); TOC(a, 2);
        if (s0.live) lookback_prefetch(LookArgs{a.desc + (int64_t)s0.c * a.T, a.blocks + (int64_t)s0.c * a.TB, s0.t}, sc.lb);
        const bool more = fetch(it, buf);
        if (more && !it.is_b) { TIC(a, 1); pass_a(a, it, buf, sc); TOC(a, 1); }
        TIC(a, 3); stage2<EMIT>(a, sc, 0, s0); TOC(a, 3);
        if (!more) break;
        if (it.is_b) {                    // cannot happen while G is odd; kept so that no item is ever dropped
            stage1<EMIT>(a, it, buf, sc, 0, s0);
            stage2<EMIT>(a, sc, 0, s0);
        }
    }
#endif
    if ((a.dbg & 32) && threadIdx.x == 0) {
        sc.tacc[6] += gtime();
#ifdef DME_TIMERS
        if (blockIdx.x == 0) { sc.tacc[5] = g_polls; g_polls = 0; }
#endif
        for (int q = 0; q < 8; ++q) atomicAdd(reinterpret_cast<unsigned long long *>(a.hdr->pad + 1) + q, sc.tacc[q]);
    }
}

static int g_sms = 0, g_occ[2] = {0, 0};

int launch_stream(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, const WsLayout &L, void *ws,
                  const float *x_inject, const float *l1_inject, uint64_t seed, uint64_t client0,
                  int32_t *k_out, uint8_t *sgn_out, float *deq_out, int64_t ld_out,
                  uint32_t *codes, int64_t codes_bytes, uint64_t *dir, float *l1_out, cudaStream_t st, bool packed) {
    char *base = (char *)ws;
    StreamArgs a;
    a.X = X; a.d = d; a.ld = ld; a.T = L.T; a.n = n; a.m = m;
    a.consts = (RowConst *)(base + L.off_consts);
    a.desc = (TileRec *)(base + L.off_desc);                       // tile records: first 16 bytes per tile of the region
    a.TB = (L.T + 31) / 32;                                        // blocks of 32 tiles per row
    a.blocks = (Rec *)(base + L.off_desc + 16 * n * L.T);         // block records: in the second half of the region
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
    const size_t dyn = (size_t)kRing * kTile * sizeof(float);
    if (g_sms == 0) {
        int dev = 0;
        DME_CUDA(cudaGetDevice(&dev));
        DME_CUDA(cudaFuncSetAttribute(quantize_stream_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        DME_CUDA(cudaFuncSetAttribute(quantize_stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        DME_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_occ[0], quantize_stream_kernel<0>, kThreads, dyn));
        DME_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_occ[1], quantize_stream_kernel<1>, kThreads, dyn));
        DME_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
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
    a.dbg = 0;
    if (const char *e = getenv("DME_DBG")) a.dbg = atoi(e);
    void *args[] = {&a};
    const void *fn = packed ? (const void *)quantize_stream_kernel<1> : (const void *)quantize_stream_kernel<0>;
    DME_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)G), dim3(kThreads), args, dyn, st));
    count_launch();
    return DME_OK;
}

}  // namespace dme
