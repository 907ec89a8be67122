// quantize_fx.cu -- the unbiased type quantizer (AS:609-641) for all clients of one GPU in ONE persistent launch that reads
// the input ONCE from HBM: per-client L1 norm (pass A), scale to m, floor + systematic-sampling allocation of the fractional
// mass (pass B), sign/magnitude packing (or the dequantised output of the drop-in API).
//
// Schedule.  Tiles (4096 coordinates = 16 KB) of all client rows are numbered row-major.  One CTA per SM (cooperative launch);
// SM b owns the positions b, b + S, b + 2S, ... (S = number of SMs) and its four 128-thread GROUPS draw them in order from a
// ticket in shared memory.  At position p a group runs pass A of tile p (first touch, from HBM) and pass B of tile p - Lg.
// Lg = T + lead is a multiple of S, so THE SAME SM runs both passes of a tile: on B200 the two halves of L2 each cache what
// their own SMs touch, and a tile re-read by the SM that streamed it is an L2 hit where a re-read by an arbitrary SM is not
// (tools/ubench3.cu: 8 GiB of 64 MiB rows in 1.46 ms with this mapping, 1.69 ms with global tickets, 2.47 ms for two passes
// over HBM).  Inside an SM the assignment is dynamic: a group that waits for a look-back does not hold up the SM's next
// position -- with a static assignment per CTA every wait delayed the waiter's next aggregate, which delayed its successors
// (measured: 44 % of the tiles polled, 8.7 ms; profiles/r02_fx_history.md).  Every wait is on a smaller position and all CTAs
// are resident: no deadlock.
//
// Pass A.  Coalesced 128-bit loads, |x| widened to fp64 by ONE integer multiply-add (the fp32 bit pattern times 2^29 is
// the fp64 pattern up to the exponent bias; the conversion unit is kept for pass B), one fp64 sum per tile (fixed association).
// Tile sums are added into a per-row superaccumulator of twelve 64-bit bins (32 bits of payload each) with integer atomics:
// integer addition is associative, so the row's norm does not depend on which group streamed which tile or when.  Every group
// derives a row's constants from the bins when it first needs them (same arithmetic, same bits everywhere).
//
// Pass B, in integers.  y = m|x|/D (AS:625-629: Markstein division x*rcp corrected by two FMAs = the IEEE quotient, in packed
// f32x2) is converted ONCE to 64-bit fixed point with 2^-32 resolution, F = RN(y 2^32): the high word is floor(y) (AS:630),
// the low word the fractional part (AS:631).  Prefixes of the low words are integer sums -- associative, so thread / warp /
// tile boundaries and timing cannot change a result -- and AS:636's t = floor(RN32(RN32(c) - X)) is, while the prefix stays
// inside one binade [2^e + 1, 2^(e+1)), 2 <= e <= 22, t = floor((P - U_e) / 2^32) with U_e = g (a - 1/2) (+ 1 unit for odd a),
// g = 2^(e-23), a = ceil(X/g - 1/2): r_j = [t_j - t_(j-1) == 1] (AS:637) is the CARRY of a 32-bit running sum: one add with
// carry-out and one add-with-carry per coordinate, no fp64, no conversion.  Threads whose prefixes cross a binade (or sit
// below 5.5) evaluate AS:636 literally on integers (round to 24 significant bits twice).  tests/test_fixed_point_scan.py
// restates both in Python and checks them against the oracle.
//   B-phase (tile i)  : chain, conversion, low words parked in place of x, thread sums -> warp scan -> tile aggregate, published
//                       at once as a 16-byte record + added to the records of its block (32 tiles) and super-block (1024);
//   C-phase (tile i-1): exclusive prefix from the look-back window (copied by cp.async right before the barrier), carry walk,
//                       emit (2- and 4-bit tiles straight from registers).
//
// Rows outside the proven operand range of the fast chain (D outside [2^-20, 2^100] or with an all-ones mantissa, X off the
// 2^-32 grid or outside [0, 1)) are listed in the workspace and left to literal_rows_kernel (quantize_literal.cu).
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "type_quantize.cuh"

namespace dme {

typedef unsigned long long u64;
constexpr int kFxThreads = 128, kFxWarps = kFxThreads / 32;
constexpr int kFxMaxG = 1024;             // partial slots per row in the workspace

// ---- records of the decoupled look-back
struct __align__(16) Rec { u64 v; uint32_t flag; uint32_t pad; };     // per tile: flag 1 = v is the tile aggregate
struct __align__(16) Rec2 { u64 lo, hi; };                            // per block / super-block: two 31-bit halves + counts
constexpr int kCntShift = 44;
constexpr u64 kSumMask = (1ull << kCntShift) - 1ull;

__device__ __forceinline__ void rec_store(Rec *p, u64 v, uint32_t flag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((uint32_t)v), "r"((uint32_t)(v >> 32)), "r"(flag), "r"(0u) : "memory");
}
__device__ __forceinline__ uint32_t rec_load(const Rec *p, u64 &v) {
    uint32_t a, b, f, z;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(f), "=r"(z) : "l"(p) : "memory");
    (void)z;
    v = ((u64)b << 32) | a;
    return f;
}
__device__ __forceinline__ void rec2_load(const Rec2 *p, u64 &lo, u64 &hi) {
    uint32_t a, b, c, d;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
    lo = ((u64)b << 32) | a;
    hi = ((u64)d << 32) | c;
}
__device__ __forceinline__ void red_add_u64(u64 *p, u64 v) { asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

// ---- async-copy / mbarrier / named-barrier primitives
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"      // suspend-time hint: sleep in hardware, do not spin
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// one 16 KB box {32 floats, 128 rows, 1 client} at (0, row0, client) of the 3-D tensor map
__device__ __forceinline__ void tma_tile_g2s(uint32_t dst, const CUtensorMap *map, int row0, int client, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(dst), "l"(map), "r"(0), "r"(row0), "r"(client), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t policy_evict_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
// pass-A load: 128 bits, not cached in L1, kept in L2 (the CTA comes back for this tile Lg positions later)
__device__ __forceinline__ uint4 ldg_keep_u4(const float *p, uint64_t pol) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    return r;
}

// packed f32x2 arithmetic (sm_100: FMUL2 / FFMA2, one issue slot for two coordinates)
typedef unsigned long long f2;
__device__ __forceinline__ f2 f2_pack(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(f2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 f2u64_abs(float v) { u64 r; asm("cvt.rni.u64.f32 %0, %1;" : "=l"(r) : "f"(fabsf(v))); return r; }


#ifndef FX_GROUPS
#define FX_GROUPS 4
#define FX_AWARPS 4
#define FX_CTAS 1
#endif
constexpr int kFxGroups = FX_GROUPS;      // independent 128-thread groups per CTA
constexpr int kFxAWarps = FX_AWARPS;        // pass-A warps per CTA (they only stream and add)
constexpr int kFxCtas = FX_CTAS;            // CTAs per SM
constexpr int kFxBlock = kFxGroups * kFxThreads + kFxAWarps * 32;
constexpr int kBins = 12;                 // 64-bit bins of a row's |x| superaccumulator (bin q weighs 2^(32 q - 224))
constexpr int kBinBias = 224;

#ifdef DME_TIMERS
__device__ unsigned long long g_fx_trace[1 << 16];      // development: {globaltimer, code << 56 | tile position}
__device__ unsigned int g_fx_trace_n;
__device__ __forceinline__ void fx_trace(int code, long long pos) {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned int i = atomicAdd(&g_fx_trace_n, 1u);
    if (i < (1u << 15)) { g_fx_trace[2 * i] = t; g_fx_trace[2 * i + 1] = ((unsigned long long)code << 56) | (unsigned long long)pos; }
}
#define FXTR(code, c, t) do { if (gtid == 0 && (c) == 40) fx_trace(code, (long long)(t)); } while (0)
#define FXW(code) do { if (lane == 0 && g == 0 && blockIdx.x == 5 && it >= 200 && it < 232) fx_trace(code, (long long)it * 8 + warp); } while (0)
#define FXT(k) do { if (gtid == 0) { const long long _t = clock64(); tacc[k] += _t - tlast; tlast = _t; } } while (0)
#else
#define FXT(k) do { } while (0)
#define FXTR(code, c, t) do { } while (0)
#define FXW(code) do { } while (0)
#endif
struct FxArgs {
    const float *X; int64_t d, ld, T, n, m;
    int64_t rows32;                        // full 128-byte rows per client vector (the part the tensor map covers)
    RowConst *consts; Rec *desc; Rec2 *blocks; Rec2 *supers; int64_t TB, TS; WsHeader *hdr;
    u64 *bins; uint32_t *pub_count; uint32_t *row_ready; int32_t *exotic_rows;
    const float *x_inject; const float *l1_inject; uint64_t seed, client0; float *l1_out;
    int64_t nT, Lg;                        // tiles in all rows; pass B runs Lg positions behind pass A (a multiple of S)
    int S;                                 // CTAs = SMs: SM b owns positions b, b + S, ...
    int tiles_tma, has_tail;
    int32_t *k_out; uint8_t *sgn_out; float *deq_out; int64_t ld_out;     // array outputs
    PackTarget pack;                                                    // packed output
    int hack;
    uint32_t one;                          // = 1, as a register operand: keeps the 64-bit accumulate one IMAD.WIDE
};

// one drawn position: pass-B tile (c, t) and pass-A tile (ca, ta)
struct __align__(16) FxItem { int c, t; uint32_t flags; uint32_t pad; };
constexpr uint32_t kItValid = 1u, kItTma = 4u, kItTail = 8u, kItEnd = 16u;

struct FxScratch {               // per group
    Rec win[2][96];              // look-back window of the C tile (tile / block / super-block records), by iteration parity
    u64 wtot[2][kFxWarps];       // warp totals of the B tile, by iteration parity
    uint32_t fmw[2][kFxWarps];   // per-warp largest floor
    FxItem item[3];
    uint64_t mbar[3];
    uint64_t winbar[2];          // completion of the look-back window copies (32 arrivals: the lanes of warp 1), by iteration parity
    unsigned int hit[2];
    int rc_row[2];
    u64 off16;
    long long pfb;               // look-back result of the fallback path
    RowConst rc[2];
    u64 tab[2][kBinades];        // U_e of the closed form, per binade
};
constexpr size_t kGroupBytes = (size_t)3 * kTile * sizeof(float) + ((sizeof(FxScratch) + 1023) / 1024) * 1024;

// swizzled shared-memory offset of the 16-byte piece q (0..3) of 16-coordinate chunk `ch` inside a 16 KB tile
// (SWIZZLE_128B: piece index within the 128-byte row is XORed with row & 7); q enters as an XOR of (q << 4)
__device__ __forceinline__ uint32_t blocked_off_of(uint32_t ch) {
    const uint32_t row = ch >> 1;
    return row * 128u + ((((ch & 1u) << 2) ^ (row & 7u)) << 4);
}

// ------------------------------------------------------------------ row constants (cold: once per group and client row)
// U_e of the closed form in 2^-32 units: a = ceil(X / g - 1/2), g = 2^(e-23); U = g (a - 1/2) + [a odd]
__device__ __forceinline__ u64 binade_offset(uint32_t Xi, int e) {
    const u64 a = ((u64)Xi + (1ull << (e + 8)) - 1ull) >> (e + 9);
    return (a << (e + 9)) - (1ull << (e + 8)) + (a & 1ull);
}
// The constants of a row, written once by the pass-A lane that counted the row's last tile.
__device__ void make_row_const_fx(const FxArgs &a, int64_t c, double l1sum) {
    RowConst rc;
    bool exotic = false;
    // pass A widens |x| by an integer multiply-add that maps +-0 to 2^-127: a row of zeros sums to d * 2^-127 exactly
    if (l1sum < 5.421010862427522e-20) {                                          // 2^-64
        if (l1sum == (double)a.d * 5.877471754111438e-39) l1sum = 0.0;          // all zero
        else exotic = true;                                                       // denormal-scale row: L1 is redone literally
    }
    rc.L1f = a.l1_inject ? a.l1_inject[c] : (float)l1sum;           // AS:624
    rc.D = __fadd_rn(rc.L1f, 1e-12f);                               // AS:625
    rc.mf = (float)a.m;
    rc.X = a.x_inject ? a.x_inject[c] : philox_client_uniform(a.seed, a.client0 + (uint64_t)c);   // AS:634
    rc.rcpD = __frcp_rn(rc.D);
    // The fast chain (Markstein division) is proven for these operand ranges only (DESIGN.md "Exactness of the fast chain");
    // a row of zeros is fine with any D.
    if (!(rc.D >= 9.5367431640625e-07f && rc.D <= 1.2676506e30f) && !(rc.L1f == 0.0f && !a.l1_inject)) exotic = true;     // 2^-20 .. 2^100
    if ((__float_as_uint(rc.D) & 0x7fffffu) == 0x7fffffu) exotic = true;                                                 // 1/D rounding exception
    const float xs = __fmul_rn(rc.X, 4294967296.0f);
    if (!(rc.X >= 0.0f && rc.X < 1.0f) || floorf(xs) != xs) exotic = true;                                               // X on the 2^-32 grid
    if (!(rc.mf < 2147483648.0f)) exotic = true;
    rc.flags = exotic ? kRowExotic : 0u;
    rc.Xi = exotic ? 0u : (uint32_t)xs;
    rc.mfs = __fmul_rn(rc.mf, 4294967296.0f);
    rc.qshift = 0; rc.q_up = rc.q_dn = 0.0;
    rc.pad2[0] = rc.pad2[1] = rc.pad2[2] = 0.0f;
    a.consts[c] = rc;
    if (exotic) a.exotic_rows[atomicAdd(&a.hdr->pad[1], 1u)] = (int32_t)c;
    else if (a.l1_out) a.l1_out[c] = rc.L1f;
}

// ------------------------------------------------------------------ pass A
// |x| as fp64 by integer moves: bits(|x|) * 2^29 + (896 << 52).  0 maps to 2^-127 (see make_row_const_fx).
__device__ __forceinline__ double abs_to_double(uint32_t bits) {
    u64 r;
    asm("mad.wide.u32 %0, %1, 0x20000000, %2;" : "=l"(r) : "r"(bits & 0x7fffffffu), "l"(0x3800000000000000ull));
    return __longlong_as_double((long long)r);
}
__device__ __forceinline__ void acc_abs4(const uint4 v, double (&s)[4]) {
    s[0] += abs_to_double(v.x); s[1] += abs_to_double(v.y); s[2] += abs_to_double(v.z); s[3] += abs_to_double(v.w);
}
// Add a tile's fp64 sum (>= 0, finite) to the row's superaccumulator.  One thread; fire and forget.
// v = mant * 2^ex: mant << ((ex + bias) % 32) is spread over three 32-bit limbs that go to consecutive bins.
__device__ __forceinline__ void publish_tile_sum(const FxArgs &a, int c, double v) {
    u64 *bins = a.bins + (int64_t)c * kBins;
    const u64 bits = (u64)__double_as_longlong(v);
    const int be = (int)(bits >> 52) & 0x7ff;
    if (be != 0) {                                       // tile sums are 0 or >= 2^-127: never fp64-subnormal
        const u64 mant = (bits & 0xfffffffffffffull) | (1ull << 52);
        const int sh = be - 1075 + kBinBias;             // v = mant * 2^(sh - kBinBias), sh >= 22
        const int q = sh >> 5, r = sh & 31;
        const u64 lo = mant << r;                        // low 64 bits of the 85-bit value
        const u64 hi = r ? (mant >> (64 - r)) : 0ull;
        const u64 l0 = lo & 0xffffffffull, l1 = lo >> 32, l2 = hi;
        if (l0) red_add_u64(bins + q, l0);
        if (l1) red_add_u64(bins + q + 1, l1);
        if (l2) red_add_u64(bins + q + 2, l2);
    }
}
// Count the tile whose sum went to the bins earlier in the iteration (release: the bins first).  Called a C-phase later, when
// the thread's reductions have long been acknowledged and the fence costs next to nothing.
__device__ __forceinline__ void publish_tile_count(const FxArgs &a, int c) {
    uint32_t old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(a.pub_count + c) : "memory");
    if (old == (uint32_t)a.T - 1u) {           // the row's last tile: evaluate the bins from the top (fixed order) and publish the constants
        const u64 *bins = a.bins + (int64_t)c * kBins;
        double t = 0.0;
#pragma unroll
        for (int q = kBins - 1; q >= 0; --q) t = t * 4294967296.0 + (double)__ldcg(bins + q);
        make_row_const_fx(a, c, t * 3.7092061506874214e-68);      // 2^-224
        __threadfence();
        st_release_u32(&a.row_ready[c], 1u);
    }
}
__device__ __forceinline__ void grp_sync(int g) { bar_sync(1 + g, kFxThreads); }
// The constants of row c into cache entry e (+ the binade offsets).  Whole group.
__device__ __noinline__ void load_row(const FxArgs &a, FxScratch &sc, int g, int c, int e) {
    const int gtid = threadIdx.x & (kFxThreads - 1);
    grp_sync(g);                           // the previous users of rc[e] / tab[e] are done
    if (gtid < 32) {
        if (a.hack & 2) { if (gtid == 0) { make_row_const_fx(a, c, 1.3387e7); __threadfence(); } __syncwarp(); }
        else while (ld_acquire_u32(&a.row_ready[c]) == 0u) __nanosleep(100);
        uint4 part = make_uint4(0u, 0u, 0u, 0u);
        if (gtid < (int)(sizeof(RowConst) / 16)) {
            part = __ldcg(reinterpret_cast<const uint4 *>(&a.consts[c]) + gtid);
            reinterpret_cast<uint4 *>(&sc.rc[e])[gtid] = part;
        }
        const uint32_t Xi = __shfl_sync(0xffffffffu, part.w, 1);       // RowConst::Xi is word 7
        if (gtid < kBinades) sc.tab[e][gtid] = (gtid >= 2 && gtid <= 22) ? binade_offset(Xi, gtid) : 0ull;
        if (gtid == 0) sc.rc_row[e] = c;
    }
    grp_sync(g);
}

// ------------------------------------------------------------------ pass B
// the 16 coordinates of chunk `ch` of a tile, from the staged tile (+ the row tail straight from global)
__device__ __forceinline__ void fx_load_x(const FxArgs &a, uint32_t flags, int c, int t, uint32_t buf, uint32_t boff, int ch, float (&x)[kEpt]) {
    if (flags & kItTma) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = lds128((buf + boff) ^ (uint32_t)(q << 4));
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kEpt; ++j) x[j] = 0.0f;
    }
    if (flags & kItTail) {          // the last d % 32 coordinates of the row are not covered by the tensor map
        const int64_t i0 = (int64_t)t * kTile + (int64_t)ch * kEpt, lo = a.rows32 * 32;
        const float *row = a.X + (int64_t)c * a.ld;
        if (i0 + kEpt > lo && i0 < a.d) {
#pragma unroll
            for (int j = 0; j < kEpt; ++j)
                if (i0 + j >= lo && i0 + j < a.d) x[j] = row[i0 + j];
        }
    }
}
// the same coordinates straight from global memory (cold paths: array outputs, wide tiles)
__device__ __noinline__ void fx_load_x_global(const FxArgs &a, int c, int t, float (&x)[kEpt], int ch) {
    const int64_t i0 = (int64_t)t * kTile + (int64_t)ch * kEpt;
    const float *row = a.X + (int64_t)c * a.ld;
    if (i0 + kEpt <= a.d) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = *reinterpret_cast<const float4 *>(row + i0 + 4 * q);
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kEpt; ++j) x[j] = (i0 + j < a.d) ? row[i0 + j] : 0.0f;
    }
}
// AS:625-631 for one coordinate: F = RN(m |x / D| 2^32); q = |x / D| (its zero test is sign(v) == 0 of AS:640)
__device__ __forceinline__ u64 fx_chain(float x, const RowConst &rc, float &q) {
    const float ax = fabsf(x);
    const float q0 = __fmul_rn(ax, rc.rcpD);
    const float rem = __fmaf_rn(-q0, rc.D, ax);
    q = __fmaf_rn(rem, rc.rcpD, q0);
    return f2u64_abs(__fmul_rn(rc.mfs, q));
}

// B-phase of one chunk: signs, conversion, low words parked in place of x, floor masks, running sum of the low words
__device__ __forceinline__ void fx_chunk_b(const FxArgs &a, const FxItem &it, uint32_t buf, uint32_t boff, int ch, const RowConst &rc,
                                           uint32_t &sgw, uint32_t &flm, uint32_t &f4a, uint32_t &f4b, uint32_t &mxh, u64 &run) {
    float x[kEpt];
    fx_load_x(a, it.flags, it.c, it.t, buf, boff, ch, x);
#pragma unroll
    for (int j = kEpt - 1; j >= 0; --j) sgw = __funnelshift_l(__float_as_uint(x[j]), sgw, 2);
    const f2 R2 = f2_pack(rc.rcpD, rc.rcpD), ND = f2_pack(-rc.D, -rc.D), M2 = f2_pack(rc.mfs, rc.mfs);
    uint32_t lo[kEpt], hi[kEpt];
#pragma unroll
    for (int j = 0; j < kEpt; j += 2) {
        const f2 xx = f2_pack(x[j], x[j + 1]);
        const f2 q0 = f2_mul(xx, R2);
        const f2 rem = f2_fma(q0, ND, xx);
        const f2 pq = f2_fma(rem, R2, q0);
        const f2 mp = f2_mul(M2, pq);
        float ma, mb;
        f2_unpack(mp, ma, mb);
        const u64 Fa = f2u64_abs(ma), Fb = f2u64_abs(mb);
        lo[j] = (uint32_t)Fa; lo[j + 1] = (uint32_t)Fb;
        hi[j] = (uint32_t)(Fa >> 32); hi[j + 1] = (uint32_t)(Fb >> 32);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) sts128u((buf + boff) ^ (uint32_t)(q << 4), make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]));
    uint32_t hor = 0;
#pragma unroll
    for (int j = 0; j < kEpt; j += 2) hor |= hi[j] | hi[j + 1];
#pragma unroll
    for (int j = 0; j < kEpt; ++j) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(run) : "r"(lo[j]), "r"(a.one));
    if (hor != 0u) {               // some floor is not zero (rare at R = 1)
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            mxh = max(mxh, hi[j]);
            flm |= (hi[j] != 0u ? 1u : 0u) << (2 * j);
        }
#pragma unroll
        for (int j = 0; j < kEpt / 2; ++j) {
            f4a |= min(hi[j], 15u) << (4 * j);
            f4b |= min(hi[j + kEpt / 2], 15u) << (4 * j);
        }
    }
}

// AS:636 on integers: floor(RN32(RN32(P 2^-32) - X)); rn24 = round to 24 significant bits, ties to even
__device__ __forceinline__ u64 rn24(u64 F) {
    if (F < (1ull << 24)) return F;
    const int s = 40 - __clzll((long long)F);
    u64 q = F >> s;
    const u64 rem = F & ((1ull << s) - 1ull), half = 1ull << (s - 1);
    if (rem > half || (rem == half && (q & 1ull))) ++q;
    return q << s;
}
__device__ __forceinline__ long long t_literal(u64 P, uint32_t Xi) {
    long long g = (long long)rn24(P) - (long long)Xi;
    g = g >= 0 ? (long long)rn24((u64)g) : -(long long)rn24((u64)(-g));
    return g >> 32;
}
// r bits of one chunk the literal way (threads whose prefixes cross a binade): compact, first coordinate on top
__device__ __noinline__ uint32_t walk_literal(uint32_t addr, u64 &c, long long &tp, uint32_t Xi) {
    uint32_t rb = 0;
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
        const uint4 v = lds128u(addr ^ (uint32_t)(q << 4));
        const uint32_t l[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            c += l[j];
            const long long t = t_literal(c, Xi);
            rb = (rb << 1) | ((t - tp == 1) ? 1u : 0u);
            tp = t;
        }
    }
    return rb;
}
// r bits of one chunk by the carries of the 32-bit running sum (closed form)
__device__ __forceinline__ uint32_t walk_carry(uint32_t addr, uint32_t &acc) {
    uint32_t rb = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 v = lds128u(addr ^ (uint32_t)(q << 4));
        asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %1;" : "+r"(acc), "+r"(rb) : "r"(v.x));
        asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %1;" : "+r"(acc), "+r"(rb) : "r"(v.y));
        asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %1;" : "+r"(acc), "+r"(rb) : "r"(v.z));
        asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %1;" : "+r"(acc), "+r"(rb) : "r"(v.w));
    }
    return rb;
}
// 16 compact bits (first coordinate on top) -> bit 2j = coordinate j
__device__ __forceinline__ uint32_t spread_rev16(uint32_t rb) {
    uint32_t v = __brev(rb) >> 16;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}
// eight 2-bit pairs [s | r] (low 16 bits of v) -> eight nibbles [s 0 0 r]
__device__ __forceinline__ uint32_t spread_pairs_to_nibbles(uint32_t v) {
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    return (v & 0x11111111u) | ((v & 0x22222222u) << 2);
}
__device__ __forceinline__ int width_of_u(uint32_t kmax) { return kmax < 2u ? 2 : kmax < 8u ? 4 : kmax < 128u ? 8 : kmax < 32768u ? 16 : 32; }

// ------------------------------------------------------------------ look-back
// warp 1 of a group: start the copies of tile t's look-back window (earlier tiles of its block, earlier blocks of its
// super-block, earlier super-blocks) into shared memory; entries that do not exist are filled with complete neutral records
__device__ __forceinline__ void window_prefetch(const FxArgs &a, Rec *win, int c, int t, int lane) {
    const Rec *tiles = a.desc + (int64_t)c * a.T;
    const Rec2 *blocks = a.blocks + (int64_t)c * a.TB, *supers = a.supers + (int64_t)c * a.TS;
    const int b = t >> 5, pos = t & 31, sb = b >> 5, bpos = b & 31;
    if (lane < pos) cp_async16(smem_u32(&win[lane]), tiles + (t - 1 - lane));
    else { Rec r; r.v = 0; r.flag = 1u; r.pad = 0; win[lane] = r; }
    Rec2 *w2 = reinterpret_cast<Rec2 *>(win);
    if (lane < bpos) cp_async16(smem_u32(&w2[32 + lane]), blocks + (sb * 32 + lane));
    else { Rec2 r; r.lo = r.hi = 32ull << kCntShift; w2[32 + lane] = r; }
    if (lane < sb) cp_async16(smem_u32(&w2[64 + lane]), supers + lane);
    else { Rec2 r; r.lo = r.hi = 1024ull << kCntShift; w2[64 + lane] = r; }
}
// any warp: exclusive prefix of tile t from the window; false when a record was not complete yet
__device__ __forceinline__ bool window_eval(const Rec *win, int lane, u64 &P) {
    const uint4 tr = *reinterpret_cast<const uint4 *>(&win[lane]);
    const Rec2 *w2 = reinterpret_cast<const Rec2 *>(win);
    const Rec2 br = w2[32 + lane], sr = w2[64 + lane];
    const bool ok = tr.z != 0u && (br.lo >> kCntShift) == 32ull && (br.hi >> kCntShift) == 32ull && (sr.lo >> kCntShift) == 1024ull &&
                    (sr.hi >> kCntShift) == 1024ull;
    u64 x = (((u64)tr.y << 32) | tr.x) + ((((br.hi & kSumMask) + (sr.hi & kSumMask)) << 31) + (br.lo & kSumMask) + (sr.lo & kSumMask));
    if (!__all_sync(0xffffffffu, ok)) return false;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    P = x;
    return true;
}
// the same from global memory (fallback: polls until every record is complete)
__device__ __noinline__ u64 lookback_poll(const FxArgs &a, int c, int t, int lane) {
    const Rec *tiles = a.desc + (int64_t)c * a.T;
    const Rec2 *blocks = a.blocks + (int64_t)c * a.TB, *supers = a.supers + (int64_t)c * a.TS;
    const int b = t >> 5, pos = t & 31, sb = b >> 5, bpos = b & 31;
    for (;;) {
        u64 tv = 0, blo = 32ull << kCntShift, bhi = blo, x = 0;
        uint32_t tf = 1u;
        if (lane < pos) tf = rec_load(tiles + (t - 1 - lane), tv);
        if (lane < bpos) rec2_load(blocks + (sb * 32 + lane), blo, bhi);
        bool ok = tf != 0u && (blo >> kCntShift) == 32ull && (bhi >> kCntShift) == 32ull;
        x = tv + (((bhi & kSumMask) << 31) + (blo & kSumMask));
        for (int s = lane; s < sb; s += 32) {
            u64 lo, hi;
            rec2_load(supers + s, lo, hi);
            ok = ok && (lo >> kCntShift) == 1024ull && (hi >> kCntShift) == 1024ull;
            x += ((hi & kSumMask) << 31) + (lo & kSumMask);
        }
        if (__all_sync(0xffffffffu, ok)) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            return x;
        }
        __nanosleep(100);
    }
}

// one thread of a group: draw the SM's next position, decode its pass-B tile into item slot `slot` and start the tile's copy
__device__ __forceinline__ void fx_take(const FxArgs &a, const CUtensorMap *tmap, FxScratch &sc, unsigned int *ticket, int slot, uint32_t buf0, uint64_t pol,
                                        bool draw) {
    FxItem it; it.c = 0; it.t = 0; it.flags = kItEnd; it.pad = 0;
    if (draw) {
        long long pb;
        if (a.hack & 4) { pb = sc.pfb; sc.pfb += (long long)kFxGroups * gridDim.x; }
        else pb = (long long)atomicAdd(ticket, 1u);
        it.flags = 0;
        if (pb >= a.nT) it.flags = kItEnd;
        else if (pb >= 0) {
            it.c = (int)((unsigned int)pb / (unsigned int)a.T); it.t = (int)((unsigned int)pb - (unsigned int)it.c * (unsigned int)a.T);
            it.flags = kItValid;
            if (it.t < a.tiles_tma) it.flags |= kItTma;
            if (it.t == (int)a.T - 1 && a.has_tail) it.flags |= kItTail;
            if (it.flags & kItTma) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&sc.mbar[slot], (uint32_t)kTile * 4u);
                tma_tile_g2s(buf0 + (uint32_t)slot * kTile * 4u, tmap, it.t * (kTile / 32), it.c, &sc.mbar[slot], pol);
            } else {
                mbar_arrive(&sc.mbar[slot]);
            }
        }
    }
    sc.item[slot] = it;
}

// ------------------------------------------------------------------ pass A (its own warps: they stream, add and publish; nobody waits for them inside the CTA)
__device__ __forceinline__ unsigned int ldg_volatile_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ void pass_a_warp(const FxArgs &a, unsigned int *ticketA, const unsigned int *ticketB, int lane) {
    const uint64_t polA = policy_evict_last();
    for (;;) {
        unsigned int kA = 0;
        if (lane == 0) kA = atomicAdd(ticketA, 1u);
        kA = __shfl_sync(0xffffffffu, kA, 0);
        const long long pos = (long long)kA;
        if (pos >= a.nT) break;
        // stay within Lg tiles of the pass-B draws: what pass A streams has to survive in L2 until pass B comes for it
        while ((long long)ldg_volatile_u32(ticketB) + a.Lg < pos) __nanosleep(2000);
        const int c = (int)((unsigned int)pos / (unsigned int)a.T), t = (int)((unsigned int)pos - (unsigned int)c * (unsigned int)a.T);
        const float *pa = a.X + (int64_t)c * a.ld + (int64_t)t * kTile;
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        if ((int64_t)(t + 1) * kTile <= a.d) {
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                uint4 v[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = ldg_keep_u4(pa + ((h * 16 + q) * 32 + lane) * 4, polA);
#pragma unroll
                for (int q = 0; q < 16; ++q) acc_abs4(v[q], s);
            }
        } else {
            const int64_t i0 = (int64_t)t * kTile;
            for (int i = lane; i < kTile && i0 + i < a.d; i += 32) s[0] += abs_to_double(__float_as_uint(pa[i]));
        }
        const double ws = warp_sum_f64((s[0] + s[1]) + (s[2] + s[3]));      // fixed association
        if (lane == 0) {
            publish_tile_sum(a, c, ws);
            publish_tile_count(a, c);
        }
    }
}

template <int EMIT>
__global__ void __launch_bounds__(kFxBlock, kFxCtas)
quantize_fx_kernel(const __grid_constant__ FxArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char dyn_smem[];      // per group: three tile buffers + FxScratch; then the ticket
    const int g = threadIdx.x / kFxThreads, gtid = threadIdx.x & (kFxThreads - 1), lane = gtid & 31, warp = gtid >> 5;
    unsigned int *ticket = &a.hdr->ticket;             // pass-B draws (pass A: ticket2); zeroed by the host before the launch
    if (kFxAWarps > 0 && g == kFxGroups) {                  // the pass-A warps
        __syncthreads();
        pass_a_warp(a, &a.hdr->ticket2, ticket, lane);
        return;
    }
    unsigned char *gbase = dyn_smem + (size_t)g * kGroupBytes;
    FxScratch &sc = *reinterpret_cast<FxScratch *>(gbase + (size_t)3 * kTile * sizeof(float));
    const uint32_t buf0 = smem_u32(gbase);
    const int ch0 = 2 * gtid, ch1 = 2 * gtid + 1;
    const uint64_t polB = policy_evict_first();
    const uint32_t boff0 = blocked_off_of((uint32_t)ch0), boff1 = blocked_off_of((uint32_t)ch1);
    const bool window_ok = a.TS <= 32;
    const int kBarFree = 1 + kFxGroups + g;
    if (gtid == 0) {
        for (int q = 0; q < 3; ++q) mbar_init(&sc.mbar[q], 1);
        mbar_init(&sc.winbar[0], 32); mbar_init(&sc.winbar[1], 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sc.rc_row[0] = sc.rc_row[1] = -1;
        sc.hit[0] = sc.hit[1] = 0;
        sc.pfb = (long long)blockIdx.x * kFxGroups + g;
    }
    __syncthreads();
    if (gtid == 0) {
        fx_take(a, &tmap, sc, ticket, 0, buf0, polB, true);
        fx_take(a, &tmap, sc, ticket, 1, buf0, polB, !(sc.item[0].flags & kItEnd));
        fx_take(a, &tmap, sc, ticket, 2, buf0, polB, false);
    }
    grp_sync(g);
    // state of the tile whose C-phase is pending (one iteration behind its B-phase); index 0 / 1 = the thread's chunks
    uint32_t sgw0P = 0, sgw1P = 0, flm0P = 0, flm1P = 0, f4a0P = 0, f4b0P = 0, f4a1P = 0, f4b1P = 0, mxhP = 0, fmP = 0;
    u64 exclP = 0, runP = 0, wbaseP = 0;
    uint32_t run0P = 0;              // low word of the first chunk's sum
    bool liveC = false;
    FxItem iC; iC.c = 0; iC.t = 0; iC.flags = 0;
    int sB = 0, sC = 2;
    uint32_t phase = 0;             // bit s: parity of the next completion of mbar[s]
#ifdef DME_TIMERS
    long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif
    for (int it = 0;; ++it) {
        const FxItem iB = sc.item[sB];
        if ((iB.flags & kItEnd) && !liveC) break;
        const bool validB = iB.flags & kItValid;
        const int e = it & 1;
        FXW(20);
        // The look-back window of tile C: its copies take an L2 round trip (well over a microsecond under load), so they start
        // here and land during the B-phase; records that are not complete yet are polled for later.
        if (warp == 1) {
            if (liveC && window_ok) window_prefetch(a, sc.win[e], iC.c, iC.t, lane);
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&sc.winbar[e])) : "memory");
        }
        // ---------------------------------------------------------------- B-phase of tile iB
        uint32_t sgw0 = 0, sgw1 = 0, flm0 = 0, flm1 = 0, f4a0 = 0, f4b0 = 0, f4a1 = 0, f4b1 = 0, mxh = 0;
        u64 run = 0, incl = 0;
        uint32_t run0 = 0;
        bool liveB = false;
        if (validB) {
            if (sc.rc_row[e] != iB.c) {          // group-uniform: the row's constants and binade offsets into shared memory
                FXT(0);
                load_row(a, sc, g, iB.c, e);
                FXT(10);
            }
            const RowConst &rc = sc.rc[e];
            const uint32_t buf = buf0 + (uint32_t)sB * kTile * 4u;
            FXT(0);
            mbar_wait(smem_u32(&sc.mbar[sB]), (phase >> sB) & 1u);
            phase ^= 1u << sB;
            FXT(1);
            liveB = !(rc.flags & kRowExotic);
            FXW(21);
            FXTR(1, iB.c, iB.t);
            if (liveB) {
                fx_chunk_b(a, iB, buf, boff0, ch0, rc, sgw0, flm0, f4a0, f4b0, mxh, run);
                run0 = (uint32_t)run;
                fx_chunk_b(a, iB, buf, boff1, ch1, rc, sgw1, flm1, f4a1, f4b1, mxh, run);
                incl = run;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const u64 up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                if (lane == 31) sc.wtot[e][warp] = incl;
                const uint32_t wmx = __reduce_max_sync(0xffffffffu, mxh);
                if (lane == 0) sc.fmw[e][warp] = wmx;
            }
        }
        FXT(2);
        // The serial jobs of an iteration are spread over the warps (window copies: warp 1, aggregate: warp 2,
        // next draw + copy: warp 3): warp w of every group runs on scheduler w.  The look-back window of tile C is copied as
        // LATE as possible -- the later, the more of the earlier tiles' aggregates are there -- and nobody waits for the copies
        // here: they signal winbar[e].
        FXW(22);
        FXW(23);
        grp_sync(g);
        FXW(24);
        FXT(4);
        FXT(5);
        // ---------------------------------------------------------------- every warp: warp base of tile B, its aggregate
        u64 wbase = 0;
        uint32_t fm = 0;
        if (liveB) {
            u64 Aq = 0;
#pragma unroll
            for (int w = 0; w < kFxWarps; ++w) {
                const u64 v = sc.wtot[e][w];
                if (w < warp) wbase += v;
                Aq += v;
                fm = max(fm, sc.fmw[e][w]);
            }
            FXTR(2, iB.c, iB.t);
            if (gtid == 64) {
                rec_store(a.desc + (int64_t)iB.c * a.T + iB.t, Aq, 1u);
                const u64 lo = (Aq & 0x7fffffffull) + (1ull << kCntShift), hi = (Aq >> 31) + (1ull << kCntShift);
                Rec2 *br = a.blocks + (int64_t)iB.c * a.TB + (iB.t >> 5), *sr = a.supers + (int64_t)iB.c * a.TS + (iB.t >> 10);
                red_add_u64(&br->lo, lo); red_add_u64(&br->hi, hi);
                red_add_u64(&sr->lo, lo); red_add_u64(&sr->hi, hi);
            }
        }
        if (gtid == 0) sc.hit[e ^ 1] = 0;
        // ---------------------------------------------------------------- C-phase of tile iC: AS:635-637
        uint32_t kw0 = 0, kw1 = 0;
        const bool need_hit = liveC && EMIT == 1 && fmP < 0x7fffffffu && width_of_u(fmP) != width_of_u(fmP + 1u);      // group-uniform
        if (liveC) {
            const RowConst &rc = sc.rc[e ^ 1];
            const uint32_t buf = buf0 + (uint32_t)sC * kTile * 4u;
            u64 P = 0;
            FXT(6);
            FXTR(3, iC.c, iC.t);
            mbar_wait(smem_u32(&sc.winbar[e]), (uint32_t)((it >> 1) & 1));          // the window copies have landed
            FXT(7);
            if (a.hack & 1) P = (u64)iC.t * (877ull << 32);
            if (!(a.hack & 1) && iC.t > 0 && !(window_ok && window_eval(sc.win[e], lane, P))) {
#ifdef DME_TIMERS
                if (gtid == 0) tacc[11] += 1;
#endif
                if (warp == 0) {
                    const u64 pp = lookback_poll(a, iC.c, iC.t, lane);
                    if (lane == 0) sc.pfb = (long long)pp;
                }
                grp_sync(g);
                P = (u64)sc.pfb;
            }
            FXT(8);
            FXW(25);
            FXTR(4, iC.c, iC.t);
            const u64 E = P + wbaseP + exclP, En = E + runP;
            const int eb = 31 - __clzll((long long)(E | 1ull));                     // E in [2^eb, 2^(eb+1))
            // every prefix of the thread, and their fp32 roundings, stay inside [2^eb + 1, 2^(eb+1))
            const bool fast = eb >= 2 && eb <= 22 && ((E - (3ull << 31)) >> (eb + 32)) == 1ull && ((En + (1ull << 32)) >> (eb + 32)) == 1ull;
            uint32_t rb0, rb1;
            if (fast) {
                uint32_t acc0 = (uint32_t)(E - sc.tab[e ^ 1][eb]), acc1 = acc0 + run0P;      // the second chunk starts from the first chunk's sum: two independent carry chains
                rb0 = walk_carry(buf + boff0, acc0);
                rb1 = walk_carry(buf + boff1, acc1);
            } else {
                u64 c = E;
                long long tp = t_literal(c, rc.Xi);
                rb0 = walk_literal(buf + boff0, c, tp, rc.Xi);
                rb1 = walk_literal(buf + boff1, c, tp, rc.Xi);
            }
            kw0 = spread_rev16(rb0); kw1 = spread_rev16(rb1);
            if (need_hit && mxhP == fmP) {
                // the largest floor sits right below a width boundary: did one of those coordinates also receive a unit?
                bool hit;
                if (fmP == 1u) hit = ((flm0P & kw0) | (flm1P & kw1)) != 0u;
                else {
                    hit = false;
#pragma unroll 1
                    for (int h = 0; h < 2; ++h) {
                        float x[kEpt];
                        fx_load_x_global(a, iC.c, iC.t, x, h ? ch1 : ch0);
                        const uint32_t kw = h ? kw1 : kw0;
#pragma unroll
                        for (int j = 0; j < kEpt; ++j) {
                            float q;
                            const u64 F = fx_chain(x[j], rc, q);
                            hit |= ((uint32_t)(F >> 32) == fmP) && ((kw >> (2 * j)) & 1u);
                        }
                    }
                }
                if (hit) atomicOr(&sc.hit[e], 1u);
            }
        }
        // Buffer sC is free once every warp has read its low words: only the warp that issues the next copy into it has
        // to wait for that, the others just signal.  When the tile's width depends on sc.hit, everybody waits.
        FXT(9);
        FXW(26);
        if (warp == 3 || need_hit) bar_sync(kBarFree, kFxThreads);
        else bar_arrive(kBarFree, kFxThreads);
        if (gtid == 96) fx_take(a, &tmap, sc, ticket, sC, buf0, polB, !(iB.flags & kItEnd));
        // ---------------------------------------------------------------- emit tile iC
        if (liveC) {
            const RowConst &rc = sc.rc[e ^ 1];
            if (EMIT == 0) {
                bool ovf = false;
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    float x[kEpt];
                    const int ch = h ? ch1 : ch0;
                    const uint32_t kw = h ? kw1 : kw0, sgwP = h ? sgw1P : sgw0P;
                    fx_load_x_global(a, iC.c, iC.t, x, ch);
                    const int64_t i0 = (int64_t)iC.t * kTile + (int64_t)ch * kEpt;
#pragma unroll
                    for (int j = 0; j < kEpt; ++j) {
                        float q;
                        const u64 F = fx_chain(x[j], rc, q);
                        const u64 k = (F >> 32) + (u64)((kw >> (2 * j)) & 1u);
                        const int64_t i = i0 + j;
                        if (i >= a.d) continue;
                        const uint32_t sbit = (sgwP >> (2 * j + 1)) & 1u;
                        if (a.deq_out) {
                            // sign(v) of AS:640 is zero exactly when v = x / D is
                            const float sgf = (q == 0.0f) ? 0.0f : (sbit ? -1.0f : 1.0f);
                            a.deq_out[(int64_t)iC.c * a.ld_out + i] = __fdiv_rn(__fmul_rn(__fmul_rn(rc.L1f, sgf), __ull2float_rn(k)), rc.mf);
                        }
                        if (a.k_out) {
                            if (k >= 2147483648ull) { ovf = true; a.k_out[(int64_t)iC.c * a.ld_out + i] = 0x7fffffff; }
                            else a.k_out[(int64_t)iC.c * a.ld_out + i] = (int32_t)k;
                        }
                        if (a.sgn_out) a.sgn_out[(int64_t)iC.c * a.ld_out + i] = (uint8_t)sbit;
                    }
                }
                if (ovf) atomicOr(&a.hdr->status, 1u);
            } else {
                const int W = sc.hit[e] ? width_of_u(fmP + 1u) : width_of_u(fmP);
                if (fmP >= 0x7fffffffu && gtid == 0) atomicOr(&a.hdr->status, 1u);
                const int64_t slot_id = (int64_t)iC.c * a.T + iC.t;
                u64 off16;
                if (W <= a.pack.W0) {
                    off16 = primary_off16(a.pack, iC.c, iC.t);
                    if (gtid == 0) a.pack.dir[slot_id] = (off16 << 8) | (u64)W;
                } else {
                    if (gtid == 0) {
                        const u64 units = 32ull * W;
                        u64 o = a.pack.arena_base16 + atomicAdd(&a.hdr->arena_top, units);
                        if ((long long)((o + units) * 16ull) > a.pack.codes_bytes) { atomicOr(&a.hdr->status, 2u); o = ~0ull; }
                        sc.off16 = o;
                        a.pack.dir[slot_id] = (o == ~0ull) ? 0ull : ((o << 8) | (u64)W);
                    }
                    grp_sync(g);
                    off16 = sc.off16;
                }
                if (off16 != ~0ull) {
                    uint32_t *tw = a.pack.codes + off16 * 4ull;
                    if (W == 2) {
                        // fields [sign | magnitude bit]: k = floor + r <= 1; the thread's two words are adjacent
                        uint2 w2;
                        w2.x = kw0 | flm0P | (sgw0P & 0xaaaaaaaau);
                        w2.y = kw1 | flm1P | (sgw1P & 0xaaaaaaaau);
                        *reinterpret_cast<uint2 *>(tw + ch0) = w2;
                    } else if (W == 4) {
                        const uint32_t c0 = (kw0 & 0x55555555u) | (sgw0P & 0xaaaaaaaau), c1 = (kw1 & 0x55555555u) | (sgw1P & 0xaaaaaaaau);
                        uint2 lo, hi;
                        lo.x = spread_pairs_to_nibbles(c0 & 0xffffu) + f4a0P; lo.y = spread_pairs_to_nibbles(c1 & 0xffffu) + f4a1P;
                        hi.x = spread_pairs_to_nibbles(c0 >> 16) + f4b0P;     hi.y = spread_pairs_to_nibbles(c1 >> 16) + f4b1P;
                        *reinterpret_cast<uint2 *>(tw + ch0) = lo;
                        *reinterpret_cast<uint2 *>(tw + kThreads + ch0) = hi;
                    } else {
#pragma unroll 1
                        for (int h = 0; h < 2; ++h) {
                            float x[kEpt];
                            fx_load_x_global(a, iC.c, iC.t, x, h ? ch1 : ch0);
                            uint32_t k[kEpt], sg[kEpt];
                            const uint32_t kw = h ? kw1 : kw0, sgwP = h ? sgw1P : sgw0P;
#pragma unroll
                            for (int j = 0; j < kEpt; ++j) {
                                float q;
                                const u64 F = fx_chain(x[j], rc, q);
                                k[j] = (uint32_t)min((F >> 32) + (u64)((kw >> (2 * j)) & 1u), 0x7fffffffull);      // overflow already reported
                                sg[j] = (sgwP >> (2 * j + 1)) & 1u;
                            }
                            const int ch = h ? ch1 : ch0;
                            switch (W) {
                                case 8: pack_store<8>(k, sg, tw, ch); break;
                                case 16: pack_store<16>(k, sg, tw, ch); break;
                                default: pack_store<32>(k, sg, tw, ch); break;
                            }
                        }
                    }
                }
            }
        }
        sgw0P = sgw0; sgw1P = sgw1; flm0P = flm0; flm1P = flm1; f4a0P = f4a0; f4b0P = f4b0; f4a1P = f4a1; f4b1P = f4b1;
        mxhP = mxh; exclP = incl - run; runP = run; run0P = run0; wbaseP = wbase; fmP = fm; liveC = liveB; iC = iB;
        FXW(27);
        if (liveC) FXTR(5, iC.c, iC.t);
        const int nB = sB == 2 ? 0 : sB + 1;
        sC = sB; sB = nB;
        FXT(10);
    }
#ifdef DME_TIMERS
    if (gtid == 0)
        for (int q = 0; q < 12; ++q) atomicAdd(reinterpret_cast<unsigned long long *>(&a.hdr->pad[5]) + q, (unsigned long long)tacc[q]);
#endif
}

// ------------------------------------------------------------------ host side
struct FxDevice { bool ready = false; int sms = 0; };
static FxDevice g_fx_dev[64];
static std::mutex g_fx_mu;
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_fx_encode = nullptr;
static int g_fx_lead = -1;       // DME_FX_LEAD (development): minimum distance, in tiles, between the end of a row's pass A and the start of its pass B
constexpr size_t kFxDynSmem = (size_t)kFxGroups * kGroupBytes + 16;

static int fx_device(FxDevice **out) {
    int dev = 0;
    DME_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return DME_ECUDA; }
    std::lock_guard<std::mutex> lock(g_fx_mu);
    FxDevice &D = g_fx_dev[dev];
    if (!D.ready) {
        // function attributes are per device: set once for each device the library is used on
        DME_CUDA(cudaFuncSetAttribute(quantize_fx_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFxDynSmem));
        DME_CUDA(cudaFuncSetAttribute(quantize_fx_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFxDynSmem));
        int occ = 0;
        DME_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, quantize_fx_kernel<1>, kFxBlock, kFxDynSmem));
        if (occ < 1) { set_error("quantize_fx_kernel does not fit on an SM of device %d", dev); return DME_ECUDA; }
        DME_CUDA(cudaDeviceGetAttribute(&D.sms, cudaDevAttrMultiProcessorCount, dev));
        if (g_fx_encode == nullptr) {
            cudaDriverEntryPointQueryResult qres;
            void *fn = nullptr;
            DME_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
            if (fn == nullptr || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return DME_ECUDA; }
            g_fx_encode = (EncodeTiledFn)fn;
            if (const char *e = getenv("DME_FX_LEAD")) g_fx_lead = atoi(e);
        }
        D.ready = true;
    }
    *out = &D;
    return DME_OK;
}

int launch_literal_rows(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, const WsLayout &L, void *ws,
                        const float *x_inject, const float *l1_inject, uint64_t seed, uint64_t client0,
                        int32_t *k_out, uint8_t *sgn_out, float *deq_out, int64_t ld_out,
                        uint32_t *codes, int64_t codes_bytes, uint64_t *dir, float *l1_out, cudaStream_t st, bool packed, bool all_rows);   // quantize_literal.cu

int launch_quantize_fx(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, const WsLayout &L, void *ws,
                       const float *x_inject, const float *l1_inject, uint64_t seed, uint64_t client0,
                       int32_t *k_out, uint8_t *sgn_out, float *deq_out, int64_t ld_out,
                       uint32_t *codes, int64_t codes_bytes, uint64_t *dir, float *l1_out, cudaStream_t st, bool packed) {
    FxDevice *D = nullptr;
    int rc = fx_device(&D);
    if (rc) return rc;
    char *base = (char *)ws;
    FxArgs a;
    a.X = X; a.d = d; a.ld = ld; a.T = L.T; a.n = n; a.m = m;
    a.rows32 = d / 32;
    a.consts = (RowConst *)(base + L.off_consts);
    a.TB = (L.T + 31) / 32;                                        // blocks of 32 tiles per row
    a.TS = (a.TB + 31) / 32;                                       // super-blocks of 32 blocks per row
    a.desc = (Rec *)(base + L.off_desc);                           // tile records, then block records, then super-block records
    a.blocks = (Rec2 *)(base + L.off_desc + 16 * n * L.T);
    a.supers = (Rec2 *)(base + L.off_desc + 16 * n * (L.T + a.TB));
    a.hdr = (WsHeader *)base;
    a.bins = (u64 *)(base + L.off_bins);
    a.pub_count = (uint32_t *)(base + L.off_done);
    a.row_ready = (uint32_t *)(base + L.off_ready);
    a.exotic_rows = (int32_t *)(base + L.off_exotic);
    a.x_inject = x_inject; a.l1_inject = l1_inject; a.seed = seed; a.client0 = client0; a.l1_out = l1_out;
    a.k_out = k_out; a.sgn_out = sgn_out; a.deq_out = deq_out; a.ld_out = ld_out;
    a.pack.codes = codes; a.pack.codes_bytes = codes_bytes; a.pack.dir = dir; a.pack.hdr = a.hdr; a.pack.n = n; a.pack.T = L.T;
    a.pack.W0 = expected_width(m > 0 ? m : 1, d);
    a.nT = n * L.T;
    a.pack.arena_base16 = (unsigned long long)a.nT * 32ull * (unsigned long long)a.pack.W0;
    if (packed && (long long)(a.pack.arena_base16 * 16ull) > codes_bytes) {
        set_error("code arena too small for the primary slots: %lld < %llu bytes", (long long)codes_bytes, a.pack.arena_base16 * 16ull);
        return DME_EWORKSPACE;
    }
    int64_t S = (int64_t)D->sms * kFxCtas;
    if (const char *e = getenv("DME_FX_G")) S = atoll(e);        // development
    if (S > a.nT) S = a.nT;
    // pass B of a tile runs Lg = T + lead positions after its pass A, on the same SM (Lg a multiple of S); `lead` covers
    // the time it takes until a row's norm is known after its last tile has been streamed
    const int64_t lead = g_fx_lead >= 0 ? g_fx_lead : 512;
    a.S = (int)S;
    a.Lg = L.T + lead;
    a.tiles_tma = (int)((a.rows32 + kTile / 32 - 1) / (kTile / 32));
    a.has_tail = (d & 31) ? 1 : 0;
    a.one = 1u;
    a.hack = getenv("DME_FX_HACK") ? atoi(getenv("DME_FX_HACK")) : 0;
    // 3-D view of the client rows: {32 floats, full 128-byte rows of a client, clients}; the last d % 32 coordinates
    // of every row are read directly by the kernel
    CUtensorMap tmap;
    {
        const cuuint64_t dims[3] = {32, (cuuint64_t)(a.rows32 > 0 ? a.rows32 : 1), (cuuint64_t)n};
        const cuuint64_t strides[2] = {128, (cuuint64_t)ld * 4};
        const cuuint32_t box[3] = {32, kTile / 32, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = g_fx_encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)X, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for n=%lld d=%lld ld=%lld", (int)r, (long long)n, (long long)d, (long long)ld); return DME_ECUDA; }
    }
    void *args[] = {&a, &tmap};
    const void *fn = packed ? (const void *)quantize_fx_kernel<1> : (const void *)quantize_fx_kernel<0>;
    DME_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)S), dim3(kFxBlock), args, kFxDynSmem, st));
    count_launch();
    // rows outside the proven range of the fast chain (none on ordinary inputs: the kernel exits at once)
    return launch_literal_rows(X, n, d, ld, m, L, ws, x_inject, l1_inject, seed, client0, k_out, sgn_out, deq_out, ld_out, codes, codes_bytes, dir,
                               l1_out, st, packed, false);
}

}  // namespace dme

#ifdef DME_TIMERS
extern "C" __attribute__((visibility("default"))) int dme_debug_fx_trace(unsigned long long *out, int cap) {
    unsigned int n = 0;
    cudaMemcpyFromSymbol(&n, dme::g_fx_trace_n, sizeof(n));
    if ((int)n > cap) n = cap;
    if (n > (1u << 15)) n = 1u << 15;
    cudaMemcpyFromSymbol(out, dme::g_fx_trace, (size_t)n * 16);
    unsigned int z = 0;
    cudaMemcpyToSymbol(dme::g_fx_trace_n, &z, sizeof(z));
    return (int)n;
}
#endif
