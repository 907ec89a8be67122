// quantize_warp.cu -- the quantize kernel of the unbiased type quantizer (AS:609-641): scale to m, floor + systematic-sampling
// allocation of the fractional mass, sign/magnitude packing (or the dequantised output of the drop-in API), for all clients of
// one GPU in ONE launch.  The per-client L1 norms and row constants come from l1_kernel (type_quantize.cu).
//
// Unit of work = one WARP and one code tile (1024 coordinates = one directory entry of the packed code).  Warps are independent:
// no CTA barrier, no shared state between warps; a CTA is only a container for four of them.  Warp g works on tiles g, g + G,
// g + 2G, ... of the client-major tile order (G = resident warps; dynamic tickets let the look-back form convoys),
// stages its tile with one TMA tensor copy (3-D map {32 floats, 128-byte rows, client}, SWIZZLE_128B, 4 KB box: lane l owns
// row l = coordinates [32 l, 32 l + 32), its eight 16-byte pieces are conflict-free) into a private ring of three buffers, and
// runs two phases that are one tile apart:
//     B-phase (tile i)  : x / D (Markstein), * m, floor (magic add toward zero), fraction -- packed f32x2; the exact fractions are
//                         parked in place of x and enter the lane's fp64 sum; lane sums -> 2^-43 fixed point -> warp aggregate
//                         (integer REDUX) -> published at once: 8-byte tile record, one 64-bit atomic on the record of its block
//                         of 32 tiles (value + count in one word); the tile that completes a block forwards the block total to
//                         the super-block record;
//     C-phase (tile i-1): look-back window (<= 31 tile + 31 block + S super-block records, one load per lane) -> exclusive
//                         prefix P -> every lane's start phase phi = frac(E - Xp) as a 32-bit integer -> the allocation of
//                         AS:635-637 is the CARRY chain of phi + sum of trunc(fraction * 2^32): one add-with-carry-out and one
//                         add-with-carry-in (shift the carry into a mask) per coordinate; emit.
// Why the carry chain is AS:636.  floor(RN32(RN32(c) - X)) = floor(c - Xp) or ceil(c - Xp) - 1 while the fp32 prefix stays in
// one binade [2^e + 1, 2^(e+1)), Xp = 2^(e-24) (2a - 1), a = ceil(X / 2^(e-23) - 1/2) (tests/test_closed_form_floor.py): both
// count the integers crossed by c - Xp and differ only when c - Xp IS an integer.  The integer chain carries truncation errors
// (< 2^-26 per lane); a lane whose running phase ever comes within the band 2^-24 of an integer, whose tile straddles a binade,
// or whose row is outside the proven operand range evaluates AS:636 literally in fp64 from the parked fractions.
// Integer addition is associative: prefixes do not depend on timing, results are run-to-run deterministic.
// Every wait is on a smaller tile whose B-phase never waits: no deadlock while all CTAs are resident.
#include <cuda.h>

#include <cmath>
#include <mutex>

#include "type_quantize.cuh"

namespace dme {

constexpr int kCpl = 32;                          // coordinates per lane
constexpr int kLag = 1;                           // the C-phase of a tile runs kLag iterations after its B-phase: 1 (default) or 2 (measured: no gain)
constexpr int kRing = 2 + kLag;                   // tile buffers per warp: parked (C; kLag of them), current (B), in flight
constexpr int kQWarps = 4;                        // warps per CTA
constexpr int kQThreads = kQWarps * 32;
constexpr int kTileBytes = kCodeTile * 4;
constexpr int kTicketBatch = 1;                   // consecutive tickets per atomic
constexpr uint32_t kBand = 256u;                  // ambiguity band of the carry chain, 2^-32 units (2^-24)
constexpr unsigned long long kBlkCnt1 = 1ull << 58;       // block record: tiles published so far in bits 58..63,
constexpr unsigned long long kBlkVal = kBlkCnt1 - 1ull;   // sum of their aggregates (2^-qshift units, qshift <= 43) below
constexpr int kSupShift = 44;                             // super-block record halves: blocks forwarded so far in bits 44..,
constexpr unsigned long long kSupVal = (1ull << kSupShift) - 1ull;   // low / high 31-bit halves of the block totals summed below
struct __align__(16) SupRec { unsigned long long lo, hi; };

// ---- memory primitives
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
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
// one 4 KB box {32 floats, 32 rows, 1 client} at (0, row0, client) of the 3-D tensor map
__device__ __forceinline__ void tma_tile_g2s(uint32_t dst, const CUtensorMap *map, int row0, int client, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(dst), "l"(map), "r"(0), "r"(row0), "r"(client), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_rec(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ld_rec2(const SupRec *p, unsigned long long &lo, unsigned long long &hi) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_rec(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long atom_add_u64(unsigned long long *p, unsigned long long v) {
    unsigned long long old;
    asm volatile("atom.relaxed.gpu.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
    return old;
}
__device__ __forceinline__ void red_add_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- packed f32x2 arithmetic (sm_100: FMUL2 / FFMA2, one issue slot for two coordinates)
typedef unsigned long long f2;
__device__ __forceinline__ f2 f2_pack(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(f2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

struct WarpArgs {
    const float *X; int64_t d, ld, n;
    uint32_t T4;                               // code tiles per client row
    int64_t rows32;                            // full 128-byte rows per client vector (the part the tensor map covers)
    int tail;                                  // d % 32: coordinates of the last, partial 128-byte row (read directly)
    uint32_t total;                            // n * T4
    uint32_t recipT4;                          // floor(2^32 / T4) (T4 = 1: 2^32 - 1): ticket -> (client, tile)
    int qshift; double q_dn;                   // look-back records are 2^-qshift fixed point (the same for every row: make_row_const)
    const RowConst *consts;
    unsigned long long *trec;                  // [n][T4]  (aggregate << 1) | 1
    unsigned long long *brec;                  // [n][TB]  count << 58 | sum of the block's aggregates
    SupRec *srec;                              // [n][TS]  two halves, each count << 44 | sum of 31-bit halves of the block totals
    uint32_t TB, TS;
    WsHeader *hdr;
    int32_t *k_out; uint8_t *sgn_out; float *deq_out; int64_t ld_out;     // array outputs (EMIT == 0)
    PackTarget pack;                                                    // packed output (EMIT == 1)
};

#ifdef DME_TIMERS
__device__ unsigned long long *g_dbg = nullptr;       // per ticket: draw, B start, publish, C start, C end (globaltimer ns), smid << 8 | warp slot
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t smid() { uint32_t r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t warpid_hw() { uint32_t r; asm volatile("mov.u32 %0, %%warpid;" : "=r"(r)); return r; }
#define DBG_MARK(tk, f) do { if (g_dbg && lane == 0) g_dbg[(size_t)(tk) * 12 + (f)] = gtime(); } while (0)
#else
#define DBG_MARK(tk, f) do { } while (0)
#endif
struct Item { int c, w; uint32_t tk; bool valid; };       // client, tile of its row, ticket = c * T4 + w

// AS:625-631 for one coordinate, literal: IEEE division + floorf.
__device__ __forceinline__ void chain_exact(float x, float D, float mf, float &flf, float &fr) {
    const float v = __fdiv_rn(x, D);
    const float mp = __fmul_rn(mf, fabsf(v));
    flf = floorf(mp);
    fr = __fsub_rn(mp, flf);
}
// double of a non-negative fp32 fraction by integer moves (no conversion unit).  0 maps to 2^-127, denormals to values below
// 2^-126: both vanish in every fixed-point conversion they enter.
__device__ __forceinline__ double frac_to_double(float f) {
    const uint32_t b = __float_as_uint(f);
    return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
}

__device__ __forceinline__ uint32_t spread16(uint32_t v) {       // bit j -> bit 2j
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}
__device__ __forceinline__ int width_of(float kmax) {
    return kmax < 2.0f ? 2 : kmax < 8.0f ? 4 : kmax < 128.0f ? 8 : kmax < 32768.0f ? 16 : 32;
}
// sum over the warp of a 63-bit value: three 21-bit limbs, one integer REDUX each
__device__ __forceinline__ unsigned long long warp_sum_u63(unsigned long long x) {
    const uint32_t l0 = (uint32_t)x & 0x1fffffu, l1 = (uint32_t)(x >> 21) & 0x1fffffu, l2 = (uint32_t)(x >> 42);
    const uint32_t r0 = __reduce_add_sync(0xffffffffu, l0), r1 = __reduce_add_sync(0xffffffffu, l1), r2 = __reduce_add_sync(0xffffffffu, l2);
    return (unsigned long long)r0 + ((unsigned long long)r1 << 21) + ((unsigned long long)r2 << 42);
}
__device__ __forceinline__ uint32_t warp_excl_scan_u32(uint32_t v, int lane) {
    uint32_t s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += up;
    }
    return s - v;
}

// the lane's 32 coordinates of tile w of client c straight from global memory (cold paths)
__device__ __noinline__ void load_lane_global(const WarpArgs &a, int c, int w, int lane, float (&x)[kCpl]) {
    const int64_t i0 = (int64_t)w * kCodeTile + (int64_t)lane * kCpl;
    const float *row = a.X + (int64_t)c * a.ld;
    if (i0 + kCpl <= a.d) {
#pragma unroll
        for (int q = 0; q < kCpl / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4 *>(row + i0 + 4 * q);
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kCpl; ++j) x[j] = (i0 + j < a.d) ? row[i0 + j] : 0.0f;
    }
}

// State of a tile between its B-phase and its C-phase (per lane unless noted).
struct Carry {
    uint32_t sgw0, sgw1;      // signs of x, interleaved: bit 2j+1 of word h = sign of coordinate 16h + j
    uint32_t flm;             // bit j: floor_j != 0
    uint32_t fl4[4];          // min(floor, 15) of the lane's coordinates as nibbles, eight per word (0 unless flm != 0)
    float mxl;                // largest floor of the lane
    uint32_t L0, L1;          // the lane sum of the fractions in 2^-43 units, 24-bit limbs
    uint32_t ex32;            // exclusive scan over the lanes of the lane sums in 2^-32 units (truncated, mod 2^32)
    float fmf;                // warp-uniform: largest floor of the tile
    unsigned long long Aq;    // warp-uniform: the published aggregate (2^-qshift units)
    uint32_t AI;              // warp-uniform: integer part of the tile's sum of fractions
};

// ---------------------------------------------------------------------------------------------------------------- B-phase
// AS:625-631 for the lane's 32 coordinates (swizzled row at `rowb`): fractions parked as fixed-point words in place of x.
// `exact`: IEEE division + floorf (rows / lanes outside the proven range of the fast chain); same values.
template <bool EXACT>
__device__ __forceinline__ void b_group(const float (&x)[8], float D, float rcpD, float mf, float (&fl)[8], float (&fr)[8]) {
    if (EXACT) {
#pragma unroll
        for (int j = 0; j < 8; ++j) chain_exact(x[j], D, mf, fl[j], fr[j]);
    } else {
        const f2 R2 = f2_pack(rcpD, rcpD), ND = f2_pack(-D, -D), M2 = f2_pack(mf, mf);
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            const f2 xx = f2_pack(x[j], x[j + 1]);
            const f2 q0 = f2_mul(xx, R2);
            const f2 rem = f2_fma(q0, ND, xx);
            const f2 pq = f2_fma(rem, R2, q0);             // x / D, correctly rounded (Markstein), sign kept
            const f2 mp = f2_mul(M2, pq);
            float m0, m1;
            f2_unpack(mp, m0, m1);
            const float t0 = __fadd_rz(fabsf(m0), 8388608.0f), t1 = __fadd_rz(fabsf(m1), 8388608.0f);
            fl[j] = __fsub_rn(t0, 8388608.0f); fl[j + 1] = __fsub_rn(t1, 8388608.0f);     // floor(m |v|), exact below 2^23
            fr[j] = __fsub_rn(fabsf(m0), fl[j]); fr[j + 1] = __fsub_rn(fabsf(m1), fl[j + 1]);
        }
    }
}
template <bool EXACT>
__device__ __forceinline__ double b_phase_body(uint32_t rowb, uint32_t swz, float D, float rcpD, float mf, Carry &cy) {
    uint32_t s0 = 0, s1 = 0, flm = 0;
    float mxl = 0.0f;
    double ra = 0.0, rb = 0.0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint32_t a0 = rowb + (((uint32_t)(2 * g) << 4) ^ swz), a1 = rowb + (((uint32_t)(2 * g + 1) << 4) ^ swz);
        const uint4 va = lds128u(a0), vb = lds128u(a1);
        const uint32_t xb[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
        float x[8], fl[8], fr[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            x[j] = __uint_as_float(xb[j]);
            if (g < 2) s0 = __funnelshift_l(xb[j], s0, 2);          // top two bits of x; field order reversed, fixed below
            else s1 = __funnelshift_l(xb[j], s1, 2);
        }
        b_group<EXACT>(x, D, rcpD, mf, fl, fr);
        float gm = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double f64 = frac_to_double(fr[j]);
            if (j & 1) rb += f64; else ra += f64;
            gm = fmaxf(gm, fl[j]);
        }
        uint32_t nib = 0;
        if (gm != 0.0f) {                                            // rare at R = 1
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                nib |= (uint32_t)fminf(fl[j], 15.0f) << (4 * j);
                flm |= ((fl[j] != 0.0f) ? 1u : 0u) << (8 * g + j);
            }
            mxl = fmaxf(mxl, gm);
        }
        cy.fl4[g] = nib;
        sts128u(a0, make_uint4(__float_as_uint(fr[0]), __float_as_uint(fr[1]), __float_as_uint(fr[2]), __float_as_uint(fr[3])));
        sts128u(a1, make_uint4(__float_as_uint(fr[4]), __float_as_uint(fr[5]), __float_as_uint(fr[6]), __float_as_uint(fr[7])));
    }
    cy.sgw0 = (__brev(s0) & 0x55555555u) << 1;
    cy.sgw1 = (__brev(s1) & 0x55555555u) << 1;
    cy.flm = flm;
    cy.mxl = mxl;
    return ra + rb;
}
// cold: results go through the lane's 16-byte scratch slot so that the caller's state stays in registers
__device__ __noinline__ double b_phase_exact(uint32_t rowb, uint32_t swz, float D, float rcpD, float mf, uint32_t scratch) {
    Carry cy;
    const double run = b_phase_body<true>(rowb, swz, D, rcpD, mf, cy);
    sts128u(scratch, make_uint4(cy.sgw0, cy.sgw1, cy.flm, __float_as_uint(cy.mxl)));
    sts128u(scratch + 16u, make_uint4(cy.fl4[0], cy.fl4[1], cy.fl4[2], cy.fl4[3]));
    return run;
}

// ---------------------------------------------------------------------------------------------------------------- C-phase
// The parked fractions as 32-bit fixed point, truncated: p = trunc(frac * 2^32) (done while the look-back loads are in flight).
__device__ __forceinline__ void fx8(const uint4 &a, const uint4 &b, uint32_t (&p)[8]) {
    const f2 K = f2_pack(4294967296.0f, 4294967296.0f);
    float s0, s1, s2, s3, s4, s5, s6, s7;
    f2_unpack(f2_mul(f2_pack(__uint_as_float(a.x), __uint_as_float(a.y)), K), s0, s1);
    f2_unpack(f2_mul(f2_pack(__uint_as_float(a.z), __uint_as_float(a.w)), K), s2, s3);
    f2_unpack(f2_mul(f2_pack(__uint_as_float(b.x), __uint_as_float(b.y)), K), s4, s5);
    f2_unpack(f2_mul(f2_pack(__uint_as_float(b.z), __uint_as_float(b.w)), K), s6, s7);
    p[0] = __float2uint_rz(s0); p[1] = __float2uint_rz(s1); p[2] = __float2uint_rz(s2); p[3] = __float2uint_rz(s3);
    p[4] = __float2uint_rz(s4); p[5] = __float2uint_rz(s5); p[6] = __float2uint_rz(s6); p[7] = __float2uint_rz(s7);
}
// Carry chain over eight of them: acc += p (carry out), bits = 2 bits + carry; running minimum of acc for the band test.
__device__ __forceinline__ void carry8(uint32_t &acc, uint32_t &bits, uint32_t &mn, const uint32_t (&p)[8]) {
    uint32_t t0, t1, t2, t3, t4, t5, t6;
    asm("add.cc.u32 %0, %8, %9;\n\taddc.u32 %7, %7, %7;\n\t"
        "add.cc.u32 %1, %0, %10;\n\taddc.u32 %7, %7, %7;\n\t"
        "add.cc.u32 %2, %1, %11;\n\taddc.u32 %7, %7, %7;\n\t"
        "add.cc.u32 %3, %2, %12;\n\taddc.u32 %7, %7, %7;\n\t"
        "add.cc.u32 %4, %3, %13;\n\taddc.u32 %7, %7, %7;\n\t"
        "add.cc.u32 %5, %4, %14;\n\taddc.u32 %7, %7, %7;\n\t"
        "add.cc.u32 %6, %5, %15;\n\taddc.u32 %7, %7, %7;\n\t"
        "add.cc.u32 %8, %6, %16;\n\taddc.u32 %7, %7, %7;"
        : "=&r"(t0), "=&r"(t1), "=&r"(t2), "=&r"(t3), "=&r"(t4), "=&r"(t5), "=&r"(t6), "+r"(bits), "+r"(acc)
        : "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]));
    mn = min(mn, min(t0, t1));
    mn = min(mn, min(t2, t3));
    mn = min(mn, min(t4, t5));
    mn = min(mn, min(t6, acc));
}
// Xp of binade e (2 <= e <= 22) in 2^-32 units (mod 2^32): Xp = 2^(e-24) (2a - 1), a = ceil(X / 2^(e-23) - 1/2); any fp32 X in [0, 1)
__device__ __forceinline__ uint32_t xp32_of(float X, int e) {
    const double av = ceil(fma((double)X, __hiloint2double((1023 + 23 - e) << 20, 0), -0.5));       // exact: X has 24 bits
    return (uint32_t)(2 * (int)av - 1) << (e + 8);
}
// AS:636 literally, for one prefix value
__device__ __forceinline__ int floor_ref(double c, float X) { return __float2int_rd(__fsub_rn(__double2float_rn(c), X)); }

// Slow lanes: AS:635-637 as written, from the parked (exact) fractions and the exact fp64 prefixes E (before the lane's first
// coordinate) and En (at its last one).
__device__ __noinline__ uint32_t lane_literal(uint32_t rowb, uint32_t swz, float X, double E, double En) {
    uint32_t rm = 0;
    double cp = E;
    int tp = floor_ref(cp, X);
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
        const uint4 v = lds128u(rowb + (((uint32_t)q << 4) ^ swz));
        const float fr[4] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            cp = (4 * q + j < kCpl - 1) ? cp + (double)fr[j] : En;    // AS:635
            const int t = floor_ref(cp, X);                           // AS:636
            rm |= ((t - tp == 1) ? 1u : 0u) << (4 * q + j);           // AS:637
            tp = t;
        }
    }
    return rm;
}
// does a coordinate whose floor equals fm receive a unit?  (fm = 7, 127, 32767: the tile's width depends on it)
__device__ __noinline__ bool lane_hit_wide(const WarpArgs &a, int c, int w, int lane, float D, float mf, float fm, uint32_t rm) {
    float x[kCpl];
    load_lane_global(a, c, w, lane, x);
    bool hit = false;
#pragma unroll 4
    for (int j = 0; j < kCpl; ++j) {
        float fl, fr;
        chain_exact(x[j], D, mf, fl, fr);
        hit |= (fl == fm) && ((rm >> j) & 1u);
    }
    return hit;
}
// fields of 8 / 16 / 32 bits (cold)
__device__ __noinline__ void emit_wide(const WarpArgs &a, int c, int w, int lane, float D, float mf, uint32_t rm, int W, uint32_t *tw) {
    float x[kCpl];
    load_lane_global(a, c, w, lane, x);
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        uint32_t k[kEpt], sg[kEpt];
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            float fl, fr;
            chain_exact(x[16 * h + j], D, mf, fl, fr);
            k[j] = (uint32_t)fminf(fl, 2147483520.0f) + ((rm >> (16 * h + j)) & 1u);       // overflow already reported
            sg[j] = __float_as_uint(x[16 * h + j]) >> 31;
        }
        pack_store_w(W, k, sg, tw, 2 * lane + h);
    }
}
// array outputs of the drop-in API (k, sign bit, dequantised value)
__device__ __noinline__ void emit_arrays(const WarpArgs &a, int c, int w, int lane, float L1f, float D, float mf, uint32_t rm) {
    float x[kCpl];
    load_lane_global(a, c, w, lane, x);
    const int64_t i0 = (int64_t)w * kCodeTile + (int64_t)lane * kCpl;
    bool ovf = false;
#pragma unroll 4
    for (int j = 0; j < kCpl; ++j) {
        const int64_t i = i0 + j;
        if (i >= a.d) break;
        float fl, fr;
        chain_exact(x[j], D, mf, fl, fr);
        const float kf = __fadd_rn(fl, (float)((rm >> j) & 1u));
        const uint32_t sbit = __float_as_uint(x[j]) >> 31;
        const int64_t o = (int64_t)c * a.ld_out + i;
        if (a.deq_out) {
            // sign(v) of AS:640: v = x / D is zero exactly when m * |v| is (floor and fraction both zero, m > 0)
            const float sgf = (fl == 0.0f && fr == 0.0f) ? 0.0f : (sbit ? -1.0f : 1.0f);
            a.deq_out[o] = __fdiv_rn(__fmul_rn(__fmul_rn(L1f, sgf), kf), mf);
        }
        if (a.k_out) {
            if (kf >= 2147483648.0f) { ovf = true; a.k_out[o] = 0x7fffffff; }
            else a.k_out[o] = (int32_t)kf;
        }
        if (a.sgn_out) a.sgn_out[o] = (uint8_t)sbit;
    }
    if (ovf) atomicOr(&a.hdr->status, 1u);
}

// ---------------------------------------------------------------------------------------------------------------- the kernel
__device__ __forceinline__ Item item_of(const WarpArgs &a, uint32_t tk) {
    Item it; it.c = 0; it.w = 0; it.tk = tk; it.valid = tk < a.total;
    if (it.valid) {
        uint32_t c = __umulhi(tk, a.recipT4), w = tk - c * a.T4;                 // c is the quotient or one less
        if (w >= a.T4) { ++c; w -= a.T4; }
        it.c = (int)c; it.w = (int)w;
    }
    return it;
}
// lane 0: start the copy of the item's tile into `buf` (or complete the barrier's phase when no full row is left to copy)
__device__ __forceinline__ void issue_tile(const WarpArgs &a, const CUtensorMap *tmap, const Item &it, uint32_t buf, uint32_t bar, uint64_t pol) {
    if ((int64_t)it.w * 32 < a.rows32) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, (uint32_t)kTileBytes);
        tma_tile_g2s(buf, tmap, it.w * 32, it.c, bar, pol);
    } else {
        mbar_arrive(bar);
    }
}

template <int EMIT>
__global__ void __launch_bounds__(kQThreads, kLag == 2 ? 3 : 4)
quantize_warp_kernel(const __grid_constant__ WarpArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char dyn_smem[];      // per warp: three 4 KB tile buffers; then the mbarriers
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ring = smem_u32(dyn_smem) + (uint32_t)warp * (kRing * kTileBytes);
    const uint32_t bars = smem_u32(dyn_smem) + (uint32_t)(kQWarps * kRing * kTileBytes) + (uint32_t)warp * 32u;
    const uint32_t scratch = smem_u32(dyn_smem) + (uint32_t)(kQWarps * kRing * kTileBytes) + (uint32_t)kQWarps * 32u + (uint32_t)threadIdx.x * 32u;
    const uint32_t swz = ((uint32_t)lane & 7u) << 4, rowoff = (uint32_t)lane * 128u;
    const uint64_t pol = policy_evict_first();
    if (lane == 0) {
        for (int q = 0; q < kRing; ++q) mbar_init(bars + 8u * q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // Tickets.  In steady state a warp draws, at the end of every iteration, the tile it will work on two iterations later; the
    // first two rounds are dealt statically the same way (round r, warp g: ticket r * G + g), so that the look-back never
    // starts out with neighbouring tiles one iteration apart.  After that: one atomic per tile.
    const uint32_t G = gridDim.x * (uint32_t)kQWarps, gw = blockIdx.x * (uint32_t)kQWarps + (uint32_t)warp;
    // static round-robin: round r, warp g works on ticket r * G + g
    uint32_t tk_round = 2;
    bool tk_more = true;
    auto next_item = [&]() -> Item {
        Item it; it.c = 0; it.w = 0; it.tk = 0; it.valid = false;
        if (!tk_more) return it;
        const unsigned long long t = (unsigned long long)tk_round * G + gw;
        ++tk_round;
        if (t < (unsigned long long)a.total) it = item_of(a, (uint32_t)t);
        if (!it.valid) tk_more = false;
        if (it.valid) DBG_MARK(it.tk, 0);
        return it;
    };
    Item iB = item_of(a, gw), iN = item_of(a, G + gw), iC, iC1;      // kLag == 2: iC1 = the tile parked between its B- and C-phase
    iC1.c = 0; iC1.w = 0; iC1.tk = 0; iC1.valid = false;
    if (!iN.valid) tk_more = false;
    iC.c = 0; iC.w = 0; iC.tk = 0; iC.valid = false;
    if (lane == 0) {
        if (iB.valid) issue_tile(a, &tmap, iB, ring, bars, pol);
        if (iN.valid) issue_tile(a, &tmap, iN, ring + kTileBytes, bars + 8u, pol);
    }
    Carry cy, cy1;                              // of tile iC (and of iC1)
    cy1.sgw0 = cy1.sgw1 = cy1.flm = cy1.L0 = cy1.L1 = cy1.ex32 = 0; cy1.mxl = 0.0f; cy1.fmf = 0.0f; cy1.Aq = 0; cy1.AI = 0;
    cy1.fl4[0] = cy1.fl4[1] = cy1.fl4[2] = cy1.fl4[3] = 0;
    cy.sgw0 = cy.sgw1 = cy.flm = cy.L0 = cy.L1 = cy.ex32 = 0; cy.mxl = 0.0f; cy.fmf = 0.0f; cy.Aq = 0; cy.AI = 0;
    cy.fl4[0] = cy.fl4[1] = cy.fl4[2] = cy.fl4[3] = 0;
    const int qshift = a.qshift, sh = 43 - a.qshift;
    int sB = 0;
    for (uint32_t it = 0;; ++it) {
        if (!iB.valid && !iC.valid && !(kLag == 2 && iC1.valid)) break;
        const int sC = (sB + kRing - kLag) % kRing;
        Carry nb;                               // of tile iB
        nb.sgw0 = nb.sgw1 = nb.flm = nb.L0 = nb.L1 = nb.ex32 = 0; nb.mxl = 0.0f; nb.fmf = 0.0f; nb.Aq = 0; nb.AI = 0;
        nb.fl4[0] = nb.fl4[1] = nb.fl4[2] = nb.fl4[3] = 0;
        unsigned long long blk_old = 0;
        // Look-back loads of tile iC (earlier tiles of its block, earlier blocks of its super-block, earlier super-blocks): issued
        // after tile iB has been published; the scan of tile iB runs under their round trip to L2.  (Issued before the B-phase:
        // 4.4 ms instead of 3.4, right after its arithmetic: 3.5 -- the later, the more of the records are complete.)
        const int pos = iC.w & 31, bpos = (iC.w >> 5) & 31, S = iC.w >> 10;
        const unsigned long long *tp = a.trec + (iC.tk - 1u - (uint32_t)lane);
        const unsigned long long *bp = a.brec + ((uint32_t)iC.c * a.TB + (uint32_t)(S * 32 + lane));
        const SupRec *sp = a.srec + (uint32_t)iC.c * a.TS;
        unsigned long long tv = 1ull, bv = 32ull << 58, slo = 32ull << kSupShift, shi = 32ull << kSupShift;
        auto window_issue = [&]() {
            if (iC.valid) {
                if (lane < pos) tv = ld_rec(tp);
                if (lane < bpos) bv = ld_rec(bp);
                if (lane < S) ld_rec2(sp + lane, slo, shi);
            }
        };
        // kLag == 2: tile iC was published a whole iteration ago, and so were its predecessors (the warps move in lock-step rounds):
        // the window is complete now, and its round trip to L2 hides behind the B-phase
        if (kLag == 2) window_issue();
        // ---------------------------------------------------------------- B-phase of tile iB
        if (iB.valid) {
            const RowConst *rcp = a.consts + iB.c;
            const float D = __ldg(&rcp->D), rcpD = __ldg(&rcp->rcpD), mf = __ldg(&rcp->mf);
            const uint32_t rflags = __ldg(&rcp->flags);
            const uint32_t buf = ring + (uint32_t)sB * kTileBytes, rowb = buf + rowoff;
            DBG_MARK(iB.tk, 1);
            mbar_wait(bars + 8u * sB, (it / kRing) & 1u);
#ifdef DME_TIMERS
            if (g_dbg && lane == 0) g_dbg[(size_t)iB.tk * 12 + 5] = gtime();
#endif
            if ((int64_t)iB.w * 32 >= a.rows32) {                   // nothing was copied: the tile is the row's tail only
#pragma unroll
                for (int q = 0; q < 8; ++q) sts128u(rowb + ((uint32_t)q << 4), make_uint4(0u, 0u, 0u, 0u));
            }
            if (a.tail && iB.w == (int)a.T4 - 1) {
                // the last d % 32 coordinates of the row are not covered by the tensor map (their row arrived as zeros)
                __syncwarp();
                if (lane == (int)(a.rows32 - (int64_t)iB.w * 32)) {
                    const float *src = a.X + (int64_t)iB.c * a.ld + a.rows32 * 32;
                    for (int j = 0; j < a.tail; ++j)
                        sts32(rowb + ((((uint32_t)(j >> 2)) << 4) ^ swz) + 4u * (uint32_t)(j & 3), __float_as_uint(src[j]));
                }
                __syncwarp();
            }
            bool exact = (rflags & kRowExact) != 0u;
            if (!exact && (rflags & kRowGuardFloor)) {
                // m * p can reach 2^23 in this row: lanes that actually see such a value use floorf
                float mx = 0.0f;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint4 v = lds128u(rowb + (((uint32_t)q << 4) ^ swz));
                    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(__uint_as_float(v.x)), fabsf(__uint_as_float(v.y))), fmaxf(fabsf(__uint_as_float(v.z)), fabsf(__uint_as_float(v.w)))));
                }
                exact = !(__fmul_rn(mf, __fmul_rn(mx, rcpD)) < 4194304.0f);
            }
            double run;
            if (exact) {
                run = b_phase_exact(rowb, swz, D, rcpD, mf, scratch);
                const uint4 r = lds128u(scratch), r4 = lds128u(scratch + 16u);
                nb.sgw0 = r.x; nb.sgw1 = r.y; nb.flm = r.z; nb.mxl = __uint_as_float(r.w);
                nb.fl4[0] = r4.x; nb.fl4[1] = r4.y; nb.fl4[2] = r4.z; nb.fl4[3] = r4.w;
            } else {
                run = b_phase_body<false>(rowb, swz, D, rcpD, mf, nb);
            }
            DBG_MARK(iB.tk, 9);
            // lane sum -> 2^-43 fixed point (run < 32): mantissa of run + 1.5 * 2^9
            const double tfx = __dadd_rn(run, 768.0);
            const uint32_t flo = (uint32_t)__double2loint(tfx), fhi = (uint32_t)__double2hiint(tfx) & 0x7ffffu;
            nb.L0 = flo & 0xffffffu;
            nb.L1 = __funnelshift_r(flo, fhi, 24);
            const uint32_t r0 = __reduce_add_sync(0xffffffffu, nb.L0), r1 = __reduce_add_sync(0xffffffffu, nb.L1);
            const unsigned long long A43 = (unsigned long long)r0 + ((unsigned long long)r1 << 24);
            nb.AI = (uint32_t)(A43 >> 43);
            nb.Aq = sh > 0 ? (A43 + (1ull << (sh - 1))) >> sh : A43;
            nb.fmf = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(nb.mxl)));
            if (lane == 0) {
                st_rec(a.trec + iB.tk, (nb.Aq << 1) | 1ull);
                DBG_MARK(iB.tk, 2);
                blk_old = atom_add_u64(a.brec + ((uint32_t)iB.c * a.TB + ((uint32_t)iB.w >> 5)), nb.Aq | kBlkCnt1);
            }
        }
        // ---------------------------------------------------------------- C-phase of tile iC: AS:635-637, emit
        if (kLag == 1) window_issue();
        if (iB.valid) nb.ex32 = warp_excl_scan_u32((nb.L0 >> 11) | (nb.L1 << 13), lane);      // lane sums in 2^-32 units (mod 2^32)
        // The tile that completed its block forwards the block total to the super-block record: at the end of the iteration (the
        // atomic's round trip hides behind the C-phase), but BEFORE this warp starts to wait for anything -- the look-back of a
        // later super-block depends on it.
        bool fwd_pending = iB.valid;
        auto forward_block = [&]() {
            asm volatile("" : "+l"(blk_old));                       // keep the use of the atomic's result HERE
            if (lane == 0 && (blk_old >> 58) == 31ull) {
                const unsigned long long Bq = (blk_old & kBlkVal) + nb.Aq;
                SupRec *sr = a.srec + ((uint32_t)iB.c * a.TS + ((uint32_t)iB.w >> 10));
                red_add_u64(&sr->lo, (Bq & 0x7fffffffull) + (1ull << kSupShift));
                red_add_u64(&sr->hi, (Bq >> 31) + (1ull << kSupShift));
            }
            fwd_pending = false;
        };
        if (iC.valid) {
            const RowConst *rcp = a.consts + iC.c;
            const float X = __ldg(&rcp->X);
            const uint32_t rflags = __ldg(&rcp->flags);
            const uint32_t buf = ring + (uint32_t)sC * kTileBytes, rowb = buf + rowoff;
            // the parked fractions become fixed-point words while the look-back loads are in flight
            uint32_t pw[4][8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint4 va = lds128u(rowb + (((uint32_t)(2 * g) << 4) ^ swz)), vb = lds128u(rowb + (((uint32_t)(2 * g + 1) << 4) ^ swz));
                fx8(va, vb, pw[g]);
            }
            unsigned long long P;
            DBG_MARK(iC.tk, 3);
            for (;;) {
                const bool okl = (tv & 1ull) && (bv >> 58) == 32ull;
                bool oks = (slo >> kSupShift) == 32ull && (shi >> kSupShift) == 32ull;
                unsigned long long xs = ((shi & kSupVal) << 31) + (slo & kSupVal);
                for (int s = lane + 32; s < S; s += 32) {             // rows longer than 2^25 coordinates
                    unsigned long long lo2, hi2;
                    ld_rec2(sp + s, lo2, hi2);
                    oks = oks && (lo2 >> kSupShift) == 32ull && (hi2 >> kSupShift) == 32ull;
                    xs += ((hi2 & kSupVal) << 31) + (lo2 & kSupVal);
                }
                const unsigned long long xl = (tv >> 1) + (bv & kBlkVal);
                const uint32_t miss_l = __ballot_sync(0xffffffffu, !okl), miss_s = __ballot_sync(0xffffffffu, !oks);
                if ((miss_l | miss_s) == 0u) { P = warp_sum_u63(xl + xs); break; }
                if (miss_l == 0u && S >= 1 && S <= 32 && miss_s == (1u << (S - 1))) {
                    // only the newest super-block record is incomplete (its last blocks have not been forwarded yet): its 32 block
                    // records, complete as soon as their tiles are published, stand in for it
                    const unsigned long long v2 = ld_rec(a.brec + ((uint32_t)iC.c * a.TB + (uint32_t)((S - 1) * 32 + lane)));
                    if (__all_sync(0xffffffffu, (v2 >> 58) == 32ull)) { P = warp_sum_u63(xl + (lane == S - 1 ? 0ull : xs) + (v2 & kBlkVal)); break; }
                }
#ifdef DME_TIMERS
                if (lane == 0) {
                    atomicAdd(&a.hdr->pad[0], 1u);
                    if (miss_l) atomicAdd(&a.hdr->pad[2], 1u);
                    if (miss_s) atomicAdd(&a.hdr->pad[4], 1u);
                }
#endif
                if (fwd_pending) forward_block();
                __nanosleep(64);
                if (lane < pos) tv = ld_rec(tp);
                if (lane < bpos) bv = ld_rec(bp);
                if (lane < S) ld_rec2(sp + lane, slo, shi);
            }
            // The tile's prefixes lie in [P, P + A]: when that range (with margins) stays inside one binade of the fp32 prefix and
            // the row is inside the proven operand range, every lane starts from phi = frac(E - Xp) in 2^-32 units (mod 2^32).
            const uint32_t IP = (uint32_t)(P >> qshift);
            const uint32_t tlo = IP - 2u, thi = IP + cy.AI + 4u;
            const bool tile_fast = !(rflags & kRowExact) && IP >= 6u && thi < (1u << 23) && (__clz((int)tlo) == __clz((int)thi));
            bool slow = !tile_fast;
            uint32_t phi = 0;
            if (tile_fast) {
                const int e = 31 - __clz((int)tlo);
                const uint32_t Xp32 = xp32_of(X, e);                              // Xp = 2^(e-24) (2a - 1) in 2^-32 units
                phi = (uint32_t)((P << sh) >> 11) + cy.ex32 - Xp32;
            } else if (!(rflags & kRowExact)) {
                // cold: the tile straddles a binade (or starts below 6): every lane decides for itself from its own prefixes
                const uint32_t e0 = warp_excl_scan_u32(cy.L0, lane), e1 = warp_excl_scan_u32(cy.L1, lane);
                const unsigned long long ex43 = (unsigned long long)e0 + ((unsigned long long)e1 << 24);
                const unsigned long long in43 = ex43 + (unsigned long long)cy.L0 + ((unsigned long long)cy.L1 << 24);
                const uint32_t IElo = IP + (uint32_t)(ex43 >> 43), IEnhi = IP + (uint32_t)(in43 >> 43) + 2u;
                const uint32_t lo = IElo - 2u, hi = IEnhi + 2u;
                if (IElo >= 6u && hi < (1u << 23) && (__clz((int)lo) == __clz((int)hi))) {
                    const int e = 31 - __clz((int)lo);
                    const uint32_t Xp32 = xp32_of(X, e);
                    phi = (uint32_t)((P << sh) >> 11) + cy.ex32 - Xp32;
                    slow = false;
                }
            }
            uint32_t acc = phi + kBand, bits = 0, mn = acc;
#pragma unroll
            for (int g = 0; g < 4; ++g) carry8(acc, bits, mn, pw[g]);
            uint32_t rm = __brev(bits);
            slow = slow || mn < 2u * kBand;
            if (__any_sync(0xffffffffu, slow)) {
                // exact fp64 prefixes of the lanes: E before the first coordinate, En at the last one (= the next lane's E, bit for bit)
                const uint32_t e0 = warp_excl_scan_u32(cy.L0, lane), e1 = warp_excl_scan_u32(cy.L1, lane);
                const unsigned long long ex43 = (unsigned long long)e0 + ((unsigned long long)e1 << 24);
                const unsigned long long in43 = ex43 + (unsigned long long)cy.L0 + ((unsigned long long)cy.L1 << 24);
                const double Pd = __ll2double_rn((long long)P) * a.q_dn;
                const double k43 = 1.1368683772161603e-13;                         // 2^-43
                const double E = Pd + __ll2double_rn((long long)ex43) * k43;
                double En = Pd + __ll2double_rn((long long)in43) * k43;
                if (lane == 31) En = __ll2double_rn((long long)(P + cy.Aq)) * a.q_dn;      // = the next tile's first prefix
                if (slow) rm = lane_literal(rowb, swz, X, E, En);
            }
            __syncwarp();
            if (EMIT == 0) {
                emit_arrays(a, iC.c, iC.w, lane, __ldg(&rcp->L1f), __ldg(&rcp->D), __ldg(&rcp->mf), rm);
            } else {
                const float fm = cy.fmf;
                int W = 2;
                if (fm != 0.0f) {
                    W = width_of(fm);
                    const int Wh = width_of(__fadd_rn(fm, 1.0f));
                    if (W != Wh) {
                        // the largest floor sits right below a width boundary: did one of those coordinates also receive a unit?
                        bool hit = false;
                        if (cy.mxl == fm) {
                            if (fm == 1.0f) hit = (cy.flm & rm) != 0u;
                            else if (fm == 7.0f) {
#pragma unroll
                                for (int g = 0; g < 4; ++g)
#pragma unroll
                                    for (int j = 0; j < 8; ++j) hit |= ((cy.fl4[g] >> (4 * j)) & 15u) == 7u && ((rm >> (8 * g + j)) & 1u);
                            } else hit = lane_hit_wide(a, iC.c, iC.w, lane, __ldg(&rcp->D), __ldg(&rcp->mf), fm, rm);
                        }
                        if (__any_sync(0xffffffffu, hit)) W = Wh;
                    }
                    if (fm >= 2147483520.0f && lane == 0) atomicOr(&a.hdr->status, 1u);
                }
                unsigned long long off16;
                if (W <= a.pack.W0) {
                    off16 = primary_off16(a.pack, iC.c, iC.w);
                    if (lane == 0) a.pack.dir[iC.tk] = (off16 << 8) | (unsigned long long)W;
                } else {
                    off16 = 0;
                    if (lane == 0) off16 = place_code_tile(a.pack, iC.c, iC.w, W);
                    off16 = __shfl_sync(0xffffffffu, off16, 0);
                }
                if (off16 != ~0ull) {
                    uint32_t *tw = a.pack.codes + off16 * 4ull;
                    if (W == 2) {
                        // fields [sign | magnitude bit]: k = floor + r <= 1; the lane's two words are adjacent
                        uint2 w2;
                        w2.x = spread16(rm & 0xffffu) | cy.sgw0;
                        w2.y = spread16(rm >> 16) | cy.sgw1;
                        if (cy.flm) { w2.x |= spread16(cy.flm & 0xffffu); w2.y |= spread16(cy.flm >> 16); }
                        *reinterpret_cast<uint2 *>(tw + 2 * lane) = w2;
                    } else if (W == 4) {
                        // fields [sign | 3-bit magnitude]: k = floor + r as nibbles, signs at bit 3
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const uint32_t sg = (g < 2 ? cy.sgw0 : cy.sgw1) >> (16 * (g & 1));
                            uint32_t rs = 0;                  // nibble j = r_j | sign_j << 3
#pragma unroll
                            for (int j = 0; j < 8; ++j) rs |= (((rm >> (8 * g + j)) & 1u) | (((sg >> (2 * j + 1)) & 1u) << 3)) << (4 * j);
                            tw[(g & 1) * kCodeChunks + 2 * lane + (g >> 1)] = cy.fl4[g] + rs;       // k <= 7: no carry between nibbles
                        }
                    } else {
                        emit_wide(a, iC.c, iC.w, lane, __ldg(&rcp->D), __ldg(&rcp->mf), rm, W, tw);
                    }
                }
            }
        }
        if (iC.valid) DBG_MARK(iC.tk, 4);
        if (fwd_pending) forward_block();
        if (iC.valid) DBG_MARK(iC.tk, 6);
        // buffer sC is free: the tile after the one in flight goes there
        __syncwarp();
        Item iNN = iB.valid ? next_item() : Item{0, 0, 0u, false};
        if (iC.valid) DBG_MARK(iC.tk, 7);
        if (lane == 0 && iNN.valid) issue_tile(a, &tmap, iNN, ring + (uint32_t)sC * kTileBytes, bars + 8u * sC, pol);
        if (iC.valid) DBG_MARK(iC.tk, 8);
        if (kLag == 2) { iC = iC1; iC1 = iB; cy = cy1; cy1 = nb; }
        else { iC = iB; cy = nb; }
        iB = iN; iN = iNN;
        sB = sB == kRing - 1 ? 0 : sB + 1;
    }
}

// ------------------------------------------------------------------ host side
// Function attributes, occupancy and the SM count are per device: set / queried once for every device the library is used on
// (one process may drive several GPUs), under a mutex.
struct WarpDevice { bool ready = false; int sms = 0; int occ[2] = {0, 0}; };
static WarpDevice g_warp_dev[64];
static std::mutex g_warp_mu;
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
constexpr size_t kWarpDynSmem = (size_t)kQWarps * kRing * kTileBytes + (size_t)kQWarps * 32 + (size_t)kQThreads * 32;

static int warp_device(WarpDevice **out) {
    int dev = 0;
    DME_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return DME_ECUDA; }
    std::lock_guard<std::mutex> lock(g_warp_mu);
    WarpDevice &D = g_warp_dev[dev];
    if (!D.ready) {
        DME_CUDA(cudaFuncSetAttribute(quantize_warp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWarpDynSmem));
        DME_CUDA(cudaFuncSetAttribute(quantize_warp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWarpDynSmem));
        DME_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.occ[0], quantize_warp_kernel<0>, kQThreads, kWarpDynSmem));
        DME_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.occ[1], quantize_warp_kernel<1>, kQThreads, kWarpDynSmem));
        DME_CUDA(cudaDeviceGetAttribute(&D.sms, cudaDevAttrMultiProcessorCount, dev));
        if (g_encode == nullptr) {
            cudaDriverEntryPointQueryResult qres;
            void *fn = nullptr;
            DME_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
            if (fn == nullptr || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return DME_ECUDA; }
            g_encode = (EncodeTiledFn)fn;
        }
        D.ready = true;
    }
    *out = &D;
    return DME_OK;
}

// The row constants are there already (l1_kernel, launched by the caller on the same stream).
int launch_quantize_warp(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, const WsLayout &L, void *ws,
                         int32_t *k_out, uint8_t *sgn_out, float *deq_out, int64_t ld_out,
                         uint32_t *codes, int64_t codes_bytes, uint64_t *dir, cudaStream_t st, bool packed) {
    WarpDevice *D = nullptr;
    int rc = warp_device(&D);
    if (rc) return rc;
    char *base = (char *)ws;
    WarpArgs a;
    a.X = X; a.d = d; a.ld = ld; a.n = n;
    a.T4 = (uint32_t)((d + kCodeTile - 1) / kCodeTile);
    a.rows32 = d / 32;
    a.tail = (int)(d & 31);
    a.TB = (a.T4 + 31) / 32;                                       // blocks of 32 tiles per row
    a.TS = (a.TB + 31) / 32;                                       // super-blocks of 32 blocks per row
    const int64_t nT = n * (int64_t)a.T4;
    if (nT >= ((int64_t)1 << 32) - (1 << 20)) { set_error("n * tiles = %lld does not fit the 32-bit ticket counter", (long long)nT); return DME_EINVAL; }
    a.total = (uint32_t)nT;
    a.recipT4 = a.T4 == 1 ? 0xffffffffu : (uint32_t)(((uint64_t)1 << 32) / (uint64_t)a.T4);
    {
        int lg = 0;
        while (((int64_t)1 << lg) < d) ++lg;
        a.qshift = 62 - lg < 43 ? 62 - lg : 43;
        a.q_dn = ldexp(1.0, -a.qshift);
    }
    a.consts = (const RowConst *)(base + L.off_consts);
    a.trec = (unsigned long long *)(base + L.off_desc);            // tile records, then block records, then super-block records
    a.brec = a.trec + nT;
    a.srec = (SupRec *)(((uintptr_t)(a.brec + n * (int64_t)a.TB) + 15u) & ~(uintptr_t)15u);
    a.hdr = (WsHeader *)base;
    a.k_out = k_out; a.sgn_out = sgn_out; a.deq_out = deq_out; a.ld_out = ld_out;
    init_pack_target(a.pack, codes, codes_bytes, dir, a.hdr, n, d, m);
    if (packed && (long long)(a.pack.arena_base16 * 16ull) > codes_bytes) {
        set_error("code arena too small for the primary slots: %lld < %llu bytes", (long long)codes_bytes, a.pack.arena_base16 * 16ull);
        return DME_EWORKSPACE;
    }
    // 3-D view of the client rows: {32 floats, full 128-byte rows of a client, clients}; the last d % 32 coordinates
    // of every row are read directly by the kernel
    CUtensorMap tmap;
    {
        const cuuint64_t dims[3] = {32, (cuuint64_t)(a.rows32 > 0 ? a.rows32 : 1), (cuuint64_t)n};
        const cuuint64_t strides[2] = {128, (cuuint64_t)ld * 4};
        const cuuint32_t box[3] = {32, kCodeTile / 32, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = g_encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)X, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for n=%lld d=%lld ld=%lld", (int)r, (long long)n, (long long)d, (long long)ld); return DME_ECUDA; }
    }
    const int occ = D->occ[packed ? 1 : 0];
    if (occ < 1) { set_error("quantize_warp_kernel does not fit on an SM"); return DME_ECUDA; }
    int64_t G = (int64_t)D->sms * occ;             // every CTA resident: a look-back never waits on a warp that has not started
    const int64_t need = (nT + kQWarps - 1) / kQWarps;
    if (G > need) G = need;
    if (packed) quantize_warp_kernel<1><<<(unsigned)G, kQThreads, kWarpDynSmem, st>>>(a, tmap);
    else quantize_warp_kernel<0><<<(unsigned)G, kQThreads, kWarpDynSmem, st>>>(a, tmap);
    DME_LAUNCH_CHECK("quantize_warp_kernel");
    return DME_OK;
}

}  // namespace dme
#ifdef DME_TIMERS
extern "C" __attribute__((visibility("default"))) int dme_debug_buffer(void *p) {
    return (int)cudaMemcpyToSymbol(dme::g_dbg, &p, sizeof(p));
}
#endif
