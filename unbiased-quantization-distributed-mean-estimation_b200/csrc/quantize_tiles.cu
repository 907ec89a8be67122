// quantize_tiles.cu -- the quantize kernel of the unbiased type quantizer (AS:609-641): scale to m, floor + systematic-sampling
// allocation of the fractional mass, sign/magnitude packing (or the dequantised output of the drop-in API), for all clients of
// one GPU in ONE launch.  The per-client L1 norms and row constants come from l1_kernel (type_quantize.cu).
//
// Schedule.  128-thread CTAs (4 per SM) draw (client, tile) tickets in client-major order.  A tile goes through two phases
// that are one ticket apart while it stays in its shared-memory buffer (ring of three 16 KB buffers filled by TMA tensor
// copies: 3-D map {32 floats, rows of 128 B, client}, SWIZZLE_128B -- the blocked read "thread t owns coordinates
// [32t, 32t + 32)" is free of bank conflicts, rows past the end of a client vector arrive as zeros):
//     B-phase (tile i)  : division, floor, fractional parts (parked in place of x) -> thread sums -> warp scan -> tile aggregate,
//                         published at once;
//     C-phase (tile i-1): exclusive prefix of the tile (look-back window) -> floor(c - X) for every coordinate, emit.
// Every wait is on a smaller ticket held by a resident CTA that does not wait on a larger one: no deadlock.
//
// C-phase in closed form.  AS:636 evaluates t = floor(RN32(RN32(c) - X)) with c the fp64 prefix.  While c32 stays
// inside one binade [2^e + 1, 2^(e+1)), 2 <= e <= 22 (fp32 grid g = 2^(e-23)), this equals floor(c - Xp) when
// a = ceil(X/g - 1/2) is even and ceil(c - Xp) - 1 when a is odd, with Xp = g (a - 1/2) (proof in DESIGN.md
// section 3.1; ties of both roundings included; tests/test_closed_form_floor.py checks it exhaustively on the CPU).  So per
// coordinate: one DFMA (running sigma (c - Xp)) and one DADD.RM with the magic constant 1.5 * 2^52 whose low word is the
// floor -- no conversions; the 0/1 differences of consecutive floors telescope into one IMAD per coordinate that builds the
// 2-bit fields directly.  Threads whose prefixes cross a binade (or sit below 5.5) evaluate AS:636 literally.
//
// Look-back.  Every tile publishes its aggregate (int64 fixed point) as a 16-byte record and adds it, split in two
// 31-bit halves that each carry a contribution count, to the record of its block (32 tiles) and of its super-block
// (1024 tiles) with fire-and-forget 64-bit reductions.  A record is complete when both counts are full, so the
// exclusive prefix of a tile needs nothing but the B-phase of the earlier tiles of its row: one round of independent
// 16-byte loads (tiles of its block, blocks of its super-block, earlier super-blocks), prefetched into shared memory with
// cp.async right before the barrier that ends the B-phase.  Integer addition is associative: the result does not depend
// on timing.
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "type_quantize.cuh"

namespace dme {

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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"      // suspend-time hint: sleep in hardware, do not spin
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
constexpr int kBarFree = 1;
// one 16 KB box {32 floats, 128 rows, 1 client} at (0, row0, client) of the 3-D tensor map
__device__ __forceinline__ void tma_tile_g2s(uint32_t dst, const CUtensorMap *map, int row0, int client, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(dst), "l"(map), "r"(0), "r"(row0), "r"(client), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}


struct StreamArgs {
    const float *X; int64_t d, ld, T, n, m;
    int64_t rows32;                        // full 128-byte rows per client vector (the part the tensor map covers)
    RowConst *consts; BinadeEntry *tabs; TileRec *desc; Rec2 *blocks; Rec2 *supers; int64_t TB, TS; WsHeader *hdr;
    int tiles_tma, has_tail;
    int32_t *k_out; uint8_t *sgn_out; float *deq_out; int64_t ld_out;     // array outputs
    PackTarget pack;                                                    // packed output
};
constexpr uint32_t kItValid = 1u, kItTma = 4u, kItTail = 8u;


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



// swizzled shared-memory offset of the 16-byte chunk q (0..3) of thread tid's 16 coordinates inside a 16 KB tile
// (SWIZZLE_128B: chunk index within the 128-byte row is XORed with row & 7); q enters as an XOR of (q << 4)
__device__ __forceinline__ uint32_t blocked_off_of(uint32_t tid) {
    const uint32_t row = tid >> 1;
    return row * 128u + ((((tid & 1u) << 2) ^ (row & 7u)) << 4);
}

// AS:625-631 for the thread's 16 coordinates: floors and fractional parts of m |x| / D.  The fast chain (pairs of |x|:
// x/D by Markstein's correction of x * rcp, floor by adding 2^23 toward zero) and the literal one (rows / threads outside
// the proven operand range) give the same values.  `big` is decided per thread from its largest |x|.
__device__ __forceinline__ void floors_and_fracs(const float (&x)[kEpt], const RowConst &rc, float (&flf)[kEpt], float (&fr)[kEpt]) {
    const bool exact = rc.flags & kRowExact;
    bool big = false;
    if (!exact && (rc.flags & kRowGuardFloor)) {
        // m*p can reach 2^23 in this row: threads that actually see such a value use floorf
        float mx = 0.0f;
#pragma unroll
        for (int j = 0; j < kEpt; ++j) mx = fmaxf(mx, fabsf(x[j]));
        big = !(__fmul_rn(rc.mf, __fmul_rn(mx, rc.rcpD)) < 4194304.0f);
    }
    if (exact || big) {
#pragma unroll
        for (int j = 0; j < kEpt; ++j) chain_exact(x[j], rc, flf[j], fr[j]);
    } else {
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
            f2_unpack(fr2, fr[j], fr[j + 1]);
        }
    }
}
// the tile's coordinates owned by this thread, from the staged tile (+ the row tail straight from global)
__device__ __forceinline__ void load_x(const StreamArgs &a, uint32_t flags, int c, int t, uint32_t buf, float (&x)[kEpt], int chunk) {
    const uint32_t off = blocked_off_of((uint32_t)chunk);
    if (flags & kItTma) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = lds128((buf + off) ^ (uint32_t)(q << 4));
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kEpt; ++j) x[j] = 0.0f;
    }
    if (flags & kItTail) {          // the last d % 32 coordinates of the row are not covered by the tensor map
        const int64_t i0 = (int64_t)t * kTile + (int64_t)chunk * kEpt, lo = a.rows32 * 32;
        const float *row = a.X + (int64_t)c * a.ld;
        if (i0 + kEpt > lo && i0 < a.d) {
#pragma unroll
            for (int j = 0; j < kEpt; ++j)
                if (i0 + j >= lo && i0 + j < a.d) x[j] = row[i0 + j];
        }
    }
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

__device__ __forceinline__ uint32_t spread16(uint32_t v) {       // bit j -> bit 2j
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}
// r_j = [floor(c_j - X) - floor(c_{j-1} - X) == 1] (AS:636-637) for the thread's 16 coordinates as 2-bit interleaved
// fields (bit 2j = r_j).  Closed form: the 0/1 differences telescope into one IMAD per coordinate.
__device__ __forceinline__ uint32_t rbits_interleaved(const Geo &g, const float (&fr)[kEpt], float X) {
    if (g.fast) {
        const int sgi = g.sig > 0.0 ? 1 : -1;
        double u = g.sE;
        uint32_t acc = 0u - (uint32_t)floor_lo(u);
        int L14 = 0;
#pragma unroll
        for (int j = 0; j < kEpt - 1; ++j) {
            u = fma(frac_to_double(fr[j]), g.sig, u);
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
    uint32_t rb = 0;
    double c = g.E;
    int tp = floor_ref(c, X);
#pragma unroll
    for (int j = 0; j < kEpt; ++j) {
        c = (j < kEpt - 1) ? c + (double)fr[j] : g.En;
        const int t = floor_ref(c, X);
        rb |= ((t - tp == 1) ? 1u : 0u) << j;
        tp = t;
    }
    return spread16(rb);
}

// fields of 4 / 8 / 16 / 32 bits (cold: kept out of line)
__device__ __noinline__ void emit_wide_x(const float (&x)[kEpt], const RowConst &rc, uint32_t kw, uint32_t sgw, int W, uint32_t *tw, int chunk) {
    float fl[kEpt], fr[kEpt];
    floors_and_fracs(x, rc, fl, fr);                       // the floors again: same function of the same inputs
    uint32_t k[kEpt], sg[kEpt];
#pragma unroll
    for (int j = 0; j < kEpt; ++j) {
        const float ff = fminf(fl[j], 2147483520.0f);       // overflow already reported
        k[j] = (uint32_t)ff + ((kw >> (2 * j)) & 1u);
        sg[j] = (sgw >> (2 * j + 1)) & 1u;
    }
    switch (W) {
        case 4: pack_store<4>(k, sg, tw, chunk); break;
        case 8: pack_store<8>(k, sg, tw, chunk); break;
        case 16: pack_store<16>(k, sg, tw, chunk); break;
        default: pack_store<32>(k, sg, tw, chunk); break;
    }
}
// the thread's 16 coordinates of tile t of client c straight from global memory (cold paths of the tiles kernel)
__device__ __noinline__ void load_x_global(const StreamArgs &a, int c, int t, float (&x)[kEpt], int chunk) {
    const int64_t i0 = (int64_t)t * kTile + (int64_t)chunk * kEpt;
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


// Exclusive fixed-point prefix of tile t of a row: earlier tiles of its block + earlier blocks of its super-block +
// earlier super-blocks (see the header).  The records are loaded in one round (lookback_load) and evaluated later
// (lookback_eval), so that the warp can work while the loads are in flight; a round that finds an incomplete record
// is repeated.
struct LookRegs { unsigned long long tv, blo, bhi, slo, shi; uint32_t tf; };
__device__ __forceinline__ void lookback_load(const TileRec *tiles, const Rec2 *blocks, const Rec2 *supers, int t, int lane, LookRegs &r) {
    const int b = t >> 5, pos = t & 31, sb = b >> 5, bpos = b & 31;
    r.tv = 0; r.tf = 1u; r.blo = r.bhi = 32ull << kCntShift; r.slo = r.shi = 1024ull << kCntShift;
    if (lane < pos) r.tf = rec_load(tiles + (t - 1 - lane), r.tv);
    if (lane < bpos) rec2_load(blocks + (sb * 32 + lane), r.blo, r.bhi);
    if (lane < sb) rec2_load(supers + lane, r.slo, r.shi);
}
__device__ __forceinline__ bool lookback_eval(const Rec2 *supers, int t, int lane, const LookRegs &r, long long &P) {
    const int sb = t >> 10;
    bool ok = r.tf != 0u && (r.blo >> kCntShift) == 32ull && (r.bhi >> kCntShift) == 32ull && (r.slo >> kCntShift) == 1024ull &&
              (r.shi >> kCntShift) == 1024ull;
    long long x = (long long)r.tv + (long long)((((r.bhi & kSumMask) + (r.shi & kSumMask)) << 31) + (r.blo & kSumMask) + (r.slo & kSumMask));
    for (int s = lane + 32; s < sb; s += 32) {          // rows longer than 2^27 coordinates
        unsigned long long lo, hi;
        rec2_load(supers + s, lo, hi);
        ok = ok && (lo >> kCntShift) == 1024ull && (hi >> kCntShift) == 1024ull;
        x += (long long)(((hi & kSumMask) << 31) + (lo & kSumMask));
    }
    if (!__all_sync(0xffffffffu, ok)) return false;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    P = x;
    return true;
}
__device__ __forceinline__ int width_of(float kmax) {
    return kmax < 2.0f ? 2 : kmax < 8.0f ? 4 : kmax < 128.0f ? 8 : kmax < 32768.0f ? 16 : 32;
}
// ------------------------------------------------------------------ service warp: copies, row constants, look-backs
// Two-kernel path (the default): l1_kernel (type_quantize.cu) has published every row's constants; quantize_tiles_kernel
// makes ONE pass over the tiles in ticket order (client-major) with the same decoupled look-back records as above.
// Per CTA (128 threads, a thread owns 32 coordinates) and iteration:
//   B-phase of tile i   : floors, fractional parts (parked in place of x in the tile buffer), thread sums, scans;
//   one warp            : copies the look-back window of tile i-1 into shared memory (cp.async, completion on an mbarrier)
//                         as late as possible, i.e. right before the barrier that ends the B-phase;
//   every warp          : warp bases and aggregate of tile i; one thread publishes the aggregate;
//   C-phase of tile i-1 : window -> exclusive prefix -> floor(c - X) -> type vector, emit;
//   then the next ticket is taken and its tile copied into the buffer just freed (ring of three).
// No aggregate ever waits for a look-back.  Every wait is on a smaller ticket held by a resident CTA that does not wait
// on a larger one: no deadlock.  The fallback (a window record not complete yet: 0.3 % of the tiles at d = 2^24, n = 128)
// polls global memory.
struct __align__(16) TItem { int c, t; uint32_t flags; uint32_t ticket; };
struct TScratch {
    Rec win[2][96];              // look-back window of the C tile (tile / block / super-block records), by iteration parity
    double wtot[2][kWarps];      // warp totals of the B tile, by iteration parity
    uint32_t flmaxw[2][kWarps];
    TItem item[3];
    uint64_t mbar[3];
    uint64_t winbar[2];          // completion of the look-back window copies (32 arrivals: the lanes of warp 0), by iteration parity
    unsigned int hit[2];
    int rc_row[2];
    unsigned long long off16;
    long long pfb;               // look-back result of the fallback path
    RowConst rc[2];
    BinadeEntry tab[2][kBinades];
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// thread 0: decode ticket tk (0xffffffff: none) into item slot `slot` and start the tile's copy into buffer `slot`.
// (Drawing the ticket one iteration early to hide the atomic's round trip was measured: no gain, slightly slower.)
__device__ __forceinline__ unsigned int tiles_ticket(const StreamArgs &a) { return atomicAdd(&a.hdr->ticket, 1u); }
__device__ __forceinline__ void tiles_take(const StreamArgs &a, const CUtensorMap *tmap, TScratch &sc, int slot, uint32_t buf0, uint64_t pol, unsigned int tk) {
    TItem it; it.c = 0; it.t = 0; it.flags = 0; it.ticket = 0xffffffffu;
    if (tk != 0xffffffffu && (long long)tk < a.n * a.T) {
        it.ticket = tk;
        it.c = (int)(tk / (unsigned int)a.T); it.t = (int)(tk - (unsigned int)it.c * (unsigned int)a.T);
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
    sc.item[slot] = it;
}

// warp 0: start the copies of tile t's look-back window (earlier tiles of its block, earlier blocks of its super-block,
// earlier super-blocks) into shared memory; entries that do not exist are filled with complete neutral records
__device__ __forceinline__ void window_prefetch(const StreamArgs &a, Rec *win, int c, int t, int lane) {
    const TileRec *tiles = a.desc + (int64_t)c * a.T;
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
// any warp: exclusive fixed-point prefix of tile t from the window; false when a record was not complete yet
__device__ __forceinline__ bool window_eval(const Rec *win, int lane, long long &P) {
    const uint4 tr = *reinterpret_cast<const uint4 *>(&win[lane]);
    const Rec2 *w2 = reinterpret_cast<const Rec2 *>(win);
    const Rec2 br = w2[32 + lane], sr = w2[64 + lane];
    const bool ok = tr.z != 0u && (br.lo >> kCntShift) == 32ull && (br.hi >> kCntShift) == 32ull && (sr.lo >> kCntShift) == 1024ull &&
                    (sr.hi >> kCntShift) == 1024ull;
    long long x = (long long)(((unsigned long long)tr.y << 32) | tr.x) +
                  (long long)((((br.hi & kSumMask) + (sr.hi & kSumMask)) << 31) + (br.lo & kSumMask) + (sr.lo & kSumMask));
    if (!__all_sync(0xffffffffu, ok)) return false;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    P = x;
    return true;
}

// eight 2-bit pairs [s | r] (low 16 bits of v) -> eight nibbles [s 0 0 r]
__device__ __forceinline__ uint32_t spread_pairs_to_nibbles(uint32_t v) {
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    return (v & 0x11111111u) | ((v & 0x22222222u) << 2);
}

// ====================================================================================================================
// 128 threads per tile, TWO chunks per thread: thread t owns coordinates [32t, 32t + 32) = one 128-byte swizzle row of the
// staged tile (its eight 16-byte pieces fall into distinct banks for the eight lanes of a quarter warp).  The per-thread
// work that does not depend on the number of coordinates (warp scans, prefix geometry, look-back evaluation, control) is
// paid once per 32 coordinates, and a CTA-wide barrier joins 4 warps.  The prefix after a thread's first chunk is DEFINED
// as E + (sum of the first 16 fractional parts, left to right); the second chunk starts from that very value.
// (A 256-thread variant with one chunk per thread was measured: 4.26 ms vs 4.05 ms at d = 2^24, n = 128 with the same
// look-back timing.)
constexpr int kThreads2 = kThreads / 2, kWarps2 = kThreads2 / 32;
// B-phase of one chunk: signs, floors / fractions (parked in place of x), floor masks, running fp64 sum of the fractions
__device__ __forceinline__ void tiles_chunk_b(const StreamArgs &a, const TItem &iB, uint32_t buf, uint32_t boff, int ch, const RowConst &rc,
                                               uint32_t &sgw, uint32_t &flm, uint32_t &fl4a, uint32_t &fl4b, float &mx, double &run, bool first) {
    float x[kEpt], flf[kEpt], fr[kEpt];
    load_x(a, iB.flags, iB.c, iB.t, buf, x, ch);
#pragma unroll
    for (int j = kEpt - 1; j >= 0; --j) sgw = __funnelshift_l(__float_as_uint(x[j]), sgw, 2);
    floors_and_fracs(x, rc, flf, fr);
#pragma unroll
    for (int q = 0; q < 4; ++q)
        sts128((buf + boff) ^ (uint32_t)(q << 4), make_float4(fr[4 * q], fr[4 * q + 1], fr[4 * q + 2], fr[4 * q + 3]));
    mx = flf[0];
#pragma unroll
    for (int j = 1; j < kEpt; ++j) mx = fmaxf(mx, flf[j]);
    if (mx != 0.0f) {
#pragma unroll
        for (int j = 0; j < kEpt; ++j) flm |= ((flf[j] != 0.0f) ? 1u : 0u) << (2 * j);
#pragma unroll
        for (int j = 0; j < kEpt / 2; ++j) {
            fl4a |= (uint32_t)fminf(flf[j], 15.0f) << (4 * j);
            fl4b |= (uint32_t)fminf(flf[j + kEpt / 2], 15.0f) << (4 * j);
        }
    }
    if (first) run = (double)fr[0];
    else run += (double)fr[0];
#pragma unroll
    for (int j = 1; j < kEpt; ++j) run += (double)fr[j];
}
__device__ __forceinline__ uint32_t tiles_chunk_c(const BinadeEntry *tab, uint32_t buf, uint32_t boff, double E, double En, float X) {
    float fr[kEpt];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 v = lds128((buf + boff) ^ (uint32_t)(q << 4));
        fr[4 * q] = v.x; fr[4 * q + 1] = v.y; fr[4 * q + 2] = v.z; fr[4 * q + 3] = v.w;
    }
    const Geo g = make_geo(tab, E, En);
    return rbits_interleaved(g, fr, X);
}
template <int EMIT>
__global__ void __launch_bounds__(kThreads2, 4)
quantize_tiles_kernel(const __grid_constant__ StreamArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char dyn_smem[];      // three tile buffers, TScratch
    TScratch &sc = *reinterpret_cast<TScratch *>(dyn_smem + (size_t)3 * kTile * sizeof(float));
    const uint32_t buf0 = smem_u32(dyn_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ch0 = 2 * tid, ch1 = 2 * tid + 1;
    const uint64_t pol = policy_evict_first();
    const uint32_t boff0 = blocked_off_of((uint32_t)ch0), boff1 = blocked_off_of((uint32_t)ch1);
    const bool window_ok = a.TS <= 32;
    if (tid == 0) {
        for (int q = 0; q < 3; ++q) mbar_init(&sc.mbar[q], 1);
        mbar_init(&sc.winbar[0], 32); mbar_init(&sc.winbar[1], 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned int t0 = tiles_ticket(a), t1 = tiles_ticket(a);
        tiles_take(a, &tmap, sc, 0, buf0, pol, t0);
        tiles_take(a, &tmap, sc, 1, buf0, pol, t1);
        tiles_take(a, &tmap, sc, 2, buf0, pol, 0xffffffffu);
        sc.rc_row[0] = sc.rc_row[1] = -1;
        sc.hit[0] = sc.hit[1] = 0;
    }
    __syncthreads();
    // state of the tile whose C-phase is pending (one iteration behind its B-phase); index 0 / 1 = the thread's chunks
    uint32_t sgw0P = 0, sgw1P = 0, flm0P = 0, flm1P = 0, f4a0P = 0, f4b0P = 0, f4a1P = 0, f4b1P = 0;
    float mxfP = 0.0f;
    double inclP = 0.0, run0P = 0.0;
    double wbaseP = 0.0, wnextP = 0.0;
    uint32_t fmP = 0;
    long long AqP = 0;
    int sB = 0, sC = 2, useB = 0;
    for (int it = 0;; ++it) {
        const TItem iB = sc.item[sB];
        TItem iC = sc.item[sC];
        if (it == 0) iC.flags = 0;
        const bool validB = iB.flags & kItValid, validC = iC.flags & kItValid;
        if (!validB && !validC) break;
        const int e = it & 1;
        // ---------------------------------------------------------------- B-phase of tile iB
        uint32_t sgw0 = 0, sgw1 = 0, flm0 = 0, flm1 = 0, f4a0 = 0, f4b0 = 0, f4a1 = 0, f4b1 = 0;
        float mxf = 0.0f;
        double incl = 0.0, run0 = 0.0;
        if (validB) {
            if (sc.rc_row[e] != iB.c) {          // CTA-uniform: the row's constants and binade table into shared memory
                __syncthreads();
                if (tid < (int)(sizeof(RowConst) / 16))
                    reinterpret_cast<uint4 *>(&sc.rc[e])[tid] = __ldg(reinterpret_cast<const uint4 *>(&a.consts[iB.c]) + tid);
                else if (tid >= 32 && tid < 32 + kBinades)
                    reinterpret_cast<uint4 *>(sc.tab[e])[tid - 32] = __ldg(reinterpret_cast<const uint4 *>(a.tabs + (int64_t)iB.c * kBinades) + (tid - 32));
                if (tid == 64) sc.rc_row[e] = iB.c;
                __syncthreads();
            }
            const RowConst &rc = sc.rc[e];
            const uint32_t buf = buf0 + (uint32_t)sB * kTile * 4u;
            mbar_wait(smem_u32(&sc.mbar[sB]), (uint32_t)(useB & 1));
            float mx0, mx1;
            double run;
            tiles_chunk_b(a, iB, buf, boff0, ch0, rc, sgw0, flm0, f4a0, f4b0, mx0, run, true);
            run0 = run;
            tiles_chunk_b(a, iB, buf, boff1, ch1, rc, sgw1, flm1, f4a1, f4b1, mx1, run, false);
            mxf = fmaxf(mx0, mx1);
            incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            if (lane == 31) sc.wtot[e][warp] = incl;
            const uint32_t wmx = __reduce_max_sync(0xffffffffu, __float_as_uint(mxf));
            if (lane == 0) sc.flmaxw[e][warp] = wmx;
        }
        // The serial jobs of an iteration are spread over the four warps (window copies: warp 1, publish: warp 2, next ticket and
        // copy: warp 3, directory / arena: warp 0): warp w of every CTA runs on scheduler w, so giving them all to warp 0
        // overloads one scheduler while the other three wait at the barrier.
        if (warp == 1) {
            // The look-back window of tile C is copied as LATE as possible -- the later, the more of the earlier tiles'
            // aggregates are there (issued at the start of the B-phase a third of the tiles found it incomplete and had to
            // poll) -- and nobody waits for the copies here: they signal winbar[e], which the C-phase checks.
            if (validC && window_ok) window_prefetch(a, sc.win[e], iC.c, iC.t, lane);
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&sc.winbar[e])) : "memory");
        }
        __syncthreads();
        // ---------------------------------------------------------------- every warp: warp bases of tile B (fixed order), its aggregate
        double wbase = 0.0, wnext = 0.0;
        uint32_t fm = 0;
        long long Aq = 0;
        if (validB) {
            double wi = lane < kWarps2 ? sc.wtot[e][lane] : 0.0;
#pragma unroll
            for (int o = 1; o < kWarps2; o <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += up;
            }
            wnext = __shfl_sync(0xffffffffu, wi, warp);
            wbase = __shfl_sync(0xffffffffu, wi, warp > 0 ? warp - 1 : 0);
            if (warp == 0) wbase = 0.0;
            const double A = __shfl_sync(0xffffffffu, wi, kWarps2 - 1);
            fm = __reduce_max_sync(0xffffffffu, lane < kWarps2 ? sc.flmaxw[e][lane] : 0u);
            Aq = __double2ll_rn(A * sc.rc[e].q_up);          // fixed point, 2^-qshift resolution
            if (tid == 64) {
                rec_store(a.desc + (int64_t)iB.c * a.T + iB.t, (unsigned long long)Aq, 1u);
                const unsigned long long lo = ((unsigned long long)Aq & 0x7fffffffull) + (1ull << kCntShift);
                const unsigned long long hi = ((unsigned long long)Aq >> 31) + (1ull << kCntShift);
                Rec2 *br = a.blocks + (int64_t)iB.c * a.TB + (iB.t >> 5), *sr = a.supers + (int64_t)iB.c * a.TS + (iB.t >> 10);
                red_add_u64(&br->lo, lo); red_add_u64(&br->hi, hi);
                red_add_u64(&sr->lo, lo); red_add_u64(&sr->hi, hi);
            }
        }
        if (tid == 0) sc.hit[e ^ 1] = 0;
        // ---------------------------------------------------------------- C-phase of tile iC: AS:635-637
        uint32_t kw0 = 0, kw1 = 0;
        const float fmf = __uint_as_float(fmP);
        const bool need_hit = validC && EMIT == 1 && width_of(fmf) != width_of(__fadd_rn(fmf, 1.0f));      // CTA-uniform
        if (validC) {
            const RowConst &rc = sc.rc[e ^ 1];
            const uint32_t buf = buf0 + (uint32_t)sC * kTile * 4u;
            long long P = 0;
            mbar_wait(smem_u32(&sc.winbar[e]), (uint32_t)((it >> 1) & 1));          // the window copies have landed
            if (iC.t > 0 && !(window_ok && window_eval(sc.win[e], lane, P))) {
                if (warp == 0) {
                    const TileRec *tiles = a.desc + (int64_t)iC.c * a.T;
                    const Rec2 *blocks = a.blocks + (int64_t)iC.c * a.TB, *supers = a.supers + (int64_t)iC.c * a.TS;
                    LookRegs r;
                    for (;;) {
                        lookback_load(tiles, blocks, supers, iC.t, lane, r);
                        if (lookback_eval(supers, iC.t, lane, r, P)) break;
                        __nanosleep(200);
                    }
                    if (lane == 0) sc.pfb = P;
                }
                __syncthreads();
                P = sc.pfb;
            }
            const double Pd = __ll2double_rn(P) * rc.q_dn;
            double excl = __shfl_up_sync(0xffffffffu, inclP, 1);
            if (lane == 0) excl = 0.0;
            const double Pw = Pd + wbaseP;
            const double E = Pw + excl;
            double En = Pw + inclP;
            if (lane == 31) En = Pd + wnextP;                    // = the next warp's first prefix, bit for bit
            if (tid == kThreads2 - 1) En = __ll2double_rn(P + AqP) * rc.q_dn;      // = the next tile's first prefix
            const double Em = E + run0P;                         // prefix after the first chunk
            kw0 = tiles_chunk_c(sc.tab[e ^ 1], buf, boff0, E, Em, rc.X);
            kw1 = tiles_chunk_c(sc.tab[e ^ 1], buf, boff1, Em, En, rc.X);
            if (need_hit && mxfP == fmf) {
                // the largest floor sits right below a width boundary: did one of those coordinates also receive a unit?
                bool hit;
                if (fmf == 1.0f) hit = ((flm0P & kw0) | (flm1P & kw1)) != 0u;
                else {
                    hit = false;
#pragma unroll 1
                    for (int h = 0; h < 2; ++h) {
                        float x[kEpt], fl[kEpt], fr2[kEpt];
                        load_x_global(a, iC.c, iC.t, x, h ? ch1 : ch0);
                        floors_and_fracs(x, rc, fl, fr2);
                        const uint32_t kw = h ? kw1 : kw0;
#pragma unroll
                        for (int j = 0; j < kEpt; ++j) hit |= (fl[j] == fmf) && ((kw >> (2 * j)) & 1u);
                    }
                }
                if (hit) atomicOr(&sc.hit[e], 1u);
            }
        }
        // Buffer sC is free once every warp has read its fractional parts: only the warp that issues the next copy into it has
        // to wait for that, the others just signal.  When the tile's width depends on sc.hit, everybody waits.
        if (warp == 3 || need_hit) bar_sync(kBarFree, kThreads2);
        else bar_arrive(kBarFree, kThreads2);
        if (tid == 96) tiles_take(a, &tmap, sc, sC, buf0, pol, validB ? tiles_ticket(a) : 0xffffffffu);
        // ---------------------------------------------------------------- emit tile iC
        if (validC) {
            const RowConst &rc = sc.rc[e ^ 1];
            const float fm_c = fmf;
            if (EMIT == 0) {
                bool ovf = false;
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    float x[kEpt], fl[kEpt], fr2[kEpt];
                    const int ch = h ? ch1 : ch0;
                    const uint32_t kw = h ? kw1 : kw0, sgwP = h ? sgw1P : sgw0P;
                    load_x_global(a, iC.c, iC.t, x, ch);
                    floors_and_fracs(x, rc, fl, fr2);
                    const int64_t i0 = (int64_t)iC.t * kTile + (int64_t)ch * kEpt;
#pragma unroll
                    for (int j = 0; j < kEpt; ++j) {
                        const float kf = __fadd_rn(fl[j], (float)((kw >> (2 * j)) & 1u));
                        const int64_t i = i0 + j;
                        if (i >= a.d) continue;
                        const uint32_t sbit = (sgwP >> (2 * j + 1)) & 1u;
                        if (a.deq_out) {
                            // sign(v) of AS:640: v = x / D is zero exactly when m * |v| is (floor and fraction both zero, m > 0)
                            const float sgf = (fl[j] == 0.0f && fr2[j] == 0.0f) ? 0.0f : (sbit ? -1.0f : 1.0f);
                            a.deq_out[(int64_t)iC.c * a.ld_out + i] = __fdiv_rn(__fmul_rn(__fmul_rn(rc.L1f, sgf), kf), rc.mf);
                        }
                        if (a.k_out) {
                            if (kf >= 2147483648.0f) { ovf = true; a.k_out[(int64_t)iC.c * a.ld_out + i] = 0x7fffffff; }
                            else a.k_out[(int64_t)iC.c * a.ld_out + i] = (int32_t)kf;
                        }
                        if (a.sgn_out) a.sgn_out[(int64_t)iC.c * a.ld_out + i] = (uint8_t)sbit;
                    }
                }
                if (ovf) atomicOr(&a.hdr->status, 1u);
            } else {
                const int W = sc.hit[e] ? width_of(__fadd_rn(fm_c, 1.0f)) : width_of(fm_c);
                if (fm_c >= 2147483520.0f && tid == 0) atomicOr(&a.hdr->status, 1u);
                const int64_t slot_id = (int64_t)iC.c * a.T + iC.t;
                unsigned long long off16;
                if (W <= a.pack.W0) {
                    off16 = primary_off16(a.pack, iC.c, iC.t);
                    if (tid == 0) a.pack.dir[slot_id] = (off16 << 8) | (unsigned long long)W;
                } else {
                    if (tid == 0) {
                        const unsigned long long units = 32ull * W;
                        unsigned long long o = a.pack.arena_base16 + atomicAdd(&a.hdr->arena_top, units);
                        if ((long long)((o + units) * 16ull) > a.pack.codes_bytes) { atomicOr(&a.hdr->status, 2u); o = ~0ull; }
                        sc.off16 = o;
                        a.pack.dir[slot_id] = (o == ~0ull) ? 0ull : ((o << 8) | (unsigned long long)W);
                    }
                    __syncthreads();
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
                            load_x_global(a, iC.c, iC.t, x, h ? ch1 : ch0);
                            emit_wide_x(x, rc, h ? kw1 : kw0, h ? sgw1P : sgw0P, W, tw, h ? ch1 : ch0);
                        }
                    }
                }
            }
        }
        sgw0P = sgw0; sgw1P = sgw1; flm0P = flm0; flm1P = flm1; f4a0P = f4a0; f4b0P = f4b0; f4a1P = f4a1; f4b1P = f4b1;
        mxfP = mxf; inclP = incl; run0P = run0; wbaseP = wbase; wnextP = wnext; fmP = fm; AqP = Aq;
        const int nB = sB == 2 ? 0 : sB + 1;
        sC = sB; sB = nB;
        if (sB == 0) ++useB;
    }
}

// ------------------------------------------------------------------ host side
// Function attributes, occupancy and the SM count are per device: set / queried once for every device the library is used on
// (one process may drive several GPUs), under a mutex.
struct TilesDevice { bool ready = false; int sms = 0; int occ[2] = {0, 0}; };
static TilesDevice g_tiles_dev[64];
static std::mutex g_tiles_mu;
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
constexpr size_t kTilesDynSmem = (size_t)3 * kTile * sizeof(float) + sizeof(TScratch);

static int tiles_device(TilesDevice **out) {
    int dev = 0;
    DME_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return DME_ECUDA; }
    std::lock_guard<std::mutex> lock(g_tiles_mu);
    TilesDevice &D = g_tiles_dev[dev];
    if (!D.ready) {
        DME_CUDA(cudaFuncSetAttribute(quantize_tiles_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTilesDynSmem));
        DME_CUDA(cudaFuncSetAttribute(quantize_tiles_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTilesDynSmem));
        DME_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.occ[0], quantize_tiles_kernel<0>, kThreads2, kTilesDynSmem));
        DME_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.occ[1], quantize_tiles_kernel<1>, kThreads2, kTilesDynSmem));
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
int launch_quantize_tiles(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, const WsLayout &L, void *ws,
                          int32_t *k_out, uint8_t *sgn_out, float *deq_out, int64_t ld_out,
                          uint32_t *codes, int64_t codes_bytes, uint64_t *dir, cudaStream_t st, bool packed) {
    TilesDevice *D = nullptr;
    int rc = tiles_device(&D);
    if (rc) return rc;
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
    a.k_out = k_out; a.sgn_out = sgn_out; a.deq_out = deq_out; a.ld_out = ld_out;
    a.pack.codes = codes; a.pack.codes_bytes = codes_bytes; a.pack.dir = dir; a.pack.hdr = a.hdr; a.pack.n = n; a.pack.T = L.T;
    a.pack.W0 = expected_width(m > 0 ? m : 1, d);
    const int64_t nT = n * L.T;
    a.pack.arena_base16 = (unsigned long long)nT * 32ull * (unsigned long long)a.pack.W0;
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
        const cuuint32_t box[3] = {32, kTile / 32, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = g_encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)X, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for n=%lld d=%lld ld=%lld", (int)r, (long long)n, (long long)d, (long long)ld); return DME_ECUDA; }
    }
    a.tiles_tma = (int)((a.rows32 + kTile / 32 - 1) / (kTile / 32));
    a.has_tail = (d & 31) ? 1 : 0;
    const int occ = D->occ[packed ? 1 : 0];
    if (occ < 1) { set_error("quantize_tiles_kernel does not fit on an SM"); return DME_ECUDA; }
    if (nT >= ((int64_t)1 << 32) - 65536) { set_error("n * tiles = %lld does not fit the 32-bit ticket counter", (long long)nT); return DME_EINVAL; }
    int64_t G = (int64_t)D->sms * occ;             // every CTA resident: a look-back never waits on a CTA that has not started
    if (G > nT) G = nT;
    if (packed) quantize_tiles_kernel<1><<<(unsigned)G, kThreads2, kTilesDynSmem, st>>>(a, tmap);
    else quantize_tiles_kernel<0><<<(unsigned)G, kThreads2, kTilesDynSmem, st>>>(a, tmap);
    DME_LAUNCH_CHECK("quantize_tiles_kernel");
    return DME_OK;
}

}  // namespace dme
