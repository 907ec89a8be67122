// type_quantize.cuh -- data structures shared by the type-quantizer kernels (unbiased scan, Reznik select,
// packed-code emit, decode-mean).
#pragma once
#include "common.cuh"

namespace dme {

// Per-client constants, written by the kernel that finishes the client's L1 reduction.
struct __align__(64) RowConst {
    float L1f;   // fp32(sum |x|)  (AS:624) or injected
    float D;     // L1f + 1e-12f   (AS:625)
    float mf;    // float(m)
    float X;     // the client's uniform (AS:634)
    float rcpD;  // RN(1 / D) for the Markstein division  x/D = fma(fma(-q0, D, x), rcp, q0), q0 = x*rcp
    uint32_t flags;      // kRowExact: use IEEE div / floorf (operands outside the proven range of the fast chain)
                         // kRowGuardFloor: mp may reach 2^23 -> per-thread check before the magic floor
    int32_t qshift;      // tile aggregates are exchanged as int64 fixed point with 2^-qshift resolution
    uint32_t pad0;
    double q_up, q_dn;   // 2^qshift, 2^-qshift
    double pad1[2];
};
constexpr uint32_t kRowExact = 1u, kRowGuardFloor = 2u;

// Decoupled look-back record of one (client, tile).  state: 0 = nothing, 1 = aggregate valid,
// 2 = aggregate + inclusive valid.  a_state: 1 = a_last valid.
struct __align__(32) TileDesc {
    long long aggregate;   // sum of fractional parts inside the tile, fixed point (exact associativity =>
    long long inclusive;   // the look-back result does not depend on timing); inclusive = sum of aggregates 0..t
    int a_last;         // floor(c - X) of the tile's last coordinate
    uint32_t state;
    uint32_t a_state;
    uint32_t pad;
};

struct __align__(256) WsHeader {
    uint32_t ticket;          // scan-order ticket dispenser
    uint32_t status;          // sticky error bits: 1 = magnitude overflow, 2 = arena exhausted
    unsigned long long arena_top;   // bump pointer of the code arena, in 16-byte units
    uint32_t ticket2;         // second dispenser (Reznik passes)
    uint32_t pad[59];
};

// Closed form of AS:636 inside one binade of the fp32 prefix (quantize_tiles.cu): for c32 in [2^e + 1, 2^(e+1)),
// floor(RN32(RN32(c) - X)) = floor(c - Xp) when sigma = +1, ceil(c - Xp) - 1 when sigma = -1; sigma = 0: no closed form.
struct __align__(16) BinadeEntry { double Xp; double sigma; };
constexpr int kBinades = 24;

struct WsLayout {
    int64_t T;            // tiles per row
    int64_t off_done;     // uint32 a_done[n]
    int64_t off_ready;    // uint32 row_ready[n]
    int64_t off_consts;   // RowConst consts[n]
    int64_t off_partial;  // double partial[n*T]
    int64_t off_desc;     // look-back records: 16 bytes per tile, per block of 32 tiles, per super-block of 1024 tiles
    int64_t desc_bytes;
    int64_t off_sel;      // RowSelect sel[n] (biased mode)
    int64_t off_tab;      // BinadeEntry tab[n][kBinades] (quantize_tiles_kernel: closed-form floor(c - X) per binade)
    int64_t zero_bytes;   // prefix that must be zeroed before each call (header + a_done)
    int64_t total;
};
inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// Radix-select state of one client row (biased / Reznik mode).
struct __align__(16) RowSelect {
    unsigned long long mprime;   // sum of round-to-nearest k'
    long long Delta;             // m' - m
    uint32_t prefix;             // selected key prefix so far
    uint32_t remaining;          // how many still to pick among keys matching the prefix
    uint32_t tie_key;            // final threshold key
    uint32_t tie_take;           // how many threshold-equal elements (lowest index first) are adjusted
    uint32_t hist[4][256];       // per-pass histograms
    uint32_t done[4];            // tiles finished per pass
    uint32_t tie_seen;           // running count for the ordered tie scan
    uint32_t pad[3];
};

// Field width the packed code is expected to need for rate m/d on light-tailed data (largest magnitude in a
// 4096-tile ~ 6 m/d + 1).  Tiles up to this width use fixed primary slots of the arena; wider tiles overflow.
inline int expected_width(int64_t m, int64_t d) {
    const double ell = (double)m / (double)d;
    int w = 2;
    while (w < 32 && 4.0 * ell + 1.0 >= (double)(1u << (w - 1))) w <<= 1;
    return w;
}

inline WsLayout ws_layout(int64_t n, int64_t d) {
    WsLayout L;
    L.T = (d + kTile - 1) / kTile;
    int64_t o = (int64_t)sizeof(WsHeader);
    L.off_done = o; o = align_up(o + 4 * n, 256);
    L.off_ready = o; o = align_up(o + 4 * n, 256);
    L.zero_bytes = o;
    L.off_consts = o; o = align_up(o + (int64_t)sizeof(RowConst) * n, 256);
    L.off_partial = o; o = align_up(o + 8 * n * L.T, 256);
    {
        const int64_t TB = (L.T + 31) / 32, TS = (TB + 31) / 32;
        L.desc_bytes = 16 * n * (L.T + TB + TS);
    }
    L.off_desc = o; o = align_up(o + L.desc_bytes, 256);
    L.off_sel = o; o = align_up(o + (int64_t)sizeof(RowSelect) * n, 256);
    L.off_tab = o; o = align_up(o + (int64_t)sizeof(BinadeEntry) * kBinades * n, 256);
    L.total = o;
    return L;
}

#ifdef __CUDACC__
// Row constants + the binade table of AS:636's closed form (cold: once per client row).
struct RowConstIn { int64_t m, d; const float *x_inject, *l1_inject; uint64_t seed, client0; RowConst *consts; BinadeEntry *tabs; float *l1_out; };
__device__ inline void make_row_const(const RowConstIn &a, int64_t c, double l1sum) {
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
    if (a.l1_out) a.l1_out[c] = rc.L1f;
}
// Entry e of the row's binade table (closed form of AS:636): threads 0 .. kBinades-1 of the finishing CTA, one entry each.
__device__ inline void make_binade_entry(const RowConstIn &a, int64_t c, int e) {
    const RowConst &rc = a.consts[c];
    BinadeEntry b; b.Xp = 0.0; b.sigma = 0.0;
    if (!(rc.flags & kRowExact) && e >= 2 && e <= 22) {
        const double Xd = (double)rc.X;
        const double g = __longlong_as_double((long long)(1023 + e - 23) << 52), ginv = __longlong_as_double((long long)(1023 + 23 - e) << 52);
        const double av = ceil(Xd * ginv - 0.5);                 // exact: X has 24 bits, X >= 2^-24 or X == 0
        b.sigma = (((long long)av) & 1) ? -1.0 : 1.0;
        b.Xp = -b.sigma * (g * (av - 0.5));                      // stored as -sigma * Xp: sigma (c - Xp) = fma(c, sigma, b.Xp)
    }
    a.tabs[c * kBinades + e] = b;
}

// ------------------------------------------------------------------ packed-code emit shared by both quantizer modes
struct PackTarget { uint32_t *codes; int64_t codes_bytes; uint64_t *dir; WsHeader *hdr; int W0; unsigned long long arena_base16; int64_t n, T; };
// Primary slots are laid out TILE-major (tile t of all n clients is one contiguous run of n * 512 * W0 bytes): the
// decoder walks a tile's clients in order, so its reads are long sequential runs instead of 1 KB pieces 4 MB apart.
__device__ __forceinline__ unsigned long long primary_off16(const PackTarget &p, int64_t c, int64_t t) {
    return (unsigned long long)(t * p.n + c) * (32ull * (unsigned long long)p.W0);
}
struct PackScratch { uint32_t u32[kWarps]; unsigned long long off16; };

// Pack 16 (magnitude, sign) pairs of one thread with field width W into W/2 words.
template <int W>
__device__ __forceinline__ void pack_store(const uint32_t (&k)[kEpt], const uint32_t (&sg)[kEpt], uint32_t *tile_words, int chunk) {
    constexpr int kPerWord = 32 / W;
#pragma unroll
    for (int q = 0; q < W / 2; ++q) {
        uint32_t word = 0;
#pragma unroll
        for (int e = 0; e < kPerWord; ++e) {
            const int j = q * kPerWord + e;
            const uint32_t field = (W == 32) ? ((sg[j] << 31) | k[j]) : ((sg[j] << (W - 1)) | k[j]);
            word |= field << ((W * e) & 31);
        }
        tile_words[q * kThreads + chunk] = word;
    }
}
template <int W>
__device__ __forceinline__ void pack_store(const uint32_t (&k)[kEpt], const uint32_t (&sg)[kEpt], uint32_t *tile_words) {
    pack_store<W>(k, sg, tile_words, (int)threadIdx.x);       // one chunk per thread
}

// Whole CTA: choose the tile's minimal field width, place it (fixed primary slot when W <= W0, bump-allocated
// overflow space otherwise), write the directory entry and the words.  slot = client * T + tile.
__device__ __forceinline__ void emit_packed_tile(const PackTarget &p, int64_t slot, const uint32_t (&k)[kEpt],
                                                 const uint32_t (&sg)[kEpt], bool ovf, PackScratch &ps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t kmax = 0;
#pragma unroll
    for (int j = 0; j < kEpt; ++j) kmax = max(kmax, k[j]);
    if (ovf) atomicOr(&p.hdr->status, 1u);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0) ps.u32[warp] = kmax;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kWarps; ++w) kmax = max(kmax, ps.u32[w]);
    int W = 2;
    while (W < 32 && kmax >= (1u << (W - 1))) W <<= 1;
    unsigned long long off16;
    if (W <= p.W0) {
        off16 = primary_off16(p, slot / p.T, slot % p.T);
        if (threadIdx.x == 0) p.dir[slot] = (off16 << 8) | (unsigned long long)W;
    } else {
        if (threadIdx.x == 0) {
            const unsigned long long units = 32ull * W;                  // 512*W bytes / 16
            unsigned long long off = p.arena_base16 + atomicAdd(&p.hdr->arena_top, units);
            if ((long long)((off + units) * 16ull) > p.codes_bytes) { atomicOr(&p.hdr->status, 2u); off = ~0ull; }
            ps.off16 = off;
            p.dir[slot] = (off == ~0ull) ? 0ull : ((off << 8) | (unsigned long long)W);
        }
        __syncthreads();
        off16 = ps.off16;
    }
    if (off16 != ~0ull) {
        uint32_t *tw = p.codes + off16 * 4ull;
        switch (W) {
            case 2: pack_store<2>(k, sg, tw); break;
            case 4: pack_store<4>(k, sg, tw); break;
            case 8: pack_store<8>(k, sg, tw); break;
            case 16: pack_store<16>(k, sg, tw); break;
            default: pack_store<32>(k, sg, tw); break;
        }
    }
    __syncthreads();        // ps is reused by the caller's next tile
}
#endif  // __CUDACC__

}  // namespace dme
