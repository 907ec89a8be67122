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

struct __align__(256) WsHeader {
    uint32_t ticket;          // scan-order ticket dispenser
    uint32_t status;          // sticky error bits: 1 = magnitude overflow, 2 = arena exhausted
    unsigned long long arena_top;   // bump pointer of the code arena, in 16-byte units
    uint32_t ticket2;         // second dispenser (Reznik passes)
    uint32_t pad[59];
};

struct WsLayout {
    int64_t T;            // tiles per row
    int64_t off_done;     // uint32 a_done[n]
    int64_t off_consts;   // RowConst consts[n]
    int64_t off_lut;      // float lut[n][16]: the decoder's per-client value tables (+-q(k) / n_total, k < 8), fused call only
    int64_t off_partial;  // double partial[n*T]
    int64_t off_desc;     // look-back records of quantize_warp_kernel: 8 bytes per code tile, 8 per block of 32 tiles, 16 per
    int64_t desc_bytes;   // super-block of 1024 tiles
    int64_t off_sel;      // RowSelect sel[n] (biased mode)
    int64_t off_lin;      // biased mode, linear selection: uint32 hist[n][lin_bins], then uint2 cand[n][lin_cap]
    int64_t lin_hist_bytes, lin_bytes;
    int64_t lin_bins, lin_cap;
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
    // linear-histogram selection (reznik.cu, the default path)
    int32_t lin_bstar;           // threshold bin: bins beyond it are adjusted entirely
    uint32_t lin_need;           // how many of the threshold bin's coordinates are adjusted
    uint32_t lin_ncand;          // coordinates of the threshold bin appended to the candidate list so far
    uint32_t lin_cut;            // threshold-equal candidates with index < lin_cut are adjusted
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

// Linear selection of the biased mode: bins of the residual histogram (about 8 coordinates per bin for short rows, at most 2048)
// and capacity of a row's candidate list (the threshold bin holds d / bins coordinates on average: eight times that, plus slack).
constexpr int kLinBinsMax = 2048;
inline int64_t lin_bins_for(int64_t d) { int64_t b = 64; while (b < kLinBinsMax && b * 8 < d) b <<= 1; return b; }
inline int64_t lin_cap_for(int64_t d) { const int64_t c = 8 * (d / lin_bins_for(d)) + 512; return c < d ? c : d; }

inline WsLayout ws_layout(int64_t n, int64_t d) {
    WsLayout L;
    L.T = (d + kTile - 1) / kTile;
    int64_t o = (int64_t)sizeof(WsHeader);
    L.off_done = o; o = align_up(o + 4 * n, 256);
    L.zero_bytes = o;
    L.off_consts = o; o = align_up(o + (int64_t)sizeof(RowConst) * n, 256);
    L.off_lut = o; o = align_up(o + 64 * n, 256);
    L.off_partial = o; o = align_up(o + 8 * n * L.T, 256);
    {
        const int64_t T4 = (d + kCodeTile - 1) / kCodeTile, TB = (T4 + 31) / 32, TS = (TB + 31) / 32;
        L.desc_bytes = n * (8 * T4 + 8 * TB + 16 * TS) + 16;
    }
    L.off_desc = o; o = align_up(o + L.desc_bytes, 256);
    L.off_sel = o; o = align_up(o + (int64_t)sizeof(RowSelect) * n, 256);
    L.lin_bins = lin_bins_for(d);
    L.lin_cap = lin_cap_for(d);
    L.lin_hist_bytes = align_up(n * L.lin_bins * 4, 256);
    L.lin_bytes = L.lin_hist_bytes + n * L.lin_cap * 8;
    L.off_lin = o; o = align_up(o + L.lin_bytes, 256);
    L.total = o;
    return L;
}

#ifdef __CUDACC__
// Row constants + the binade table of AS:636's closed form (cold: once per client row).
struct RowConstIn { int64_t m, d; const float *x_inject, *l1_inject; uint64_t seed, client0; RowConst *consts; float *l1_out; };
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
    rc.qshift = min(43, 62 - lg);                                   // a block of 32 tile aggregates + its count fit one 64-bit word
    rc.pad0 = 0;
    rc.q_up = __longlong_as_double((long long)(1023 + rc.qshift) << 52);
    rc.q_dn = __longlong_as_double((long long)(1023 - rc.qshift) << 52);
    rc.pad1[0] = rc.pad1[1] = 0.0;
    a.consts[c] = rc;
    if (a.l1_out) a.l1_out[c] = rc.L1f;
}
// ------------------------------------------------------------------ packed-code emit shared by both quantizer modes
// Code tiles are kCodeTile = 1024 coordinates (64 chunks of 16); T4 = code tiles per client row.
struct PackTarget { uint32_t *codes; int64_t codes_bytes; uint64_t *dir; WsHeader *hdr; int W0; unsigned long long arena_base16; int64_t n, T4; };
// Primary slots are laid out TILE-major (code tile t of all n clients is one contiguous run of n * 128 * W0 bytes): the
// decoder walks a tile's clients in order, so its reads are long sequential runs.
__device__ __forceinline__ unsigned long long primary_off16(const PackTarget &p, int64_t c, int64_t t4) {
    return (unsigned long long)(t4 * p.n + c) * (8ull * (unsigned long long)p.W0);
}
struct PackScratch { uint32_t u32[kWarps]; unsigned long long off16[kTile / kCodeTile]; };

// Pack 16 (magnitude, sign) pairs of one thread with field width W into W/2 words of chunk `chunk` (0..63) of a code tile.
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
        tile_words[q * kCodeChunks + chunk] = word;
    }
}
__device__ __forceinline__ void pack_store_w(int W, const uint32_t (&k)[kEpt], const uint32_t (&sg)[kEpt], uint32_t *tw, int chunk) {
    switch (W) {
        case 2: pack_store<2>(k, sg, tw, chunk); break;
        case 4: pack_store<4>(k, sg, tw, chunk); break;
        case 8: pack_store<8>(k, sg, tw, chunk); break;
        case 16: pack_store<16>(k, sg, tw, chunk); break;
        default: pack_store<32>(k, sg, tw, chunk); break;
    }
}
// Place one code tile of width W: the fixed primary slot when W <= W0, bump-allocated overflow space otherwise; writes the
// directory entry.  One thread per code tile calls it.  ~0 = the arena is exhausted (status bit 2 set, entry 0).
__device__ __forceinline__ unsigned long long place_code_tile(const PackTarget &p, int64_t c, int64_t t4, int W) {
    unsigned long long off16;
    if (W <= p.W0) off16 = primary_off16(p, c, t4);
    else {
        const unsigned long long units = 8ull * W;                  // 128 * W bytes / 16
        off16 = p.arena_base16 + atomicAdd(&p.hdr->arena_top, units);
        if ((long long)((off16 + units) * 16ull) > p.codes_bytes) { atomicOr(&p.hdr->status, 2u); off16 = ~0ull; }
    }
    p.dir[c * p.T4 + t4] = (off16 == ~0ull) ? 0ull : ((off16 << 8) | (unsigned long long)W);
    return off16;
}

// Whole CTA (kThreads threads, thread t owns coordinates [16t, 16t + 16) of CTA tile `t` of client c = four code tiles of
// 64 threads each): every code tile gets its minimal field width, is placed and written.
__device__ __forceinline__ void emit_packed_tile(const PackTarget &p, int64_t c, int64_t t, const uint32_t (&k)[kEpt],
                                                 const uint32_t (&sg)[kEpt], bool ovf, PackScratch &ps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int quarter = threadIdx.x >> 6, chunk = threadIdx.x & (kCodeChunks - 1);
    uint32_t kmax = 0;
#pragma unroll
    for (int j = 0; j < kEpt; ++j) kmax = max(kmax, k[j]);
    if (ovf) atomicOr(&p.hdr->status, 1u);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0) ps.u32[warp] = kmax;
    __syncthreads();
    kmax = max(ps.u32[2 * quarter], ps.u32[2 * quarter + 1]);
    int W = 2;
    while (W < 32 && kmax >= (1u << (W - 1))) W <<= 1;
    const int64_t t4 = t * (kTile / kCodeTile) + quarter;
    const bool live = t4 < p.T4;                                     // the row ends before this quarter of the CTA tile
    if (chunk == 0 && live) ps.off16[quarter] = place_code_tile(p, c, t4, W);
    __syncthreads();
    if (live) {
        const unsigned long long off16 = ps.off16[quarter];
        if (off16 != ~0ull) pack_store_w(W, k, sg, p.codes + off16 * 4ull, chunk);
    }
    __syncthreads();        // ps is reused by the caller's next tile
}
inline void init_pack_target(PackTarget &p, uint32_t *codes, int64_t codes_bytes, uint64_t *dir, WsHeader *hdr, int64_t n, int64_t d, int64_t m) {
    p.codes = codes; p.codes_bytes = codes_bytes; p.dir = dir; p.hdr = hdr; p.n = n;
    p.T4 = (d + kCodeTile - 1) / kCodeTile;
    p.W0 = expected_width(m > 0 ? m : 1, d);
    p.arena_base16 = (unsigned long long)(n * p.T4) * 8ull * (unsigned long long)p.W0;
}
#endif  // __CUDACC__

}  // namespace dme
