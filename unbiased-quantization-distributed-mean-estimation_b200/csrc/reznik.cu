// reznik.cu -- biased type quantizer: nearest-type rounding with mass repair (Reznik, AS:644-666) behind
// Type_biased_quantize (AS:669-687).
//
// The reference repairs the mass with torch.topk(delta, |Delta|); |Delta| is Theta(d), so this is an
// order-statistic problem, solved here without sorting:
//   rz_sum     : k' = floor(m p + 1/2), m' = sum k' (exact, integer atomics)             one pass
//   rz_hist x4 : MSB-first radix select of the |Delta|-th largest residual key            four passes, 8 bits each
//                (key = order-preserving uint32 of v, v = delta for Delta>0, -delta for Delta<0)
//   rz_tiecnt  : per-tile count of residuals equal to the threshold                        one pass
//   rz_apply   : k = k' -/+ 1 for keys above the threshold and for the FIRST tie_take threshold-equal
//                coordinates in index order (torch.topk leaves ties unspecified; documented deviation), emit
// Everything that decides the result is integer arithmetic, so the outcome is independent of scheduling.
#include "type_quantize.cuh"

namespace dme {

__device__ __forceinline__ void load_tile_blocked(const float *__restrict__ row, int64_t d, int64_t tile0, float (&x)[kEpt]) {
    const int64_t i0 = tile0 + (int64_t)threadIdx.x * kEpt;
    if (i0 + kEpt <= d) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = *reinterpret_cast<const float4 *>(row + i0 + 4 * q);
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kEpt; ++j) x[j] = (i0 + j < d) ? row[i0 + j] : 0.0f;
    }
}

// striped: fully coalesced 128-bit loads for the passes where the order inside the tile does not matter (sums, histograms,
// counts); element e of load q of thread t is coordinate tile0 + 1024 q + 4 t + e (zero past the end of the row)
__device__ __forceinline__ void load_tile_striped(const float *__restrict__ row, int64_t d, int64_t tile0, float (&x)[kEpt]) {
    if (tile0 + kTile <= d) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = ldg_stream_f4(row + tile0 + q * 1024 + 4 * threadIdx.x);
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
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
__device__ __forceinline__ int64_t striped_index(int64_t tile0, int j) { return tile0 + (j >> 2) * 1024 + 4 * threadIdx.x + (j & 3); }

// AS:648 + AS:654 for one coordinate: k' and the residual delta' = k' - m p (fp32, one rounding per op).
// Rows inside the proven operand range of the fast chain (RowConst flags, type_quantize.cuh) get the same values from
// Markstein's correction of |x| * RN(1/D) (= the correctly rounded quotient) and a magic-number floor (exact below 2^23):
// 8 instructions instead of an IEEE division and a floorf per coordinate and pass.
template <bool FAST>
__device__ __forceinline__ void rz_round_t(float x, const RowConst &rc, float &kp, float &delta) {
    float mp;
    if (FAST) {
        const float ax = fabsf(x);
        const float q0 = __fmul_rn(ax, rc.rcpD);
        const float p = __fmaf_rn(__fmaf_rn(-q0, rc.D, ax), rc.rcpD, q0);          // AS:683
        mp = __fmul_rn(rc.mf, p);
        kp = __fsub_rn(__fadd_rz(__fadd_rn(mp, 0.5f), 8388608.0f), 8388608.0f);    // AS:648: floor(mp + 1/2), mp + 1/2 < 2^23
    } else {
        const float p = __fdiv_rn(fabsf(x), rc.D);             // AS:683
        mp = __fmul_rn(rc.mf, p);
        kp = floorf(__fadd_rn(mp, 0.5f));                      // AS:648
    }
    delta = __fsub_rn(kp, mp);                                 // AS:654
}
__device__ __forceinline__ bool rz_fast_row(const RowConst &rc) { return !(rc.flags & (kRowExact | kRowGuardFloor)); }
__device__ __forceinline__ void rz_round(float x, const RowConst &rc, float &kp, float &delta) {
    if (rz_fast_row(rc)) rz_round_t<true>(x, rc, kp, delta); else rz_round_t<false>(x, rc, kp, delta);
}
// Order-preserving key of the value that torch.topk ranks: delta (Delta > 0) or -delta (Delta < 0).
__device__ __forceinline__ uint32_t rz_key(float delta, bool neg) {
    float v = neg ? -delta : delta;
    v = __fadd_rn(v, 0.0f);                                    // -0 -> +0: equal floats must get equal keys
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(kThreads)
rz_sum_kernel(const float *__restrict__ X, int64_t d, int64_t ld, const RowConst *__restrict__ consts, RowSelect *sel) {
    __shared__ unsigned long long s_red[kWarps];
    const int64_t c = blockIdx.y, t = blockIdx.x;
    float x[kEpt];
    load_tile_striped(X + c * ld, d, t * kTile, x);            // issued first: the row constants' round trip overlaps with it
    const RowConst rc = consts[c];
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < kEpt; ++j) {
        float kp, dl;
        rz_round(x[j], rc, kp, dl);
        s += (unsigned long long)kp;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tot = 0;
        for (int w = 0; w < kWarps; ++w) tot += s_red[w];
        if (tot) atomicAdd(&sel[c].mprime, tot);
    }
}

__global__ void rz_init_kernel(RowSelect *sel, int64_t n, int64_t m, int64_t d) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const long long Delta = (long long)sel[c].mprime - (long long)m;       // AS:655
    sel[c].Delta = Delta;
    long long need = Delta < 0 ? -Delta : Delta;
    if (need > d) need = d;            // the reference's topk raises when |Delta| > d; here every coordinate is adjusted once
    sel[c].remaining = (uint32_t)need;
    sel[c].prefix = 0;
}

// One radix digit: histogram of byte `pass` (0 = most significant) over keys that match the prefix so far.
__global__ void __launch_bounds__(kThreads)
rz_hist_kernel(const float *__restrict__ X, int64_t d, int64_t ld, const RowConst *__restrict__ consts, RowSelect *sel, int pass) {
    __shared__ uint32_t s_hist[256];
    const int64_t c = blockIdx.y, t = blockIdx.x;
    float x[kEpt];
    load_tile_striped(X + c * ld, d, t * kTile, x);            // issued first: the selection state's round trips overlap with it
    const long long Delta = sel[c].Delta;
    if (Delta == 0 || sel[c].remaining == 0) return;
    const RowConst rc = consts[c];
    const uint32_t prefix = sel[c].prefix;
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int shift = 24 - 8 * pass;
    const bool neg = Delta < 0, full = (t + 1) * kTile <= d;
#pragma unroll
    for (int j = 0; j < kEpt; ++j) {
        if (!full && striped_index(t * kTile, j) >= d) continue;
        float kp, dl;
        rz_round(x[j], rc, kp, dl);
        const uint32_t key = rz_key(dl, neg);
        const bool match = pass == 0 ? true : ((key >> (shift + 8)) == prefix);
        if (match) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    const uint32_t v = s_hist[threadIdx.x];
    if (v) atomicAdd(&sel[c].hist[pass][threadIdx.x], v);
}

// Choose the digit: walk the bins from the top until the wanted rank falls inside one.
__global__ void rz_pick_kernel(RowSelect *sel, int pass) {
    RowSelect &s = sel[blockIdx.x];
    if (threadIdx.x != 0 || s.Delta == 0 || s.remaining == 0) return;
    uint32_t rem = s.remaining, cum = 0;
    int b = 255;
    for (; b > 0; --b) {
        const uint32_t h = s.hist[pass][b];
        if (cum + h >= rem) break;
        cum += h;
    }
    s.prefix = (s.prefix << 8) | (uint32_t)b;
    s.remaining = rem - cum;                     // keys in higher bins are all selected
    if (pass == 3) { s.tie_key = s.prefix; s.tie_take = s.remaining; }
}

__global__ void __launch_bounds__(kThreads)
rz_tiecnt_kernel(const float *__restrict__ X, int64_t d, int64_t ld, int64_t T, const RowConst *__restrict__ consts,
                 const RowSelect *__restrict__ sel, uint32_t *__restrict__ tie_cnt) {
    __shared__ uint32_t s_red[kWarps];
    const int64_t c = blockIdx.y, t = blockIdx.x;
    float x[kEpt];
    load_tile_striped(X + c * ld, d, t * kTile, x);
    const long long Delta = sel[c].Delta;
    uint32_t cnt = 0;
    if (Delta != 0) {
        const RowConst rc = consts[c];
        const uint32_t tie = sel[c].tie_key;
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            if (striped_index(t * kTile, j) >= d) continue;
            float kp, dl;
            rz_round(x[j], rc, kp, dl);
            cnt += (rz_key(dl, Delta < 0) == tie) ? 1u : 0u;
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < kWarps; ++w) tot += s_red[w];
        tie_cnt[c * T + t] = tot;
    }
}

// Exclusive prefix of the per-tile tie counts of one row (one CTA per row, in place).
__global__ void __launch_bounds__(kThreads) rz_tiescan_kernel(uint32_t *tie_cnt, int64_t T) {
    __shared__ uint32_t s_w[kWarps];
    __shared__ uint32_t s_carry;
    uint32_t *row = tie_cnt + (int64_t)blockIdx.x * T;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < T; base += kThreads) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = i < T ? row[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o) inc += up;
        }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += s_w[w];
        const uint32_t carry = s_carry;
        if (i < T) row[i] = carry + wbase + inc - v;
        __syncthreads();
        if (threadIdx.x == kThreads - 1) s_carry = carry + wbase + inc;
        __syncthreads();
    }
}

struct RzEmit {
    int32_t *k_out; uint8_t *sgn_out; float *deq_out; int64_t ld_out;
    PackTarget pack; int packed;
};

__global__ void __launch_bounds__(kThreads, 4)
rz_apply_kernel(const float *__restrict__ X, int64_t d, int64_t ld, int64_t T, const RowConst *__restrict__ consts,
                const RowSelect *__restrict__ sel, const uint32_t *__restrict__ tie_base, RzEmit e) {
    __shared__ uint32_t s_w[kWarps];
    __shared__ PackScratch s_pack;
    const int64_t c = blockIdx.y, t = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float x[kEpt];
    load_tile_blocked(X + c * ld, d, t * kTile, x);
    const RowConst rc = consts[c];
    const long long Delta = sel[c].Delta;
    const uint32_t tie = sel[c].tie_key, take = sel[c].tie_take;
    const int64_t i0 = t * kTile + (int64_t)threadIdx.x * kEpt;
    float kp[kEpt];
    uint32_t key[kEpt];
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < kEpt; ++j) {
        float dl;
        rz_round(x[j], rc, kp[j], dl);
        key[j] = rz_key(dl, Delta < 0);
        if (Delta != 0 && i0 + j < d && key[j] == tie) ++mine;
    }
    // ordered rank of this thread's first threshold-equal coordinate inside the row
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += up;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    uint32_t rank = (Delta != 0 ? tie_base[c * T + t] : 0u) + inc - mine;
    for (int w = 0; w < warp; ++w) rank += s_w[w];
    const float adj = Delta > 0 ? -1.0f : 1.0f;
#pragma unroll
    for (int j = 0; j < kEpt; ++j) {
        if (Delta == 0 || i0 + j >= d) continue;
        bool selct = key[j] > tie;
        if (key[j] == tie) { selct = rank < take; ++rank; }
        if (selct) kp[j] = __fadd_rn(kp[j], adj);                              // AS:660 / AS:664
    }
    if (!e.packed) {
        bool ovf = false;
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            const int64_t i = i0 + j;
            if (i >= d) break;
            if (e.deq_out) {
                const float sg = (x[j] > 0.0f) ? 1.0f : ((x[j] < 0.0f) ? -1.0f : 0.0f);     // AS:682 signs of the input
                e.deq_out[c * e.ld_out + i] = __fmul_rn(__fmul_rn(rc.L1f, sg), __fdiv_rn(kp[j], rc.mf));   // AS:666, AS:687
            }
            if (e.k_out) {
                if (kp[j] >= 2147483648.0f) { ovf = true; e.k_out[c * e.ld_out + i] = 0x7fffffff; }
                else e.k_out[c * e.ld_out + i] = (int32_t)kp[j];
            }
            if (e.sgn_out) e.sgn_out[c * e.ld_out + i] = (uint8_t)(__float_as_uint(x[j]) >> 31);
        }
        if (ovf) atomicOr(&e.pack.hdr->status, 1u);
    } else {
        uint32_t k[kEpt], sg[kEpt];
        bool ovf = false;
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            const float kk = (i0 + j < d) ? kp[j] : 0.0f;
            if (kk >= 2147483648.0f) { ovf = true; k[j] = 0x7fffffffu; } else k[j] = (uint32_t)kk;
            sg[j] = __float_as_uint(x[j]) >> 31;
        }
        emit_packed_tile(e.pack, c, t, k, sg, ovf, s_pack);
    }
}

// ================================================================== linear selection (default path)
// The residuals delta' = k' - m p lie in [-1/2, 1/2] and are spread evenly there, so ONE histogram over B linear bins
// (bin = floor((delta' + 1/2) B), monotone in delta') locates the |Delta|-th largest (smallest) residual up to a bin that
// holds about d / B coordinates.  Passes over the rows:
//   rz_sumhist : k', m' = sum k' and the bin histogram, a run of tiles per CTA (one flush of the bins per run)
//   rz_linpick : Delta, the threshold bin b* (bins beyond it are adjusted entirely) and how many of b*'s coordinates are needed
//   rz_compact : the coordinates of bin b* -> candidate list {exact order-preserving key, index} of the row
//   rz_linselect (one CTA per row, on the list): the threshold key, how many threshold-equal coordinates are needed, and the
//                index below which they are taken (ties: lowest index first, as in the radix path)
//   rz_linapply: k = k' -/+ 1 for the selected coordinates, emit
// A row whose threshold bin does not fit the list (many equal residuals: a constant or mostly-zero row with a large |Delta|)
// sets status bit 4; the host reruns the call on the radix path (dme_set_biased_path(1)).
__device__ __forceinline__ int lin_bin(float delta, float Bf, float Bm1) {          // Bf = float(B), Bm1 = float(B - 1)
    float t = __fmul_rn(__fadd_rn(delta, 0.5f), Bf);
    t = fminf(fmaxf(t, 0.0f), Bm1);                            // NaN -> 0
    return (int)t;
}

__global__ void __launch_bounds__(kThreads)
rz_sumhist_kernel(const float *__restrict__ X, int64_t d, int64_t ld, int64_t T, int tiles_per_cta, int B, const RowConst *__restrict__ consts,
                  RowSelect *sel, uint32_t *__restrict__ hist) {
    extern __shared__ uint32_t s_lin[];                        // B bins
    __shared__ unsigned long long s_red[kWarps];
    const int64_t c = blockIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * tiles_per_cta, t1 = t0 + tiles_per_cta < T ? t0 + tiles_per_cta : T;
    for (int b = threadIdx.x; b < B + 2; b += kThreads) s_lin[b] = 0;
    const RowConst rc = consts[c];
    __syncthreads();
    unsigned long long s = 0;
    const float Bf = (float)B, Bm1 = (float)(B - 1);
    const bool small_k = rc.mf <= 16777216.0f && !(rc.flags & kRowGuardFloor);      // k' <= m p + 1/2 < 2^25: sixteen of them fit 32 bits
    const bool fast = rz_fast_row(rc);
    for (int64_t t = t0; t < t1; ++t) {
        float x[kEpt];
        load_tile_striped(X + c * ld, d, t * kTile, x);
        if ((t + 1) * kTile <= d && small_k && fast) {
            // fast rows hold no NaN and 0 <= (delta' + 1/2) B <= B (+ one rounding): no clamp -- the rare bin B (delta' = 1/2) has
            // its own counter and is folded into bin B - 1 at the flush
            uint32_t s32 = 0;
#pragma unroll
            for (int j = 0; j < kEpt; ++j) {
                float kp, dl;
                rz_round_t<true>(x[j], rc, kp, dl);
                s32 += (uint32_t)kp;
                atomicAdd(&s_lin[(int)__fmul_rn(__fadd_rn(dl, 0.5f), Bf)], 1u);
            }
            s += s32;
        } else if ((t + 1) * kTile <= d && small_k) {
            uint32_t s32 = 0;
#pragma unroll
            for (int j = 0; j < kEpt; ++j) {
                float kp, dl;
                rz_round(x[j], rc, kp, dl);
                s32 += (uint32_t)kp;
                atomicAdd(&s_lin[lin_bin(dl, Bf, Bm1)], 1u);
            }
            s += s32;
        } else {
#pragma unroll
            for (int j = 0; j < kEpt; ++j) {
                if (striped_index(t * kTile, j) >= d) continue;
                float kp, dl;
                rz_round(x[j], rc, kp, dl);
                s += (unsigned long long)kp;
                atomicAdd(&s_lin[lin_bin(dl, Bf, Bm1)], 1u);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tot = 0;
        for (int w = 0; w < kWarps; ++w) tot += s_red[w];
        if (tot) atomicAdd(&sel[c].mprime, tot);
    }
    for (int b = threadIdx.x; b < B; b += kThreads) {
        const uint32_t v = s_lin[b] + (b == B - 1 ? s_lin[B] + s_lin[B + 1] : 0u);
        if (v) atomicAdd(&hist[c * B + b], v);
    }
}

// One warp per row: Delta, then the walk over the bins from the end the adjustment starts at (position q of the walk is bin
// B - 1 - q for Delta > 0, bin q for Delta < 0): every lane sums a run of B / 32 positions, a warp scan finds the run in which
// the wanted rank falls, that lane walks its run.
__global__ void rz_linpick_kernel(RowSelect *sel, const uint32_t *__restrict__ hist, int64_t n, int64_t m, int64_t d, int B) {
    const int64_t c = blockIdx.x;
    const int lane = threadIdx.x;
    RowSelect &s = sel[c];
    const long long Delta = (long long)s.mprime - (long long)m;            // AS:655
    long long need = Delta < 0 ? -Delta : Delta;
    if (need > d) need = d;            // the reference's topk raises when |Delta| > d; here every coordinate is adjusted once
    if (lane == 0) { s.Delta = Delta; s.lin_ncand = 0; s.lin_cut = 0; s.tie_key = 0; s.tie_take = 0; }
    if (need == 0) {                   // Delta == 0: nothing is adjusted
        if (lane == 0) { s.lin_bstar = -1; s.lin_need = 0; }
        return;
    }
    const uint32_t *h = hist + c * B;
    const int per = B / 32;            // B is a power of two >= 64
    const bool down = Delta > 0;
    unsigned long long run = 0;
    for (int q = lane * per; q < (lane + 1) * per; ++q) run += h[down ? B - 1 - q : q];
    unsigned long long incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const uint32_t hit = __ballot_sync(0xffffffffu, incl >= (unsigned long long)need);      // non-empty: the total is d >= need
    if (hit == 0u || lane != __ffs(hit) - 1) return;
    unsigned long long cum = incl - run;
    int q = lane * per;
    for (; q < (lane + 1) * per - 1; ++q) {
        const uint32_t v = h[down ? B - 1 - q : q];
        if (cum + v >= (unsigned long long)need) break;
        cum += v;
    }
    s.lin_bstar = down ? B - 1 - q : q;
    s.lin_need = (uint32_t)((unsigned long long)need - cum);               // >= 1, <= the threshold bin's count
}

__global__ void __launch_bounds__(kThreads)
rz_compact_kernel(const float *__restrict__ X, int64_t d, int64_t ld, int64_t T, int tiles_per_cta, int B, int64_t cap,
                  const RowConst *__restrict__ consts, RowSelect *sel, uint2 *__restrict__ cand, WsHeader *hdr) {
    const int64_t c = blockIdx.y;
    const long long Delta = sel[c].Delta;
    if (Delta == 0) return;
    const int bstar = sel[c].lin_bstar;
    const RowConst rc = consts[c];
    const bool neg = Delta < 0;
    const int64_t t0 = (int64_t)blockIdx.x * tiles_per_cta, t1 = t0 + tiles_per_cta < T ? t0 + tiles_per_cta : T;
    const float Bf = (float)B, Bm1 = (float)(B - 1);
    const bool fast = rz_fast_row(rc);
    const float blo = (float)bstar, bhi = bstar >= B - 1 ? __int_as_float(0x7f800000) : (float)(bstar + 1);
    for (int64_t t = t0; t < t1; ++t) {
        float x[kEpt];
        load_tile_striped(X + c * ld, d, t * kTile, x);
        const bool full = (t + 1) * kTile <= d;
        uint32_t hits = 0;                                     // bit j: coordinate j of this thread lies in the threshold bin
        float dls[kEpt];
        if (fast) {
            // bin == b* <=> b* <= t < b* + 1 with t = (delta' + 1/2) B (b* = B - 1 takes the clamped t >= B as well)
#pragma unroll
            for (int j = 0; j < kEpt; ++j) {
                float kp;
                rz_round_t<true>(x[j], rc, kp, dls[j]);
                const float tb = __fmul_rn(__fadd_rn(dls[j], 0.5f), Bf);
                hits |= ((tb >= blo && tb < bhi) ? 1u : 0u) << j;
            }
        } else {
#pragma unroll
            for (int j = 0; j < kEpt; ++j) {
                float kp;
                rz_round_t<false>(x[j], rc, kp, dls[j]);
                hits |= (lin_bin(dls[j], Bf, Bm1) == bstar ? 1u : 0u) << j;
            }
        }
        if (hits == 0u) continue;                              // the threshold bin holds one coordinate in B
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            if (!((hits >> j) & 1u)) continue;
            const int64_t i = striped_index(t * kTile, j);
            if (!full && i >= d) continue;
            const uint32_t pos = atomicAdd(&sel[c].lin_ncand, 1u);
            if ((int64_t)pos < cap) cand[c * cap + pos] = make_uint2(rz_key(dls[j], neg), (uint32_t)i);
            else atomicOr(&hdr->status, 4u);
        }
    }
}

// CTA-wide radix select on a list in global memory: the `want`-th largest (descending) / smallest (!descending) value of field
// .x (keys) or .y (indices) among the entries that pass the filter (key == fkey when filter_on).  Returns the value and how many
// entries lie strictly beyond it in the chosen direction.  want >= 1 and <= the number of entries that pass.
__device__ uint32_t cta_select(const uint2 *list, uint32_t n, uint32_t want, bool descending, bool field_y, bool filter_on, uint32_t fkey,
                               uint32_t *s_hist /* 256 */, uint32_t *s_misc /* 2 */, uint32_t &beyond) {
    uint32_t prefix = 0, rem = want, passed = 0;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        s_hist[threadIdx.x] = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += kThreads) {
            const uint2 e = list[i];
            if (filter_on && e.x != fkey) continue;
            const uint32_t v = field_y ? e.y : e.x;
            if (pass == 0 || (v >> (shift + 8)) == prefix) atomicAdd(&s_hist[(v >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t cum = 0;
            int b = descending ? 255 : 0;
            for (;;) {
                const uint32_t h = s_hist[b];
                if (cum + h >= rem) break;
                cum += h;
                b += descending ? -1 : 1;
                if (b < 0 || b > 255) { b = descending ? 0 : 255; break; }      // cannot happen for a valid `want`
            }
            s_misc[0] = (uint32_t)b; s_misc[1] = cum;
        }
        __syncthreads();
        prefix = (prefix << 8) | s_misc[0];
        passed += s_misc[1];
        rem -= s_misc[1];
        __syncthreads();
    }
    beyond = passed;
    return prefix;
}

__global__ void __launch_bounds__(kThreads)
rz_linselect_kernel(RowSelect *sel, const uint2 *__restrict__ cand, int64_t cap) {
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_misc[2];
    const int64_t c = blockIdx.x;
    RowSelect &s = sel[c];
    if (s.Delta == 0 || s.lin_need == 0) return;
    const uint32_t n = s.lin_ncand < (uint32_t)cap ? s.lin_ncand : (uint32_t)cap;
    if (s.lin_ncand > (uint32_t)cap || s.lin_need > n) return;             // overflow: status bit 4 is set, the host reruns the call
    const uint2 *list = cand + c * cap;
    uint32_t above = 0, before = 0;
    const uint32_t tie = cta_select(list, n, s.lin_need, true, false, false, 0u, s_hist, s_misc, above);
    const uint32_t take = s.lin_need - above;                               // >= 1 threshold-equal candidates are adjusted
    const uint32_t cut = cta_select(list, n, take, false, true, true, tie, s_hist, s_misc, before);      // the take-th smallest index
    if (threadIdx.x == 0) { s.tie_key = tie; s.tie_take = take; s.lin_cut = cut + 1u; }
}

// FAST: the row is inside the proven operand range of the fast chain and k' < 2^25 (no overflow test); FULL: the tile lies inside
// the row (no index tests).  One instantiation per combination: the per-coordinate loops carry no row- or tile-uniform branches.
// The largest float below a positive integer-valued float.
__device__ __forceinline__ float pred_pos(float v) { return __uint_as_float(__float_as_uint(v) - 1u); }      // v >= 1
template <bool FAST, bool FULL>
__device__ __forceinline__ void rz_linapply_tile(const float (&x)[kEpt], const RowConst &rc, long long Delta, int bstar, uint32_t tie, uint32_t cut,
                                                 int B, int64_t d, int64_t c, int64_t t, const RzEmit &e, PackScratch &s_pack) {
    const int64_t i0 = t * kTile + (int64_t)threadIdx.x * kEpt;
    const bool neg = Delta < 0;
    const float adj = Delta > 0 ? -1.0f : 1.0f;
    const float Bf = (float)B, Bm1 = (float)(B - 1);
    const int sdir = neg ? -1 : 1;                             // bins beyond the threshold bin: (b - b*) * sdir > 0
    // coordinates past the end of the row read as zero: k' = 0, and they are never written
    float kp[kEpt];
    uint32_t rare = 0;                                         // bit j: coordinate j lies in the threshold bin (one in B does)
    const float adj0 = Delta != 0 ? adj : 0.0f;
    if (FAST) {
        // No NaN in a fast row, and t = (delta' + 1/2) B >= 0: bin = floor(min(t, B - 1)), so "bin beyond b*" and "bin == b*" are two
        // comparisons of u = +-t (the sign of the walk folded into B) with constants -- no clamp, no conversion, no integer
        // arithmetic.  Delta > 0: beyond <=> t >= b* + 1 (never for b* = B - 1), in or beyond <=> t >= b*.  Delta < 0: beyond <=>
        // t < b* <=> -t >= -pred(b*), in or beyond <=> t < b* + 1 <=> -t >= -pred(b* + 1) (always for b* = B - 1: the clamp).
        const float inf = __int_as_float(0x7f800000);
        const float lo = (float)bstar, hi = bstar >= B - 1 ? inf : (float)(bstar + 1);      // bins are < 2^11: exact
        const float sB = neg ? -Bf : Bf;
        const float A = neg ? (bstar <= 0 ? inf : -pred_pos(lo)) : hi;      // u >= A: beyond the threshold bin (b* = 0, Delta < 0: never)
        const float C = neg ? (bstar >= B - 1 ? -inf : -pred_pos(hi)) : lo;      // u >= C: in the threshold bin or beyond it
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            float dl;
            rz_round_t<true>(x[j], rc, kp[j], dl);
            const float u = __fmul_rn(__fadd_rn(dl, 0.5f), sB);
            const bool beyond = u >= A;
            rare |= ((u >= C && !beyond) ? 1u : 0u) << j;
            kp[j] = __fadd_rn(kp[j], beyond ? adj0 : 0.0f);                // AS:660 / AS:664
        }
    } else {
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            float dl;
            rz_round_t<FAST>(x[j], rc, kp[j], dl);
            const int rel = (lin_bin(dl, Bf, Bm1) - bstar) * sdir;
            rare |= (rel == 0 ? 1u : 0u) << j;
            kp[j] = __fadd_rn(kp[j], rel > 0 ? adj0 : 0.0f);               // AS:660 / AS:664
        }
    }
    if (rare != 0u && Delta != 0) {
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            if (!((rare >> j) & 1u)) continue;
            float k0, dl;
            rz_round_t<FAST>(x[j], rc, k0, dl);
            const uint32_t key = rz_key(dl, neg);
            if (key > tie || (key == tie && (uint32_t)(i0 + j) < cut)) kp[j] = __fadd_rn(k0, adj);
        }
    }
    if (!e.packed) {
        bool ovf = false;
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            const int64_t i = i0 + j;
            if (!FULL && i >= d) break;
            if (e.deq_out) {
                const float sg = (x[j] > 0.0f) ? 1.0f : ((x[j] < 0.0f) ? -1.0f : 0.0f);     // AS:682 signs of the input
                e.deq_out[c * e.ld_out + i] = __fmul_rn(__fmul_rn(rc.L1f, sg), __fdiv_rn(kp[j], rc.mf));   // AS:666, AS:687
            }
            if (e.k_out) {
                if (!FAST && kp[j] >= 2147483648.0f) { ovf = true; e.k_out[c * e.ld_out + i] = 0x7fffffff; }
                else e.k_out[c * e.ld_out + i] = (int32_t)kp[j];
            }
            if (e.sgn_out) e.sgn_out[c * e.ld_out + i] = (uint8_t)(__float_as_uint(x[j]) >> 31);
        }
        if (ovf) atomicOr(&e.pack.hdr->status, 1u);
    } else {
        uint32_t k[kEpt], sg[kEpt];
        bool ovf = false;
#pragma unroll
        for (int j = 0; j < kEpt; ++j) {
            const float kk = (FULL || i0 + j < d) ? kp[j] : 0.0f;
            if (!FAST && kk >= 2147483648.0f) { ovf = true; k[j] = 0x7fffffffu; } else k[j] = (uint32_t)kk;
            sg[j] = __float_as_uint(x[j]) >> 31;
        }
        emit_packed_tile(e.pack, c, t, k, sg, ovf, s_pack);
    }
}

__global__ void __launch_bounds__(kThreads, 4)
rz_linapply_kernel(const float *__restrict__ X, int64_t d, int64_t ld, int64_t T, int B, const RowConst *__restrict__ consts,
                   const RowSelect *__restrict__ sel, RzEmit e) {
    __shared__ PackScratch s_pack;
    const int64_t c = blockIdx.y, t = blockIdx.x;
    float x[kEpt];
    load_tile_blocked(X + c * ld, d, t * kTile, x);
    const RowConst rc = consts[c];
    const long long Delta = sel[c].Delta;
    const int bstar = sel[c].lin_bstar;
    const uint32_t tie = sel[c].tie_key, cut = sel[c].lin_cut;
    const bool fast = rz_fast_row(rc) && rc.mf <= 16777216.0f, full = (t + 1) * kTile <= d;
    if (fast && full) rz_linapply_tile<true, true>(x, rc, Delta, bstar, tie, cut, B, d, c, t, e, s_pack);
    else if (fast) rz_linapply_tile<true, false>(x, rc, Delta, bstar, tie, cut, B, d, c, t, e, s_pack);
    else rz_linapply_tile<false, false>(x, rc, Delta, bstar, tie, cut, B, d, c, t, e, s_pack);
}

static int g_biased_path = 0;
void set_biased_path(int path) { g_biased_path = path; }

static int biased_quantize_linear(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, const WsLayout &L, void *ws, RzEmit e,
                                  cudaStream_t st) {
    char *base = (char *)ws;
    const RowConst *consts = (const RowConst *)(base + L.off_consts);
    RowSelect *sel = (RowSelect *)(base + L.off_sel);
    uint32_t *hist = (uint32_t *)(base + L.off_lin);
    uint2 *cand = (uint2 *)(base + L.off_lin + L.lin_hist_bytes);
    const int B = (int)L.lin_bins;
    // a run of tiles per CTA: one flush of the bins per run, and still a few thousand CTAs
    int64_t tp = (n * L.T) / 4096;
    tp = tp < 1 ? 1 : (tp > 64 ? 64 : tp);
    const dim3 grun((unsigned)((L.T + tp - 1) / tp), (unsigned)n), grid((unsigned)L.T, (unsigned)n);
    rz_sumhist_kernel<<<grun, kThreads, (size_t)(B + 2) * 4, st>>>(X, d, ld, L.T, (int)tp, B, consts, sel, hist);
    DME_LAUNCH_CHECK("rz_sumhist_kernel");
    rz_linpick_kernel<<<(unsigned)n, 32, 0, st>>>(sel, hist, n, m, d, B);
    DME_LAUNCH_CHECK("rz_linpick_kernel");
    rz_compact_kernel<<<grun, kThreads, 0, st>>>(X, d, ld, L.T, (int)tp, B, L.lin_cap, consts, sel, cand, (WsHeader *)base);
    DME_LAUNCH_CHECK("rz_compact_kernel");
    rz_linselect_kernel<<<(unsigned)n, kThreads, 0, st>>>(sel, cand, L.lin_cap);
    DME_LAUNCH_CHECK("rz_linselect_kernel");
    rz_linapply_kernel<<<grid, kThreads, 0, st>>>(X, d, ld, L.T, B, consts, sel, e);
    DME_LAUNCH_CHECK("rz_linapply_kernel");
    return DME_OK;
}

// Called after l1_kernel has published the row constants.
int biased_quantize(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, const WsLayout &L, void *ws,
                    int32_t *k_out, uint8_t *sgn_out, float *deq_out, int64_t ld_out,
                    uint32_t *codes, int64_t codes_bytes, uint64_t *dir, cudaStream_t st) {
    char *base = (char *)ws;
    if (g_biased_path == 0) {
        RzEmit e0;
        e0.k_out = k_out; e0.sgn_out = sgn_out; e0.deq_out = deq_out; e0.ld_out = ld_out;
        e0.packed = codes != nullptr;
        init_pack_target(e0.pack, codes, codes_bytes, dir, (WsHeader *)base, n, d, m);
        if (e0.packed && (long long)(e0.pack.arena_base16 * 16ull) > codes_bytes) {
            set_error("code arena too small for the primary slots: %lld bytes", (long long)codes_bytes);
            return DME_EWORKSPACE;
        }
        return biased_quantize_linear(X, n, d, ld, m, L, ws, e0, st);
    }
    const RowConst *consts = (const RowConst *)(base + L.off_consts);
    RowSelect *sel = (RowSelect *)(base + L.off_sel);
    uint32_t *tie_cnt = (uint32_t *)(base + L.off_partial);         // the L1 partials are dead by now
    const dim3 grid((unsigned)L.T, (unsigned)n);
    rz_sum_kernel<<<grid, kThreads, 0, st>>>(X, d, ld, consts, sel);
    DME_LAUNCH_CHECK("rz_sum_kernel");
    rz_init_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(sel, n, m, d);
    DME_LAUNCH_CHECK("rz_init_kernel");
    for (int pass = 0; pass < 4; ++pass) {
        rz_hist_kernel<<<grid, kThreads, 0, st>>>(X, d, ld, consts, sel, pass);
        DME_LAUNCH_CHECK("rz_hist_kernel");
        rz_pick_kernel<<<(unsigned)n, 32, 0, st>>>(sel, pass);
        DME_LAUNCH_CHECK("rz_pick_kernel");
    }
    rz_tiecnt_kernel<<<grid, kThreads, 0, st>>>(X, d, ld, L.T, consts, sel, tie_cnt);
    DME_LAUNCH_CHECK("rz_tiecnt_kernel");
    rz_tiescan_kernel<<<(unsigned)n, kThreads, 0, st>>>(tie_cnt, L.T);
    DME_LAUNCH_CHECK("rz_tiescan_kernel");
    RzEmit e;
    e.k_out = k_out; e.sgn_out = sgn_out; e.deq_out = deq_out; e.ld_out = ld_out;
    e.packed = codes != nullptr;
    init_pack_target(e.pack, codes, codes_bytes, dir, (WsHeader *)base, n, d, m);
    if (e.packed && (long long)(e.pack.arena_base16 * 16ull) > codes_bytes) {
        set_error("code arena too small for the primary slots: %lld bytes", (long long)codes_bytes);
        return DME_EWORKSPACE;
    }
    rz_apply_kernel<<<grid, kThreads, 0, st>>>(X, d, ld, L.T, consts, sel, tie_cnt, e);
    DME_LAUNCH_CHECK("rz_apply_kernel");
    return DME_OK;
}

}  // namespace dme
