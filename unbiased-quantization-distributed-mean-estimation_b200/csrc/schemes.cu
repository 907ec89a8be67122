// schemes.cu -- the comparison quantizers of the reference: DRIVE (AS:707-752), EDEN (AS:323-426), the QUIC-FL
// receiver (AS:526-535) and the scalar stochastic quantizer (AS:755-790).  Rotations come from hadamard.cu.
// Reductions whose fp32 order ATen leaves unspecified (norms, dot products, abs-sums) accumulate in fp64 in a
// fixed order and round once, like the oracle.
#include <cmath>

#include "common.cuh"

namespace dme {

int fwht_rows(const float *src, int64_t src_d, int64_t src_ld, float *dst, int64_t d, int64_t dst_ld, int64_t n,
              const float *diag, uint64_t seed, uint64_t seed_stride, int pre_diag, int post_diag, cudaStream_t st);   // hadamard.cu

__device__ __forceinline__ float sgn0(float v) { return (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f); }

// ------------------------------------------------------------------ DRIVE: one CTA per 2048-chunk
__global__ void __launch_bounds__(256)
drive_kernel(const float *__restrict__ X, int64_t d, int64_t ld, float *__restrict__ out, int64_t ld_out, uint64_t seed,
             const float *__restrict__ dsign, int64_t dsign_row, int compat) {
    __shared__ float buf[2048];
    __shared__ float Dg[2048];
    __shared__ double s_red[kWarps];
    const int64_t c = blockIdx.y, chunk = blockIdx.x;
    const int64_t s0 = chunk * 2048;
    const int len = (int)((d - s0 < 2048) ? (d - s0) : 2048);
    int np2 = 1, lg = 0;
    while (np2 < len) { np2 <<= 1; ++lg; }
    const float *row = X + c * ld;
    // the chunk's 2048 sign bits = 16 Philox blocks of 128 bits (philox_sign's mapping: coordinate i is bit i & 31 of word
    // (i >> 5) & 3 of block i >> 7): computed once per chunk by 16 threads instead of once per coordinate
    __shared__ uint32_t s_bits[64];
    if (!dsign && threadIdx.x < 16) {
        const uint64_t blk = (uint64_t)(s0 >> 7) + threadIdx.x;
        const Philox4 p = philox4x32_10(seed, (uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)c, kStreamDrive);
        s_bits[4 * threadIdx.x] = p.x; s_bits[4 * threadIdx.x + 1] = p.y; s_bits[4 * threadIdx.x + 2] = p.z; s_bits[4 * threadIdx.x + 3] = p.w;
    }
    __syncthreads();
    double n2 = 0.0;
    for (int j = threadIdx.x; j < np2; j += 256) {
        const float xv = j < len ? row[s0 + j] : 0.0f;
        const float Dv = dsign ? dsign[c * dsign_row + s0 + j]
                               : (((s_bits[j >> 5] >> (j & 31)) & 1u) ? 1.0f : -1.0f);                   // AS:735
        Dg[j] = Dv;
        buf[j] = __fmul_rn(Dv, xv);                                                                    // AS:737
        n2 += (double)xv * (double)xv;
    }
    n2 = block_sum_f64(n2, s_red);
    __syncthreads();
    auto transform = [&]() {
        if (compat == 0) {
            // AS:37-59 as executed: lg stages on every adjacent pair
            for (int j = threadIdx.x; 2 * j + 1 < np2; j += 256) {
                float a = buf[2 * j], b = buf[2 * j + 1];
                for (int s = 0; s < lg; ++s) { const float t = __fadd_rn(a, b); b = __fsub_rn(t, b); a = t; }
                buf[2 * j] = a; buf[2 * j + 1] = b;
            }
            __syncthreads();
        } else {
            for (int h = 1; h < np2; h <<= 1) {
                for (int j = threadIdx.x; j < np2 / 2; j += 256) {
                    const int i = ((j / h) * 2 * h) + (j % h);
                    const float a = buf[i], b = buf[i + h];
                    buf[i] = __fadd_rn(a, b); buf[i + h] = __fsub_rn(a, b);
                }
                __syncthreads();
            }
        }
    };
    transform();                                                                                       // AS:738
    double l1 = 0.0;
    for (int j = threadIdx.x; j < np2; j += 256) l1 += (double)fabsf(buf[j]);
    l1 = block_sum_f64(l1, s_red);
    const float nrm = (float)sqrt(n2);
    const float S = __fdiv_rn(__fmul_rn(nrm, nrm), __fadd_rn((float)l1, 1e-12f));                      // AS:741
    __syncthreads();
    for (int j = threadIdx.x; j < np2; j += 256) buf[j] = __fmul_rn(S, sgn0(buf[j]));                  // AS:743
    __syncthreads();
    transform();                                                                                       // AS:746
    for (int j = threadIdx.x; j < len; j += 256) out[c * ld_out + s0 + j] = __fmul_rn(buf[j], Dg[j]);  // AS:747-750
}

// ------------------------------------------------------------------ row reductions (fp64, fixed order)
// partial[c * nb + b] = sum over the b-th slice of row c of f(v).  kind 0: v^2, 1: centroid[bin]*v (EDEN dot)
struct EdenTab { float cent[4]; float bnd[3]; int nc; };

__global__ void __launch_bounds__(256)
row_sumsq_kernel(const float *__restrict__ V, int64_t dpad, int64_t ld, int nb, double *__restrict__ partial) {
    __shared__ double s_red[kWarps];
    const int64_t c = blockIdx.y;
    const int64_t per = (dpad + nb - 1) / nb, lo = blockIdx.x * per, hi = min(dpad, lo + per);
    double s = 0.0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += 256) { const double v = V[c * ld + i]; s += v * v; }
    s = block_sum_f64(s, s_red);
    if (threadIdx.x == 0) partial[c * nb + blockIdx.x] = s;
}
__global__ void norm_finalize_kernel(const double *__restrict__ partial, int nb, int64_t n, const float *__restrict__ inject,
                                     float *__restrict__ nrm) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += partial[c * nb + b];
    nrm[c] = inject ? inject[c] : (float)sqrt(s);                                 // torch.norm(vec, 2)
}
__global__ void __launch_bounds__(256)
eden_bucket_kernel(const float *__restrict__ V, int64_t dpad, int64_t ld, int nb, const float *__restrict__ nrm, float sq, EdenTab tab,
                   uint8_t *__restrict__ bins, double *__restrict__ partial) {
    __shared__ double s_red[kWarps];
    const int64_t c = blockIdx.y;
    const int64_t per = (dpad + nb - 1) / nb, lo = blockIdx.x * per, hi = min(dpad, lo + per);
    const float nr = nrm[c];
    double s = 0.0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
        const float v = V[c * ld + i];
        const float z = __fdiv_rn(__fmul_rn(v, sq), nr);                           // AS:343
        int b = 0;
        while (b < tab.nc - 1 && tab.bnd[b] < z) ++b;                              // torch.bucketize, right=False
        bins[c * dpad + i] = (uint8_t)b;
        s += (double)tab.cent[b] * (double)v;
    }
    s = block_sum_f64(s, s_red);
    if (threadIdx.x == 0) partial[c * nb + blockIdx.x] = s;
}
__global__ void eden_scale_kernel(const double *__restrict__ partial, int nb, int64_t n, const float *__restrict__ nrm, float *__restrict__ scale) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += partial[c * nb + b];
    scale[c] = __fdiv_rn(__fmul_rn(nrm[c], nrm[c]), (float)s);                     // AS:348
}
// Fractional rates (AS:352-368): coordinate i uses the high-rate table where mask_i is set, the low-rate table elsewhere.
// mask_i = injected byte, or [u_i < p_high] with u_i from Philox keyed by the client's mask seed (AS:389: seed * 7 + 13).
constexpr uint32_t kStreamEdenMask = 0x45444Du;  // "EDM"
__device__ __forceinline__ bool eden_mask_bit(const uint8_t *__restrict__ mask, int64_t c, int64_t dpad, int64_t i, uint64_t mseed, float p_high) {
    if (mask) return mask[c * dpad + i] != 0;
    const Philox4 p = philox4x32_10(mseed, (uint32_t)(i >> 2), (uint32_t)((uint64_t)i >> 34), 0u, kStreamEdenMask);
    const uint32_t w = (i & 3) == 0 ? p.x : (i & 3) == 1 ? p.y : (i & 3) == 2 ? p.z : p.w;
    return u24_to_unit(w) < p_high;
}
__global__ void __launch_bounds__(256)
eden_bucket_frac_kernel(const float *__restrict__ V, int64_t dpad, int64_t ld, int nb, const float *__restrict__ nrm, float sq, EdenTab tlo, EdenTab thi,
                        const uint8_t *__restrict__ mask, uint64_t seed, uint64_t seed_stride, float p_high, uint8_t *__restrict__ bins,
                        double *__restrict__ partial) {
    __shared__ double s_red[kWarps];
    const int64_t c = blockIdx.y;
    const int64_t per = (dpad + nb - 1) / nb, lo = blockIdx.x * per, hi = min(dpad, lo + per);
    const float nr = nrm[c];
    const uint64_t mseed = (seed + seed_stride * (uint64_t)c) * 7ull + 13ull;
    double s = 0.0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
        const float v = V[c * ld + i];
        const float z = __fdiv_rn(__fmul_rn(v, sq), nr);                           // AS:354-355
        const EdenTab &tab = eden_mask_bit(mask, c, dpad, i, mseed, p_high) ? thi : tlo;     // AS:360-364
        int b = 0;
        while (b < tab.nc - 1 && tab.bnd[b] < z) ++b;
        bins[c * dpad + i] = (uint8_t)b;
        s += (double)tab.cent[b] * (double)v;                                      // AS:366
    }
    s = block_sum_f64(s, s_red);
    if (threadIdx.x == 0) partial[c * nb + blockIdx.x] = s;
}
// AS:400-421: centroid look-up by the coordinate's table, then the optional drop (zero the coordinate, rescale the rest)
__global__ void eden_lookup_frac_kernel(const uint8_t *__restrict__ bins, int64_t n, int64_t dpad, EdenTab tlo, EdenTab thi, const uint8_t *__restrict__ mask,
                                        uint64_t seed, uint64_t seed_stride, float p_high, const uint8_t *__restrict__ drop, float inv_keep,
                                        float *__restrict__ work) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n * dpad) return;
    const int64_t c = k / dpad, i = k - c * dpad;
    const uint64_t mseed = (seed + seed_stride * (uint64_t)c) * 7ull + 13ull;
    const bool high = p_high >= 1.0f || (p_high > 0.0f && eden_mask_bit(mask, c, dpad, i, mseed, p_high));
    float v = (high ? thi : tlo).cent[bins[k] & 3];
    if (drop) v = drop[k] ? 0.0f : __fdiv_rn(v, inv_keep);                         // AS:419-421 (inv_keep = 1 - pdrop)
    work[k] = v;
}
// four bins -> four centroids per thread (total is a multiple of 4: rows are padded to a power of two >= 4 or handled by the tail)
__global__ void eden_lookup_kernel(const uint8_t *__restrict__ bins, int64_t total, EdenTab tab, float *__restrict__ work) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, i = 4 * q;
    if (i + 4 <= total && ((reinterpret_cast<uintptr_t>(bins) & 3u) == 0) && ((reinterpret_cast<uintptr_t>(work) & 15u) == 0)) {
        const uchar4 b = *reinterpret_cast<const uchar4 *>(bins + i);
        *reinterpret_cast<float4 *>(work + i) = make_float4(tab.cent[b.x & 3], tab.cent[b.y & 3], tab.cent[b.z & 3], tab.cent[b.w & 3]);   // AS:400
    } else {
        for (int64_t k = i; k < total && k < i + 4; ++k) work[k] = tab.cent[bins[k] & 3];
    }
}
// grid (blocks along the row, client): out[c][j] = scale[c] * work[c][j] for j < d, four coordinates per thread
__global__ void scale_rows_kernel(const float *__restrict__ work, int64_t dpad, const float *__restrict__ scale, int64_t n, int64_t d,
                                  float *__restrict__ out, int64_t ld_out) {
    const int64_t c = blockIdx.y;
    const int64_t j = 4 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (j >= d) return;
    const float sc = scale ? scale[c] : 1.0f;
    const float *w = work + c * dpad;
    float *o = out + c * ld_out;
    const bool vec = j + 4 <= d && ((dpad & 3) == 0) && ((ld_out & 3) == 0) && (((reinterpret_cast<uintptr_t>(work) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0);
    if (vec) {
        float4 v = *reinterpret_cast<const float4 *>(w + j);
        if (scale) { v.x = __fmul_rn(sc, v.x); v.y = __fmul_rn(sc, v.y); v.z = __fmul_rn(sc, v.z); v.w = __fmul_rn(sc, v.w); }   // AS:426 / AS:535
        *reinterpret_cast<float4 *>(o + j) = v;
    } else {
        for (int64_t k = j; k < d && k < j + 4; ++k) o[k] = scale ? __fmul_rn(sc, w[k]) : w[k];
    }
}

static EdenTab eden_tab(int nbits) {
    EdenTab t{};
    if (nbits == 1) {
        t.nc = 2; t.cent[0] = -(float)0.7978845608028654; t.cent[1] = (float)0.7978845608028654;   // AS:303
        t.bnd[0] = (t.cent[0] + t.cent[1]) / 2.0f;
    } else {
        t.nc = 4;
        t.cent[0] = -(float)1.5104176087114887; t.cent[1] = -(float)0.4527800398860679;            // AS:304
        t.cent[2] = (float)0.4527800398860679; t.cent[3] = (float)1.5104176087114887;
        for (int i = 0; i < 3; ++i) t.bnd[i] = (t.cent[i] + t.cent[i + 1]) / 2.0f;                 // AS:311-312
    }
    return t;
}

// ------------------------------------------------------------------ QUIC-FL receiver: one CTA per row
__global__ void __launch_bounds__(256)
quicfl_gather_kernel(const int32_t *__restrict__ Xq, const int32_t *__restrict__ h, int64_t dpad, int h_len,
                     const float *__restrict__ table, int table_len, const uint8_t *__restrict__ mask, const float *__restrict__ exact_vals,
                     const int64_t *__restrict__ exact_off, const float *__restrict__ scale, float *__restrict__ work) {
    __shared__ uint32_t s_w[kWarps];
    __shared__ uint32_t s_carry;
    const int64_t c = blockIdx.x;
    const float sc = scale[c];
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < dpad; base += 256) {
        const int64_t i = base + threadIdx.x;
        const bool live = i < dpad;
        const uint32_t mk = (live && mask) ? (mask[c * dpad + i] != 0) : 0u;
        uint32_t inc = mk;
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
        if (live) {
            int64_t idx = (int64_t)Xq[c * dpad + i] * h_len + h[c * dpad + i];
            idx = idx < 0 ? 0 : (idx >= table_len ? table_len - 1 : idx);
            float val = table[idx];                                                               // AS:530
            if (mk) val = exact_off ? exact_vals[exact_off[c] + carry + wbase + inc - 1] : exact_vals[c * dpad + i];   // AS:531 (or dense)
            work[c * dpad + i] = __fdiv_rn(val, sc);                                               // AS:532
        }
        __syncthreads();
        if (threadIdx.x == 255) s_carry = carry + wbase + inc;
        __syncthreads();
    }
}

// ------------------------------------------------------------------ scalar stochastic quantizer
__global__ void __launch_bounds__(256)
row_minmax_kernel(const float *__restrict__ X, int64_t d, int64_t ld, int nb, float *__restrict__ pmin, float *__restrict__ pmax) {
    __shared__ float s_mn[kWarps], s_mx[kWarps];
    const int64_t c = blockIdx.y;
    const int64_t per = (d + nb - 1) / nb, lo = blockIdx.x * per, hi = min(d, lo + per);
    float mn = INFINITY, mx = -INFINITY;
    const float *xr = X + c * ld;
    int64_t i = lo + threadIdx.x;
    if (((uintptr_t)(xr + lo) & 15u) == 0) {          // 128-bit loads over the aligned body of the slice
        const int64_t nv = (hi - lo) / 4;
        for (int64_t q = threadIdx.x; q < nv; q += 256) {
            const float4 v = ldg_stream_f4(xr + lo + 4 * q);
            mn = fminf(fminf(mn, fminf(v.x, v.y)), fminf(v.z, v.w)); mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
        }
        i = lo + 4 * nv + threadIdx.x;
    }
    for (; i < hi; i += 256) { const float v = xr[i]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kWarps; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
        pmin[c * nb + blockIdx.x] = mn; pmax[c * nb + blockIdx.x] = mx;
    }
}
// One coordinate of AS:768-788 (u: the uniform of AS:783).
__device__ __forceinline__ float scalar_one(float x, float mn, float mx, float denom, float nlevels, float u) {
    float q = __fdiv_rn(__fsub_rn(x, mn), denom);                                                    // AS:768
    q = fminf(fmaxf(q, 0.0f), 1.0f);                                                                 // AS:776
    const float t = __fmul_rn(q, nlevels);
    const float bf = floorf(t);                                                                      // AS:779
    const float fr = __fsub_rn(t, bf);
    const float bi = __fadd_rn(bf, (u < fr) ? 1.0f : 0.0f);                                          // AS:783-784
    q = __fdiv_rn(bi, nlevels);                                                                      // AS:787
    return __fadd_rn(__fmul_rn(q, __fsub_rn(mx, mn)), mn);                                           // AS:788
}
// grid (blocks along the row, client).  The CTA first folds the row's partial minima / maxima (fminf / fmaxf: any order
// gives the same value), then every thread takes groups of 4 consecutive coordinates: one 128-bit load and store and ONE
// Philox block per group (coordinate i uses component i & 3 of block i >> 2 of the client's stream).
constexpr int kScalarGroups = 4;      // groups of 4 coordinates per thread
__global__ void __launch_bounds__(256)
scalar_kernel(const float *__restrict__ X, int64_t n, int64_t d, int64_t ld, int nb, const float *__restrict__ pmin,
              const float *__restrict__ pmax, float nlevels, uint64_t seed, uint64_t client0, const float *__restrict__ u_inject,
              float *__restrict__ out, int64_t ld_out, int vec_ok) {
    __shared__ float s_mn[kWarps], s_mx[kWarps];
    const int64_t c = blockIdx.y;
    float mn = INFINITY, mx = -INFINITY;
    for (int b = threadIdx.x; b < nb; b += 256) { mn = fminf(mn, pmin[c * nb + b]); mx = fmaxf(mx, pmax[c * nb + b]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    mn = s_mn[0]; mx = s_mx[0];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
    const float denom = __fsub_rn(mx, mn);
    const bool copy = (denom == 0.0f || nlevels < 1.0f);                                            // AS:763-765, AS:772-773
    const uint64_t cl = client0 + (uint64_t)c;
    const uint64_t key = seed ^ (cl * 0x9E3779B97F4A7C15ull);
    const float *xr = X + c * ld;
    float *orow = out + c * ld_out;
    const int64_t g0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * kScalarGroups;       // first group of this thread
#pragma unroll
    for (int k = 0; k < kScalarGroups; ++k) {
        const int64_t g = g0 + k, i0 = g * 4;
        if (i0 >= d) break;
        float x[4], u[4];
        const bool full = vec_ok && i0 + 4 <= d;
        if (full) { const float4 v = *reinterpret_cast<const float4 *>(xr + i0); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
        else { for (int j = 0; j < 4; ++j) x[j] = (i0 + j < d) ? xr[i0 + j] : 0.0f; }
        if (!copy) {
            if (u_inject) { for (int j = 0; j < 4; ++j) u[j] = (i0 + j < d) ? u_inject[c * d + i0 + j] : 0.0f; }
            else {
                const Philox4 p = philox4x32_10(key, (uint32_t)g, (uint32_t)((uint64_t)g >> 32), (uint32_t)cl, kStreamScalar);
                u[0] = u24_to_unit(p.x); u[1] = u24_to_unit(p.y); u[2] = u24_to_unit(p.z); u[3] = u24_to_unit(p.w);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = scalar_one(x[j], mn, mx, denom, nlevels, u[j]);
        }
        if (full) *reinterpret_cast<float4 *>(orow + i0) = make_float4(x[0], x[1], x[2], x[3]);
        else { for (int j = 0; j < 4; ++j) if (i0 + j < d) orow[i0 + j] = x[j]; }
    }
}

static int pow2_ceil(int64_t v) { int64_t p = 1; while (p < v) p <<= 1; return (int)p; }
static int slices(int64_t d) { int64_t nb = (d + 16383) / 16384; return (int)(nb < 1 ? 1 : (nb > 64 ? 64 : nb)); }
static int slices_minmax(int64_t d) { int64_t nb = (d + 16383) / 16384; return (int)(nb < 1 ? 1 : (nb > 256 ? 256 : nb)); }   // min / max: any grouping gives the same value

}  // namespace dme

using namespace dme;

extern "C" int dme_drive(const float *X, int64_t n, int64_t d, int64_t ld, float *out, int64_t ld_out, uint64_t seed,
                         const float *dsign_inject, int compat, dme_stream_t stream) {
    DME_REQUIRE(X && out && n >= 1 && n <= 65535 && d >= 1 && ld >= d && ld_out >= d, "bad argument");
    DME_REQUIRE(compat == 0 || compat == 1, "compat must be 0 (reference transform) or 1 (true WHT)");
    const int64_t chunks = (d + 2047) / 2048;
    const int64_t tail = d - (chunks - 1) * 2048;
    const int64_t row_pad = (chunks - 1) * 2048 + pow2_ceil(tail);
    dim3 grid((unsigned)chunks, (unsigned)n);
    drive_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, d, ld, out, ld_out, seed, dsign_inject, row_pad, compat);
    DME_LAUNCH_CHECK("drive_kernel");
    return DME_OK;
}

// stream-ordered scratch that is released on every exit path
struct AsyncScratch {
    void *p = nullptr; cudaStream_t st;
    explicit AsyncScratch(cudaStream_t s) : st(s) {}
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes, st); }
    ~AsyncScratch() { if (p) cudaFreeAsync(p, st); }
};

static int eden_encode_impl(const float *X, int64_t n, int64_t d, int64_t ld, int64_t dpad, int nbits_low, int nbits_high, float p_high,
                            const uint8_t *mask_inject, uint64_t seed, uint64_t seed_stride, const float *diag_inject, const float *norm_inject,
                            float *rot, uint8_t *bins, float *scale, cudaStream_t st) {
    DME_REQUIRE(X && rot && bins && scale && n >= 1 && n <= 65535, "bad argument");
    DME_REQUIRE((nbits_low == 1 || nbits_low == 2) && (nbits_high == 1 || nbits_high == 2), "EDEN centroids exist for 1 and 2 bits only (AS:301-320)");
    int rc = fwht_rows(X, d, ld, rot, dpad, dpad, n, diag_inject, seed, seed_stride, 1, 0, st);                   // AS:378-380
    if (rc) return rc;
    const int nb = slices(dpad);
    AsyncScratch partial(st), nrm(st);
    DME_CUDA(partial.alloc(sizeof(double) * (size_t)(n * nb)));
    DME_CUDA(nrm.alloc(sizeof(float) * (size_t)n));
    dim3 grid((unsigned)nb, (unsigned)n);
    row_sumsq_kernel<<<grid, 256, 0, st>>>(rot, dpad, dpad, nb, (double *)partial.p);
    DME_LAUNCH_CHECK("row_sumsq_kernel");
    norm_finalize_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>((double *)partial.p, nb, n, norm_inject, (float *)nrm.p);
    DME_LAUNCH_CHECK("norm_finalize_kernel");
    const float sq = (float)std::pow((double)dpad, 0.5);                                             // vec.numel() ** 0.5
    if (nbits_low == nbits_high) {
        eden_bucket_kernel<<<grid, 256, 0, st>>>(rot, dpad, dpad, nb, (float *)nrm.p, sq, eden_tab(nbits_low), bins, (double *)partial.p);
        DME_LAUNCH_CHECK("eden_bucket_kernel");
    } else {
        eden_bucket_frac_kernel<<<grid, 256, 0, st>>>(rot, dpad, dpad, nb, (float *)nrm.p, sq, eden_tab(nbits_low), eden_tab(nbits_high), mask_inject,
                                                     seed, seed_stride, p_high, bins, (double *)partial.p);
        DME_LAUNCH_CHECK("eden_bucket_frac_kernel");
    }
    eden_scale_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>((double *)partial.p, nb, n, (float *)nrm.p, scale);
    DME_LAUNCH_CHECK("eden_scale_kernel");
    return DME_OK;
}

extern "C" int dme_eden_encode(const float *X, int64_t n, int64_t d, int64_t ld, int64_t dpad, int nbits, uint64_t seed,
                               uint64_t seed_stride, const float *diag_inject, const float *norm_inject, float *rot, uint8_t *bins, float *scale,
                               dme_stream_t stream) {
    return eden_encode_impl(X, n, d, ld, dpad, nbits, nbits, 0.0f, nullptr, seed, seed_stride, diag_inject, norm_inject, rot, bins, scale,
                            (cudaStream_t)stream);
}
extern "C" int dme_eden_encode_frac(const float *X, int64_t n, int64_t d, int64_t ld, int64_t dpad, int nbits_low, int nbits_high, float p_high,
                                    const uint8_t *mask_inject, uint64_t seed, uint64_t seed_stride, const float *diag_inject,
                                    const float *norm_inject, float *rot, uint8_t *bins, float *scale, dme_stream_t stream) {
    DME_REQUIRE(p_high >= 0.0f && p_high <= 1.0f, "p_high=%f outside [0, 1]", (double)p_high);
    return eden_encode_impl(X, n, d, ld, dpad, nbits_low, nbits_high, p_high, mask_inject, seed, seed_stride, diag_inject, norm_inject, rot, bins,
                            scale, (cudaStream_t)stream);
}

extern "C" int dme_eden_decode(const uint8_t *bins, const float *scale, int64_t n, int64_t d, int64_t dpad, int nbits, uint64_t seed,
                               uint64_t seed_stride, const float *diag_inject, float *work, float *out, int64_t ld_out, dme_stream_t stream) {
    DME_REQUIRE(bins && scale && work && out && n >= 1 && d >= 1 && dpad >= d && ld_out >= d, "bad argument");
    DME_REQUIRE(nbits == 1 || nbits == 2, "EDEN centroids exist for 1 and 2 bits only (AS:301-320)");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = n * dpad;
    eden_lookup_kernel<<<(unsigned)((total + 1023) / 1024), 256, 0, st>>>(bins, total, eden_tab(nbits), work);
    DME_LAUNCH_CHECK("eden_lookup_kernel");
    int rc = fwht_rows(work, dpad, dpad, work, dpad, dpad, n, diag_inject, seed, seed_stride, 0, 1, st);   // AS:425
    if (rc) return rc;
    scale_rows_kernel<<<dim3((unsigned)((d + 1023) / 1024), (unsigned)n), 256, 0, st>>>(work, dpad, scale, n, d, out, ld_out);
    DME_LAUNCH_CHECK("scale_rows_kernel");
    return DME_OK;
}
// Fractional rates and drops (AS:401-421).  drop (n * dpad bytes, nullable): coordinates the receiver zeroes; the rest is divided
// by keep = 1 - pdrop.
extern "C" int dme_eden_decode_frac(const uint8_t *bins, const float *scale, int64_t n, int64_t d, int64_t dpad, int nbits_low, int nbits_high,
                                    float p_high, const uint8_t *mask_inject, const uint8_t *drop, float keep, uint64_t seed, uint64_t seed_stride,
                                    const float *diag_inject, float *work, float *out, int64_t ld_out, dme_stream_t stream) {
    DME_REQUIRE(bins && scale && work && out && n >= 1 && d >= 1 && dpad >= d && ld_out >= d, "bad argument");
    DME_REQUIRE((nbits_low == 1 || nbits_low == 2) && (nbits_high == 1 || nbits_high == 2), "EDEN centroids exist for 1 and 2 bits only (AS:301-320)");
    DME_REQUIRE(p_high >= 0.0f && p_high <= 1.0f && (!drop || keep > 0.0f), "p_high / keep out of range");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = n * dpad;
    eden_lookup_frac_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(bins, n, dpad, eden_tab(nbits_low), eden_tab(nbits_high), mask_inject, seed,
                                                                            seed_stride, nbits_low == nbits_high ? 0.0f : p_high, drop, keep, work);
    DME_LAUNCH_CHECK("eden_lookup_frac_kernel");
    int rc = fwht_rows(work, dpad, dpad, work, dpad, dpad, n, diag_inject, seed, seed_stride, 0, 1, st);   // AS:425
    if (rc) return rc;
    scale_rows_kernel<<<dim3((unsigned)((d + 1023) / 1024), (unsigned)n), 256, 0, st>>>(work, dpad, scale, n, d, out, ld_out);
    DME_LAUNCH_CHECK("scale_rows_kernel");
    return DME_OK;
}

extern "C" int dme_quicfl_decode(const int32_t *Xq, const int32_t *h, int64_t n, int64_t d, int64_t dpad, int h_len,
                                 const float *recv_table, int table_len, const uint8_t *exact_mask, const float *exact_vals,
                                 const int64_t *exact_off, const float *scale, uint64_t rotation_seed, const float *diag_inject,
                                 float *work, float *out, int64_t ld_out, dme_stream_t stream) {
    DME_REQUIRE(Xq && h && recv_table && scale && work && out, "null pointer argument");
    DME_REQUIRE(n >= 1 && d >= 1 && dpad >= d && h_len >= 1 && table_len >= h_len && ld_out >= d, "bad geometry");
    DME_REQUIRE(!exact_mask || exact_vals, "exact_mask needs exact_vals (compacted with exact_off, or dense n x dpad without)");
    cudaStream_t st = (cudaStream_t)stream;
    quicfl_gather_kernel<<<(unsigned)n, 256, 0, st>>>(Xq, h, dpad, h_len, recv_table, table_len, exact_mask, exact_vals, exact_off, scale, work);
    DME_LAUNCH_CHECK("quicfl_gather_kernel");
    int rc = fwht_rows(work, dpad, dpad, work, dpad, dpad, n, diag_inject, rotation_seed, 0, 0, 1, st);  // AS:534
    if (rc) return rc;
    scale_rows_kernel<<<dim3((unsigned)((d + 1023) / 1024), (unsigned)n), 256, 0, st>>>(work, dpad, nullptr, n, d, out, ld_out);
    DME_LAUNCH_CHECK("scale_rows_kernel");
    return DME_OK;
}

// ------------------------------------------------------------------ QUIC-FL sender (AS:455-503)
// After the rotation and the scaling to unit variance: coordinates beyond the tail threshold are sent exactly; the others are
// rounded stochastically to the grid of step delta, and the sender table of the (grid point, shared randomness h) pair gives
// the index X and the probability of X + 1.  Four coordinates per thread; two Philox blocks per group: word j of the first
// gives coordinate j its h (low bits) and the uniform of the grid rounding (top 24 bits), word j of the second the uniform of
// the X / X + 1 choice.  (The reference draws h from a torch generator that sender and receiver seed alike and the last
// Bernoulli from the global generator, AS:457-490; here both sides get h from the sender's output.)
constexpr uint32_t kStreamQuic = 0x514643u;   // "QFC"
__global__ void __launch_bounds__(256)
quicfl_encode_kernel(const float *__restrict__ rot, int64_t dpad, const float *__restrict__ nrm, float sqd, int h_len, int x_len, float delta,
                     float thr, const int8_t *__restrict__ send_X, const float *__restrict__ send_p, uint64_t seed, uint64_t client0,
                     int32_t *__restrict__ Xq, int32_t *__restrict__ h_out, uint8_t *__restrict__ exact_mask, float *__restrict__ exact_dense,
                     float *__restrict__ scale_out) {
    const int64_t c = blockIdx.y;
    const float scale = __fdiv_rn(sqd, nrm[c]);                                                       // AS:466 / AS:470
    if (blockIdx.x == 0 && threadIdx.x == 0) scale_out[c] = scale;
    const uint64_t cl = client0 + (uint64_t)c;
    const uint64_t key = seed ^ (cl * 0x9E3779B97F4A7C15ull);
    const int half = (x_len - 1) / 2;
    const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x, i0 = g * 4;
    if (i0 >= dpad) return;
    const float4 v4 = *reinterpret_cast<const float4 *>(rot + c * dpad + i0);
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
    const Philox4 pa = philox4x32_10(key, (uint32_t)g, (uint32_t)((uint64_t)g >> 32), 0u, kStreamQuic);
    const Philox4 pb = philox4x32_10(key, (uint32_t)g, (uint32_t)((uint64_t)g >> 32), 1u, kStreamQuic);
    const uint32_t wa[4] = {pa.x, pa.y, pa.z, pa.w}, wb[4] = {pb.x, pb.y, pb.z, pb.w};
    int xs[4], hs[4];
    uint32_t em = 0;
    float ev[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float z = __fmul_rn(v[j], scale);                                                       // AS:472
        const bool exact = z > thr || z < -thr;                                                       // AS:478
        float q = __fdiv_rn(z, delta);                                                                // AS:480
        if (exact) q = 0.0f;                                                                          // AS:481
        const float fl = floorf(q), p = __fsub_rn(q, fl);                                             // AS:483
        int iq = (int)fl + ((u24_to_unit(wa[j]) < p) ? 1 : 0);                                        // AS:484
        iq = iq < -half ? -half : (iq > half ? half : iq);
        const int h = (int)(wa[j] & 0xffu) % h_len;
        const int64_t idx = (int64_t)(iq + half) * h_len + h;                                         // AS:486
        const int X = (int)send_X[idx] + ((u24_to_unit(wb[j]) < send_p[idx]) ? 1 : 0);                // AS:486-489
        xs[j] = X; hs[j] = h; ev[j] = exact ? z : 0.0f;
        em |= (exact ? 1u : 0u) << (8 * j);
    }
    *reinterpret_cast<int4 *>(Xq + c * dpad + i0) = make_int4(xs[0], xs[1], xs[2], xs[3]);
    *reinterpret_cast<int4 *>(h_out + c * dpad + i0) = make_int4(hs[0], hs[1], hs[2], hs[3]);
    *reinterpret_cast<uint32_t *>(exact_mask + c * dpad + i0) = em;
    *reinterpret_cast<float4 *>(exact_dense + c * dpad + i0) = make_float4(ev[0], ev[1], ev[2], ev[3]);
}

extern "C" int dme_quicfl_encode(const float *X, int64_t n, int64_t d, int64_t ld, int64_t dpad, int h_len, int x_len, float delta,
                                 float exact_threshold, const int8_t *send_X, const float *send_p, uint64_t seed, uint64_t client0,
                                 uint64_t rotation_seed, const float *diag_inject, float *rot, int32_t *Xq, int32_t *h_out,
                                 uint8_t *exact_mask, float *exact_dense, float *scale_out, dme_stream_t stream) {
    DME_REQUIRE(X && send_X && send_p && rot && Xq && h_out && exact_mask && exact_dense && scale_out, "null pointer argument");
    DME_REQUIRE(n >= 1 && n <= 65535 && d >= 1 && dpad >= d && dpad >= 4, "bad geometry");
    DME_REQUIRE(h_len >= 1 && h_len <= 256 && x_len >= 3 && (x_len & 1) && delta > 0.0f, "bad table geometry");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = fwht_rows(X, d, ld, rot, dpad, dpad, n, diag_inject, rotation_seed, 0, 1, 0, st);       // AS:464 / AS:468 (shared rotation)
    if (rc) return rc;
    const int nb = slices(dpad);
    AsyncScratch partial(st), nrm(st);
    DME_CUDA(partial.alloc(sizeof(double) * (size_t)(n * nb)));
    DME_CUDA(nrm.alloc(sizeof(float) * (size_t)n));
    row_sumsq_kernel<<<dim3((unsigned)nb, (unsigned)n), 256, 0, st>>>(rot, dpad, dpad, nb, (double *)partial.p);
    DME_LAUNCH_CHECK("row_sumsq_kernel");
    norm_finalize_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>((double *)partial.p, nb, n, nullptr, (float *)nrm.p);
    DME_LAUNCH_CHECK("norm_finalize_kernel");
    const float sq = (float)std::sqrt((double)dpad);
    quicfl_encode_kernel<<<dim3((unsigned)((dpad / 4 + 255) / 256), (unsigned)n), 256, 0, st>>>(
        rot, dpad, (const float *)nrm.p, sq, h_len, x_len, delta, exact_threshold, send_X, send_p, seed, client0, Xq, h_out, exact_mask,
        exact_dense, scale_out);
    DME_LAUNCH_CHECK("quicfl_encode_kernel");
    return DME_OK;
}

extern "C" int dme_scalar_quantize(const float *X, int64_t n, int64_t d, int64_t ld, float nlevels, uint64_t seed, uint64_t client0,
                                   const float *u_inject, float *out, int64_t ld_out, dme_stream_t stream) {
    DME_REQUIRE(X && out && n >= 1 && n <= 65535 && d >= 1 && ld >= d && ld_out >= d, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = slices_minmax(d);
    AsyncScratch pms(st);
    DME_CUDA(pms.alloc(sizeof(float) * (size_t)(2 * n * nb)));
    float *pm = (float *)pms.p;
    dim3 grid((unsigned)nb, (unsigned)n);
    row_minmax_kernel<<<grid, 256, 0, st>>>(X, d, ld, nb, pm, pm + n * nb);
    DME_LAUNCH_CHECK("row_minmax_kernel");
    const int vec_ok = (ld % 4 == 0) && (ld_out % 4 == 0) && (((uintptr_t)X | (uintptr_t)out) & 15u) == 0;
    const int64_t per_cta = 256 * kScalarGroups * 4;
    dim3 grid2((unsigned)((d + per_cta - 1) / per_cta), (unsigned)n);
    scalar_kernel<<<grid2, 256, 0, st>>>(X, n, d, ld, nb, pm, pm + n * nb, nlevels, seed, client0, u_inject, out, ld_out, vec_ok);
    DME_LAUNCH_CHECK("scalar_kernel");
    return DME_OK;
}
