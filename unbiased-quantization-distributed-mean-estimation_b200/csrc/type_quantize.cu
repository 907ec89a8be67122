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
//   quantize_warp_kernel (quantize_warp.cu): single pass over the rows in ticket order with a decoupled look-back.
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
// 8 CTAs per SM (32 registers): 1.37 -> 1.20 ms for 8 GiB (7.2 TB/s)
__global__ void __launch_bounds__(kThreads, 8)
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
    RowConstIn in;
    in.m = m; in.d = d; in.x_inject = x_inject; in.l1_inject = l1_inject; in.seed = seed; in.client0 = client0;
    in.consts = consts; in.l1_out = l1_out;
    if (threadIdx.x == 0) make_row_const(in, c, acc);
}

// ------------------------------------------------------------------ K7: decode + mean (tile-major)
// Per-client table of the values a coordinate can take: lut[c][k] = q(k) / n for k = 1..7 with
// q(k) = ((L1 * 1) * k) / m (AS:640) or (L1 * 1) * (k / m) (AS:687); the sign is applied by flipping the sign bit,
// which is exact.  Magnitudes above 7 are rare (heavy tails) and computed on the spot.
constexpr int kLut = 8;
__device__ __forceinline__ float deq_over_n(float L1f, float kf, float mf, float nf, int biased) {
    const float q = biased ? __fmul_rn(L1f, __fdiv_rn(kf, mf)) : __fdiv_rn(__fmul_rn(L1f, kf), mf);
    return __fdiv_rn(q, nf);                                                                       // ND:137
}
// Add one client's chunk with field width W (words already in registers).
template <int W>
__device__ __forceinline__ void add_fields(const uint32_t (&words)[W / 2], const float *__restrict__ lutc, const float (&lv)[kLut], float L1f,
                                           float mf, float nf, int biased, float (&acc)[kEpt]) {
    constexpr int kPerWord = 32 / W;
#pragma unroll
    for (int q = 0; q < W / 2; ++q) {
#pragma unroll
        for (int e = 0; e < kPerWord; ++e) {
            const int j = q * kPerWord + e;
            const uint32_t field = (W == 32) ? words[q] : ((words[q] >> (W * e)) & ((1u << W) - 1u));
            const uint32_t mag = (W == 32) ? (field & 0x7fffffffu) : (field & ((1u << (W - 1)) - 1u));
            const uint32_t sgn = (field >> (W - 1)) << 31;
            float v;
            if (W == 2) v = mag ? lv[1] : 0.0f;
            else if (W == 4) v = lv[mag];
            else v = mag < (uint32_t)kLut ? lutc[mag] : deq_over_n(L1f, (float)mag, mf, nf, biased);
            acc[j] = __fadd_rn(acc[j], __uint_as_float(__float_as_uint(v) ^ (mag ? sgn : 0u)));   // zero magnitudes add +0
        }
    }
}
template <int W>
__device__ __forceinline__ void decode_generic(const uint32_t *__restrict__ tw, int chunk, const float *__restrict__ lutc, float L1f,
                                               float mf, float nf, int biased, float (&acc)[kEpt]) {
    uint32_t words[W / 2];
#pragma unroll
    for (int q = 0; q < W / 2; ++q) words[q] = __ldg(tw + q * kCodeChunks + chunk);
    float lv[kLut];
#pragma unroll
    for (int k = 0; k < kLut; ++k) lv[k] = (W <= 4) ? lutc[k] : 0.0f;
    add_fields<W>(words, lutc, lv, L1f, mf, nf, biased, acc);
}

// One thread = one 16-coordinate chunk of a tile, all clients in order (est += q / n, ND:133-147), fp32.
// 64-thread CTAs, one per code tile, so that short rows still fill the GPU.  Clients are taken in batches of 8 whose code words
// are loaded together, one batch ahead of the adds.
constexpr int kBatch = 8;
constexpr int kLutClients = 128;     // clients whose tables are staged in shared memory at a time
constexpr int kSignLutBytes = 8192;  // dynamic shared memory: a 4 KB table at a 4 KB-aligned address

// packed f32x2 (sm_100: one FFMA2 for two coordinates)
typedef unsigned long long pf2;
__device__ __forceinline__ pf2 pf2_pack(float lo, float hi) { pf2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void pf2_unpack(pf2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ pf2 pf2_fma(pf2 a, pf2 b, pf2 c) { pf2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ pf2 pf2_lds(uint32_t addr) { pf2 v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr)); return v; }
__device__ __forceinline__ float f32_lds(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }

// Fast paths (a tile's field width is the same for the whole CTA, so the choice does not diverge):
//  * W = 2, fields [sign | magnitude bit]: a coordinate adds 0 or +-v, v = q(1) / n.  Four bits of the code word (two
//    coordinates) index a table of sign pairs (s0, s1), s in {0, +1, -1}; one FFMA2 does acc += s * v for both -- exact:
//    s * v is exact and the sum rounds once, like the FADD of ND:137 (0 * v = +0 leaves acc unchanged; acc is never -0).
//    The table has one copy per lane (entry i of lane l at i * 256 + l * 8 bytes): a warp's 64-bit loads never conflict
//    whatever the code bits are, and the 4 KB alignment lets one LOP3 form the address.  Needs a finite v.
//  * W = 4, fields [sign | 3-bit magnitude]: the field indexes the client's 16 signed values q(k) / n directly (64 bytes
//    per client: 16 entries in 16 banks, equal entries broadcast -- conflict-free), one FADD per coordinate.
//  * anything else (wide fields of heavy-tailed rows, m = 0): the batch is decoded field by field (decode_generic).
constexpr uint32_t kWSlow = 255u;     // staged width code of a client whose tile needs decode_generic
struct DecBatch { uint32_t wa[kBatch], wb[kBatch], wc[kBatch]; };
// Start the loads of one batch: word 0 of 2- and 4-bit tiles, word 1 of 4-bit tiles (sinfo = {offset / 16, width code}).
__device__ __forceinline__ void dec_fetch(DecBatch &b, const uint2 *sinfo, const uint32_t *__restrict__ cptr) {
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
        const uint2 inf = sinfo[u];
        const uint32_t *tw = cptr + (unsigned long long)inf.x * 4ull;
        b.wc[u] = inf.y;
        b.wa[u] = (inf.y == 2u || inf.y == 4u) ? __ldg(tw) : 0u;
        b.wb[u] = (inf.y == 4u) ? __ldg(tw + kCodeChunks) : 0u;
    }
}

// Entry i = 8 c + k of a block of clients starting at cb: table[c][k] = q(k) / n, table[c][8 + k] = -q(k) / n (k = 0: +0 both).
__device__ __forceinline__ void decode_lut_entry(const float *__restrict__ l1, int64_t cb, int i, float mf, float nf, int biased, float *table) {
    const int k = i & (kLut - 1), c = i >> 3;
    const float v = k == 0 ? 0.0f : deq_over_n(__ldg(l1 + cb + c), (float)k, mf, nf, biased);
    table[c * 16 + k] = v;
    table[c * 16 + 8 + k] = k == 0 ? 0.0f : __uint_as_float(__float_as_uint(v) ^ 0x80000000u);
}
// The same tables for all n clients, once per fused call (workspace, WsLayout::off_lut): the decode kernels' CTAs (one or four per code
// tile) otherwise each repeat the 2 n IEEE divisions per entry -- a sixth of decode_mean_kernel at n = 128, half of
// decode_mean_short_kernel at n = 1000.
__global__ void decode_lut_kernel(const float *__restrict__ l1, int64_t n, float mf, float nf, int biased, float *__restrict__ lut) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * kLut) return;
    const int64_t cb = (i >> 3) & ~(int64_t)127;            // blocks of 128 clients, so that the in-block entry index fits an int
    decode_lut_entry(l1, cb, (int)(i - cb * kLut), mf, nf, biased, lut + cb * 16);
}

__global__ void __launch_bounds__(64, 8)
decode_mean_kernel(const uint32_t *__restrict__ codes, const uint64_t *__restrict__ dir, const float *__restrict__ l1,
                   int64_t n, int64_t d, int64_t T, float mf, float nf, int biased, float *__restrict__ mean, int accumulate,
                   int64_t tile0, const float *__restrict__ glut) {
    __shared__ __align__(64) float lut16[kLutClients * 16];    // [client][sign << 3 | k] = +-q(k) / n for k < 8
    __shared__ __align__(8) uint2 sinfo[kLutClients + kBatch]; // per staged client: {code offset / 16, width code}
    __shared__ uint32_t sslow[kLutClients / kBatch];           // per batch: some client needs decode_generic
    extern __shared__ unsigned char dec_dyn[];
    const uint32_t slut = ((uint32_t)__cvta_generic_to_shared(dec_dyn) + 4095u) & ~4095u;
    for (int i = threadIdx.x; i < 512; i += 64) {
        const int idx = i >> 5, f0 = idx & 3, f1 = idx >> 2;
        const float s0 = (f0 & 1) ? ((f0 & 2) ? -1.0f : 1.0f) : 0.0f, s1 = (f1 & 1) ? ((f1 & 2) ? -1.0f : 1.0f) : 0.0f;
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(slut + (uint32_t)i * 8u), "f"(s0), "f"(s1) : "memory");
    }
    const uint32_t lanebase = slut | ((threadIdx.x & 31u) << 3);
    const uint32_t lut16_base = (uint32_t)__cvta_generic_to_shared(lut16);
    const int64_t t = tile0 + blockIdx.x;                  // one CTA = one code tile (64 chunks of 16 coordinates)
    const int chunk = threadIdx.x;
    const uint32_t *cptr = codes + chunk;
    const int64_t i0 = t * kCodeTile + (int64_t)chunk * kEpt;
    const bool live = i0 < d;
    pf2 acc2[kEpt / 2];
#pragma unroll
    for (int p = 0; p < kEpt / 2; ++p)
        acc2[p] = pf2_pack((accumulate && i0 + 2 * p < d) ? mean[i0 + 2 * p] : 0.0f, (accumulate && i0 + 2 * p + 1 < d) ? mean[i0 + 2 * p + 1] : 0.0f);
    for (int64_t cb = 0; cb < n; cb += kLutClients) {
        const int nc = (int)((cb + kLutClients < n) ? kLutClients : n - cb);
        __syncthreads();
        if (glut) {
            // the fused call computed the tables once (decode_lut_kernel): every CTA copies instead of dividing
            const float4 *src = reinterpret_cast<const float4 *>(glut + cb * 16);
            for (int i = threadIdx.x; i < nc * 4; i += 64) reinterpret_cast<float4 *>(lut16)[i] = __ldg(src + i);
        } else {
            for (int i = threadIdx.x; i < nc * kLut; i += 64) decode_lut_entry(l1, cb, i, mf, nf, biased, lut16);
        }
        for (int i = threadIdx.x; i < kLutClients / kBatch; i += 64) sslow[i] = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < kLutClients + kBatch; i += 64) {
            uint2 inf = make_uint2(0u, 0u);
            if (i < nc) {
                const uint64_t e = __ldg(dir + (cb + i) * T + t);
                uint32_t W = (uint32_t)(e & 0xffu);
                // the sign-table path multiplies: it needs a finite q(1) / n (m = 0 gives 0/0: decode_generic adds it as is);
                // offsets that do not fit 32 bits take the generic path too
                if (W > 4u || (e >> 40) != 0ull || (W == 2u && !(fabsf(lut16[i * 16 + 1]) <= 3.0e38f))) { W = kWSlow; atomicOr(&sslow[i / kBatch], 1u); }
                inf = make_uint2((uint32_t)(e >> 8), W);
            }
            sinfo[i] = inf;
        }
        __syncthreads();
        if (!live) continue;
        // software pipeline over batches of kBatch clients: the code words of the next batch are in flight while this
        // batch is added (one round trip to L2 / HBM per batch would otherwise sit on every thread's critical path)
        DecBatch nx;
        dec_fetch(nx, sinfo, cptr);
        for (int c0 = 0; c0 < nc; c0 += kBatch) {
            const DecBatch cur = nx;
            dec_fetch(nx, sinfo + c0 + kBatch, cptr);          // past the last client: {0, 0} entries, no loads
            if (sslow[c0 / kBatch] == 0u) {
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const uint32_t cbase = lut16_base + (uint32_t)(c0 + u) * 64u;
                    if (cur.wc[u] == 2u) {
                        const uint32_t w = cur.wa[u];
                        const float v = f32_lds(cbase + 4u);
                        const pf2 vv = pf2_pack(v, v);
#pragma unroll
                        for (int p = 0; p < kEpt / 2; ++p) {
                            const uint32_t sh = (p < 2) ? (w << (8 - 4 * p)) : (w >> (4 * p - 8));
                            acc2[p] = pf2_fma(pf2_lds((sh & 0xf00u) | lanebase), vv, acc2[p]);
                        }
                    } else if (cur.wc[u] == 4u) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const uint32_t w = q ? cur.wb[u] : cur.wa[u];
#pragma unroll
                            for (int p = 0; p < 4; ++p) {
                                float lo, hi;
                                pf2_unpack(acc2[4 * q + p], lo, hi);
                                const uint32_t s0 = (p == 0) ? (w << 2) : (w >> (8 * p - 2)), s1 = w >> (8 * p + 2);
                                lo = __fadd_rn(lo, f32_lds((s0 & 0x3cu) | cbase));
                                hi = __fadd_rn(hi, f32_lds((s1 & 0x3cu) | cbase));
                                acc2[4 * q + p] = pf2_pack(lo, hi);
                            }
                        }
                    }
                }
            } else {
                float acc[kEpt];
#pragma unroll
                for (int p = 0; p < kEpt / 2; ++p) pf2_unpack(acc2[p], acc[2 * p], acc[2 * p + 1]);
                for (int u = 0; u < kBatch; ++u) {
                    if (c0 + u >= nc) break;
                    const uint64_t e = __ldg(dir + (cb + c0 + u) * T + t);
                    const int W = (int)(e & 0xffu);
                    const uint32_t *tw = codes + (e >> 8) * 4ull;
                    const float *lutc = lut16 + (c0 + u) * 16;
                    const float L1c = __ldg(l1 + cb + c0 + u);
                    if (W == 2) decode_generic<2>(tw, chunk, lutc, 0.0f, mf, nf, biased, acc);
                    else if (W == 4) decode_generic<4>(tw, chunk, lutc, 0.0f, mf, nf, biased, acc);
                    else if (W == 8) decode_generic<8>(tw, chunk, lutc, L1c, mf, nf, biased, acc);
                    else if (W == 16) decode_generic<16>(tw, chunk, lutc, L1c, mf, nf, biased, acc);
                    else if (W == 32) decode_generic<32>(tw, chunk, lutc, L1c, mf, nf, biased, acc);
                    // W == 0: the tile was dropped (arena exhausted; status bit 2 is set)
                }
#pragma unroll
                for (int p = 0; p < kEpt / 2; ++p) acc2[p] = pf2_pack(acc[2 * p], acc[2 * p + 1]);
            }
        }
    }
    if (!live) return;
    float acc[kEpt];
#pragma unroll
    for (int p = 0; p < kEpt / 2; ++p) pf2_unpack(acc2[p], acc[2 * p], acc[2 * p + 1]);
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

// Short rows (few code tiles, many clients: BASELINE config 2): the same walk over the clients in order, but FOUR coordinates per
// thread -- four threads share a chunk's code word (one broadcast load) and four 64-thread CTAs share a tile, so d / 1024 tiles
// give 4x the CTAs.  Same arithmetic per coordinate as decode_mean_kernel: bit-identical results.
constexpr int kCptS = 4;
__global__ void __launch_bounds__(64, 8)
decode_mean_short_kernel(const uint32_t *__restrict__ codes, const uint64_t *__restrict__ dir, const float *__restrict__ l1,
                         int64_t n, int64_t d, int64_t T, float mf, float nf, int biased, float *__restrict__ mean, int accumulate,
                         int64_t tile0, const float *__restrict__ glut) {
    __shared__ __align__(64) float lut16[kLutClients * 16];
    __shared__ __align__(8) uint2 sinfo[kLutClients + kBatch];
    __shared__ uint32_t sslow[kLutClients / kBatch];
    extern __shared__ unsigned char dec_dyn[];
    const uint32_t slut = ((uint32_t)__cvta_generic_to_shared(dec_dyn) + 4095u) & ~4095u;
    for (int i = threadIdx.x; i < 512; i += 64) {
        const int idx = i >> 5, f0 = idx & 3, f1 = idx >> 2;
        const float s0 = (f0 & 1) ? ((f0 & 2) ? -1.0f : 1.0f) : 0.0f, s1 = (f1 & 1) ? ((f1 & 2) ? -1.0f : 1.0f) : 0.0f;
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(slut + (uint32_t)i * 8u), "f"(s0), "f"(s1) : "memory");
    }
    const uint32_t lanebase = slut | ((threadIdx.x & 31u) << 3);
    const uint32_t lut16_base = (uint32_t)__cvta_generic_to_shared(lut16);
    const int64_t t = tile0 + (blockIdx.x >> 2);
    const int chunk = (int)(blockIdx.x & 3) * 16 + (int)(threadIdx.x >> 2), sub = threadIdx.x & 3;     // fields [4 sub, 4 sub + 4) of the chunk
    const uint32_t *cptr = codes + chunk;
    const int64_t i0 = t * kCodeTile + (int64_t)chunk * kEpt + sub * kCptS;
    const bool live = i0 < d;
    pf2 acc2[2];
#pragma unroll
    for (int p = 0; p < 2; ++p)
        acc2[p] = pf2_pack((accumulate && i0 + 2 * p < d) ? mean[i0 + 2 * p] : 0.0f, (accumulate && i0 + 2 * p + 1 < d) ? mean[i0 + 2 * p + 1] : 0.0f);
    for (int64_t cb = 0; cb < n; cb += kLutClients) {
        const int nc = (int)((cb + kLutClients < n) ? kLutClients : n - cb);
        __syncthreads();
        if (glut) {
            // the fused call computed the tables once (decode_lut_kernel): every CTA copies instead of dividing
            const float4 *src = reinterpret_cast<const float4 *>(glut + cb * 16);
            for (int i = threadIdx.x; i < nc * 4; i += 64) reinterpret_cast<float4 *>(lut16)[i] = __ldg(src + i);
        } else {
            for (int i = threadIdx.x; i < nc * kLut; i += 64) decode_lut_entry(l1, cb, i, mf, nf, biased, lut16);
        }
        for (int i = threadIdx.x; i < kLutClients / kBatch; i += 64) sslow[i] = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < kLutClients + kBatch; i += 64) {
            uint2 inf = make_uint2(0u, 0u);
            if (i < nc) {
                const uint64_t e = __ldg(dir + (cb + i) * T + t);
                uint32_t W = (uint32_t)(e & 0xffu);
                if (W > 4u || (e >> 40) != 0ull || (W == 2u && !(fabsf(lut16[i * 16 + 1]) <= 3.0e38f))) { W = kWSlow; atomicOr(&sslow[i / kBatch], 1u); }
                inf = make_uint2((uint32_t)(e >> 8), W);
            }
            sinfo[i] = inf;
        }
        __syncthreads();
        if (!live) continue;
        DecBatch nx;
        dec_fetch(nx, sinfo, cptr);
        for (int c0 = 0; c0 < nc; c0 += kBatch) {
            const DecBatch cur = nx;
            dec_fetch(nx, sinfo + c0 + kBatch, cptr);
            if (sslow[c0 / kBatch] == 0u) {
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const uint32_t cbase = lut16_base + (uint32_t)(c0 + u) * 64u;
                    if (cur.wc[u] == 2u) {
                        const uint32_t w = cur.wa[u] >> (8 * sub);          // the thread's four 2-bit fields
                        const float v = f32_lds(cbase + 4u);
                        const pf2 vv = pf2_pack(v, v);
                        acc2[0] = pf2_fma(pf2_lds(((w << 8) & 0xf00u) | lanebase), vv, acc2[0]);
                        acc2[1] = pf2_fma(pf2_lds(((w << 4) & 0xf00u) | lanebase), vv, acc2[1]);
                    } else if (cur.wc[u] == 4u) {
                        const uint32_t w = ((sub & 2) ? cur.wb[u] : cur.wa[u]) >> (16 * (sub & 1));      // the thread's four nibbles
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            float lo, hi;
                            pf2_unpack(acc2[p], lo, hi);
                            lo = __fadd_rn(lo, f32_lds((((w >> (8 * p)) & 0xfu) << 2) | cbase));
                            hi = __fadd_rn(hi, f32_lds((((w >> (8 * p + 4)) & 0xfu) << 2) | cbase));
                            acc2[p] = pf2_pack(lo, hi);
                        }
                    }
                }
            } else {
                // wide fields: decode the whole chunk with this thread's accumulators in place, keep its four coordinates
                float acc[kEpt];
#pragma unroll
                for (int j = 0; j < kEpt; ++j) acc[j] = 0.0f;
                float m4[4];
                pf2_unpack(acc2[0], m4[0], m4[1]); pf2_unpack(acc2[1], m4[2], m4[3]);
                for (int u = 0; u < kBatch; ++u) {
                    if (c0 + u >= nc) break;
                    const uint64_t e = __ldg(dir + (cb + c0 + u) * T + t);
                    const int W = (int)(e & 0xffu);
                    const uint32_t *tw = codes + (e >> 8) * 4ull;
                    const float *lutc = lut16 + (c0 + u) * 16;
                    const float L1c = __ldg(l1 + cb + c0 + u);
#pragma unroll
                    for (int j = 0; j < kEpt; ++j) acc[j] = 0.0f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {                 // place the running sums where decode_generic adds this thread's fields
                        if (sub == 0) acc[j] = m4[j]; else if (sub == 1) acc[4 + j] = m4[j]; else if (sub == 2) acc[8 + j] = m4[j]; else acc[12 + j] = m4[j];
                    }
                    if (W == 2) decode_generic<2>(tw, chunk, lutc, 0.0f, mf, nf, biased, acc);
                    else if (W == 4) decode_generic<4>(tw, chunk, lutc, 0.0f, mf, nf, biased, acc);
                    else if (W == 8) decode_generic<8>(tw, chunk, lutc, L1c, mf, nf, biased, acc);
                    else if (W == 16) decode_generic<16>(tw, chunk, lutc, L1c, mf, nf, biased, acc);
                    else if (W == 32) decode_generic<32>(tw, chunk, lutc, L1c, mf, nf, biased, acc);
#pragma unroll
                    for (int j = 0; j < 4; ++j) m4[j] = sub == 0 ? acc[j] : sub == 1 ? acc[4 + j] : sub == 2 ? acc[8 + j] : acc[12 + j];
                }
                acc2[0] = pf2_pack(m4[0], m4[1]); acc2[1] = pf2_pack(m4[2], m4[3]);
            }
        }
    }
    if (!live) return;
    float m4[4];
    pf2_unpack(acc2[0], m4[0], m4[1]); pf2_unpack(acc2[1], m4[2], m4[3]);
    if (i0 + kCptS <= d) *reinterpret_cast<float4 *>(mean + i0) = make_float4(m4[0], m4[1], m4[2], m4[3]);
    else {
#pragma unroll
        for (int j = 0; j < kCptS; ++j)
            if (i0 + j < d) mean[i0 + j] = m4[j];
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
    if (need_desc) {
        DME_CUDA(cudaMemsetAsync(base + L.off_desc, 0, (size_t)L.desc_bytes, st));
    }
    if (need_sel) {
        DME_CUDA(cudaMemsetAsync(base + L.off_sel, 0, sizeof(RowSelect) * (size_t)n, st));
        DME_CUDA(cudaMemsetAsync(base + L.off_lin, 0, (size_t)L.lin_hist_bytes, st));
    }
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
void set_biased_path(int path);                                                                        // reznik.cu
int launch_quantize_warp(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, const WsLayout &L, void *ws,
                          int32_t *k_out, uint8_t *sgn_out, float *deq_out, int64_t ld_out,
                          uint32_t *codes, int64_t codes_bytes, uint64_t *dir, cudaStream_t st, bool packed);   // quantize_warp.cu
int launch_literal_rows(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, const WsLayout &L, void *ws,
                        const float *x_inject, const float *l1_inject, uint64_t seed, uint64_t client0,
                        int32_t *k_out, uint8_t *sgn_out, float *deq_out, int64_t ld_out,
                        uint32_t *codes, int64_t codes_bytes, uint64_t *dir, float *l1_out, cudaStream_t st, bool packed);   // quantize_literal.cu
// Which implementation quantises the unbiased mode: 0 = l1_kernel + quantize_warp_kernel (the product path),
// 1 = literal_rows_kernel (tests: an independent implementation).  dme_set_unbiased_path.
static int g_unbiased_path = 0;

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
    const bool literal = mode == DME_MODE_UNBIASED && g_unbiased_path == 1;
    rc = ws_prepare(ws, ws_bytes, n, d, st, &L, !literal, mode == DME_MODE_BIASED);
    if (rc) return rc;
    if (literal) {
        rc = launch_literal_rows(X, n, d, ld, m, L, ws, x_inject, l1_inject, seed, client0, k_out, sgn_out, deq_out, ld_out, codes, codes_bytes,
                                 dir, l1_out, st, packed);
        if (rc) return rc;
            return DME_OK;
    }
    rc = launch_l1(X, n, d, ld, m, L, ws, x_inject, l1_inject, seed, client0, l1_out, st);
    if (rc) return rc;
    if (mode == DME_MODE_BIASED)
        return biased_quantize(X, n, d, ld, m, L, ws, k_out, sgn_out, deq_out, ld_out, codes, codes_bytes, dir, st);
    rc = launch_quantize_warp(X, n, d, ld, m, L, ws, k_out, sgn_out, deq_out, ld_out, codes, codes_bytes, dir, st, packed);
    if (rc) return rc;
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

static int decode_launch(const void *codes, const uint64_t *dir, const float *l1, int64_t n, int64_t d, int64_t m, int mode,
                         int64_t n_total, float *mean, int accumulate, int64_t tile0, int64_t tiles, dme_stream_t stream, const float *glut) {
    DME_REQUIRE(codes && dir && l1 && mean, "null pointer argument");
    DME_REQUIRE(n >= 1 && d >= 1 && m >= 1 && n_total >= 1, "n, d, m, n_total must be >= 1");
    DME_REQUIRE(((uintptr_t)mean & 15u) == 0 && ((uintptr_t)codes & 15u) == 0, "mean and codes must be 16-byte aligned");
    const int64_t T = (d + kCodeTile - 1) / kCodeTile;
    DME_REQUIRE(tile0 >= 0 && tiles >= 0 && tile0 + tiles <= T, "tile range [%lld, %lld) outside [0, %lld)", (long long)tile0,
                (long long)(tile0 + tiles), (long long)T);
    if (tiles > 0 && tiles <= 256) {
        // few tiles: four coordinates per thread, four CTAs per tile (the client walk is what takes the time: spread it wider)
        decode_mean_short_kernel<<<(unsigned)(4 * tiles), 64, kSignLutBytes, (cudaStream_t)stream>>>(
            (const uint32_t *)codes, dir, l1, n, d, T, (float)m, (float)n_total, mode == DME_MODE_BIASED, mean, accumulate, tile0, glut);
        DME_LAUNCH_CHECK("decode_mean_short_kernel");
    } else if (tiles > 0) {
        decode_mean_kernel<<<(unsigned)tiles, kCodeChunks, kSignLutBytes, (cudaStream_t)stream>>>(
            (const uint32_t *)codes, dir, l1, n, d, T, (float)m, (float)n_total, mode == DME_MODE_BIASED, mean, accumulate, tile0, glut);
        DME_LAUNCH_CHECK("decode_mean_kernel");
    }
    return DME_OK;
}

extern "C" int dme_decode_mean_tiles(const void *codes, const uint64_t *dir, const float *l1, int64_t n, int64_t d, int64_t m, int mode,
                                     int64_t n_total, float *mean, int accumulate, int64_t tile0, int64_t tiles, dme_stream_t stream) {
    return decode_launch(codes, dir, l1, n, d, m, mode, n_total, mean, accumulate, tile0, tiles, stream, nullptr);
}

extern "C" int dme_decode_mean(const void *codes, const uint64_t *dir, const float *l1, int64_t n, int64_t d, int64_t m, int mode,
                               int64_t n_total, float *mean, int accumulate, dme_stream_t stream) {
    return dme_decode_mean_tiles(codes, dir, l1, n, d, m, mode, n_total, mean, accumulate, 0, d >= 1 ? (d + kCodeTile - 1) / kCodeTile : 0, stream);
}

extern "C" int dme_set_biased_path(int path) {
    DME_REQUIRE(path == 0 || path == 1, "path=%d unknown", path);
    set_biased_path(path);
    return DME_OK;
}

extern "C" int dme_set_unbiased_path(int path) {
    DME_REQUIRE(path == 0 || path == 1, "path=%d unknown", path);
    g_unbiased_path = path;
    return DME_OK;
}

extern "C" int dme_quantize_mean(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, int mode, const float *x_inject,
                                 uint64_t seed, uint64_t client0, int64_t n_total, float *mean, int accumulate, void *codes,
                                 int64_t codes_bytes, uint64_t *dir, float *l1_out, void *ws, int64_t ws_bytes, dme_stream_t stream) {
    int rc = dme_type_encode(X, n, d, ld, m, mode, x_inject, nullptr, seed, client0, codes, codes_bytes, dir, l1_out, ws, ws_bytes, stream);
    if (rc) return rc;
    DME_REQUIRE(n_total >= 1, "n_total must be >= 1");
    // the decoder's per-client value tables, once for all of its CTAs (the workspace has room for them: WsLayout::off_lut)
    float *lut = (float *)((char *)ws + ws_layout(n, d).off_lut);
    decode_lut_kernel<<<(unsigned)((n * kLut + 255) / 256), 256, 0, (cudaStream_t)stream>>>(l1_out, n, (float)m, (float)n_total, mode == DME_MODE_BIASED, lut);
    DME_LAUNCH_CHECK("decode_lut_kernel");
    return decode_launch(codes, dir, l1_out, n, d, m, mode, n_total, mean, accumulate, 0, (d + kCodeTile - 1) / kCodeTile, stream, lut);
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
    if (st & 4u) { set_error("biased selection: a row's threshold bin overflowed its candidate list (tie-heavy row): rerun with dme_set_biased_path(1)"); return DME_ERETRY; }
    if (st & 2u) { set_error("code arena exhausted: enlarge codes_bytes (dme_codes_bytes(..., expect=0) is always enough)"); return DME_EWORKSPACE; }
    if (st & 1u) { set_error("a magnitude does not fit the requested integer output"); return DME_EOVERFLOW; }
    return DME_OK;
}
