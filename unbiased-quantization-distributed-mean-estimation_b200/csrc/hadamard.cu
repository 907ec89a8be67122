// hadamard.cu -- normalised natural-order Walsh-Hadamard transform (AS:100-115), the randomized transform and its
// inverse (AS:127-156), the Rademacher diagonal (AS:117-120, Philox instead of torch.Generator), and the
// reference's pair transform (AS:37-59 as it executes).
//
// Bit-exactness: the reference runs stages h = 2, 4, .., d; each butterfly is a' = a + b, b' = a' - 2b (one
// rounding), then divides by float(sqrt(d)).  Butterflies of one stage are independent, so any schedule that
// keeps the stage ORDER and these two operations reproduces the reference bit for bit:
//   fwht_contig_kernel : stages with distance 1 .. 2048 on contiguous 4096-blocks, radix-16 in registers, two
//                        shared-memory transposes (padded, conflict-free)
//   fwht_strided_wide_kernel: 5 to 8 further stages per pass (distance 2^s .. 2^(s+P-1)) on tiles of 2^P strided rows x
//                        4096 / 2^P contiguous columns, one shared-memory transpose
//   fwht_strided_kernel: up to 4 further stages (what is left when fewer than 5 remain), 16 strided elements per thread
// The +-1 diagonal, the zero padding and the final 1/sqrt(d) are fused into the first load / last store.
#include <cmath>

#include "common.cuh"

namespace dme {

__device__ __forceinline__ void bfly(float &a, float &b) {
    const float s = __fadd_rn(a, b);          // AS:110
    b = __fmaf_rn(-2.0f, b, s);               // AS:111: (a + b) - 2b, 2b exact -> one rounding
    a = s;
}
template <int NB>   // NB butterfly stages over the 2^NB-point register array, distance 1 first
__device__ __forceinline__ void reg_stages(float (&v)[16], int nb) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (b < nb) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (!(i & (1 << b))) bfly(v[i], v[i | (1 << b)]);
        }
    }
}

struct FwhtIo {
    const float *src; int64_t src_d, src_ld;   // source rows (length src_d <= d, zero padded)
    float *dst; int64_t d, dst_ld;             // destination rows (length d = power of two)
    int logd;                                  // d = 2^logd: rows and columns of a flattened index by shift and mask (a 64-bit division is a ~100-instruction call)
    const float *diag;                         // injected +-1 diagonal (d entries) or null
    uint64_t seed, seed_stride; int use_philox; // Philox diagonal of row r is keyed by seed + r * seed_stride
    int pre_diag, post_diag;                   // multiply by the diagonal before the first stage / after the scaling
    int finalize; float sq;                    // divide by sq = float(sqrt(d)) after the last stage
};

__device__ __forceinline__ float diag_at(const FwhtIo &io, int64_t col, int64_t row) {
    if (io.diag) return io.diag[col];
    return philox_sign(io.seed + (uint64_t)row * io.seed_stride, (uint64_t)col, kStreamDiag);
}

// Signs of 16 consecutive columns col .. col+15 (col a multiple of 16) of one row as a bit mask (bit j set: +1).  With the
// Philox diagonal they are 16 bits of ONE 32-bit word of one block (philox_sign's mapping: column i = bit i & 31 of word
// (i >> 5) & 3 of block i >> 7), so one Philox evaluation serves the thread's whole chunk.
__device__ __forceinline__ uint32_t diag_bits16(const FwhtIo &io, int64_t col, int64_t row) {
    if (io.diag) {
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) m |= (io.diag[col + j] > 0.0f ? 1u : 0u) << j;
        return m;
    }
    const uint64_t blk = (uint64_t)col >> 7;
    const Philox4 p = philox4x32_10(io.seed + (uint64_t)row * io.seed_stride, (uint32_t)blk, (uint32_t)(blk >> 32), 0u, kStreamDiag);
    const int wi = (int)((col >> 5) & 3);
    const uint32_t word = wi == 0 ? p.x : wi == 1 ? p.y : wi == 2 ? p.z : p.w;
    return (word >> (col & 31)) & 0xffffu;
}

constexpr int kPad = 4096 + 4096 / 32;   // +1 float every 32: transposes are conflict-free
__device__ __forceinline__ int padded(int e) { return e + (e >> 5); }

// One CTA = one 4096-element block of the flattened (row-major, dense) index space of n rows x d.
// logL = min(log2 d, 12) stages are done here.
__global__ void __launch_bounds__(256, 5) fwht_contig_kernel(FwhtIo io, int64_t n, int logL) {
    __shared__ float sm[kPad];
    const int t = threadIdx.x;
    // padded(e) = e + (e >> 5) for the four layouts of the block, as base + compile-time offset (the address arithmetic was 40 % of
    // the kernel's instructions): A: element 16 t + j; B: 256 hi + 16 j + lo; C: 256 j + t; S (coalesced pieces): 1024 q + 4 t + k
    float *const pA = sm + 16 * t + (t >> 1);                   // [j]
    float *const pB = sm + (t >> 4) * 264 + (t & 15);           // [16 j + (j >> 1)]
    float *const pC = sm + t + (t >> 5);                        // [264 j]
    float *const pS = sm + 4 * t + (t >> 3);                    // [1056 q + k]
    const int64_t g0 = (int64_t)blockIdx.x * 4096 + 16 * t;      // first flattened element of this thread
    const int64_t total = n * io.d;
    float v[16];
    // ---- load (+ zero padding, + diagonal)
    if (io.d >= 16) {
        const int64_t row = g0 >> io.logd, col = g0 & (io.d - 1);
        const bool live = g0 < total;
        const bool vec = ((io.src_ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(io.src) & 15) == 0);
        if (vec && io.d >= 4096 && io.src_d == io.d) {
            // the block is 4096 consecutive columns of ONE dense row: coalesced 128-bit loads (piece q * 256 + t of the block),
            // handed to their owners (thread t: elements 16 t .. 16 t + 15) through the padded shared-memory array
            const int64_t b0 = (int64_t)blockIdx.x * 4096;
            const int64_t brow = b0 >> io.logd;
            const float *bsrc = io.src + brow * io.src_ld + (b0 & (io.d - 1));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 x = *reinterpret_cast<const float4 *>(bsrc + 4 * (q * 256 + t));
                float *dstp = pS + 1056 * q;                  // the piece's four elements share one 32-block: no pad inside
                dstp[0] = x.x; dstp[1] = x.y; dstp[2] = x.z; dstp[3] = x.w;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = pA[j];
            __syncthreads();
        } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t c0 = col + 4 * q;
            if (live && c0 + 4 <= io.src_d && vec) {
                const float4 x = *reinterpret_cast<const float4 *>(io.src + row * io.src_ld + c0);
                v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[4 * q + j] = (live && c0 + j < io.src_d) ? io.src[row * io.src_ld + c0 + j] : 0.0f;
            }
        }
        }
        if (io.pre_diag && live) {
            const uint32_t sb = diag_bits16(io, col, row);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __fmul_rn(v[j], ((sb >> j) & 1u) ? 1.0f : -1.0f);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int64_t g = g0 + j;
            const int64_t row = g >> io.logd, col = g & (io.d - 1);
            float x = (g < total && col < io.src_d) ? io.src[row * io.src_ld + col] : 0.0f;
            if (io.pre_diag && g < total) x = __fmul_rn(x, diag_at(io, col, row));
            v[j] = x;
        }
    }
    // ---- phase A: element bits 0..3 (thread-local)
    reg_stages<4>(v, logL);
    if (logL > 4) {
        // ---- phase B: bits 4..7.  thread (lo = t & 15, hi = t >> 4) takes e = hi*256 + j*16 + lo
#pragma unroll
        for (int j = 0; j < 16; ++j) pA[j] = v[j];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = pB[16 * j + (j >> 1)];
        reg_stages<4>(v, logL - 4);
        if (logL > 8) {
            // ---- phase C: bits 8..11.  thread t takes e = j*256 + t
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 16; ++j) pB[16 * j + (j >> 1)] = v[j];
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = pC[264 * j];
            reg_stages<4>(v, logL - 8);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 16; ++j) pC[264 * j] = v[j];
        } else {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 16; ++j) pB[16 * j + (j >> 1)] = v[j];
        }
        __syncthreads();
        if (io.d >= 4096 && !(io.finalize && io.post_diag) && ((io.dst_ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(io.dst) & 15) == 0)) {
            // no per-column sign to apply: the block (4096 consecutive columns of one row) leaves as coalesced 128-bit pieces
            const int64_t b0 = (int64_t)blockIdx.x * 4096;
            const int64_t brow = b0 >> io.logd;
            float *bdst = io.dst + brow * io.dst_ld + (b0 & (io.d - 1));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int e = 4 * (q * 256 + t);
                const float *sp = pS + 1056 * q;
                float4 y = make_float4(sp[0], sp[1], sp[2], sp[3]);
                if (io.finalize) { y.x = __fdiv_rn(y.x, io.sq); y.y = __fdiv_rn(y.y, io.sq); y.z = __fdiv_rn(y.z, io.sq); y.w = __fdiv_rn(y.w, io.sq); }   // AS:113
                *reinterpret_cast<float4 *>(bdst + e) = y;
            }
            return;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = pA[j];
    }
    // ---- store (+ scaling, + diagonal)
    if (io.d >= 16) {
        // the thread's 16 elements are consecutive columns of one row: one row / column computation, one sign word, and
        // 128-bit stores when the destination allows
        if (g0 >= total) return;
        const int64_t row = g0 >> io.logd, col = g0 & (io.d - 1);
        uint32_t sb = 0xffffu;
        if (io.finalize && io.post_diag) sb = diag_bits16(io, col, row);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (io.finalize) {
                v[j] = __fdiv_rn(v[j], io.sq);                                                          // AS:113
                if (io.post_diag) v[j] = __fmul_rn(v[j], ((sb >> j) & 1u) ? 1.0f : -1.0f);                // AS:154
            }
        }
        float *drow = io.dst + row * io.dst_ld + col;
        if (((io.dst_ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(io.dst) & 15) == 0)) {
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<float4 *>(drow + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) drow[j] = v[j];
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int64_t g = g0 + j;
        if (g >= total) break;
        const int64_t row = g >> io.logd, col = g & (io.d - 1);
        float y = v[j];
        if (io.finalize) {
            y = __fdiv_rn(y, io.sq);                                        // AS:113
            if (io.post_diag) y = __fmul_rn(y, diag_at(io, col, row));      // AS:154
        }
        io.dst[row * io.dst_ld + col] = y;
    }
}

// Stages with distance 2^s .. 2^(s+p-1), p <= 4, s >= 12, in place on dst.  One thread = one column, 2^p rows.
__global__ void __launch_bounds__(256) fwht_strided_kernel(FwhtIo io, int64_t n, int s, int p) {
    const int64_t cols_per_row = io.d >> p;                     // independent (hi, col) groups per row
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * cols_per_row) return;
    const int64_t row = gid >> (io.logd - p), w = gid & (cols_per_row - 1);
    const int64_t col = w & (((int64_t)1 << s) - 1), hi = w >> s;
    float *base = io.dst + row * io.dst_ld + (hi << (s + p)) + col;
    const int R = 1 << p;
    float v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = (r < R) ? base[(int64_t)r << s] : 0.0f;
    reg_stages<4>(v, p);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        if (r < R) {
            float y = v[r];
            if (io.finalize) {
                y = __fdiv_rn(y, io.sq);
                if (io.post_diag) y = __fmul_rn(y, diag_at(io, (hi << (s + p)) + ((int64_t)r << s) + col, row));
            }
            base[(int64_t)r << s] = y;
        }
    }
}

// P = 5..8 stages with distance 2^s .. 2^(s+P-1), s >= 12, in place on dst: one CTA = a tile of 2^P strided rows x C = 4096 / 2^P
// consecutive columns (C * 4 >= 64 contiguous bytes per row: coalesced), 16 elements per thread: the first four stages in
// registers, one shared-memory transpose, the remaining P - 4 stages in registers again.  Same stage order and butterfly as
// AS:105-112: bit-exact.  With this pass the transform of d = 2^20 takes 2 passes over memory and d = 2^24 takes 3.
template <int P>
__global__ void __launch_bounds__(256) fwht_strided_wide_kernel(FwhtIo io, int64_t n, int s) {
    constexpr int R = 1 << P, C = 4096 / R, J = R / 16, K = 16 / J;
    constexpr int kStride = C + (C == 16 ? 1 : 0);          // C = 16: two rows per warp land in different bank halves
    __shared__ float sm[R * kStride];
    const int t = threadIdx.x, col_l = t % C, q = t / C;
    const int64_t tiles_per_row = io.d >> 12;
    const int64_t row = (int64_t)blockIdx.x >> (io.logd - 12), w = (int64_t)blockIdx.x & (tiles_per_row - 1);
    const int lcb = s - 12 + P;                             // log2 of the column blocks inside one 2^s run (2^s / C)
    const int64_t hi = w >> lcb, col0 = (w & (((int64_t)1 << lcb) - 1)) * C;
    float *base = io.dst + row * io.dst_ld + (hi << (s + P)) + col0 + col_l;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = base[(int64_t)(q * 16 + j) << s];
    reg_stages<4>(v, 4);
#pragma unroll
    for (int j = 0; j < 16; ++j) sm[(q * 16 + j) * kStride + col_l] = v[j];
    __syncthreads();
    // v[k * J + j2] = element r = j2 * 16 + (q * K + k)
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j2 = 0; j2 < J; ++j2) v[k * J + j2] = sm[(j2 * 16 + q * K + k) * kStride + col_l];
#pragma unroll
    for (int b = 0; b < P - 4; ++b)
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (!((i % J) & (1 << b))) bfly(v[i], v[i | (1 << b)]);
    if (io.finalize) {
        // Philox diagonal: the columns of a warp's row segment share one 32-bit word of one block, so lane e computes the word
        // of element e once and the warp passes it around
        uint32_t myword = 0;
        const bool philox = io.post_diag && io.diag == nullptr;
        if (philox) {
            const int e = (t & 15), k = e / J, j2 = e % J;
            const int64_t col = (hi << (s + P)) + ((int64_t)(j2 * 16 + q * K + k) << s) + col0 + col_l;
            const uint64_t blk = (uint64_t)col >> 7;
            const Philox4 ph = philox4x32_10(io.seed + (uint64_t)row * io.seed_stride, (uint32_t)blk, (uint32_t)(blk >> 32), 0u, kStreamDiag);
            const int wi = (int)((col >> 5) & 3);
            myword = wi == 0 ? ph.x : wi == 1 ? ph.y : wi == 2 ? ph.z : ph.w;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int k = i / J, j2 = i % J;
            const int64_t col = (hi << (s + P)) + ((int64_t)(j2 * 16 + q * K + k) << s) + col0 + col_l;
            float y = __fdiv_rn(v[i], io.sq);                                        // AS:113
            if (io.post_diag) {                                                      // AS:154
                float sg;
                if (philox) {
                    const uint32_t word = __shfl_sync(0xffffffffu, myword, (t & 16) | i);
                    sg = ((word >> (col & 31)) & 1u) ? 1.0f : -1.0f;
                } else sg = io.diag[col];
                y = __fmul_rn(y, sg);
            }
            v[i] = y;
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int k = i / J, j2 = i % J;
        base[(int64_t)(j2 * 16 + q * K + k) << s] = v[i];
    }
}

__global__ void rademacher_kernel(float *out, int64_t d, uint64_t seed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d) out[i] = philox_sign(seed, (uint64_t)i, kStreamDiag);
}

// AS:37-59 as executed: log2(len) times (a, b) -> (a + b, (a + b) - b) on every adjacent pair.
__global__ void pair_transform_kernel(float *V, int64_t n, int64_t len, int64_t ld, int stages) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t pairs = len >> 1;
    if (gid >= n * pairs) return;
    const int64_t row = gid / pairs, j = gid - row * pairs;
    float *p = V + row * ld + 2 * j;
    float a = p[0], b = p[1];
    for (int s = 0; s < stages; ++s) {
        const float t = __fadd_rn(a, b);     // AS:51
        b = __fsub_rn(t, b);                 // AS:52 (reads the already-updated a)
        a = t;
    }
    p[0] = a; p[1] = b;
}

static int ilog2(int64_t v) { int l = 0; while (((int64_t)1 << l) < v) ++l; return l; }

// Full transform of n rows: src (length src_d, stride src_ld) -> dst (length d = 2^k, stride dst_ld).
int fwht_rows(const float *src, int64_t src_d, int64_t src_ld, float *dst, int64_t d, int64_t dst_ld, int64_t n,
              const float *diag, uint64_t seed, uint64_t seed_stride, int pre_diag, int post_diag, cudaStream_t st) {
    DME_REQUIRE(d >= 1 && (d & (d - 1)) == 0, "input numel must be a power of 2");
    DME_REQUIRE(src && dst && n >= 1 && src_d >= 1 && src_d <= d && src_ld >= src_d && dst_ld >= d, "bad row geometry");
    DME_REQUIRE(n * d < ((int64_t)1 << 40), "problem too large");
    const int logd = ilog2(d);
    FwhtIo io;
    io.src = src; io.src_d = src_d; io.src_ld = src_ld; io.dst = dst; io.d = d; io.dst_ld = dst_ld;
    io.diag = diag; io.seed = seed; io.seed_stride = seed_stride; io.use_philox = diag == nullptr; io.pre_diag = pre_diag; io.post_diag = post_diag;
    io.sq = (float)std::sqrt((double)d);
    io.logd = logd;
    const int logL = logd < 12 ? logd : 12;
    io.finalize = (logd <= 12);
    const int64_t blocks = (n * d + 4095) / 4096;
    fwht_contig_kernel<<<(unsigned)blocks, 256, 0, st>>>(io, n, logL);
    DME_LAUNCH_CHECK("fwht_contig_kernel");
    io.src = dst; io.src_d = d; io.src_ld = dst_ld; io.pre_diag = 0;
    // the remaining logd - 12 stages: one pass of up to 8 stages (two for more than 8), so d <= 2^20 takes 2 passes over
    // memory and d <= 2^28 takes 3
    int s = 12;
    while (s < logd) {
        const int rem = logd - s;
        const int p = rem <= 8 ? rem : (rem <= 16 ? (rem + 1) / 2 : 8);
        io.finalize = (s + p >= logd);
        if (p >= 5) {
            const unsigned grid = (unsigned)(n * (d >> 12));
            switch (p) {
                case 5: fwht_strided_wide_kernel<5><<<grid, 256, 0, st>>>(io, n, s); break;
                case 6: fwht_strided_wide_kernel<6><<<grid, 256, 0, st>>>(io, n, s); break;
                case 7: fwht_strided_wide_kernel<7><<<grid, 256, 0, st>>>(io, n, s); break;
                default: fwht_strided_wide_kernel<8><<<grid, 256, 0, st>>>(io, n, s); break;
            }
            DME_LAUNCH_CHECK("fwht_strided_wide_kernel");
        } else {
            const int64_t threads = n * (d >> p);
            fwht_strided_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(io, n, s, p);
            DME_LAUNCH_CHECK("fwht_strided_kernel");
        }
        s += p;
    }
    return DME_OK;
}

}  // namespace dme

using namespace dme;

extern "C" int dme_hadamard(float *V, int64_t n, int64_t dpad, int64_t ld, dme_stream_t stream) {
    DME_REQUIRE(V != nullptr, "V is null");
    return fwht_rows(V, dpad, ld, V, dpad, ld, n, nullptr, 0, 0, 0, 0, (cudaStream_t)stream);
}
extern "C" int dme_rht(const float *X, int64_t n, int64_t d, int64_t ld, float *out, int64_t dpad, int64_t ld_out, uint64_t seed,
                       uint64_t seed_stride, const float *diag_inject, dme_stream_t stream) {
    DME_REQUIRE(X && out, "null pointer argument");
    DME_REQUIRE(dpad >= d, "dpad < d");
    return fwht_rows(X, d, ld, out, dpad, ld_out, n, diag_inject, seed, seed_stride, 1, 0, (cudaStream_t)stream);
}
extern "C" int dme_irht(float *V, int64_t n, int64_t dpad, int64_t ld, uint64_t seed, uint64_t seed_stride, const float *diag_inject,
                        dme_stream_t stream) {
    DME_REQUIRE(V != nullptr, "V is null");
    return fwht_rows(V, dpad, ld, V, dpad, ld, n, diag_inject, seed, seed_stride, 0, 1, (cudaStream_t)stream);
}
extern "C" int dme_rademacher(float *diag, int64_t dpad, uint64_t seed, dme_stream_t stream) {
    DME_REQUIRE(diag && dpad >= 1, "bad argument");
    rademacher_kernel<<<(unsigned)((dpad + 255) / 256), 256, 0, (cudaStream_t)stream>>>(diag, dpad, seed);
    DME_LAUNCH_CHECK("rademacher_kernel");
    return DME_OK;
}
extern "C" int dme_pair_transform(float *V, int64_t n, int64_t len, int64_t ld, dme_stream_t stream) {
    DME_REQUIRE(V && n >= 1 && len >= 1 && (len & (len - 1)) == 0 && ld >= len, "len must be a power of two, ld >= len");
    if (len < 2) return DME_OK;
    const int64_t work = n * (len >> 1);
    pair_transform_kernel<<<(unsigned)((work + 255) / 256), 256, 0, (cudaStream_t)stream>>>(V, n, len, ld, ilog2(len));
    DME_LAUNCH_CHECK("pair_transform_kernel");
    return DME_OK;
}
