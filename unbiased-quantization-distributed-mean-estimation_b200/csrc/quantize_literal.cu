// quantize_literal.cu -- the unbiased type quantizer (AS:609-641) evaluated LITERALLY: IEEE division, floorf, fp64 prefix,
// AS:636 as written.  One CTA per client row, tiles in order, no look-back, no closed forms.  It is NOT the product path
// (quantize_warp.cu is): dme_set_unbiased_path(1) selects it so that the GPU tests can run an independent second
// implementation against the same oracle and the same goldens.  Slow by construction (a 2^24-coordinate row takes milliseconds).
#include "type_quantize.cuh"

namespace dme {

struct LitArgs {
    const float *X; int64_t d, ld, T, n, m;
    WsHeader *hdr;
    const float *x_inject; const float *l1_inject; uint64_t seed, client0; float *l1_out;
    int32_t *k_out; uint8_t *sgn_out; float *deq_out; int64_t ld_out;
    PackTarget pack; int packed;
};

// AS:636 literally, for one prefix value
__device__ __forceinline__ int floor_ref(double c, float X) { return __float2int_rd(__fsub_rn(__double2float_rn(c), X)); }

__global__ void __launch_bounds__(kThreads)
literal_rows_kernel(const LitArgs a) {
    __shared__ double s_red[kWarps];
    __shared__ double s_scan[kThreads];
    __shared__ double s_wtot[kWarps];
    __shared__ PackScratch s_ps;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int64_t c = blockIdx.x; c < a.n; c += gridDim.x) {
        const float *row = a.X + c * a.ld;
        // ---- AS:624: L1 in fp64, fixed association (thread-strided, then the block tree)
        double s = 0.0;
        for (int64_t i = tid; i < a.d; i += kThreads) s += (double)fabsf(row[i]);
        s = block_sum_f64(s, s_red);
        const float L1f = a.l1_inject ? a.l1_inject[c] : (float)s;
        const float D = __fadd_rn(L1f, 1e-12f);                                                        // AS:625
        const float mf = (float)a.m;
        const float X = a.x_inject ? a.x_inject[c] : philox_client_uniform(a.seed, a.client0 + (uint64_t)c);   // AS:634
        if (tid == 0 && a.l1_out) a.l1_out[c] = L1f;
        double base = 0.0;                      // fp64 prefix before the tile
        for (int64_t t = 0; t < a.T; ++t) {
            const int64_t i0 = t * kTile + (int64_t)tid * kEpt;
            float x[kEpt], fl[kEpt], fr[kEpt];
#pragma unroll
            for (int q = 0; q < kEpt; ++q) x[q] = (i0 + q < a.d) ? row[i0 + q] : 0.0f;
            double run = 0.0;
#pragma unroll
            for (int q = 0; q < kEpt; ++q) {
                const float v = __fdiv_rn(x[q], D);                    // AS:625
                const float mp = __fmul_rn(mf, fabsf(v));              // AS:626-629
                fl[q] = floorf(mp);                                    // AS:630
                fr[q] = __fsub_rn(mp, fl[q]);                          // AS:631
                run += (double)fr[q];
            }
            // inclusive scan of the thread sums over the CTA, fixed association
            double incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            __syncthreads();
            if (lane == 31) s_wtot[warp] = incl;
            __syncthreads();
            double wb = 0.0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w)
                if (w < warp) wb += s_wtot[w];
            s_scan[tid] = wb + incl;
            __syncthreads();
            // the prefix at a thread's last coordinate is DEFINED as base + scan value: the next thread starts from that very number
            const double E = base + (tid > 0 ? s_scan[tid - 1] : 0.0), En = base + s_scan[tid];
            base = base + s_scan[kThreads - 1];
            uint32_t k[kEpt], sg[kEpt];
            bool ovf = false;
            double cp = E;
            int tp = floor_ref(cp, X);
#pragma unroll
            for (int q = 0; q < kEpt; ++q) {
                cp = (q < kEpt - 1) ? cp + (double)fr[q] : En;         // AS:635
                const int tt = floor_ref(cp, X);                       // AS:636
                const float kf = __fadd_rn(fl[q], (tt - tp == 1) ? 1.0f : 0.0f);     // AS:637-638
                tp = tt;
                sg[q] = __float_as_uint(x[q]) >> 31;
                if (kf >= 2147483648.0f) { ovf = true; k[q] = 0x7fffffffu; }
                else k[q] = (uint32_t)kf;
                if (!a.packed && i0 + q < a.d) {
                    const int64_t o = c * a.ld_out + i0 + q;
                    if (a.deq_out) {
                        const float v = __fdiv_rn(x[q], D);
                        const float sgf = (v > 0.0f) ? 1.0f : (v < 0.0f) ? -1.0f : 0.0f;                    // sign(v), AS:640
                        a.deq_out[o] = __fdiv_rn(__fmul_rn(__fmul_rn(L1f, sgf), kf), mf);
                    }
                    if (a.k_out) a.k_out[o] = (int32_t)k[q];
                    if (a.sgn_out) a.sgn_out[o] = (uint8_t)sg[q];
                }
            }
            if (a.packed) emit_packed_tile(a.pack, c, t, k, sg, ovf, s_ps);
            else if (ovf) atomicOr(&a.hdr->status, 1u);
        }
    }
}

int launch_literal_rows(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, const WsLayout &L, void *ws,
                        const float *x_inject, const float *l1_inject, uint64_t seed, uint64_t client0,
                        int32_t *k_out, uint8_t *sgn_out, float *deq_out, int64_t ld_out,
                        uint32_t *codes, int64_t codes_bytes, uint64_t *dir, float *l1_out, cudaStream_t st, bool packed) {
    char *base = (char *)ws;
    LitArgs a;
    a.X = X; a.d = d; a.ld = ld; a.T = L.T; a.n = n; a.m = m;
    a.hdr = (WsHeader *)base;
    a.x_inject = x_inject; a.l1_inject = l1_inject; a.seed = seed; a.client0 = client0; a.l1_out = l1_out;
    a.k_out = k_out; a.sgn_out = sgn_out; a.deq_out = deq_out; a.ld_out = ld_out;
    init_pack_target(a.pack, codes, codes_bytes, dir, a.hdr, n, d, m);
    a.packed = packed ? 1 : 0;
    if (packed && (long long)(a.pack.arena_base16 * 16ull) > codes_bytes) {
        set_error("code arena too small for the primary slots: %lld < %llu bytes", (long long)codes_bytes, a.pack.arena_base16 * 16ull);
        return DME_EWORKSPACE;
    }
    const unsigned grid = (unsigned)(n < 1024 ? n : 1024);
    literal_rows_kernel<<<grid, kThreads, 0, st>>>(a);
    DME_LAUNCH_CHECK("literal_rows_kernel");
    return DME_OK;
}

}  // namespace dme
