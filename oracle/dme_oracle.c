/*
 * dme_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's DME hot path.
 *
 * This file is the *checker* for the CUDA library (libdme_b200.so).  Nothing in the product path
 * links, imports or calls it: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it.  Build: `make -C oracle` (gcc, -ffp-contract=off so every
 * fp32 operation rounds exactly once, like the ATen CPU kernels the reference runs on).
 *
 * Citations: AS = /root/reference/NMSE_Results/Codes/All_Schemes.py,
 *            ND = /root/reference/NMSE_Results/Codes/Normal_dist.py.
 *
 * Parity pin: tests/golden/make_golden.py imports the UNMODIFIED reference in the build container,
 * injects the random draws (torch.rand / rand_like / random_diagonal) and writes fixtures that
 * tests/test_oracle_golden.py checks this file against (bit-exact for the type quantizers, FWHT,
 * the reference's pair-transform, QUIC-FL receiver; 1e-6 relative for DRIVE/EDEN scales, whose
 * fp32 reductions in ATen have an unspecified summation order).
 *
 * Numerics contract, also implemented by the CUDA kernels:
 *   - L1 norm: sum of |x_i| accumulated in fp64, rounded once to fp32 (ATen's own fp32 sum order
 *     is not reproducible; tests inject the reference's fp32 L1 when comparing with it).
 *   - elementwise chain exactly as AS:625-631 in fp32: v = x / (L1 + 1e-12f); p = |v|;
 *     mp = float(m) * p; fl = floor(mp); fr = mp - fl.
 *   - prefix c_i = fp32(sum_{j<=i} fr_j) with the running sum held in fp64 (this IS what
 *     torch.cumsum does on CPU: acc_type<float> = double, rounded on store; verified).
 *   - r_i = [floor(c_i - X) - floor(c_{i-1} - X) == 1] in fp32, c_0 = 0  (AS:635-637).
 *   - output ((L1 * sign(v)) * (fl + r)) / float(m), left to right, fp32 true division (AS:640).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------- helpers */

static inline float sgnf(float v) { return (float)((v > 0.0f) - (v < 0.0f)); } /* torch.sign: sign(+-0)=0 */

/* AS:624 / AS:681 `input_vector.abs().sum()` under the fp64-accumulate contract. */
ORC_API double orc_l1_f64(const float *x, int64_t d) {
    double s = 0.0;
    for (int64_t i = 0; i < d; ++i) s += fabs((double)x[i]);
    return s;
}

/* ---------------------------------------------------------------- a1: Type_unbiased_quantize
 * AS:609-641.  l1_inject: NaN => compute (contract above); otherwise the fp32 L1 to use.
 * Outputs (any may be NULL): k (integer type vector, >=0), sgn (raw IEEE sign bit of x_i),
 * deq (the reference's return value).  Returns the fp32 L1 that was used. */
ORC_API float orc_type_unbiased(const float *x, int64_t d, int64_t m, float X, float l1_inject,
                                int64_t *k, uint8_t *sgn, float *deq) {
    float L1 = isnan(l1_inject) ? (float)orc_l1_f64(x, d) : l1_inject; /* AS:624 */
    const float D = L1 + 1e-12f;                                         /* AS:625 */
    const float mf = (float)m;                                           /* python int * fp32 tensor */
    double C = 0.0;                                                      /* AS:635 cumsum, fp64 acc */
    float a_prev = floorf(0.0f - X);                                     /* c_0 = 0 (AS:635 cat) */
    for (int64_t i = 0; i < d; ++i) {
        float v = x[i] / D;                                              /* AS:625 */
        float p = fabsf(v);                                              /* AS:626 */
        float mp = mf * p;                                               /* AS:629 */
        float fl = floorf(mp);                                           /* AS:630 */
        float fr = mp - fl;                                              /* AS:631 */
        C += (double)fr;
        float c = (float)C;                                              /* rounded on store */
        float a = floorf(c - X);                                         /* AS:636 */
        float r = ((a - a_prev) == 1.0f) ? 1.0f : 0.0f;                  /* AS:636-637 */
        a_prev = a;
        float kf = fl + r;
        if (k) k[i] = (int64_t)kf;
        if (sgn) { uint32_t b; memcpy(&b, &x[i], 4); sgn[i] = (uint8_t)(b >> 31); }
        if (deq) deq[i] = ((L1 * sgnf(v)) * kf) / mf;                    /* AS:640 */
    }
    return L1;
}

/* ---------------------------------------------------------------- a3/a4: Reznik + Type_biased_quantize
 * AS:644-687.  Tie rule (torch.topk leaves it unspecified): among equal residuals the LOWEST index
 * is adjusted first.  m' is summed exactly in int64 (the reference's fp32 sum is exact while < 2^24). */
typedef struct { float key; int64_t idx; } orc_kv;
static int cmp_desc(const void *a, const void *b) {
    const orc_kv *x = (const orc_kv *)a, *y = (const orc_kv *)b;
    if (x->key > y->key) return -1;
    if (x->key < y->key) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}
ORC_API float orc_type_biased(const float *x, int64_t d, int64_t m, float l1_inject,
                              int64_t *k, uint8_t *sgn, float *deq, int64_t *delta_out) {
    float L1 = isnan(l1_inject) ? (float)orc_l1_f64(x, d) : l1_inject;   /* AS:681 */
    const float D = L1 + 1e-12f;
    const float mf = (float)m;
    float *kp = (float *)malloc(sizeof(float) * (size_t)d);
    orc_kv *kv = (orc_kv *)malloc(sizeof(orc_kv) * (size_t)d);
    int64_t mprime = 0;
    for (int64_t i = 0; i < d; ++i) {
        float p = fabsf(x[i]) / D;                                       /* AS:683 */
        float mp = mf * p;
        kp[i] = floorf(mp + 0.5f);                                       /* AS:648 */
        mprime += (int64_t)kp[i];                                        /* AS:649 */
        kv[i].key = kp[i] - mp;                                          /* AS:654 delta_prime */
        kv[i].idx = i;
    }
    int64_t Delta = mprime - m;                                          /* AS:655 */
    if (delta_out) *delta_out = Delta;
    if (Delta > 0) {                                                     /* AS:657-660 */
        qsort(kv, (size_t)d, sizeof(orc_kv), cmp_desc);
        for (int64_t j = 0; j < Delta && j < d; ++j) kp[kv[j].idx] -= 1.0f;
    } else if (Delta < 0) {                                              /* AS:661-664 */
        for (int64_t i = 0; i < d; ++i) kv[i].key = -kv[i].key;
        qsort(kv, (size_t)d, sizeof(orc_kv), cmp_desc);
        for (int64_t j = 0; j < -Delta && j < d; ++j) kp[kv[j].idx] += 1.0f;
    }
    for (int64_t i = 0; i < d; ++i) {
        if (k) k[i] = (int64_t)kp[i];
        if (sgn) { uint32_t b; memcpy(&b, &x[i], 4); sgn[i] = (uint8_t)(b >> 31); }
        if (deq) deq[i] = (L1 * sgnf(x[i])) * (kp[i] / mf);              /* AS:666 + AS:687 */
    }
    free(kp); free(kv);
    return L1;
}

/* ---------------------------------------------------------------- a5: Hadamard.hadamard
 * AS:100-115: stages h=2,4,..,d; a' = a + b; b' = a' - 2*b (one rounding: 2*b is exact);
 * then `vec /= np.sqrt(d)` = fp32 true division by float(sqrt(d)). */
ORC_API int orc_hadamard(float *v, int64_t d) {
    if (d <= 0 || (d & (d - 1))) return -1;                              /* AS:103-104 */
    for (int64_t h = 2; h <= d; h <<= 1) {
        int64_t hf = h >> 1;
        for (int64_t base = 0; base < d; base += h)
            for (int64_t j = 0; j < hf; ++j) {
                float a = v[base + j], b = v[base + j + hf];
                float s = a + b;                                         /* AS:110 */
                v[base + j] = s;
                v[base + j + hf] = fmaf(-2.0f, b, s);                    /* AS:111 */
            }
    }
    const float sq = (float)sqrt((double)d);                             /* AS:113 */
    for (int64_t i = 0; i < d; ++i) v[i] = v[i] / sq;
    return 0;
}

/* a7: HadamardSender.randomized_hadamard_transform (AS:127-144) with the +-1 diagonal injected
 * (AS:117-120 draws it from a torch.Generator, which is torch-version specific).
 * x has d entries, out has dpad = 2^ceil(log2 d) entries, diag has dpad entries. */
ORC_API int orc_rht(const float *x, int64_t d, const float *diag, float *out, int64_t dpad) {
    for (int64_t i = 0; i < dpad; ++i) out[i] = (i < d ? x[i] : 0.0f) * diag[i];   /* AS:133-136 */
    return orc_hadamard(out, dpad);                                                /* AS:137 */
}
/* a8: HadamardReceiver.randomized_inverse_hadamard_transform (AS:151-156), in place on dpad entries. */
ORC_API int orc_irht(float *v, int64_t dpad, const float *diag) {
    int rc = orc_hadamard(v, dpad);                                                /* AS:153 */
    for (int64_t i = 0; i < dpad; ++i) v[i] = v[i] * diag[i];                      /* AS:154 */
    return rc;
}

/* ---------------------------------------------------------------- a9: fast_walsh_hadamard_transform
 * AS:37-59 as it actually executes: `a`,`b` are views, so after the first write the
 * second stores (a+b)-b.  Every stage acts on ADJACENT pairs; log2(n) stages. */
ORC_API void orc_pair_transform(float *v, int64_t n) {
    for (int64_t h = 1; h < n; h <<= 1)
        for (int64_t j = 0; j + 1 < n; j += 2) {
            float s = v[j] + v[j + 1];                                   /* AS:51 */
            v[j] = s;
            v[j + 1] = s - v[j + 1];                                     /* AS:52 (aliased a) */
        }
}
/* True unnormalised WHT, natural order (what AS:37 documents). Used by compat="correct". */
static void wht_plain(float *v, int64_t n) {
    for (int64_t h = 1; h < n; h <<= 1)
        for (int64_t base = 0; base < n; base += 2 * h)
            for (int64_t j = 0; j < h; ++j) {
                float a = v[base + j], b = v[base + j + h];
                v[base + j] = a + b; v[base + j + h] = a - b;
            }
}

/* ---------------------------------------------------------------- a10: DRIVE_quantize_Hadamard
 * AS:707-752.  dsign: injected D (+-1) laid out chunk after chunk, each chunk padded to its own
 * power of two (AS:725-735).  compat=0: reference transform (a9); compat=1: true WHT, orthonormal.
 * Reductions (norm, abs-sum) accumulate in fp64 then round to fp32 (ATen order is unspecified). */
ORC_API void orc_drive(const float *x, int64_t d, const float *dsign, int compat, float *out) {
    const int64_t B = 2048;
    float buf[2048], D[2048];
    int64_t doff = 0;
    for (int64_t s0 = 0; s0 < d; s0 += B) {
        int64_t len = (d - s0 < B) ? d - s0 : B;
        int64_t np2 = 1; while (np2 < len) np2 <<= 1;
        double nrm2 = 0.0;
        for (int64_t j = 0; j < np2; ++j) {
            float xv = j < len ? x[s0 + j] : 0.0f;
            D[j] = dsign[doff + j];
            buf[j] = D[j] * xv;                                          /* AS:737 */
            nrm2 += (double)xv * (double)xv;
        }
        if (compat == 0) orc_pair_transform(buf, np2); else wht_plain(buf, np2);   /* AS:738 */
        double l1 = 0.0;
        for (int64_t j = 0; j < np2; ++j) l1 += fabs((double)buf[j]);
        float nrm = (float)sqrt(nrm2);
        float num = nrm * nrm;                                           /* .norm(2).pow(2) */
        float S = num / ((float)l1 + 1e-12f);                            /* AS:741; with the true WHT the
                                                                            1/sqrt(n) factors cancel, so the
                                                                            same formula is real DRIVE */
        for (int64_t j = 0; j < np2; ++j) buf[j] = S * sgnf(buf[j]);     /* AS:743 */
        if (compat == 0) orc_pair_transform(buf, np2); else wht_plain(buf, np2);   /* AS:746 */
        for (int64_t j = 0; j < len; ++j) out[s0 + j] = buf[j] * D[j];   /* AS:747-750 */
        doff += np2;
    }
}

/* ---------------------------------------------------------------- a11-a13: EDEN
 * Centroids AS:301-320 (fp32 tensors), boundaries = fp32 midpoints. nbits in {1,2}. */
static int eden_tables(int nbits, float *cent, float *bnd) {
    if (nbits == 1) {
        cent[0] = -(float)0.7978845608028654; cent[1] = (float)0.7978845608028654;
        bnd[0] = (cent[0] + cent[1]) / 2.0f;
        return 2;
    } else if (nbits == 2) {
        cent[0] = -(float)1.5104176087114887; cent[1] = -(float)0.4527800398860679;
        cent[2] = (float)0.4527800398860679;  cent[3] = (float)1.5104176087114887;
        for (int i = 0; i < 3; ++i) bnd[i] = (cent[i] + cent[i + 1]) / 2.0f;
        return 4;
    }
    return 0;
}
/* EdenSender.compress + quantize (AS:335-350, AS:370-383).  v_rot: scratch/out of dpad floats (the
 * rotated vector); bins: dpad int32; returns scale.  norm_inject: NaN => fp64-accumulated norm. */
ORC_API float orc_eden_encode(const float *x, int64_t d, const float *diag, int64_t dpad, int nbits,
                              float norm_inject, float *v_rot, int32_t *bins) {
    float cent[4], bnd[3];
    int nc = eden_tables(nbits, cent, bnd);
    orc_rht(x, d, diag, v_rot, dpad);                                    /* AS:378-380 */
    double n2 = 0.0;
    for (int64_t i = 0; i < dpad; ++i) n2 += (double)v_rot[i] * (double)v_rot[i];
    float nrm = isnan(norm_inject) ? (float)sqrt(n2) : norm_inject;      /* torch.norm(vec, 2) */
    float sq = (float)pow((double)dpad, 0.5);                            /* vec.numel() ** 0.5 */
    double dot = 0.0;
    for (int64_t i = 0; i < dpad; ++i) {
        float z = (v_rot[i] * sq) / nrm;                                 /* AS:343 */
        int b = 0;
        while (b < nc - 1 && bnd[b] < z) ++b;                            /* bucketize, right=False */
        bins[i] = b;
        dot += (double)cent[b] * (double)v_rot[i];
    }
    return (nrm * nrm) / (float)dot;                                     /* AS:348 */
}
/* EdenReceiver.decompress (AS:398-426) for integer nbits, pdrop=0. out: d floats; work: dpad. */
ORC_API void orc_eden_decode(const int32_t *bins, int64_t d, int64_t dpad, const float *diag, int nbits,
                             float scale, float *work, float *out) {
    float cent[4], bnd[3];
    eden_tables(nbits, cent, bnd);
    for (int64_t i = 0; i < dpad; ++i) work[i] = cent[bins[i]];          /* AS:400 */
    orc_irht(work, dpad, diag);                                          /* AS:425 */
    for (int64_t i = 0; i < d; ++i) out[i] = scale * work[i];            /* AS:426 */
}

/* ---------------------------------------------------------------- a15: QuicFLReceiver.decompress
 * AS:526-535.  h (shared randomness, AS:527-528) injected; exact_mask/exact_vals as in the sender's
 * dict (values listed in index order of the mask, AS:531). */
ORC_API void orc_quicfl_decode(const int32_t *X, const int32_t *h, int64_t d, int64_t dpad, int h_len,
                               const float *recv_table, const uint8_t *exact_mask, const float *exact_vals,
                               float scale, const float *diag, float *work, float *out) {
    int64_t e = 0;
    for (int64_t i = 0; i < dpad; ++i) {
        float val = recv_table[(int64_t)X[i] * h_len + h[i]];            /* AS:530 */
        if (exact_mask && exact_mask[i]) val = exact_vals[e++];          /* AS:531 */
        work[i] = val / scale;                                           /* AS:532 */
    }
    orc_irht(work, dpad, diag);                                          /* AS:534 */
    for (int64_t i = 0; i < d; ++i) out[i] = work[i];                    /* AS:535 */
}

/* ---------------------------------------------------------------- Scalar_quantize (next row, AS:755-790)
 * u: injected uniforms (torch.rand_like, AS:783). bits may be fractional in the reference; nlevels
 * is passed in as float (2**bits - 1). */
ORC_API void orc_scalar(const float *x, int64_t d, float nlevels, const float *u, float *out) {
    float mn = x[0], mx = x[0];
    for (int64_t i = 1; i < d; ++i) { if (x[i] < mn) mn = x[i]; if (x[i] > mx) mx = x[i]; }
    float denom = mx - mn;
    if (denom == 0.0f || nlevels < 1.0f) { memcpy(out, x, sizeof(float) * (size_t)d); return; }
    for (int64_t i = 0; i < d; ++i) {
        float q = (x[i] - mn) / denom;                                   /* AS:768 */
        q = q < 0.0f ? 0.0f : (q > 1.0f ? 1.0f : q);                     /* AS:776 */
        float t = q * nlevels;
        float bf = floorf(t);                                            /* AS:779 */
        float fr = t - bf;
        float bi = bf + ((u[i] < fr) ? 1.0f : 0.0f);                     /* AS:783-784 */
        q = bi / nlevels;                                                /* AS:787 */
        out[i] = q * (mx - mn) + mn;                                     /* AS:788 */
    }
}

/* ---------------------------------------------------------------- a17: server mean (ND:133-147)
 * est += q / n, one client after another, fp32. */
ORC_API void orc_mean_accumulate(float *est, const float *q, int64_t d, int64_t n) {
    const float nf = (float)n;
    for (int64_t i = 0; i < d; ++i) est[i] = est[i] + q[i] / nf;
}

/* ---------------------------------------------------------------- packed format "DMEP1" (K6; no
 * reference counterpart, SURVEY F1 -- defined by this repo, see DESIGN.md "Packed code").
 * A tile is 4096 coordinates = 256 chunks of 16.  With field width w in {2,4,8,16,32} each
 * coordinate is a w-bit field [sign:1 | magnitude:w-1]; chunk c's 16 fields form a 16w-bit
 * little-endian integer (field j at bits [w*j, w*j+w)); its 32-bit word q is stored at
 * tile_words[q*256 + c].  Tile size = 128*w words.  Missing coordinates of the last tile are 0. */
ORC_API int orc_tile_width(const int64_t *k, int64_t cnt) {
    int64_t mx = 0;
    for (int64_t i = 0; i < cnt; ++i) if (k[i] > mx) mx = k[i];
    int w = 2;
    while (w < 32 && mx >= ((int64_t)1 << (w - 1))) w <<= 1;
    if (mx >= ((int64_t)1 << 31)) return -1;
    return w;
}
ORC_API void orc_pack_tile(const int64_t *k, const uint8_t *sgn, int64_t cnt, int w, uint32_t *words) {
    int wpc = w / 2;                                  /* words per chunk = 16*w/32 */
    memset(words, 0, sizeof(uint32_t) * (size_t)(64 * wpc));   /* a code tile = 1024 coordinates = 64 chunks */
    for (int64_t i = 0; i < cnt; ++i) {
        int64_t c = i >> 4, j = i & 15;
        uint64_t field = ((uint64_t)(sgn[i] & 1) << (w - 1)) | (uint64_t)k[i];
        int64_t bit = (int64_t)w * j;
        int q = (int)(bit >> 5), sh = (int)(bit & 31);
        words[q * 64 + c] |= (uint32_t)(field << sh);   /* w divides 32: a field never straddles */
    }
}
ORC_API void orc_unpack_tile(const uint32_t *words, int w, int64_t cnt, int64_t *k, uint8_t *sgn) {
    for (int64_t i = 0; i < cnt; ++i) {
        int64_t c = i >> 4, j = i & 15;
        int64_t bit = (int64_t)w * j;
        int q = (int)(bit >> 5), sh = (int)(bit & 31);
        uint64_t field = (words[q * 64 + c] >> sh);
        if (w < 32) field &= (((uint64_t)1 << w) - 1);
        sgn[i] = (uint8_t)((field >> (w - 1)) & 1);
        k[i] = (int64_t)(field & (((uint64_t)1 << (w - 1)) - 1));
    }
}
/* Dequantise one decoded coordinate the way the fused decode-mean kernel does (== AS:640 / AS:687). */
ORC_API float orc_deq_value(float L1, int64_t m, int64_t k, int sgnbit, int biased) {
    float mf = (float)m, s = k == 0 ? 0.0f : (sgnbit ? -1.0f : 1.0f), kf = (float)k;
    return biased ? (L1 * s) * (kf / mf) : ((L1 * s) * kf) / mf;
}

/* ---------------------------------------------------------------- batched CPU baseline (bench only)
 * quantize -> dequantize -> mean over n clients, the reference's server loop (ND:133-147) with
 * Type_unbiased_quantize; pthreads over clients in groups of `threads` rows, then the mean is
 * accumulated in client order (so the result does not depend on the thread count). */
#include <pthread.h>
typedef struct {
    const float *Xm; const float *Xs; float *buf; float *mean;
    int64_t c0, g, d, ld, m, n; int tid, threads;
} orc_job;
static void *job_quant(void *p) {
    orc_job *j = (orc_job *)p;
    if (j->tid < j->g)
        orc_type_unbiased(j->Xm + (j->c0 + j->tid) * j->ld, j->d, j->m, j->Xs[j->c0 + j->tid], NAN,
                          NULL, NULL, j->buf + (int64_t)j->tid * j->d);
    return NULL;
}
static void *job_accum(void *p) {
    orc_job *j = (orc_job *)p;
    int64_t lo = j->d * j->tid / j->threads, hi = j->d * (j->tid + 1) / j->threads;
    const float nf = (float)j->n;
    for (int64_t i = lo; i < hi; ++i) {
        float e = j->mean[i];
        for (int64_t q = 0; q < j->g; ++q) e = e + j->buf[q * j->d + i] / nf;
        j->mean[i] = e;
    }
    return NULL;
}
ORC_API int orc_quantize_mean_unbiased(const float *Xm, int64_t n, int64_t d, int64_t ld, int64_t m,
                                       const float *Xs, float *mean, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    float *buf = (float *)malloc(sizeof(float) * (size_t)d * (size_t)threads);
    if (!buf) return -1;
    memset(mean, 0, sizeof(float) * (size_t)d);
    pthread_t th[256]; orc_job jobs[256];
    for (int64_t c0 = 0; c0 < n; c0 += threads) {
        int64_t g = (n - c0 < threads) ? n - c0 : threads;
        for (int pass = 0; pass < 2; ++pass) {
            for (int t = 0; t < threads; ++t) {
                orc_job jb = {Xm, Xs, buf, mean, c0, g, d, ld, m, n, t, threads};
                jobs[t] = jb;
                pthread_create(&th[t], NULL, pass == 0 ? job_quant : job_accum, &jobs[t]);
            }
            for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
        }
    }
    free(buf);
    return 0;
}
