// peer_reduce.cu -- the path's only exchange step: the sum of the ranks' partial means, as our own kernels over
// NVLink peer memory instead of a library all-reduce.  Every rank's partial mean lives in a SYMMETRIC buffer (same size on
// every GPU of the box, mapped into every process; torch.distributed._symmetric_memory hands out the mappings).  Two-shot
// all-reduce: rank r owns slice r of the vector; it reduces that slice over all ranks and writes the result back into every
// rank's buffer.  Two variants of the same kernel:
//   * multicast (NVLS): one `multimem.ld_reduce` pulls the slice element from all GPUs and adds inside the NVSwitch, one
//     `multimem.st` broadcasts the sum -- each GPU moves its slice once in and once out;
//   * peer loads / stores in FIXED rank order 0, 1, .., N-1 (no multicast object, or a bit-reproducible order wanted).
// Slice r of every buffer is read and written by rank r only, so the reduction is in place.  The caller brackets the launch
// with the symmetric-memory barrier (all partial means written before, all slices broadcast after): dme_b200/distributed.py.
#include "common.cuh"

namespace dme {

// Four 16-byte pieces per thread and iteration (independent loads in flight: the loads cross NVLink).
constexpr int kPeerUnroll = 4;
__global__ void __launch_bounds__(256)
peer_sum_slice_kernel(float *const *__restrict__ bufs, int64_t off, int rank, int world, int64_t lo, int64_t hi) {
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i0 = lo + tid * 4; i0 < hi; i0 += nthr * 4 * kPeerUnroll) {
        float4 s[kPeerUnroll];
#pragma unroll
        for (int u = 0; u < kPeerUnroll; ++u) {
            const int64_t i = i0 + (int64_t)u * nthr * 4;
            s[u] = i < hi ? *reinterpret_cast<const float4 *>(bufs[0] + off + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int p = 1; p < world; ++p) {
            float4 v[kPeerUnroll];
#pragma unroll
            for (int u = 0; u < kPeerUnroll; ++u) {
                const int64_t i = i0 + (int64_t)u * nthr * 4;
                v[u] = i < hi ? *reinterpret_cast<const float4 *>(bufs[p] + off + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kPeerUnroll; ++u) {
                s[u].x = __fadd_rn(s[u].x, v[u].x); s[u].y = __fadd_rn(s[u].y, v[u].y);
                s[u].z = __fadd_rn(s[u].z, v[u].z); s[u].w = __fadd_rn(s[u].w, v[u].w);
            }
        }
        for (int p = 0; p < world; ++p) {
#pragma unroll
            for (int u = 0; u < kPeerUnroll; ++u) {
                const int64_t i = i0 + (int64_t)u * nthr * 4;
                if (i < hi) *reinterpret_cast<float4 *>(bufs[p] + off + i) = s[u];
            }
        }
    }
}

__global__ void __launch_bounds__(256)
multimem_sum_slice_kernel(float *mc, int64_t lo, int64_t hi) {
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i0 = lo + tid * 4; i0 < hi; i0 += nthr * 4 * kPeerUnroll) {
        float4 s[kPeerUnroll];
#pragma unroll
        for (int u = 0; u < kPeerUnroll; ++u) {
            const int64_t i = i0 + (int64_t)u * nthr * 4;
            if (i < hi)
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(s[u].x), "=f"(s[u].y), "=f"(s[u].z), "=f"(s[u].w) : "l"(mc + i) : "memory");
        }
#pragma unroll
        for (int u = 0; u < kPeerUnroll; ++u) {
            const int64_t i = i0 + (int64_t)u * nthr * 4;
            if (i < hi)
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                             ::"l"(mc + i), "f"(s[u].x), "f"(s[u].y), "f"(s[u].z), "f"(s[u].w) : "memory");
        }
    }
}

}  // namespace dme

using namespace dme;

// bufs: DEVICE array of `world` pointers to the ranks' symmetric ALLOCATIONS, the vector starts `offset_bytes` into each of them
// (16-byte aligned); or multicast != null: the multicast mapping of the same allocations.  Reduces slice `rank` (d rounded up to a multiple of 4 * world floats is split
// evenly) and broadcasts it.
extern "C" int dme_peer_sum_slice(float *const *bufs, float *multicast, int64_t offset_bytes, int rank, int world, int64_t d, dme_stream_t stream) {
    DME_REQUIRE((bufs != nullptr || multicast != nullptr) && world >= 1 && rank >= 0 && rank < world && d >= 1 && offset_bytes >= 0 && offset_bytes % 16 == 0, "bad argument");
    const int64_t off = offset_bytes / 4;
    const int64_t quads = (d + 3) / 4, per = (quads + world - 1) / world * 4;
    const int64_t lo = (int64_t)rank * per, hi = lo + per < (d + 3) / 4 * 4 ? lo + per : (d + 3) / 4 * 4;
    if (lo >= hi) return DME_OK;
    const int64_t threads = ((hi - lo) / 4 + kPeerUnroll - 1) / kPeerUnroll;
    int64_t blocks = (threads + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (multicast) multimem_sum_slice_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(multicast + off, lo, hi);
    else peer_sum_slice_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(bufs, off, rank, world, lo, hi);
    DME_LAUNCH_CHECK("peer_sum_slice_kernel");
    return DME_OK;
}
