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

__global__ void __launch_bounds__(256)
peer_sum_slice_kernel(float *const *__restrict__ bufs, int64_t off, int rank, int world, int64_t lo, int64_t hi) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t i = lo + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < hi; i += stride) {
        float4 s = *reinterpret_cast<const float4 *>(bufs[0] + off + i);
        for (int p = 1; p < world; ++p) {
            const float4 v = *reinterpret_cast<const float4 *>(bufs[p] + off + i);
            s.x = __fadd_rn(s.x, v.x); s.y = __fadd_rn(s.y, v.y); s.z = __fadd_rn(s.z, v.z); s.w = __fadd_rn(s.w, v.w);
        }
        for (int p = 0; p < world; ++p) *reinterpret_cast<float4 *>(bufs[p] + off + i) = s;
    }
}

__global__ void __launch_bounds__(256)
multimem_sum_slice_kernel(float *mc, int64_t lo, int64_t hi) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t i = lo + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < hi; i += stride) {
        float4 s;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(s.x), "=f"(s.y), "=f"(s.z), "=f"(s.w) : "l"(mc + i) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                     ::"l"(mc + i), "f"(s.x), "f"(s.y), "f"(s.z), "f"(s.w) : "memory");
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
    const int64_t threads = (hi - lo) / 4;
    int64_t blocks = (threads + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (multicast) multimem_sum_slice_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(multicast + off, lo, hi);
    else peer_sum_slice_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(bufs, off, rank, world, lo, hi);
    DME_LAUNCH_CHECK("peer_sum_slice_kernel");
    return DME_OK;
}
