// pinnjet_comm.cu -- the one collective of the data-parallel path: SUM of the flat [grad_theta | sum r^2]
// buffer over the ranks of one node, as ONE kernel over NVLink peer memory.
//
// NCCL's all-reduce of this 34 KB message is latency bound (measured round 1: +20 / +40 / +52 us at 2 / 4 / 8 GPUs, on the
// critical path of every step).  Here every rank owns a SYMMETRIC buffer (same layout on every GPU, peer-mapped; the host
// side obtains the peer pointers once, e.g. from torch symmetric memory):
//       [flags: PJ_AR_BLOCKS x PJ_AR_MAX_RANKS x u32][epochs: PJ_AR_BLOCKS x u32][data: 2 x n_pad floats]
// and a call is one launch of PJ_AR_BLOCKS independent CTAs.  CTA b, epoch e (its own counter, kept on the device so that a
// captured CUDA graph can be replayed):
//   1. copies its slice of the input into data[e & 1] of the LOCAL buffer;
//   2. bar.sync, then one thread per peer: fence.sys + st.release.sys of e into flag[b][my rank] of THAT peer;
//   3. one thread per peer spins (ld.acquire.sys) on the local flag[b][peer] until it reaches e;
//   4. sums the slice over the ranks in rank order (every rank gets bit-identical results) reading the peers' data[e & 1]
//      over NVLink, writes the result to `out` (may alias the input).
// data[] is double buffered by the epoch's parity: a rank can only be two epochs ahead of a peer's reads if that peer has
// signalled the epoch in between, i.e. finished reading (stream order) -- no second barrier is needed.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pinnjet.h"

namespace pj {

struct ArArgs {
    unsigned long long buf[PJ_AR_MAX_RANKS];
    int rank, world;
    long long n, n_pad;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {   // never served from a stale L1 line
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) allreduce_oneshot_kernel(const ArArgs a, const float* __restrict__ in, float* out) {
    const int b = blockIdx.x, tid = threadIdx.x;
    unsigned char* me = reinterpret_cast<unsigned char*>(a.buf[a.rank]);
    unsigned* my_flags = reinterpret_cast<unsigned*>(me) + b * PJ_AR_MAX_RANKS;
    unsigned* my_epoch = reinterpret_cast<unsigned*>(me + PJ_AR_FLAG_BYTES) + b;
    const unsigned e = *my_epoch + 1u;
    const long long data_off = PJ_AR_HEADER_BYTES / 4 + (long long)(e & 1u) * a.n_pad;    // floats from the buffer start
    const long long per = ((a.n_pad / 4 + PJ_AR_BLOCKS - 1) / PJ_AR_BLOCKS) * 4;
    const long long lo = (long long)b * per, hi = min(lo + per, a.n_pad);
    float* mine = reinterpret_cast<float*>(me) + data_off;

    for (long long i = lo + 4 * tid; i < hi; i += 4 * blockDim.x) {
        float4 v;
        v.x = i + 0 < a.n ? in[i + 0] : 0.0f;
        v.y = i + 1 < a.n ? in[i + 1] : 0.0f;
        v.z = i + 2 < a.n ? in[i + 2] : 0.0f;
        v.w = i + 3 < a.n ? in[i + 3] : 0.0f;
        *reinterpret_cast<float4*>(mine + i) = v;
    }
    __syncthreads();
    if (tid < a.world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned*>(a.buf[tid]) + b * PJ_AR_MAX_RANKS + a.rank, e);
        while ((int)(ld_acquire_sys(my_flags + tid) - e) < 0) {
        }
    }
    __syncthreads();
    for (long long i = lo + 4 * tid; i < hi; i += 4 * blockDim.x) {
        float4 v[PJ_AR_MAX_RANKS];                       // all peer loads in flight before the first add
#pragma unroll
        for (int p = 0; p < PJ_AR_MAX_RANKS; ++p)
            if (p < a.world) v[p] = ld_peer_f4(reinterpret_cast<const float*>(a.buf[p]) + data_off + i);
        float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
        for (int p = 0; p < PJ_AR_MAX_RANKS; ++p)
            if (p < a.world) {
                acc.x += v[p].x;
                acc.y += v[p].y;
                acc.z += v[p].z;
                acc.w += v[p].w;
            }
        if (i + 0 < a.n) out[i + 0] = acc.x;
        if (i + 1 < a.n) out[i + 1] = acc.y;
        if (i + 2 < a.n) out[i + 2] = acc.z;
        if (i + 3 < a.n) out[i + 3] = acc.w;
    }
    __syncthreads();
    if (tid == 0) *my_epoch = e;
}

}  // namespace pj

extern "C" {

int64_t pj_allreduce_bytes(int64_t n_floats) {
    const int64_t n_pad = (n_floats + 3) / 4 * 4;
    return PJ_AR_HEADER_BYTES + 2 * n_pad * 4;
}

int pj_allreduce_oneshot(const uint64_t* peer_buffers, int32_t rank, int32_t world, const float* in, float* out,
                         int64_t n_floats, void* stream) {
    if (!peer_buffers || !in || !out || world < 1 || world > PJ_AR_MAX_RANKS || rank < 0 || rank >= world || n_floats < 1) return -1;
    pj::ArArgs a;
    for (int p = 0; p < PJ_AR_MAX_RANKS; ++p) a.buf[p] = p < world ? peer_buffers[p] : 0ull;
    a.rank = rank;
    a.world = world;
    a.n = n_floats;
    a.n_pad = (n_floats + 3) / 4 * 4;
    pj::allreduce_oneshot_kernel<<<PJ_AR_BLOCKS, 256, 0, (cudaStream_t)stream>>>(a, in, out);
    return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // extern "C"
