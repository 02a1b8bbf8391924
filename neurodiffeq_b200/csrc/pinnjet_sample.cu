// pinnjet_sample.cu -- collocation point sampling on the device (reference generators.py:107-191 1-D,
// :194-314 2-D / 3-D grids, :572-655 spherical).  Opt-in: the host generators stay the default (BASELINE north_star).
//
// A sampler is a list of per-coordinate laws over the GLOBAL batch index i (so that under data parallelism rank k draws
// exactly rows [lo, hi) of the batch every rank would have drawn -- no host work that scales with the global batch):
//   PJ_LAW_BASE          x_i = base[i]                                 (equally-spaced, chebyshev, ... : fixed nodes)
//   PJ_LAW_BASE_NORMAL   x_i = base[i] + std * N(0, 1)                 ('*-noisy' methods: generators.py:149-158, 253-266)
//   PJ_LAW_UNIFORM       x_i = lo + (hi - lo) * U[0, 1)                ('uniform')
//   PJ_LAW_SPHERICAL     (r, theta, phi)_i of GeneratorSpherical       (r^2 or r uniform; direction = normalised random
//                                                                        octant vector, never on the poles: :603-646)
// Random numbers: Philox4x32-10 keyed by the seed, counter = (global index, law slot, call number): reproducible, no state
// per point.  The call number lives in device memory and is advanced by the last block of each launch, so a captured CUDA
// graph draws fresh points at every replay.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pinnjet.h"

namespace pj {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }            // [0, 1)
__device__ __forceinline__ float u01_open(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0, 1)

struct SampleArgs {
    PjSampler s;
    long long first, n;              // this rank's rows [first, first + n) of the global batch
    float* out[PJ_MAX_COORDS];
};

__global__ void __launch_bounds__(256) sample_kernel(const SampleArgs a, unsigned long long* state) {
    const unsigned long long call = state[0];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) {
        const unsigned long long row = (unsigned long long)(a.first + i);
        for (int l = 0; l < a.s.n_laws; ++l) {
            const PjSampleLaw& law = a.s.law[l];
            const unsigned long long g = law.mod > 0 ? (row / (unsigned long long)law.div) % (unsigned long long)law.mod : row;
            uint32_t c[4] = {(uint32_t)g, (uint32_t)(g >> 32) ^ ((uint32_t)l << 24), (uint32_t)call, (uint32_t)(call >> 32)};
            if (law.kind != PJ_LAW_BASE) philox4x32_10(c, (uint32_t)a.s.seed, (uint32_t)(a.s.seed >> 32));
            if (law.kind == PJ_LAW_BASE) {
                a.out[law.coord][i] = law.base[g];
            } else if (law.kind == PJ_LAW_BASE_NORMAL) {
                const float r = sqrtf(-2.0f * logf(u01_open(c[0])));          // Box-Muller
                a.out[law.coord][i] = fmaf(law.p0, r * cospif(2.0f * u01(c[1])), law.base[g]);
            } else if (law.kind == PJ_LAW_UNIFORM) {
                a.out[law.coord][i] = fmaf(law.p1 - law.p0, u01(c[0]), law.p0);
            } else {   // PJ_LAW_SPHERICAL: p0 = r_min, p1 = r_max, flag: 1 = r^2 uniform ('equally-spaced-noisy'), 0 = r uniform
                uint32_t d[4] = {(uint32_t)g, (uint32_t)(g >> 32) ^ ((uint32_t)(l + 64) << 24), (uint32_t)call, (uint32_t)(call >> 32)};
                philox4x32_10(d, (uint32_t)a.s.seed, (uint32_t)(a.s.seed >> 32));
                const float u = u01(c[0]);
                const float r = law.flag ? sqrtf((law.p1 * law.p1 - law.p0 * law.p0) * u + law.p0 * law.p0)
                                         : (law.p1 - law.p0) * u + law.p0;
                const float w0 = u01_open(c[1]), w1 = u01_open(c[2]), w2 = u01_open(c[3]), ws = w0 + w1 + w2;
                float v0 = sqrtf(w0 / ws) + 1e-6f, v1 = sqrtf(w1 / ws) + 1e-6f, v2 = sqrtf(w2 / ws) + 1e-6f;
                v0 = (d[0] & 1u) ? v0 : -v0;
                v1 = (d[1] & 1u) ? v1 : -v1;
                v2 = (d[2] & 1u) ? v2 : -v2;
                a.out[law.coord][i] = r;
                a.out[law.coord + 1][i] = acosf(fminf(fmaxf(v2, -1.0f), 1.0f));
                a.out[law.coord + 2][i] = 3.14159265358979f - atan2f(v1, v0);
            }
        }
    }
    // the last block to finish advances the call number (every block has read it by then) and re-arms the ticket
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(&state[1], 1ull) == (unsigned long long)gridDim.x - 1ull;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        state[1] = 0ull;
        state[0] = call + 1ull;
    }
}

}  // namespace pj

extern "C" int pj_sample(const PjSampler* sampler, int64_t first, int64_t n, float* const* out, uint64_t* state, void* stream) {
    if (!sampler || !out || !state || n < 1 || first < 0 || sampler->n_laws < 1 || sampler->n_laws > PJ_MAX_COORDS) return -1;
    pj::SampleArgs a;
    a.s = *sampler;
    a.first = first;
    a.n = n;
    for (int i = 0; i < PJ_MAX_COORDS; ++i) a.out[i] = nullptr;
    for (int l = 0; l < sampler->n_laws; ++l) {
        const PjSampleLaw& law = sampler->law[l];
        const int span = law.kind == PJ_LAW_SPHERICAL ? 3 : 1;
        if (law.coord < 0 || law.coord + span > PJ_MAX_COORDS) return -1;
        if ((law.kind == PJ_LAW_BASE || law.kind == PJ_LAW_BASE_NORMAL) && !law.base) return -1;
        for (int k = 0; k < span; ++k) {
            if (!out[law.coord + k]) return -1;
            a.out[law.coord + k] = out[law.coord + k];
        }
    }
    pj::sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, reinterpret_cast<unsigned long long*>(state));
    return cudaGetLastError() == cudaSuccess ? 0 : -5;
}
