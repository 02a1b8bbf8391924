// pinnjet_inst.cu -- one translation unit per jet-channel scheme: compiled with -DPJ_N1=.. -DPJ_N2=.. (see build.py).
// PJ_N1 = PJ_N2 = -1 builds the scheme-independent helpers (K2b reduce, loss finalize).
#include "pinnjet_k2.cuh"

#define PJ_CAT3(a, b, c) a##b##_##c
#define PJ_NAME(prefix, n1, n2) PJ_CAT3(prefix, n1, n2)

namespace pj {

#if PJ_N1 < 0

// K2b: grad_theta[i] += sum over CTAs of partial[cta][i]   (fixed order -> run-to-run reproducible)
__global__ void k2_reduce_kernel(const float* __restrict__ gpart, int n_parts, long long n_theta,
                                 float* __restrict__ grad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_theta) return;
    float s = 0.0f;
    for (int p = 0; p < n_parts; ++p) s += gpart[(size_t)p * n_theta + i];
    grad[i] += s;
}

// sum of the per-CTA sums of squared residuals (fixed order) -> *out += total
__global__ void loss_finalize_kernel(const float* __restrict__ part, int n_parts, float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float s = 0.0f;
        for (int p = 0; p < n_parts; ++p) s += part[p];
        out[0] += s;
    }
}


cudaError_t launch_reduce(const float* gpart, int n_parts, long long n_theta, float* grad, cudaStream_t s) {
    const int nt = 256;
    k2_reduce_kernel<<<(unsigned)((n_theta + nt - 1) / nt), nt, 0, s>>>(gpart, n_parts, n_theta, grad);
    return cudaGetLastError();
}
cudaError_t launch_loss_finalize(const float* part, int n_parts, float* out, cudaStream_t s) {
    loss_finalize_kernel<<<1, 32, 0, s>>>(part, n_parts, out);
    return cudaGetLastError();
}

#else

constexpr int kC = 1 + PJ_N1 + PJ_N2;
constexpr int kP = (kC <= 2) ? 4 : 2;   // must match make_plan() in pinnjet_api.cu
constexpr int kQ = 4;

cudaError_t PJ_NAME(launch_k1_, PJ_N1, PJ_N2)(const K1Args& a, int grid, int smem, cudaStream_t s) {
    auto kern = k1_forward_kernel<kP, kQ, PJ_N1, PJ_N2>;
    static int configured = -1;
    if (configured < smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
        configured = 232448;
    }
    kern<<<grid, NT_TOTAL, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t PJ_NAME(launch_k2_, PJ_N1, PJ_N2)(const K2Args& a, int grid, int smem, cudaStream_t s) {
    auto kern = k2_backward_kernel<kP, kQ, PJ_N1, PJ_N2>;
    static int configured = -1;
    if (configured < smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
        configured = 232448;
    }
    kern<<<grid, NT_TOTAL, smem, s>>>(a);
    return cudaGetLastError();
}

#endif

}  // namespace pj
