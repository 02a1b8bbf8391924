// pinnjet_optim.cu -- the Adam update of the flat parameter buffer as ONE launch (the reference's optimizer is
// torch.optim.Adam, solvers.py:182, 396: ~20 launches per step on the per-tensor path, ~10 elementwise launches on the
// flat buffers).  Part of the opt-in device loop (one CUDA-graph replay per epoch); torch.optim.Adam stays the default.
//
// Algorithm = torch.optim.Adam with amsgrad=False, weight_decay=0, maximize=False:
//   t += 1;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  theta -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The step count and the learning rate live in device memory (a captured graph replays with fresh values); optionally the
// same launch keeps the best parameters: if *loss < *best_loss the parameters BEFORE the update are copied to best_theta
// (reference solvers.py:411-418: lowest loss and best nets are taken before optimizer.step()).
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pinnjet.h"

namespace pj {

__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, double* state /*{t, lr, ticket}*/, float b1,
                                                        float b2, float eps, const float* loss, float* best_loss,
                                                        float* __restrict__ best_theta) {
    const double t = state[0] + 1.0, lr = state[1];
    const float bc1 = (float)(1.0 - pow((double)b1, t)), rbc2 = (float)(1.0 / sqrt(1.0 - pow((double)b2, t)));
    const float step = (float)(lr / (double)bc1);
    const bool better = best_theta != nullptr && *loss < *best_loss;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float g = grad[i], th = theta[i];
        if (better) best_theta[i] = th;
        const float mi = b1 * m[i] + (1.0f - b1) * g;
        const float vi = b2 * v[i] + (1.0f - b2) * g * g;
        m[i] = mi;
        v[i] = vi;
        theta[i] = th - step * (mi / (sqrtf(vi) * rbc2 + eps));
    }
    // the last block to finish commits the scalars (every block has read them by then) and re-arms the ticket
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(reinterpret_cast<unsigned long long*>(state + 2), 1ull) == (unsigned long long)gridDim.x - 1ull;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        *reinterpret_cast<unsigned long long*>(state + 2) = 0ull;
        state[0] = t;
        if (better) *best_loss = *loss;
    }
}

}  // namespace pj

extern "C" int pj_adam_step(float* theta, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double* state,
                            float beta1, float beta2, float eps, const float* loss, float* best_loss, float* best_theta,
                            void* stream) {
    if (!theta || !grad || !exp_avg || !exp_avg_sq || !state || n < 1) return -1;
    if (best_theta && (!loss || !best_loss)) return -1;
    pj::adam_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(theta, grad, exp_avg, exp_avg_sq, n, state, beta1,
                                                                                      beta2, eps, loss, best_loss, best_theta);
    return cudaGetLastError() == cudaSuccess ? 0 : -5;
}
