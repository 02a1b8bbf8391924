// tcgen05 probe no. 2 (round-2 groundwork, NOT part of the product): the weight-gradient contraction of the reverse kernel,
//     D[m][n] = sum_r Z[r][m] * A[r][n],      r = 0..255 (the (point, channel) rows of a tile),  m, n = hidden units,
// read straight from the SAME shared-memory images the forward / adjoint GEMMs use: [256 rows x 64 units] bf16, one 128-byte
// row per r, 8-row groups 1024 B apart, SWIZZLE_128B.  For this GEMM the contraction index is the ROW, so both operands are
// "MN-major" (the M / N index is the contiguous one): instruction descriptor bits 15/16 = 1, and per the canonical layout
//     MN-major SW128:  ((T,8,m),(8,k)) : ((1,T,LBO),(8T,SBO))      [T = 8 bf16 = 16 B]
// one K step is one 128-byte row, 8 K-steps form a 1024-byte group (SBO), the next 64 M/N elements sit LBO bytes away.
// Questions this probe answers on the hardware (exact integer data, so the right encoding gives error 0):
//   * (lbo_enc, sbo_enc) semantics for MN-major operands,
//   * the start-address advance per K = 16 instruction (16 rows = 2048 B),
//   * the TMEM lane layout of an M = 64 accumulator (rows 16q..16q+15 in lanes 32q..32q+15),
//   * M = 128 with the two halves of the M index taken from two images LBO bytes apart ("stacked" split terms).
// Usage: probe_wgrad <lbo_enc> <sbo_enc> <M: 64|128> <k_advance_bytes>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_enc, uint32_t sbo_enc) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)(lbo_enc & 0x3FFF) << 16;          // leading byte offset >> 4
    d |= (uint64_t)(sbo_enc & 0x3FFF) << 32;          // stride byte offset >> 4
    d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}

constexpr int ROWS = 256, H = 64, IMG = ROWS * 128;   // one image: 256 rows x 64 bf16 = 32 KB

__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* __restrict__ z_img, const uint8_t* __restrict__ a_img,
                                                       float* __restrict__ d_out, uint32_t lbo_enc, uint32_t sbo_enc, int M,
                                                       int k_adv) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* z_s = smem;                 // 2 images (the M = 128 variant stacks image 0 and image 1 along M)
    uint8_t* a_s = smem + 2 * IMG;       // 1 image
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + 3 * IMG);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < 2 * IMG / 16; i += 128) reinterpret_cast<uint4*>(z_s)[i] = reinterpret_cast<const uint4*>(z_img)[i];
    for (int i = tid; i < IMG / 16; i += 128) reinterpret_cast<uint4*>(a_s)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = *tmem_slot;

    if (tid == 0) {
        // D = F32, A = B = BF16, A and B MN-major (bits 15, 16), N >> 3 at 17, M >> 4 at 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(H >> 3) << 17) |
                               ((uint32_t)(M >> 4) << 24);
        for (int k = 0; k < ROWS / 16; ++k) {
            const uint64_t da = make_desc(smem_u32(z_s) + k * k_adv, lbo_enc, sbo_enc);
            const uint64_t db = make_desc(smem_u32(a_s) + k * k_adv, lbo_enc, sbo_enc);
            const uint32_t acc = k ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(taddr),
                "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
    }
    asm volatile(
        "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(
            smem_u32(mbar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // dump all 128 TMEM lanes x 64 columns; the host interprets the lane layout
    const uint32_t lane_addr = taddr + ((uint32_t)(warp * 32) << 16);
    const int tl = warp * 32 + lane;
#pragma unroll
    for (int c0 = 0; c0 < H; c0 += 16) {
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(lane_addr + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) d_out[tl * H + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(taddr));
}

static inline size_t sw128_offset(int r, int u) {   // bf16 element (row r, unit u) of a [rows x 64] image
    const int b = u * 2;
    return (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 + (size_t)(((b >> 4) ^ (r & 7)) << 4) + (b & 15);
}

int main(int argc, char** argv) {
    const uint32_t lbo = argc > 1 ? atoi(argv[1]) : 0, sbo = argc > 2 ? atoi(argv[2]) : 64;
    const int M = argc > 3 ? atoi(argv[3]) : 64, k_adv = argc > 4 ? atoi(argv[4]) : 2048;
    // small integers: every product and partial sum is exact in bf16 x bf16 -> fp32
    std::vector<float> Z0(ROWS * H), Z1(ROWS * H), A(ROWS * H);
    srand(3);
    for (auto& x : Z0) x = (float)(rand() % 7 - 3);
    for (auto& x : Z1) x = (float)(rand() % 5 - 2);
    for (auto& x : A) x = (float)(rand() % 7 - 3);
    std::vector<uint8_t> zi(2 * IMG, 0), ai(IMG, 0);
    for (int r = 0; r < ROWS; ++r)
        for (int u = 0; u < H; ++u) {
            *reinterpret_cast<__nv_bfloat16*>(&zi[sw128_offset(r, u)]) = __float2bfloat16(Z0[r * H + u]);
            *reinterpret_cast<__nv_bfloat16*>(&zi[IMG + sw128_offset(r, u)]) = __float2bfloat16(Z1[r * H + u]);
            *reinterpret_cast<__nv_bfloat16*>(&ai[sw128_offset(r, u)]) = __float2bfloat16(A[r * H + u]);
        }
    std::vector<double> ref(128 * H, 0.0);   // rows 0..63: Z0^T A, rows 64..127: Z1^T A
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < H; ++n) {
            double s = 0;
            for (int r = 0; r < ROWS; ++r) s += (double)(m < 64 ? Z0[r * H + m] : Z1[r * H + m - 64]) * (double)A[r * H + n];
            ref[m * H + n] = s;
        }
    uint8_t *dz, *da;
    float* dd;
    const int smem_bytes = 3 * IMG + 64;
    CK(cudaMalloc(&dz, 2 * IMG)); CK(cudaMalloc(&da, IMG)); CK(cudaMalloc(&dd, 128 * H * 4));
    CK(cudaMemcpy(dz, zi.data(), zi.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dd, 0xff, 128 * H * 4));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    probe_kernel<<<1, 128, smem_bytes>>>(dz, da, dd, lbo, sbo, M, k_adv);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> out(128 * H);
    CK(cudaMemcpy(out.data(), dd, out.size() * 4, cudaMemcpyDeviceToHost));
    // candidate lane layouts: (a) row m in lane m; (b) M = 64: row m in lane (m % 16) + 32 * (m / 16)
    auto err_for = [&](int layout) {
        double e = 0;
        for (int m = 0; m < M; ++m) {
            const int tl = layout == 0 ? m : (m % 16) + 32 * (m / 16);
            for (int n = 0; n < H; ++n) e = fmax(e, fabs((double)out[tl * H + n] - ref[m * H + n]));
        }
        return e;
    };
    const double e_lin = err_for(0), e_q16 = M == 64 ? err_for(1) : -1.0;
    printf("wgrad probe lbo=%u sbo=%u M=%d k_adv=%d : max|err| lane=row %.3g   lane=16-row quarters %.3g   D[0][0..3] = %g %g %g %g  ref %g %g %g %g\n",
           lbo, sbo, M, k_adv, e_lin, e_q16, out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
    return (e_lin == 0.0 || e_q16 == 0.0) ? 0 : 1;
}
