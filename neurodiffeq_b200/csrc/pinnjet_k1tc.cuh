// pinnjet_k1tc.cuh -- K1-TC: the forward kernel with the hidden-layer contractions on the 5th-gen tensor cores.
//
// Same contract as k1_forward_kernel (pinnjet_k1.cuh) for networks whose hidden layers are 64 wide and whose jets have
// C = 2 or 4 channels.  The contraction  z^(c)[p][:] = W a^(c)[p][:]  of ALL channels is one GEMM with 256 rows
// r = C*p + c (T = 256/C points per tile):
//   * A operand: the a-jets of the previous layer, split into THREE bf16 terms (a = a1 + a2 + a3) and written by the
//     epilogue threads straight from registers into three K-major SWIZZLE_128B shared-memory images [256 x 64];
//   * B operand: W_l, pre-split and pre-swizzled by K0, three images [64 x 64], bulk-TMA'd once per CTA;
//   * D: two fp32 accumulators [128 x 64] in TMEM; six tcgen05.mma products per K step (a1w1, a1w2, a2w1, a2w2, a1w3,
//     a3w1) reproduce the fp32 GEMM to ~5e-7 relative (experiments/tcgen05_probe), i.e. inside the parity bound and with
//     the full fp32 exponent range.
// Row r lives in TMEM lane r % 128 of accumulator r / 128, so thread (warp w, lane l) owns row 128*(w/4) + 32*(w%4) + l
// = one (point, channel) pair; the C channels of a point sit in adjacent lanes and are exchanged with warp shuffles for
// the activation-jet rule.  The epilogue thread reads its 64 columns with tcgen05.ld, applies the rule, stores the
// workspace record for K2 (train), and writes its own row of the next layer's A images -- or, after the last hidden
// layer, accumulates the output layer on the fly.  Layer 0, the residual program (service warp), the weight program of
// the combined channel and all buffers are as in the FFMA kernel.
#pragma once
#include "pinnjet_k1.cuh"

namespace pj {

constexpr int TC_ROWS = 256;             // rows of the jet GEMM per tile
constexpr int TC_H = 64;                 // hidden width
constexpr int TC_AIMG = TC_ROWS * 128;   // bytes of one A image (256 rows x 64 bf16)
constexpr int TC_WIMG = TC_H * 128;      // bytes of one W image (64 rows x 64 bf16)

__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    // K-major, 128-byte swizzle, 8-row groups 1024 B apart (SBO), descriptor version 1 (validated by the probe)
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk16) {   // byte offset of 16-byte chunk `chunk16` of `row`
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk16 ^ (row & 7)) << 4));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // d = {hi, lo}: first source -> upper half
    return r;
}
__device__ __forceinline__ float bf16_lo_f32(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi_f32(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

template <int N1, int N2, int WL>
__global__ void __launch_bounds__(320, 1) k1tc_forward_kernel(const __grid_constant__ K1Args A) {
    constexpr int C = 1 + N1 + N2;
    static_assert(C == 2 || C == 4, "tensor-core forward kernel: 2 or 4 jet channels");
    constexpr int NTC = 256, N_CWARPS = 8, NT_TOTAL = 320, NT_COMPUTE = 256;
    constexpr int T = TC_ROWS / C;
    extern __shared__ __align__(1024) unsigned char smem[];
    const PjSpec& sp = A.spec;
    const Plan& pl = A.plan;
    unsigned char* aimg = smem + pl.k1_act;                     // 3 x 32 KB, 1024-aligned
    unsigned char* wimg = smem + pl.k1_ring;                    // [hidden->hidden layer][3] x 8 KB
    float* small = reinterpret_cast<float*>(smem + pl.k1_small);
    float* ycache = reinterpret_cast<float*>(smem + pl.k1_ycache);
    float* slots = reinterpret_cast<float*>(smem + pl.k1_slots);
    int4* prog_s = reinterpret_cast<int4*>(smem + pl.k1_prog);
    int4* progw_s = reinterpret_cast<int4*>(smem + pl.k1_progw);
    float* wbuf = reinterpret_cast<float*>(smem + pl.k1_wbuf);
    float* wslots = reinterpret_cast<float*>(smem + pl.k1_wslots);
    uint64_t* wfull = reinterpret_cast<uint64_t*>(smem + pl.k1_misc);   // W images landed
    uint64_t* mma_done = wfull + 1;
    uint64_t* yfull = mma_done + 1;
    uint64_t* yempty = yfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(yempty + 2);
    const int EB = pl.epi_batch;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T2 = pl.T, RS2 = pl.RS;
    const long long ws_points = (long long)pl.n_tiles * T2;
    const int my_tiles = (pl.n_tiles1 > (int)blockIdx.x) ? (pl.n_tiles1 - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        mbar_init(wfull, 1);
        mbar_init(mma_done, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&yfull[b], 1);
            mbar_init(&yempty[b], 1);
        }
        fence_barrier_init();
    }
    if (warp == 0) {   // 2 accumulators x 64 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < pl.small_floats; i += NT_TOTAL) small[i] = __ldg(A.pack + i);
    for (int i = tid; i < A.prog_len; i += NT_TOTAL) prog_s[i] = __ldg(A.prog + i);
    if constexpr (WL > 0)
        for (int i = tid; i < A.prog_w_len; i += NT_TOTAL) progw_s[i] = __ldg(A.prog_w + i);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == N_CWARPS) {   // ---------------- producer warp: all W images, once ----------------
        if (lane == 0 && my_tiles > 0) {
            uint32_t total = 0;
            for (int n = 0; n < sp.n_nets; ++n) total += (uint32_t)(sp.net[n].n_linear - 2) * 3u * TC_WIMG;
            if (total > 0) {
                mbar_arrive_expect_tx(wfull, total);
                int slot = 0;
                for (int n = 0; n < sp.n_nets; ++n)
                    for (int l = 1; l < sp.net[n].n_linear - 1; ++l, ++slot)
                        tma_bulk_g2s(wimg + (size_t)slot * 3 * TC_WIMG, A.pack + pl.b_wimg[n][l], 3 * TC_WIMG, wfull);
            } else {
                mbar_arrive(wfull);
            }
        }
        return;
    }
    const int tiles_per_batch = EB / T;
    if (warp == N_CWARPS + 1) {   // ---------------- program warp (as in k1_forward_kernel) ----------------
        const bool train_pw = A.mode == 1;
        float my_sumsq = 0.0f;
        const int n_batches = (my_tiles + tiles_per_batch - 1) / tiles_per_batch;
        for (int b = 0; b < n_batches; ++b) {
            const int buf = b & 1;
            mbar_wait(&yfull[buf], (uint32_t)((b >> 1) & 1));
            const float* yb = ycache + (size_t)buf * sp.n_yrows * EB;
            const int first_iter = b * tiles_per_batch;
            const int npts = min(tiles_per_batch, my_tiles - first_iter) * T;
            for (int bp = lane; bp < npts; bp += 32) {
                const int tl = bp / T, pt = bp - tl * T;
                const long long btile = (long long)blockIdx.x + (long long)(first_iter + tl) * gridDim.x;
                const long long gidx = btile * T + pt;
                float* seed_tile = (train_pw && gidx < ws_points)
                                       ? A.seeds + (gidx / T2) * ((long long)sp.n_yrows * T2) + (gidx % T2) : nullptr;
                if (gidx < A.N) {
                    ProgIO io{A.coords, gidx, A.N, yb + bp, EB, A.rbar, A.loss_scale, A.u_out, A.r_out, seed_tile, T2};
                    my_sumsq += run_program<32>(prog_s, A.prog_len, slots + lane, io);
                } else if (seed_tile) {
                    for (int r = 0; r < sp.n_yrows; ++r) seed_tile[r * T2] = 0.0f;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&yempty[buf]);
        }
        my_sumsq = warp_sum(my_sumsq);
        if (lane == 0) A.loss_part[blockIdx.x] = my_sumsq;
        return;
    }

    // -------------------------------------------- compute warps: one (point, channel) row each ------------------------
    const int acc_m = warp >> 2;                                   // accumulator / 128-row half
    const int row = acc_m * 128 + (warp & 3) * 32 + lane;          // GEMM row = TMEM lane (mod 128)
    const int pt = row / C, ch = row % C;
    const int seg = lane & ~(C - 1);                               // first lane of this point's channel group
    const uint32_t tmem_row = tmem_base + (uint32_t)(acc_m * TC_H) + ((uint32_t)((warp & 3) * 32) << 16);
    const bool train = A.mode == 1;
    uint32_t mma_phase = 0;
    int bslot = 0, batch_idx = 0;
    mbar_wait(wfull, 0);

    for (int iter = 0; iter < my_tiles; ++iter) {
        const long long tile = (long long)blockIdx.x + (long long)iter * gridDim.x;
        const long long base = tile * T;
        const long long gp = base + pt;
        if (bslot == 0 && batch_idx >= 2) mbar_wait(&yempty[batch_idx & 1], (uint32_t)(((batch_idx >> 1) - 1) & 1));
        float* yb = ycache + (size_t)(batch_idx & 1) * sp.n_yrows * EB + bslot * T;
        const bool rec = train && gp < ws_points;
        float* zj_pt = rec ? A.zj + (gp / T2) * pl.zj_tile_floats + ch * T2 + (gp % T2) : nullptr;

        if constexpr (WL > 0) {   // weights of the combined second-order channel
            const int NW = sp.n_nets * WL;
            if (tid < T) {
                ProgIO io{A.coords, min(base + tid, A.N - 1), A.N, nullptr, 0, nullptr, 0.0f, nullptr, nullptr, nullptr, T2};
                io.w_out = wbuf + tid;
                io.w_stride = T;
                run_program<NTC>(progw_s, A.prog_w_len, wslots + tid, io);
            }
            bar_compute<NTC>();
            if (train)
                for (int e = tid; e < NW * T; e += NT_COMPUTE) {
                    const int wr = e / T, wp = e - wr * T;
                    const long long g2 = base + wp;
                    if (g2 < ws_points) A.wts[(g2 / T2) * ((long long)NW * T2) + wr * T2 + (g2 % T2)] = wbuf[e];
                }
        }

        int wslot = 0;   // W image slot (hidden->hidden Linears in net order)
        for (int n = 0; n < sp.n_nets; ++n) {
            const PjNet& net = sp.net[n];
            const int L = net.n_linear - 1;
            const int act_kind = net.act;
            const int n_out = net.width[net.n_linear];
            float wq[WL > 0 ? WL : 1];
#pragma unroll
            for (int d = 0; d < (WL > 0 ? WL : 1); ++d) wq[d] = WL > 0 ? wbuf[(n * WL + d) * T + pt] : 0.0f;
            float xin[PJ_MAX_COORDS];
#pragma unroll
            for (int i = 0; i < PJ_MAX_COORDS; ++i)
                xin[i] = (i < net.n_in) ? __ldg(A.coords[net.in_coord[i]] + min(gp, A.N - 1)) : 0.0f;
            const float* wt0 = small + pl.s_wt0[n];
            const float* dzt = small + pl.s_dz[n];
            const float* wl_t = small + pl.s_wlt[n];
            float yacc[PJ_MAX_NETS];
#pragma unroll
            for (int o = 0; o < PJ_MAX_NETS; ++o) yacc[o] = 0.0f;

            for (int h = 1; h <= L; ++h) {   // produce the a-jets of hidden layer h
                const float* bias = small + pl.s_b[n][h - 1];
                float* zrec = rec ? zj_pt + pl.zj_off[n][h] : nullptr;
                if (h > 1) {
                    // GEMM of Linear h-1: every row of the A images is written -> make it visible to the tensor core
                    fence_proxy_async();
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    bar_compute<NTC>();
                    if (tid == 0) {
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_H >> 3) << 17) | (8u << 24);
                        const uint32_t a0 = smem_u32(aimg), w0 = smem_u32(wimg + (size_t)wslot * 3 * TC_WIMG);
                        const int pa[6] = {0, 0, 1, 1, 0, 2}, pb[6] = {0, 1, 0, 1, 2, 0};
#pragma unroll
                        for (int m = 0; m < 2; ++m)
#pragma unroll
                            for (int pr = 0; pr < 6; ++pr)
#pragma unroll
                                for (int k = 0; k < TC_H / 16; ++k) {
                                    const uint64_t da = umma_desc_sw128(a0 + pa[pr] * TC_AIMG + m * (TC_AIMG / 2) + k * 32);
                                    const uint64_t db = umma_desc_sw128(w0 + pb[pr] * TC_WIMG + k * 32);
                                    const uint32_t accf = (pr | k) ? 1u : 0u;
                                    asm volatile(
                                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(
                                            tmem_base + (uint32_t)(m * TC_H)),
                                        "l"(da), "l"(db), "r"(idesc), "r"(accf)
                                        : "memory");
                                }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                         smem_u32(mma_done))
                                     : "memory");
                    }
                    mbar_wait(mma_done, mma_phase);
                    mma_phase ^= 1u;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    ++wslot;
                }
#pragma unroll 1
                for (int u0 = 0; u0 < TC_H; u0 += 16) {
                    float zv[16];
                    if (h == 1) {   // Linear 0 from the coordinates (first-order channels are columns of W0)
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int u = u0 + j;
                            float s;
                            if (ch == 0) {
                                s = bias[u];
#pragma unroll
                                for (int i = 0; i < PJ_MAX_COORDS; ++i)
                                    if (i < net.n_in) s = fmaf(wt0[i * TC_H + u], xin[i], s);
                            } else if (ch <= N1) {
                                s = dzt[(ch - 1) * TC_H + u];
                            } else {
                                s = 0.0f;
                            }
                            zv[j] = s;
                        }
                    } else {
                        uint32_t v[16];
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                              "=r"(v[15])
                            : "r"(tmem_row + (uint32_t)u0));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int j = 0; j < 16; ++j) zv[j] = __uint_as_float(v[j]) + (ch == 0 ? bias[u0 + j] : 0.0f);
                    }
                    // activation-jet rule across the C lanes of the point
                    float av[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float z0 = __shfl_sync(0xffffffffu, zv[j], seg);
                        float a0, s1, s2;
                        act_d2(act_kind, z0, a0, s1, s2);
                        float q = 0.0f;
                        if constexpr (WL > 0) {
#pragma unroll
                            for (int d = 0; d < WL; ++d) {
                                const float zd = __shfl_sync(0xffffffffu, zv[j], seg + 1 + d);
                                q = fmaf(wq[d] * zd, zd, q);
                            }
                        } else if constexpr (N2 > 0) {
                            const float zp = __shfl_sync(0xffffffffu, zv[j], seg + 1 + max(ch - 1 - N1, 0));
                            q = zp * zp;
                        }
                        float a;
                        if (ch == 0) a = a0;
                        else if (ch <= N1) a = s1 * zv[j];
                        else a = fmaf(s2, q, s1 * zv[j]);
                        av[j] = a;
                        if (zrec)   // record: channel 0 = tanh(z0) for tanh nets / z0 for sin nets, other channels z-jets
                            zrec[(size_t)(u0 + j) * RS2] = (ch == 0) ? (act_kind == PJ_ACT_TANH ? a0 : z0) : zv[j];
                    }
                    if (h < L) {   // own row of the next GEMM's A images: three bf16 terms, swizzled 16-byte chunks
                        uint32_t t1[8], t2[8], t3[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float x0 = av[2 * j], x1 = av[2 * j + 1];
                            t1[j] = pack_bf16x2(x0, x1);
                            const float r0 = x0 - bf16_lo_f32(t1[j]), r1 = x1 - bf16_hi_f32(t1[j]);
                            t2[j] = pack_bf16x2(r0, r1);
                            t3[j] = pack_bf16x2(r0 - bf16_lo_f32(t2[j]), r1 - bf16_hi_f32(t2[j]));
                        }
                        const int ck = u0 >> 3;
                        const uint32_t o0 = sw128_off(row, ck), o1 = sw128_off(row, ck + 1);
                        *reinterpret_cast<uint4*>(aimg + o0) = make_uint4(t1[0], t1[1], t1[2], t1[3]);
                        *reinterpret_cast<uint4*>(aimg + o1) = make_uint4(t1[4], t1[5], t1[6], t1[7]);
                        *reinterpret_cast<uint4*>(aimg + TC_AIMG + o0) = make_uint4(t2[0], t2[1], t2[2], t2[3]);
                        *reinterpret_cast<uint4*>(aimg + TC_AIMG + o1) = make_uint4(t2[4], t2[5], t2[6], t2[7]);
                        *reinterpret_cast<uint4*>(aimg + 2 * TC_AIMG + o0) = make_uint4(t3[0], t3[1], t3[2], t3[3]);
                        *reinterpret_cast<uint4*>(aimg + 2 * TC_AIMG + o1) = make_uint4(t3[4], t3[5], t3[6], t3[7]);
                    } else {       // last hidden layer: the output Linear on the fly
#pragma unroll
                        for (int j = 0; j < 16; ++j)
#pragma unroll
                            for (int o = 0; o < PJ_MAX_NETS; ++o)
                                if (o < n_out) yacc[o] = fmaf(wl_t[(u0 + j) * n_out + o], av[j], yacc[o]);
                    }
                }
                if (h > 1) {   // all TMEM reads of this layer are done before the next GEMM overwrites the accumulators
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                }
            }
            const float* bo = small + pl.s_bout[n];
#pragma unroll
            for (int o = 0; o < PJ_MAX_NETS; ++o)
                if (o < n_out) yb[(net.yrow0 + o * C + ch) * EB + pt] = yacc[o] + (ch == 0 ? bo[o] : 0.0f);
        }
        bar_compute<NTC>();   // jets of the tile are in the batch table (also orders the last TMEM reads before the next GEMM)
        if (++bslot == tiles_per_batch || iter == my_tiles - 1) {
            if (tid == 0) mbar_arrive(&yfull[batch_idx & 1]);
            ++batch_idx;
            bslot = 0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    bar_compute<NTC>();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base));
}

}  // namespace pj
