// pinnjet_k1tc2.cuh -- K1-TC v2: tensor-core forward kernel with a transposed epilogue (PINNJET_TC=2).
//
// Same GEMM formulation as pinnjet_k1tc.cuh (rows r = C*p + c, bf16x3 split operands in K-major SWIZZLE_128B images, fp32
// accumulators in TMEM), but the epilogue no longer works "one thread = one (point, channel) row":
//   * 16 compute warps.  Warp (g, q, hf) reads TMEM lanes 32q..32q+31 of accumulator g (rows 128g + 32q + lane) and the
//     32 hidden units of column half hf with ONE tcgen05.ld.32x32b.x32, parks the 32 x 32 block in shared memory (inside
//     its own, currently dead, chunks of the A images) and reads it back transposed: every thread then owns ALL channels
//     of 2 adjacent points x UG adjacent units (UG = 4 for C = 4, 8 for C = 2).  tanh/sin is evaluated once per
//     (point, unit) instead of once per channel lane, the jet rule needs no shuffles, the z-jet records leave as 8-byte
//     stores of point pairs and the next layer's A rows as 8/16-byte stores of adjacent units.
//   * Layer 0 is computed directly in that layout (no MMA); the OUTPUT Linear is one more MMA (N = 16, W_out images built
//     in the kernel prologue), so the last hidden layer goes through the same epilogue as the others.
//   * The two 128-row halves of a tile (g = 0, 1) are independent pipelines with their own named barrier, mbarrier and
//     MMA-issuing thread: the MMAs of one half run under the epilogue of the other.
#pragma once
#include "pinnjet_k1tc.cuh"

namespace pj {

constexpr int TC2_WOUT = 16 * 128;   // bytes of one output-layer image: 16 rows (outputs, zero padded) x 64 bf16

template <int C>
struct Tc2Map {
    static constexpr int PW = 32 / C;      // points per warp (32 TMEM lanes)
    static constexpr int NPP = PW / 2;     // point pairs per warp
    static constexpr int NUG = 32 / NPP;   // unit groups per 32-unit column half
    static constexpr int UG = 32 / NUG;    // adjacent units owned by a thread: 4 (C = 4) or 8 (C = 2)
};

// x = t1 + t2 + t3 (three bf16 terms) for a pair of values; the lower half of each word is the first value
__device__ __forceinline__ void split3_bf16(float x0, float x1, uint32_t& t1, uint32_t& t2, uint32_t& t3) {
    t1 = pack_bf16x2(x0, x1);
    const float r0 = x0 - bf16_lo_f32(t1), r1 = x1 - bf16_hi_f32(t1);
    t2 = pack_bf16x2(r0, r1);
    t3 = pack_bf16x2(r0 - bf16_lo_f32(t2), r1 - bf16_hi_f32(t2));
}

template <int NTHREADS>
__device__ __forceinline__ void bar_named(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NTHREADS) : "memory");
}

template <int N1, int N2, int WL>
__global__ void __launch_bounds__(576, 1) k1tc2_forward_kernel(const __grid_constant__ K1Args A) {
    constexpr int C = 1 + N1 + N2;
    static_assert(C == 2 || C == 4, "tensor-core forward kernel: 2 or 4 jet channels");
    using M = Tc2Map<C>;
    constexpr int UG = M::UG;
    constexpr int N_CWARPS = 16, NT_COMPUTE = 512, NT_TOTAL = 576, WS_STRIDE = 256;
    constexpr int T = TC_ROWS / C;
    extern __shared__ __align__(1024) unsigned char smem[];
    const PjSpec& sp = A.spec;
    const Plan& pl = A.plan;
    unsigned char* aimg = smem + pl.k1_act;                     // 3 x 32 KB, 1024-aligned
    unsigned char* wimg = smem + pl.k1_ring;                    // [hidden->hidden layer][3] x 8 KB, then [net][3] x 2 KB
    float* small = reinterpret_cast<float*>(smem + pl.k1_small);
    float* ycache = reinterpret_cast<float*>(smem + pl.k1_ycache);
    float* slots = reinterpret_cast<float*>(smem + pl.k1_slots);
    int4* prog_s = reinterpret_cast<int4*>(smem + pl.k1_prog);
    int4* progw_s = reinterpret_cast<int4*>(smem + pl.k1_progw);
    float* wbuf = reinterpret_cast<float*>(smem + pl.k1_wbuf);
    float* wslots = reinterpret_cast<float*>(smem + pl.k1_wslots);
    uint64_t* wfull = reinterpret_cast<uint64_t*>(smem + pl.k1_misc);   // W images landed
    uint64_t* mma_done = wfull + 1;                                     // [2]: one per 128-row half
    uint64_t* yfull = mma_done + 2;
    uint64_t* yempty = yfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(yempty + 2);
    const int EB = pl.epi_batch;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T2 = pl.T, RS2 = pl.RS;
    const long long ws_points = (long long)pl.n_tiles * T2;
    const int my_tiles = (pl.n_tiles1 > (int)blockIdx.x) ? (pl.n_tiles1 - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    int n_hh = 0;
    for (int n = 0; n < sp.n_nets; ++n) n_hh += sp.net[n].n_linear - 2;
    unsigned char* woutimg = wimg + (size_t)n_hh * 3 * TC_WIMG;

    if (tid == 0) {
        mbar_init(wfull, 1);
        mbar_init(&mma_done[0], 1);
        mbar_init(&mma_done[1], 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&yfull[b], 1);
            mbar_init(&yempty[b], 1);
        }
        fence_barrier_init();
    }
    if (warp == 0) {   // columns 0..127: two [128 x 64] hidden accumulators; 128..159: two [128 x 16] output accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)));
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

    // output-layer images: B operand [16 x 64] of  y = a_L W_out^T, rows >= n_out are zero
    for (int n = 0; n < sp.n_nets; ++n) {
        const int n_out = sp.net[n].width[sp.net[n].n_linear];
        const float* wl_t = small + pl.s_wlt[n];   // [unit][output]
        unsigned char* img = woutimg + (size_t)n * 3 * TC2_WOUT;
        for (int e = tid; e < 16 * 32; e += NT_TOTAL) {
            const int o = e >> 5, u = (e & 31) * 2;
            const float x0 = o < n_out ? wl_t[u * n_out + o] : 0.0f, x1 = o < n_out ? wl_t[(u + 1) * n_out + o] : 0.0f;
            uint32_t t1, t2, t3;
            split3_bf16(x0, x1, t1, t2, t3);
            const uint32_t off = sw128_off(o, u >> 3) + (uint32_t)(u & 7) * 2u;
            *reinterpret_cast<uint32_t*>(img + off) = t1;
            *reinterpret_cast<uint32_t*>(img + TC2_WOUT + off) = t2;
            *reinterpret_cast<uint32_t*>(img + 2 * TC2_WOUT + off) = t3;
        }
    }
    fence_proxy_async();
    __syncthreads();

    if (warp == N_CWARPS) {   // ---------------- producer warp: all hidden->hidden W images, once ----------------
        if (lane == 0 && my_tiles > 0) {
            const uint32_t total = (uint32_t)n_hh * 3u * TC_WIMG;
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

    // -------------------------------------------- compute warps --------------------------------------------------------
    const int hf = warp >> 3, g = (warp >> 2) & 1, q = warp & 3;
    const int rowbase = g * 128 + q * 32;                          // first GEMM row of this warp's TMEM lanes
    const int myrow = rowbase + lane;                              // the row this thread reads from TMEM
    const int ppidx = lane / M::NUG, ug = lane % M::NUG;
    const int p0 = rowbase / C + 2 * ppidx;                        // first of the two (adjacent) tile-local points owned
    const int ubase = hf * 32 + ug * UG;                           // first of the UG adjacent hidden units owned
    const uint32_t tmem_hid = tmem_base + (uint32_t)(g * TC_H + hf * 32) + ((uint32_t)(q * 32) << 16);
    const uint32_t tmem_out = tmem_base + 128u + (uint32_t)(g * 16) + ((uint32_t)(q * 32) << 16);
    // Swizzled byte offsets.  The 2C rows owned in the epilogue are R0 + j, j = C*pp + c, with R0 a multiple of 2C: R0 & 7
    // is 0 (C = 4) or 0 / 4 (C = 2, j < 4), so (R0 + j) & 7 = (R0 & 7) ^ j and the offset of 16-byte chunk k of row R0 + j is
    //   own_row + j * 128 + (((k ^ (R0 & 7)) << 4) ^ (j << 4)):   only the last term depends on j, as an immediate.
    const int R0 = rowbase + 2 * C * ppidx;
    unsigned char* own_row = aimg + (R0 >> 3) * 1024 + (R0 & 7) * 128;
    const uint32_t awr_c = (uint32_t)(((hf * 4 + (UG == 4 ? (ug >> 1) : ug)) ^ (R0 & 7)) << 4);   // chunk of the A-row store
    const uint32_t awr_b = UG == 4 ? (uint32_t)(ug & 1) * 8u : 0u;                                 // byte inside that chunk
    unsigned char* stage_wr = aimg + (myrow >> 3) * 1024 + (myrow & 7) * 128;                      // staging: own TMEM row
    const uint32_t stage_c = (uint32_t)((hf * 4) ^ (myrow & 7));
    const bool leader = (warp == 4 * g) && lane == 0;              // issues the MMAs of half g
    const bool train = A.mode == 1;
    uint32_t mma_phase = 0;
    int bslot = 0, batch_idx = 0;
    mbar_wait(wfull, 0);

    for (int iter = 0; iter < my_tiles; ++iter) {
        const long long tile = (long long)blockIdx.x + (long long)iter * gridDim.x;
        const long long base = tile * T;
        const long long gp0 = base + p0;
        if (bslot == 0 && batch_idx >= 2) mbar_wait(&yempty[batch_idx & 1], (uint32_t)(((batch_idx >> 1) - 1) & 1));
        float* yb = ycache + (size_t)(batch_idx & 1) * sp.n_yrows * EB + bslot * T;
        const bool rec = train && gp0 < ws_points;                 // both points of the pair lie in the same K2 tile
        float* zj_pair = rec ? A.zj + (gp0 / T2) * pl.zj_tile_floats + (gp0 % T2) : nullptr;

        if constexpr (WL > 0) {   // weights of the combined second-order channel
            const int NW = sp.n_nets * WL;
            if (tid < T) {
                ProgIO io{A.coords, min(base + tid, A.N - 1), A.N, nullptr, 0, nullptr, 0.0f, nullptr, nullptr, nullptr, T2};
                io.w_out = wbuf + tid;
                io.w_stride = T;
                run_program<WS_STRIDE>(progw_s, A.prog_w_len, wslots + tid, io);
            }
            bar_named<NT_COMPUTE>(1);
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
            float wq[2][WL > 0 ? WL : 1];
#pragma unroll
            for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                for (int d = 0; d < (WL > 0 ? WL : 1); ++d) wq[pp][d] = WL > 0 ? wbuf[(n * WL + d) * T + p0 + pp] : 0.0f;
            const float* wt0 = small + pl.s_wt0[n];
            const float* dzt = small + pl.s_dz[n];

            for (int h = 1; h <= L; ++h) {   // produce the a-jets of hidden layer h
                const float* bias = small + pl.s_b[n][h - 1];
                float z[2][C][UG];
                if (h == 1) {   // Linear 0 from the coordinates (first-order channels are columns of W0 . dir)
                    float xin[2][PJ_MAX_COORDS];
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                        for (int i = 0; i < PJ_MAX_COORDS; ++i)
                            xin[pp][i] = (i < net.n_in) ? __ldg(A.coords[net.in_coord[i]] + min(gp0 + pp, A.N - 1)) : 0.0f;
#pragma unroll
                    for (int k = 0; k < UG; ++k) {
                        const int u = ubase + k;
                        float s0 = bias[u], s1v = s0;
#pragma unroll
                        for (int i = 0; i < PJ_MAX_COORDS; ++i)
                            if (i < net.n_in) {
                                const float w = wt0[i * TC_H + u];
                                s0 = fmaf(w, xin[0][i], s0);
                                s1v = fmaf(w, xin[1][i], s1v);
                            }
                        z[0][0][k] = s0;
                        z[1][0][k] = s1v;
#pragma unroll
                        for (int c = 1; c < C; ++c) {
                            const float d = c <= N1 ? dzt[(c - 1) * TC_H + u] : 0.0f;
                            z[0][c][k] = d;
                            z[1][c][k] = d;
                        }
                    }
                } else {
                    mbar_wait(&mma_done[g], mma_phase);
                    mma_phase ^= 1u;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    {   // own row, 32 units: TMEM -> registers -> staging chunks inside this warp's part of the A images
                        uint32_t v[32];
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                              "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
                              "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
                              "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                            : "r"(tmem_hid));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int s = 0; s < 8; ++s)   // slot s = units hf*32 + 4s .. +3 -> image s>>2, chunk hf*4 + (s&3)
                            *reinterpret_cast<uint4*>(stage_wr + (s >> 2) * TC_AIMG + ((stage_c ^ (uint32_t)(s & 3)) << 4)) =
                                make_uint4(v[4 * s], v[4 * s + 1], v[4 * s + 2], v[4 * s + 3]);
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
#pragma unroll
                    for (int s4 = 0; s4 < UG / 4; ++s4) {
                        const int s = ug * (UG / 4) + s4;
                        const unsigned char* src = own_row + (s >> 2) * TC_AIMG;
                        const uint32_t sc = (uint32_t)(((hf * 4 + (s & 3)) ^ (R0 & 7)) << 4);
#pragma unroll
                        for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                            for (int c = 0; c < C; ++c) {
                                const int j = C * pp + c;
                                const float4 t = *reinterpret_cast<const float4*>(src + j * 128 + (sc ^ (uint32_t)(j << 4)));
                                z[pp][c][4 * s4 + 0] = t.x;
                                z[pp][c][4 * s4 + 1] = t.y;
                                z[pp][c][4 * s4 + 2] = t.z;
                                z[pp][c][4 * s4 + 3] = t.w;
                            }
                        }
                    __syncwarp();   // every lane has its block: the staging chunks may now be overwritten by A rows
#pragma unroll
                    for (int k = 0; k < UG; ++k) {
                        const float b = bias[ubase + k];
                        z[0][0][k] += b;
                        z[1][0][k] += b;
                    }
                }

                // activation-jet rule, z-jet records (8-byte stores of the point pair), a-jets left in z
                float* zrec = rec ? zj_pair + pl.zj_off[n][h] : nullptr;
#pragma unroll
                for (int k = 0; k < UG; ++k) {
                    float a0[C], a1[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        a0[c] = z[0][c][k];
                        a1[c] = z[1][c][k];
                    }
                    act_forward<N1, N2, WL>(act_kind, a0, wq[0]);
                    act_forward<N1, N2, WL>(act_kind, a1, wq[1]);
                    if (zrec) {   // record: channel 0 = tanh(z0) for tanh nets / z0 for sin nets, other channels z-jets
                        float* zr = zrec + (size_t)(ubase + k) * RS2;
                        const bool keep_a = act_kind == PJ_ACT_TANH;
                        *reinterpret_cast<float2*>(zr) = make_float2(keep_a ? a0[0] : z[0][0][k], keep_a ? a1[0] : z[1][0][k]);
#pragma unroll
                        for (int c = 1; c < C; ++c)
                            *reinterpret_cast<float2*>(zr + c * T2) = make_float2(z[0][c][k], z[1][c][k]);
                    }
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        z[0][c][k] = a0[c];
                        z[1][c][k] = a1[c];
                    }
                }

                // own part of the next GEMM's A images: three bf16 terms of UG adjacent units per (point, channel) row
#pragma unroll
                for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int j = C * pp + c;
                        uint32_t t1[UG / 2], t2[UG / 2], t3[UG / 2];
#pragma unroll
                        for (int e = 0; e < UG / 2; ++e) split3_bf16(z[pp][c][2 * e], z[pp][c][2 * e + 1], t1[e], t2[e], t3[e]);
                        unsigned char* dst = own_row + j * 128 + ((awr_c ^ (uint32_t)(j << 4)) + awr_b);
                        if constexpr (UG == 4) {
                            *reinterpret_cast<uint2*>(dst) = make_uint2(t1[0], t1[1]);
                            *reinterpret_cast<uint2*>(dst + TC_AIMG) = make_uint2(t2[0], t2[1]);
                            *reinterpret_cast<uint2*>(dst + 2 * TC_AIMG) = make_uint2(t3[0], t3[1]);
                        } else {
                            *reinterpret_cast<uint4*>(dst) = make_uint4(t1[0], t1[1], t1[2], t1[3]);
                            *reinterpret_cast<uint4*>(dst + TC_AIMG) = make_uint4(t2[0], t2[1], t2[2], t2[3]);
                            *reinterpret_cast<uint4*>(dst + 2 * TC_AIMG) = make_uint4(t3[0], t3[1], t3[2], t3[3]);
                        }
                    }

                // rows of this half are complete -> its leader issues Linear h (hidden -> hidden) or the output Linear
                fence_proxy_async();
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                bar_named<256>(2 + g);
                if (leader) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const bool out_layer = h == L;
                    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((out_layer ? 16 : TC_H) >> 3) << 17) | (8u << 24);
                    const uint32_t a_base = smem_u32(aimg) + (uint32_t)g * (TC_AIMG / 2);
                    const uint32_t b_base = out_layer ? smem_u32(woutimg + (size_t)n * 3 * TC2_WOUT)
                                                      : smem_u32(wimg + (size_t)wslot * 3 * TC_WIMG);
                    const uint32_t b_img = out_layer ? TC2_WOUT : TC_WIMG;
                    const uint32_t d_addr = out_layer ? tmem_base + 128u + (uint32_t)(g * 16) : tmem_base + (uint32_t)(g * TC_H);
                    const int pa[6] = {0, 0, 1, 1, 0, 2}, pb[6] = {0, 1, 0, 1, 2, 0};
#pragma unroll
                    for (int pr = 0; pr < 6; ++pr)
#pragma unroll
                        for (int k = 0; k < TC_H / 16; ++k) {
                            const uint64_t da = umma_desc_sw128(a_base + pa[pr] * TC_AIMG + k * 32);
                            const uint64_t db = umma_desc_sw128(b_base + pb[pr] * b_img + k * 32);
                            const uint32_t accf = (pr | k) ? 1u : 0u;
                            asm volatile(
                                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_addr),
                                "l"(da), "l"(db), "r"(idesc), "r"(accf)
                                : "memory");
                        }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     smem_u32(&mma_done[g]))
                                 : "memory");
                }
                if (h < L) ++wslot;
            }

            // output jets of this net: row = (point, channel), n_out columns
            mbar_wait(&mma_done[g], mma_phase);
            mma_phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (hf == 0) {
                uint32_t v[4];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                             : "r"(tmem_out));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int pt = myrow / C, ch = myrow % C;
                const float* bo = small + pl.s_bout[n];
#pragma unroll
                for (int o = 0; o < PJ_MAX_NETS; ++o)
                    if (o < n_out) yb[(net.yrow0 + o * C + ch) * EB + pt] = __uint_as_float(v[o]) + (ch == 0 ? bo[o] : 0.0f);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        bar_named<NT_COMPUTE>(1);   // jets of the tile are in the batch table; A images and accumulators are free again
        if (++bslot == tiles_per_batch || iter == my_tiles - 1) {
            if (tid == 0) mbar_arrive(&yfull[batch_idx & 1]);
            ++batch_idx;
            bslot = 0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    bar_named<NT_COMPUTE>(1);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base));
}

}  // namespace pj
