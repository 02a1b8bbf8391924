// pinnjet_k2tc.cuh -- EXPERIMENTAL (round-2 work in progress, compiled only with -DPJ_EXPERIMENTAL into libpinnjet_exp.so,
// selected at run time with PINNJET_LIB=.../libpinnjet_exp.so PINNJET_TC_BWD=1; NOT yet run on a GPU).
//
// K2-TC: the reverse pass with both GEMMs of every hidden->hidden Linear on the 5th-gen tensor cores.  Same contract as
// k2_backward_kernel (pinnjet_k2.cuh): reads the seeds / z-jet records / combined-channel weights K1 left in the workspace,
// writes this CTA's gradient partial (K2b sums them).  Eligible problems: every hidden layer 64 wide, 2 or 4 jet channels,
// at most 3 hidden->hidden Linears in total (TMEM columns and shared memory).
//
// Formulation (all operand encodings validated by experiments/tcgen05_probe, profiles/r01/tcgen05_probe*.log):
//   * tile = 128 GEMM rows r = C*p + c (T = 128/C points; a multiple of the record tile pl.T), one thread per TMEM lane for
//     the loads, the transposed "owner" layout of pinnjet_k1tc2.cuh (2 adjacent points x UG adjacent units x all channels
//     per thread) for all arithmetic;
//   * z_bar_h and a_{h-1} live as bf16x3 split images [128 rows x 64 units], K-major SWIZZLE_128B (ZIMG, AIMG);
//   * adjoint GEMM   a_bar_{h-1}[r][k] = sum_j z_bar_h[r][j] W_l[j][k]:  A = ZIMG (K-major), B = the FORWARD weight images
//     of W_l read MN-major (no transposed copy), D = [128 x 64] fp32 in TMEM, 6 split products;
//   * weight-gradient GEMM  W_bar_l[j][k] += sum_r z_bar_h[r][j] a_{h-1}[r][k]:  A = ZIMG and B = AIMG both read MN-major
//     (contraction over the rows: +2048 B per K = 16), M = 64 accumulator [64 x 64] per Linear that stays in TMEM for ALL
//     tiles of the CTA and is read once at the end (row j in lane j%16 + 32*(j/16));
//   * a_{h-1} depends on the record only, so its image is built -- and the weight-gradient MMAs are issued -- while the
//     adjoint MMAs of the same layer are still running;
//   * last Linear, first Linear and all bias gradients: CUDA cores in the owner layout, reduced over the point lanes with
//     shuffles and accumulated in shared memory.
#pragma once
#include "pinnjet_k2.cuh"
#include "pinnjet_k1tc2.cuh"

namespace pj {

constexpr int K2T_ROWS = 128;
constexpr int K2T_IMG = K2T_ROWS * 128;        // bytes of one split image [128 rows x 64 bf16]
constexpr int K2T_STAGE_STRIDE = 36;           // floats per staged TMEM row (32 + 4: conflict-free 16-byte accesses)
constexpr int K2T_NCW = 8;                     // compute warps

struct K2tcLayout {
    int zimg, aimg, wimg, rec, stage, small, ybar, sgrad, misc, bytes;
    int T, n_sub, n_hh, rec_sub_floats, tmem_cols;
    bool ok;
};

__host__ __device__ inline K2tcLayout k2tc_layout(const PjSpec& sp, const Plan& pl) {
    K2tcLayout L{};
    const int C = pl.C;
    L.ok = (C == 2 || C == 4) && pl.hmax == 64;
    L.n_hh = 0;
    for (int n = 0; n < sp.n_nets; ++n) {
        for (int h = 1; h < sp.net[n].n_linear; ++h) L.ok = L.ok && pl.hp[n][h] == 64;
        L.n_hh += sp.net[n].n_linear - 2;
    }
    L.T = K2T_ROWS / (C > 0 ? C : 1);
    L.ok = L.ok && L.n_hh <= 3 && pl.T > 0 && L.T % pl.T == 0;
    L.n_sub = pl.T > 0 ? L.T / pl.T : 1;
    L.rec_sub_floats = 64 * pl.RS;             // one hidden layer's record of one K2 record tile
    L.tmem_cols = 64 * (1 + L.n_hh) <= 128 ? 128 : 256;
    int o = 0;
    L.zimg = o; o += 3 * K2T_IMG;
    L.aimg = o; o += 3 * K2T_IMG;
    L.wimg = o; o += L.n_hh * 3 * TC_WIMG;
    L.rec = o; o += L.n_sub * L.rec_sub_floats * 4;
    L.stage = o; o += K2T_NCW * 32 * K2T_STAGE_STRIDE * 4;
    L.small = o; o += ((pl.small_floats * 4 + 127) / 128) * 128;
    L.ybar = o; o += ((PJ_MAX_NETS * C * L.T * 4 + 127) / 128) * 128;
    L.sgrad = o; o += ((pl.sgrad_floats * 4 + 127) / 128) * 128;
    L.misc = o; o += 128;
    L.bytes = o;
    L.ok = L.ok && L.bytes <= 232448;   // 227 KB of dynamic shared memory per CTA
    return L;
}

// a-jet of a hidden layer from its stored record (channel 0 = tanh(z0) for tanh nets, z0 for sin nets; others z-jets)
template <int N1, int N2, int WL>
__device__ __forceinline__ void act_from_record(int act, const float (&z)[1 + N1 + N2], float (&a)[1 + N1 + N2], const float* w) {
    float a0, s1, s2;
    if (act == PJ_ACT_TANH) {
        a0 = z[0];
        s1 = fmaf(-a0, a0, 1.0f);
        s2 = -2.0f * a0 * s1;
    } else {
        sincosf(z[0], &a0, &s1);
        s2 = -a0;
    }
    a[0] = a0;
#pragma unroll
    for (int f = 0; f < N1; ++f) a[1 + f] = s1 * z[1 + f];
    if constexpr (WL > 0) {
        float q = 0.0f;
#pragma unroll
        for (int d = 0; d < WL; ++d) q = fmaf(w[d] * z[1 + d], z[1 + d], q);
        a[1 + N1] = fmaf(s2, q, s1 * z[1 + N1]);
    } else {
#pragma unroll
        for (int s = 0; s < N2; ++s) a[1 + N1 + s] = fmaf(s2 * z[1 + s], z[1 + s], s1 * z[1 + N1 + s]);
    }
}

template <int N1, int N2, int WL>
__global__ void __launch_bounds__(K2T_NCW * 32, 1) k2tc_backward_kernel(const __grid_constant__ K2Args A) {
    constexpr int C = 1 + N1 + N2;
    static_assert(C == 2 || C == 4, "tensor-core reverse kernel: 2 or 4 jet channels");
    using M = Tc2Map<C>;
    constexpr int UG = M::UG;
    constexpr int NT = K2T_NCW * 32;
    constexpr int T = K2T_ROWS / C;
    extern __shared__ __align__(1024) unsigned char smem[];
    const PjSpec& sp = A.spec;
    const Plan& pl = A.plan;
    const K2tcLayout lay = k2tc_layout(sp, pl);
    unsigned char* zimg = smem + lay.zimg;
    unsigned char* aimg = smem + lay.aimg;
    unsigned char* wimg = smem + lay.wimg;
    float* rec = reinterpret_cast<float*>(smem + lay.rec);
    float* stage = reinterpret_cast<float*>(smem + lay.stage);
    float* small = reinterpret_cast<float*>(smem + lay.small);
    float* ybar = reinterpret_cast<float*>(smem + lay.ybar);
    float* sgrad = reinterpret_cast<float*>(smem + lay.sgrad);
    uint64_t* wfull = reinterpret_cast<uint64_t*>(smem + lay.misc);
    uint64_t* rec_full = wfull + 1;
    uint64_t* adj_done = wfull + 2;
    uint64_t* wg_done = wfull + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T2 = pl.T, RS2 = pl.RS;
    const long long ws_points = (long long)pl.n_tiles * T2;
    uint32_t n_sm;
    asm("mov.u32 %0, %%nsmid;" : "=r"(n_sm));
    const int eff_grid = min((int)gridDim.x, (int)n_sm);          // CTAs beyond one per SM only clear their partial
    const int n_tiles_tc = (int)((ws_points + T - 1) / T);
    const int my_tiles = ((int)blockIdx.x < eff_grid && n_tiles_tc > (int)blockIdx.x)
                             ? (n_tiles_tc - 1 - (int)blockIdx.x) / eff_grid + 1 : 0;
    float* gpart = A.gpart + (size_t)blockIdx.x * sp.n_theta;

    for (long long i = tid; i < sp.n_theta; i += NT) gpart[i] = 0.0f;
    if (my_tiles == 0) return;

    if (tid == 0) {
        mbar_init(wfull, 1);
        mbar_init(rec_full, 1);
        mbar_init(adj_done, 1);
        mbar_init(wg_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        if (lay.tmem_cols == 128)
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)));
        else
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < pl.small_floats; i += NT) small[i] = __ldg(A.pack + i);
    for (int i = tid; i < pl.sgrad_floats; i += NT) sgrad[i] = 0.0f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) {   // all hidden->hidden weight images, once
        const uint32_t total = (uint32_t)lay.n_hh * 3u * TC_WIMG;
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

    // ---- thread geometry (as in pinnjet_k1tc2.cuh, one 128-row accumulator) ----
    const int hf = warp >> 2, q = warp & 3;
    const int rowbase = q * 32;
    const int ppidx = lane / M::NUG, ug = lane % M::NUG;
    const int p0 = rowbase / C + 2 * ppidx;                        // first of the two adjacent tile-local points owned
    const int ubase = hf * 32 + ug * UG;                           // first of the UG adjacent hidden units owned
    const int R0 = rowbase + 2 * C * ppidx;
    const uint32_t own_row = (uint32_t)((R0 >> 3) * 1024 + (R0 & 7) * 128);
    const uint32_t awr_c = (uint32_t)(((hf * 4 + (UG == 4 ? (ug >> 1) : ug)) ^ (R0 & 7)) << 4);
    const uint32_t awr_b = UG == 4 ? (uint32_t)(ug & 1) * 8u : 0u;
    const uint32_t tmem_adj = tmem_base + (uint32_t)(hf * 32) + ((uint32_t)(q * 32) << 16);
    float* my_stage = stage + (size_t)warp * 32 * K2T_STAGE_STRIDE;
    const int sub = p0 / T2, win = p0 % T2;                        // record sub-tile / offset of this thread's point pair
    uint32_t ph_rec = 0, ph_adj = 0, ph_wg = 0;
    constexpr uint32_t IDESC_ADJ = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | (8u << 24);
    constexpr uint32_t IDESC_WG = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | (4u << 24);
    const int pa[6] = {0, 0, 1, 1, 0, 2}, pb[6] = {0, 1, 0, 1, 2, 0};

    // three bf16 terms of v[pp][c][0..UG) into the rows (point p0+pp, channel c) of a split image set
    auto store_rows = [&](unsigned char* img, const float (&v)[2][C][UG]) {
#pragma unroll
        for (int pp = 0; pp < 2; ++pp)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int j = C * pp + c;
                uint32_t t1[UG / 2], t2[UG / 2], t3[UG / 2];
#pragma unroll
                for (int e = 0; e < UG / 2; ++e) split3_bf16(v[pp][c][2 * e], v[pp][c][2 * e + 1], t1[e], t2[e], t3[e]);
                unsigned char* dst = img + own_row + j * 128 + ((awr_c ^ (uint32_t)(j << 4)) + awr_b);
                if constexpr (UG == 4) {
                    *reinterpret_cast<uint2*>(dst) = make_uint2(t1[0], t1[1]);
                    *reinterpret_cast<uint2*>(dst + K2T_IMG) = make_uint2(t2[0], t2[1]);
                    *reinterpret_cast<uint2*>(dst + 2 * K2T_IMG) = make_uint2(t3[0], t3[1]);
                } else {
                    *reinterpret_cast<uint4*>(dst) = make_uint4(t1[0], t1[1], t1[2], t1[3]);
                    *reinterpret_cast<uint4*>(dst + K2T_IMG) = make_uint4(t2[0], t2[1], t2[2], t2[3]);
                    *reinterpret_cast<uint4*>(dst + 2 * K2T_IMG) = make_uint4(t3[0], t3[1], t3[2], t3[3]);
                }
            }
    };
    // record of the hidden layer currently in `rec`: z[pp][c] of unit u at this thread's point pair
    auto load_record = [&](int u, float (&z0)[C], float (&z1)[C]) {
        const float* r = rec + (size_t)sub * lay.rec_sub_floats + (size_t)u * RS2 + win;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float2 v = *reinterpret_cast<const float2*>(r + c * T2);
            z0[c] = v.x;
            z1[c] = v.y;
        }
    };
    // sum over the point-pair lanes of the warp; afterwards the lanes with ppidx == 0 hold the warp's sum
    auto reduce_points = [&](float x) {
#pragma unroll
        for (int m = M::NUG; m < 32; m <<= 1) x += __shfl_xor_sync(0xffffffffu, x, m);
        return x;
    };
    mbar_wait(wfull, 0);

    for (int iter = 0; iter < my_tiles; ++iter) {
        const long long tile = (long long)blockIdx.x + (long long)iter * eff_grid;
        const long long base = tile * T;                           // first point of the tile (a multiple of T2)
        const long long gp0 = base + p0;
        const bool live = gp0 < ws_points;                         // records / seeds exist for this point pair
        int slot0 = 0;                                             // first weight-image / accumulator slot of the net

        for (int n = 0; n < sp.n_nets; ++n) {
            const PjNet& net = sp.net[n];
            const int L = net.n_linear - 1;
            const int act_kind = net.act;
            const int n_out = net.width[net.n_linear];
            float wq[2][WL > 0 ? WL : 1];
#pragma unroll
            for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                for (int d = 0; d < (WL > 0 ? WL : 1); ++d)
                    wq[pp][d] = (WL > 0 && live)
                                    ? __ldg(A.wts + (gp0 / T2) * ((long long)sp.n_nets * WL * T2) + (n * WL + d) * T2 + win + pp)
                                    : 0.0f;

            // record loader: hidden layer h of this net for all record sub-tiles of the tile (missing sub-tiles: zeros)
            auto request_record = [&](int h) {
                if (tid == 0) {
                    fence_proxy_async();
                    uint32_t bytes = 0;
                    for (int s = 0; s < lay.n_sub; ++s)
                        if (base + (long long)s * T2 < ws_points) bytes += (uint32_t)lay.rec_sub_floats * 4u;
                    mbar_arrive_expect_tx(rec_full, bytes);
                    for (int s = 0; s < lay.n_sub; ++s)
                        if (base + (long long)s * T2 < ws_points)
                            tma_bulk_g2s(rec + (size_t)s * lay.rec_sub_floats,
                                         A.zj + (base / T2 + s) * pl.zj_tile_floats + pl.zj_off[n][h],
                                         (uint32_t)lay.rec_sub_floats * 4u, rec_full);
                }
            };

            // (0) seeds of this net and the record of the last hidden layer
            request_record(L);
            for (int e = tid; e < n_out * C * T; e += NT) {
                const int row = e / T, pt = e - row * T;
                const long long g = base + pt;
                ybar[e] = g < ws_points ? __ldg(A.seeds + (g / T2) * ((long long)sp.n_yrows * T2) + (net.yrow0 + row) * T2 + (g % T2)) : 0.0f;
            }
            __syncthreads();
            mbar_wait(rec_full, ph_rec);
            ph_rec ^= 1u;

            // (1) last Linear: a_bar_L = W_out^T y_bar, reverse activation of hidden L, gradients of W_out / b_out / b_L
            float zb[2][C][UG];                                    // z_bar of the layer just processed (owner layout)
            {
                const float* wlo = small + pl.s_wlo[n];            // [n_out][64]
                float gb[UG], gw[PJ_MAX_NETS][UG];
#pragma unroll
                for (int k = 0; k < UG; ++k) {
                    const int u = ubase + k;
                    float z[2][C], ab[2][C], a[2][C], zbk[2][C];
                    load_record(u, z[0], z[1]);
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                        for (int c = 0; c < C; ++c) ab[pp][c] = 0.0f;
#pragma unroll
                    for (int o = 0; o < PJ_MAX_NETS; ++o)
                        if (o < n_out) {
                            const float w = wlo[o * 64 + u];
#pragma unroll
                            for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                                for (int c = 0; c < C; ++c) ab[pp][c] = fmaf(w, ybar[(o * C + c) * T + p0 + pp], ab[pp][c]);
                        }
                    float gbk = 0.0f;
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp) {
                        act_backward<N1, N2, WL>(act_kind, z[pp], ab[pp], a[pp], zbk[pp], wq[pp]);
                        if (!live) {
#pragma unroll
                            for (int c = 0; c < C; ++c) zbk[pp][c] = 0.0f;
                        }
                        gbk += zbk[pp][0];
#pragma unroll
                        for (int c = 0; c < C; ++c) zb[pp][c][k] = zbk[pp][c];
                    }
                    gb[k] = gbk;
#pragma unroll
                    for (int o = 0; o < PJ_MAX_NETS; ++o) {
                        float s = 0.0f;
                        if (o < n_out && live) {
#pragma unroll
                            for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                                for (int c = 0; c < C; ++c) s = fmaf(ybar[(o * C + c) * T + p0 + pp], a[pp][c], s);
                        }
                        gw[o][k] = s;
                    }
                }
#pragma unroll
                for (int k = 0; k < UG; ++k) {
                    const float sb = reduce_points(gb[k]);
                    if (ppidx == 0) atomicAdd(&sgrad[pl.g_b[n][L - 1] + ubase + k], sb);
#pragma unroll
                    for (int o = 0; o < PJ_MAX_NETS; ++o)
                        if (o < n_out) {
                            const float sw = reduce_points(gw[o][k]);
                            if (ppidx == 0) atomicAdd(&sgrad[pl.g_wl[n] + o * 64 + ubase + k], sw);
                        }
                }
                if (tid < n_out) {   // b_out gradient: sum over the points of the value-channel seed
                    float s = 0.0f;
                    for (int pt = 0; pt < T; ++pt) s += ybar[(tid * C) * T + pt];
                    atomicAdd(&sgrad[pl.g_bout[n] + tid], s);
                }
            }

            // (2) hidden layers h = L .. 2
            if (L >= 2) {
                store_rows(zimg, zb);                              // z_bar_L
                fence_proxy_async();
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();                                   // ZIMG complete; every reader is done with `rec`
            }
            for (int h = L; h >= 2; --h) {
                const int l = h - 1;
                const int slot = slot0 + (l - 1);
                const uint32_t w_base = smem_u32(wimg + (size_t)slot * 3 * TC_WIMG);
                if (tid == 0) {   // adjoint GEMM of Linear l: D_adj[r][k] = sum_j z_bar_h[r][j] W_l[j][k]
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t z_base = smem_u32(zimg);
#pragma unroll
                    for (int pr = 0; pr < 6; ++pr)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t da = umma_desc_sw128(z_base + pa[pr] * K2T_IMG + k * 32);        // K-major, K = unit j
                            const uint64_t db = umma_desc_sw128(w_base + pb[pr] * TC_WIMG + k * 2048);      // MN-major, K = row j
                            const uint32_t accf = (pr | k) ? 1u : 0u;
                            asm volatile(
                                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_base),
                                "l"(da), "l"(db), "r"(IDESC_ADJ), "r"(accf)
                                : "memory");
                        }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     smem_u32(adj_done))
                                 : "memory");
                }
                request_record(h - 1);
                mbar_wait(rec_full, ph_rec);
                ph_rec ^= 1u;

                // a_{h-1} from the record (independent of the adjoint) -> AIMG, then the weight-gradient MMAs
                float zr[2][C][UG];                                // record of hidden h-1 (kept for the reverse rule)
                {
                    float av[2][C][UG];
#pragma unroll
                    for (int k = 0; k < UG; ++k) {
                        float z[2][C], a[2][C];
                        load_record(ubase + k, z[0], z[1]);
#pragma unroll
                        for (int pp = 0; pp < 2; ++pp) {
                            act_from_record<N1, N2, WL>(act_kind, z[pp], a[pp], wq[pp]);
#pragma unroll
                            for (int c = 0; c < C; ++c) {
                                zr[pp][c][k] = z[pp][c];
                                av[pp][c][k] = live ? a[pp][c] : 0.0f;
                            }
                        }
                    }
                    store_rows(aimg, av);
                }
                fence_proxy_async();
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();
                if (tid == 0) {   // W_bar_l[j][k] += sum_r z_bar_h[r][j] a_{h-1}[r][k]: both operands MN-major, K = rows
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t z_base = smem_u32(zimg), a_base = smem_u32(aimg);
                    const uint32_t d_addr = tmem_base + 64u + (uint32_t)slot * 64u;
#pragma unroll
                    for (int pr = 0; pr < 6; ++pr)
#pragma unroll
                        for (int k = 0; k < K2T_ROWS / 16; ++k) {
                            const uint64_t da = umma_desc_sw128(z_base + pa[pr] * K2T_IMG + k * 2048);
                            const uint64_t db = umma_desc_sw128(a_base + pb[pr] * K2T_IMG + k * 2048);
                            const uint32_t accf = (iter | pr | k) ? 1u : 0u;     // the accumulator lives across tiles
                            asm volatile(
                                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_addr),
                                "l"(da), "l"(db), "r"(IDESC_WG), "r"(accf)
                                : "memory");
                        }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     smem_u32(wg_done))
                                 : "memory");
                }

                // adjoint of hidden h-1: TMEM row -> staging -> owner layout, reverse activation rule
                mbar_wait(adj_done, ph_adj);
                ph_adj ^= 1u;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                {
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
                        : "r"(tmem_adj));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        *reinterpret_cast<uint4*>(my_stage + lane * K2T_STAGE_STRIDE + 4 * s) =
                            make_uint4(v[4 * s], v[4 * s + 1], v[4 * s + 2], v[4 * s + 3]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                float gb[UG];
#pragma unroll
                for (int k = 0; k < UG; ++k) gb[k] = 0.0f;
#pragma unroll
                for (int pp = 0; pp < 2; ++pp) {
                    float ab[C][UG];
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float* src = my_stage + (C * (2 * ppidx + pp) + c) * K2T_STAGE_STRIDE + ug * UG;
#pragma unroll
                        for (int s4 = 0; s4 < UG / 4; ++s4) {
                            const float4 t = *reinterpret_cast<const float4*>(src + 4 * s4);
                            ab[c][4 * s4 + 0] = t.x;
                            ab[c][4 * s4 + 1] = t.y;
                            ab[c][4 * s4 + 2] = t.z;
                            ab[c][4 * s4 + 3] = t.w;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < UG; ++k) {
                        float z[C], abk[C], a[C], zbk[C];
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            z[c] = zr[pp][c][k];
                            abk[c] = ab[c][k];
                        }
                        act_backward<N1, N2, WL>(act_kind, z, abk, a, zbk, wq[pp]);
#pragma unroll
                        for (int c = 0; c < C; ++c) zb[pp][c][k] = live ? zbk[c] : 0.0f;
                        gb[k] += zb[pp][0][k];
                    }
                }
                __syncwarp();   // staging rows are reused by the next layer
#pragma unroll
                for (int k = 0; k < UG; ++k) {
                    const float sb = reduce_points(gb[k]);
                    if (ppidx == 0) atomicAdd(&sgrad[pl.g_b[n][h - 2] + ubase + k], sb);
                }
                // ZIMG / AIMG are free once the weight-gradient MMAs have read them
                mbar_wait(wg_done, ph_wg);
                ph_wg ^= 1u;
                if (h > 2) {
                    store_rows(zimg, zb);                          // z_bar_{h-1}
                    fence_proxy_async();
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();                                   // images complete, accumulators read, `rec` free
            }

            // (3) Linear 0: W_0 gradient from z_bar_1, the coordinates and the direction vectors
            {
                float x[2][PJ_MAX_COORDS];
#pragma unroll
                for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                    for (int i = 0; i < PJ_MAX_COORDS; ++i)
                        x[pp][i] = (i < net.n_in) ? __ldg(A.coords[net.in_coord[i]] + min(gp0 + pp, A.N - 1)) : 0.0f;
#pragma unroll
                for (int k = 0; k < UG; ++k) {
                    float sf[N1 > 0 ? N1 : 1];
#pragma unroll
                    for (int f = 0; f < N1; ++f) sf[f] = zb[0][1 + f][k] + zb[1][1 + f][k];
#pragma unroll
                    for (int i = 0; i < PJ_MAX_COORDS; ++i)
                        if (i < net.n_in) {
                            float s = fmaf(zb[0][0][k], x[0][i], zb[1][0][k] * x[1][i]);
#pragma unroll
                            for (int f = 0; f < N1; ++f) s = fmaf(sf[f], sp.dir[f][net.in_coord[i]], s);
                            s = reduce_points(s);
                            if (ppidx == 0) atomicAdd(&sgrad[pl.g_w0[n] + (ubase + k) * net.n_in + i], s);
                        }
                }
            }
            __syncthreads();   // ybar / rec are rewritten by the next net
            slot0 += L - 1;
        }
    }

    // ---- this CTA's partial: small gradients from shared memory, hidden->hidden weight gradients from TMEM ----
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        int slot = 0;
        for (int n = 0; n < sp.n_nets; ++n) {
            const PjNet& net = sp.net[n];
            const int L = net.n_linear - 1;
            const int h1 = net.width[1], hL = net.width[L], n_out = net.width[net.n_linear];
            for (int e = tid; e < h1 * net.n_in; e += NT) gpart[net.w_off[0] + e] += sgrad[pl.g_w0[n] + e];
            for (int hl = 0; hl < L; ++hl)
                for (int e = tid; e < net.width[hl + 1]; e += NT) gpart[net.b_off[hl] + e] += sgrad[pl.g_b[n][hl] + e];
            for (int e = tid; e < n_out * hL; e += NT) {
                const int o = e / hL, k = e - o * hL;
                gpart[net.w_off[L] + e] += sgrad[pl.g_wl[n] + o * 64 + k];
            }
            for (int e = tid; e < n_out; e += NT) gpart[net.b_off[L] + e] += sgrad[pl.g_bout[n] + e];
            for (int l = 1; l < L; ++l, ++slot) {   // M = 64 accumulator: row j in lane (j % 16) + 32 * (j / 16)
                const int width_j = net.width[l + 1], width_k = net.width[l];
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
                    : "r"(tmem_base + 64u + (uint32_t)slot * 64u + (uint32_t)(hf * 32) + ((uint32_t)(q * 32) << 16)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int j = 16 * q + lane;
                if (lane < 16 && j < width_j) {
                    float* gw = gpart + net.w_off[l] + (size_t)j * width_k;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int k = hf * 32 + i;
                        if (k < width_k) gw[k] += __uint_as_float(v[i]);
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        if (lay.tmem_cols == 128)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base));
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base));
    }
}

}  // namespace pj
