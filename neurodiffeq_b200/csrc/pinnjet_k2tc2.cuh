// pinnjet_k2tc2.cuh -- K2-TC: the reverse pass (dL/dtheta) with both GEMMs of every hidden->hidden Linear on tcgen05.
//
// Same contract as k2_backward_kernel (pinnjet_k2.cuh) for the problems K1-TC handles (hidden width 64, 1..8 channels):
// reads the seeds / z-jet records / combined-channel weights K1-TC left in the workspace (tensor-core layouts, see
// pinnjet_tc.cuh), writes this CTA's gradient partial (K2b sums them in fixed order).
//
// Per 128-row tile and hidden layer h = L .. 2 (operand encodings validated by experiments/tcgen05_probe):
//   * z_bar_h and a_{h-1} live as bf16x3 split images [128 rows x 64 units], K-major SWIZZLE_128B (ZIMG, AIMG);
//   * adjoint GEMM   a_bar_{h-1}[r][k] = sum_j z_bar_h[r][j] W_l[j][k]:  A = ZIMG (K-major), B = the FORWARD weight images
//     of W_l read MN-major (no transposed copy), D = [128 x 64] fp32 in TMEM, 6 split products x 4 K-steps;
//   * weight-gradient GEMM  W_bar_l[j][k] += sum_r z_bar_h[r][j] a_{h-1}[r][k]:  A = ZIMG and B = AIMG both read MN-major
//     (contraction over the rows: +2048 B per K = 16), M = 64 accumulator [64 x 64] per Linear that stays in TMEM for ALL
//     tiles of the CTA and is read once at the end (row j in lane j%16 + 32*(j/16)).
// Warp roles: 16 compute warps in the owner layout of pinnjet_tc.cuh, one warp that issues all MMAs, one warp that
// streams the z-jet record blocks (bulk TMA, one hidden layer of one tile = 512 x C*UG floats) into shared memory ahead
// of their use.  The phases of a layer overlap through mbarriers instead of CTA barriers:
//     ADJ(h) runs  while  the compute warps turn the record of layer h-1 into AIMG;
//     WG(h)  runs  while  they apply the reverse activation rule to the adjoint (TMEM -> owner layout).
// Last Linear, first Linear and all bias gradients stay on the CUDA cores: every thread sums over its own point, the PW
// point lanes of a warp are combined by a reduce-scatter (pinnjet_tc.cuh: tc_reduce_points) and added WITHOUT atomics to
// the copy of the small-gradient block that belongs to the warp's TMEM quarter (a unit block has one owner warp per
// quarter); the four copies are summed when the partial is written.
#pragma once
#include "pinnjet_tc.cuh"

namespace pj {

constexpr int K2T_THREADS = TC_NT + 64;   // 576: compute warps, MMA warp, record warp
#ifndef PJ_WG_FIRST
#define PJ_WG_FIRST 0                      // first split product of the weight-gradient GEMM (0: all six, 3: the three largest)
#endif

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
__global__ void __launch_bounds__(K2T_THREADS, 1) k2tc2_backward_kernel(const __grid_constant__ K2Args A) {
    constexpr int C = 1 + N1 + N2;
    using G = TcGeo<C>;
    constexpr int UG = G::UG, TP = G::TP;
    constexpr int WLN = WL > 0 ? WL : 1;
    constexpr uint32_t REC_BYTES = TC_NT * G::REC * 4;            // one hidden layer's record block of a tile
    extern __shared__ __align__(1024) unsigned char smem[];
    const PjSpec& sp = A.spec;
    const Plan& pl = A.plan;
    unsigned char* zimg = smem + pl.k2_g0;                        // 3 x 16 KB: z_bar of the current layer
    unsigned char* aimg = smem + pl.k2_g1;                        // 3 x 16 KB: a-jets of the layer below
    unsigned char* wimg = smem + pl.k2_ring;                      // forward weight images, [hidden->hidden Linear][3] x 8 KB
    float* stage = reinterpret_cast<float*>(smem + pl.k2_zb);
    float* wlo_s = reinterpret_cast<float*>(smem + pl.k2_small);  // [net][output][64] last Linear, out-major; then the tile info:
    float* tinfo = wlo_s + sp.n_nets * PJ_MAX_NETS * TC_H;        // [2][n_yrows seeds | n_nets*WL weights | n_coords coordinates][TP]
    float* recbuf = reinterpret_cast<float*>(smem + pl.k2_ybar);  // record block of the current step
    float* sgrad = reinterpret_cast<float*>(smem + pl.k2_sgrad);  // [4 quarters][sgrad_floats]
    uint64_t* wfull = reinterpret_cast<uint64_t*>(smem + pl.k2_misc);
    uint64_t* z_ready = wfull + 1;       // ZIMG of a layer complete          (16 warp arrivals) -> ADJ
    uint64_t* a_ready = wfull + 2;       // AIMG of the layer below complete  (16 warp arrivals) -> WG
    uint64_t* adj_done = wfull + 3;      // adjoint accumulator complete
    uint64_t* wg_done = wfull + 4;       // weight-gradient MMAs have read ZIMG / AIMG
    uint64_t* rec_full = wfull + 5;      // record block landed
    uint64_t* rec_empty = wfull + 6;     // every compute warp has copied its part (16 warp arrivals)
    uint64_t* ti_full = wfull + 7;       // [2] seeds / weights / coordinates of a tile staged
    uint64_t* ti_empty = wfull + 9;      // [2] the tile is finished (16 warp arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 11);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = pl.n_tiles1;
    const int my_tiles = (n_tiles > (int)blockIdx.x) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    float* gpart = A.gpart + (size_t)blockIdx.x * sp.n_theta;
    int n_hh = 0;
    for (int n = 0; n < sp.n_nets; ++n) n_hh += sp.net[n].n_linear - 2;
    const int tmem_cols = 64 * (1 + n_hh) <= 128 ? 128 : (64 * (1 + n_hh) <= 256 ? 256 : 512);

    pdl_launch_dependents();
    if (tid == 0) {
#ifdef PJ_TIMING
        *reinterpret_cast<unsigned long long*>(tmem_slot + 2) = clock64();
#endif
        mbar_init(wfull, 1);
        mbar_init(z_ready, TC_NCW);
        mbar_init(a_ready, TC_NCW);
        mbar_init(adj_done, 1);
        mbar_init(wg_done, 1);
        mbar_init(rec_full, 1);
        mbar_init(rec_empty, TC_NCW);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&ti_full[b], 1);
            mbar_init(&ti_empty[b], TC_NCW);
        }
        fence_barrier_init();
    }
    if (warp == 0) {
        if (tmem_cols == 128)
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)));
        else if (tmem_cols == 256)
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)));
        else
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < 4 * pl.sgrad_floats; i += K2T_THREADS) sgrad[i] = 0.0f;
    if constexpr (G::CP != C)   // rows of padded channels are never written: they must read as zero in both GEMMs
        for (int i = tid; i < 2 * 3 * TC_AIMG / 16; i += K2T_THREADS) reinterpret_cast<uint4*>(zimg)[i] = make_uint4(0, 0, 0, 0);
    // everything above is CTA-local and runs while the forward kernel drains; from here on global memory is touched: the
    // forward kernel (records, seeds), K0 before it and -- with batches back to back -- the previous K2b (reads the gradient
    // partials zeroed below) must have completed
    pdl_wait();
    for (int i = tid; i < sp.n_nets * PJ_MAX_NETS * TC_H; i += K2T_THREADS) {
        const int n = i / (PJ_MAX_NETS * TC_H), r = i - n * (PJ_MAX_NETS * TC_H);
        wlo_s[i] = r < sp.net[n].width[sp.net[n].n_linear] * TC_H ? __ldg(A.pack + pl.s_wlo[n] + r) : 0.0f;
    }
    for (long long i = tid; i < sp.n_theta; i += K2T_THREADS) gpart[i] = 0.0f;   // parameters no network of the spec owns
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
#ifdef PJ_TIMING
    const unsigned long long t0_ = *reinterpret_cast<volatile unsigned long long*>(tmem_slot + 2);
#endif

    if (warp == TC_NCW) {   // ================= MMA warp (+ the weight images) =================
        TC_TRACE(tr, A.dbg, 250, 120, t0_, blockIdx.x == 0 && lane == 0)
        if (lane == 0) {
            if (n_hh > 0) {
                mbar_arrive_expect_tx(wfull, (uint32_t)n_hh * 3u * TC_WIMG);
                int slot = 0;
                for (int n = 0; n < sp.n_nets; ++n)
                    for (int l = 1; l < sp.net[n].n_linear - 1; ++l, ++slot)
                        tma_bulk_g2s(wimg + (size_t)slot * 3 * TC_WIMG, A.pack + pl.b_wimg[n][l], 3 * TC_WIMG, wfull);
            } else {
                mbar_arrive(wfull);
            }
        }
        mbar_wait(wfull, 0);
        constexpr uint32_t IDESC_ADJ = tc_idesc(128, TC_H, false, true), IDESC_WG = tc_idesc(64, TC_H, true, true);
        const uint64_t dz = umma_desc_sw128(smem_u32(zimg)), da = umma_desc_sw128(smem_u32(aimg));
        uint32_t phz = 0, pha = 0;
#pragma unroll 1
        for (int iter = 0; iter < my_tiles; ++iter) {
            int slot0 = 0;
#pragma unroll 1
            for (int n = 0; n < sp.n_nets; ++n) {
                const int L = sp.net[n].n_linear - 1;
#pragma unroll 1
                for (int h = L; h >= 2; --h) {
                    const int slot = slot0 + (h - 2);             // Linear l = h-1 is the (l-1)-th hidden->hidden Linear
                    mbar_wait(z_ready, phz);
                    phz ^= 1u;
                    tc_fence_after();
                    TC_MARK(tr, 1 | (h << 5))
                    if (lane == 0) {   // D_adj[r][k] = sum_j z_bar_h[r][j] W_l[j][k]
                        tc_mma_split6<TC_H / 16, TC_AIMG, 32, TC_WIMG, 2048>(
                            tmem_base, dz, umma_desc_sw128(smem_u32(wimg + (size_t)slot * 3 * TC_WIMG)), IDESC_ADJ, false);
                        tc_commit(adj_done);
                    }
                    __syncwarp();
                    TC_MARK(tr, 2 | (h << 5))
                    mbar_wait(a_ready, pha);
                    pha ^= 1u;
                    tc_fence_after();
                    TC_MARK(tr, 3 | (h << 5))
                    if (lane == 0) {   // W_bar_l[j][k] += sum_r z_bar_h[r][j] a_{h-1}[r][k]; the accumulator lives across tiles
                        tc_mma_split6<TC_ROWS / 16, TC_AIMG, 2048, TC_AIMG, 2048, PJ_WG_FIRST>(tmem_base + 64u + (uint32_t)slot * 64u, dz, da,
                                                                                               IDESC_WG, iter > 0);
                        tc_commit(wg_done);
                    }
                    __syncwarp();
                    TC_MARK(tr, 4 | (h << 5))
                }
                slot0 += L - 1;
            }
        }
        return;   // the compute warps read the weight-gradient accumulators after their last wg_done wait
    }

    if (warp == TC_NCW + 1) {   // ================= record warp: one block per (tile, net, hidden layer), in the order of use ====
        uint32_t ph = 0;
        bool first = true;
        const int NW = sp.n_nets * WL, ti_rows = sp.n_yrows + NW + sp.n_coords;
        auto stage_tile_info = [&](int it) {     // seeds, combined-channel weights and coordinates of tile `it` -> tinfo[it & 1]
            if (it >= 2) mbar_wait(&ti_empty[it & 1], (uint32_t)(((it >> 1) - 1) & 1));
            const long long t = (long long)blockIdx.x + (long long)it * gridDim.x;
            float* dst = tinfo + (size_t)(it & 1) * ti_rows * TP;
            for (int e = lane; e < ti_rows * TP; e += 32) {
                const int row = e / TP, pt = e - row * TP;
                float v;
                if (row < sp.n_yrows) v = __ldg(A.seeds + t * ((long long)sp.n_yrows * TP) + e);
                else if (row < sp.n_yrows + NW) v = __ldg(A.wts + t * ((long long)NW * TP) + (e - sp.n_yrows * TP));
                else v = __ldg(A.coords[row - sp.n_yrows - NW] + min(t * TP + pt, A.N - 1));
                dst[e] = v;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ti_full[it & 1]);
        };
        if (my_tiles > 0) stage_tile_info(0);
#pragma unroll 1
        for (int iter = 0; iter < my_tiles; ++iter) {
            const long long tile = (long long)blockIdx.x + (long long)iter * gridDim.x;
            int lidx0 = 0;
#pragma unroll 1
            for (int n = 0; n < sp.n_nets; ++n) {
                const int L = sp.net[n].n_linear - 1;
#pragma unroll 1
                for (int h = L; h >= 1; --h) {
                    if (!first) {
                        mbar_wait(rec_empty, ph);
                        ph ^= 1u;
                    }
                    first = false;
                    if (lane == 0) {
                        mbar_arrive_expect_tx(rec_full, REC_BYTES);
                        tma_bulk_g2s(recbuf, A.zj + tile * pl.tc_rec_tile_floats + (long long)(lidx0 + h - 1) * pl.tc_rec_layer_floats,
                                     REC_BYTES, rec_full);
                    }
                    __syncwarp();
                    // one tile ahead, in the idle time after the tile's last block has been requested
                    if (n == sp.n_nets - 1 && h == 1 && iter + 1 < my_tiles) stage_tile_info(iter + 1);
                }
                lidx0 += L;
            }
        }
        return;
    }

    // ================================================ compute warps ====================================================
    const TcThread<C> th(tid);
    float* sg = sgrad + (size_t)th.q * pl.sgrad_floats;           // this quarter's copy of the small-gradient block
    const bool adder = (th.pt & 1) == 0;                          // after tc_reduce_points: the lane that adds value pt >> 1
    const int uadd = th.ubase + (th.pt >> 1);                     // ... which belongs to this unit
    uint32_t ph_adj = 0, ph_wg = 0, ph_rec = 0;
    bool wg_pending = false;             // a WG commit this thread has not waited for yet (ZIMG / AIMG still being read)
    TC_TRACE(tr, A.dbg, 0, 250, t0_, blockIdx.x == 0 && tid == 0)
    float zr[C][UG];
#ifdef PJ_DBG_REC_GLOBAL
    long long dbg_rec_off = 0;
#endif
    auto next_record = [&]() {           // this thread's C*UG floats of the next record block
        mbar_wait(rec_full, ph_rec);
        ph_rec ^= 1u;
#ifdef PJ_DBG_REC_GLOBAL
        tc_load_record<C>(A.zj + dbg_rec_off + (size_t)tid * G::REC, zr);
#else
        tc_load_record_smem<C>(recbuf + (size_t)tid * G::REC, zr);
#endif
        // The block may be overwritten (bulk TMA = async proxy) as soon as all 16 warps have arrived: the reads above must
        // have been PERFORMED, not just issued -- the checksum makes the arrival depend on the loaded data -- and ordered
        // against the async proxy.
        float chk = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int k = 0; k < UG; ++k) chk += zr[c][k];
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && __float_as_uint(chk) != 0xFFFFFFFFu) mbar_arrive(rec_empty);
    };

#pragma unroll 1
    for (int iter = 0; iter < my_tiles; ++iter) {
        const int NW = sp.n_nets * WL;
        const float* ti = tinfo + (size_t)(iter & 1) * (sp.n_yrows + NW + sp.n_coords) * TP + th.p;   // this point's column
        mbar_wait(&ti_full[iter & 1], (uint32_t)((iter >> 1) & 1));

#pragma unroll 1
        for (int n = 0; n < sp.n_nets; ++n) {
            const PjNet& net = sp.net[n];
            const int L = net.n_linear - 1;
            const int act_kind = net.act;
            const int n_out = net.width[net.n_linear];
            TC_MARK(tr, 1)
            float wq[WLN];
#pragma unroll
            for (int d = 0; d < WLN; ++d) wq[d] = WL > 0 ? ti[(sp.n_yrows + n * WL + d) * TP] : 0.0f;
            // seeds of this thread's point (the 16 threads of a point read the same words)
            float yb_[PJ_MAX_NETS][C];
            {
                const float* sd = ti + net.yrow0 * TP;
#pragma unroll
                for (int o = 0; o < PJ_MAX_NETS; ++o)
#pragma unroll
                    for (int c = 0; c < C; ++c) yb_[o][c] = o < n_out ? sd[(o * C + c) * TP] : 0.0f;
            }
#ifdef PJ_DBG_REC_GLOBAL
            int dbg_l0 = 0;
            for (int m = 0; m < n; ++m) dbg_l0 += sp.net[m].n_linear - 1;
            dbg_rec_off = tile * pl.tc_rec_tile_floats + (long long)(dbg_l0 + L - 1) * pl.tc_rec_layer_floats;
#endif
            next_record();               // record of the last hidden layer
            TC_MARK(tr, 2)

            // (1) last Linear: a_bar_L = W_out^T y_bar, reverse activation of hidden L, gradients of W_out / b_out / b_L
            float zb[C][UG];                                       // z_bar of the layer just processed (owner layout)
            {
                const float* wlo = wlo_s + n * PJ_MAX_NETS * TC_H;  // [n_out][64]
                float gwl[PJ_MAX_NETS][UG], gbv[UG];
#pragma unroll
                for (int k = 0; k < UG; ++k) {
                    const int u = th.ubase + k;
                    float z[C], ab[C], a[C], zbk[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        z[c] = zr[c][k];
                        ab[c] = 0.0f;
                    }
#pragma unroll
                    for (int o = 0; o < PJ_MAX_NETS; ++o)
                        if (o < n_out) {
                            const float w = wlo[o * TC_H + u];
#pragma unroll
                            for (int c = 0; c < C; ++c) ab[c] = fmaf(w, yb_[o][c], ab[c]);
                        }
                    act_backward<N1, N2, WL>(act_kind, z, ab, a, zbk, wq);
#pragma unroll
                    for (int c = 0; c < C; ++c) zb[c][k] = zbk[c];
                    gbv[k] = zbk[0];
#pragma unroll
                    for (int o = 0; o < PJ_MAX_NETS; ++o) {
                        float s = 0.0f;
#pragma unroll
                        for (int c = 0; c < C; ++c) s = fmaf(yb_[o][c], a[c], s);
                        gwl[o][k] = s;
                    }
                }
                {
                    const float s = tc_reduce_points<C>(gbv, th.pt);
                    if (adder) sg[pl.g_b[n][L - 1] + uadd] += s;
                }
#pragma unroll
                for (int o = 0; o < PJ_MAX_NETS; ++o)
                    if (o < n_out) {
                        const float s = tc_reduce_points<C>(gwl[o], th.pt);
                        if (adder) sg[pl.g_wl[n] + o * TC_H + uadd] += s;
                    }
                if (th.j == 0) {   // b_out gradient: sum over the points of the value-channel seed (one lane per point)
#pragma unroll
                    for (int o = 0; o < PJ_MAX_NETS; ++o)
                        if (o < n_out) {
                            const float s = warp_sum(th.ug == 0 ? yb_[o][0] : 0.0f);
                            if (lane == 0) sg[pl.g_bout[n] + o] += s;
                        }
                }
            }
            TC_MARK(tr, 3)

            // (2) hidden layers h = L .. 2
            if (L >= 2) {
                if (wg_pending) {        // ZIMG / AIMG are free once the previous weight-gradient MMAs have read them
                    mbar_wait(wg_done, ph_wg);
                    ph_wg ^= 1u;
                    wg_pending = false;
                }
                tc_store_rows<C>(zimg, TC_AIMG, th, zb);           // z_bar_L
                tc_publish(z_ready, th.lane);                      // -> ADJ(L)
            }
            TC_MARK(tr, 4)
#pragma unroll 1
            for (int h = L; h >= 2; --h) {
                // a_{h-1} from the record (independent of the adjoint) -> AIMG, then the weight-gradient MMAs
#ifdef PJ_DBG_REC_GLOBAL
                dbg_rec_off -= pl.tc_rec_layer_floats;
#endif
                next_record();
                {
                    float av[C][UG];
#pragma unroll
                    for (int k = 0; k < UG; ++k) {
                        float z[C], a[C];
#pragma unroll
                        for (int c = 0; c < C; ++c) z[c] = zr[c][k];
                        act_from_record<N1, N2, WL>(act_kind, z, a, wq);
#pragma unroll
                        for (int c = 0; c < C; ++c) av[c][k] = a[c];
                    }
                    tc_store_rows<C>(aimg, TC_AIMG, th, av);
                }
                tc_publish(a_ready, th.lane);                      // -> WG(h) (after ADJ(h) in the tensor pipe)
                TC_MARK(tr, 5 | (h << 5))

                // adjoint of hidden h-1: TMEM -> owner layout, reverse activation rule
                mbar_wait(adj_done, ph_adj);
                ph_adj ^= 1u;
                tc_fence_after();
                TC_MARK(tr, 6 | (h << 5))
                float ab[C][UG];
                tc_load_owner<C>(tmem_base, stage, th, ab);
                TC_MARK(tr, 7 | (h << 5))
                float gbv[UG];
#pragma unroll
                for (int k = 0; k < UG; ++k) {
                    float z[C], abk[C], a[C], zbk[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        z[c] = zr[c][k];
                        abk[c] = ab[c][k];
                    }
                    act_backward<N1, N2, WL>(act_kind, z, abk, a, zbk, wq);
#pragma unroll
                    for (int c = 0; c < C; ++c) zb[c][k] = zbk[c];
                    gbv[k] = zbk[0];
                }
                {
                    const float s = tc_reduce_points<C>(gbv, th.pt);
                    if (adder) sg[pl.g_b[n][h - 2] + uadd] += s;
                }
                TC_MARK(tr, 8 | (h << 5))
                wg_pending = true;
                if (h > 2) {             // ZIMG is rewritten: the weight-gradient MMAs of this layer must have read it
                    mbar_wait(wg_done, ph_wg);
                    ph_wg ^= 1u;
                    wg_pending = false;
                    TC_MARK(tr, 9 | (h << 5))
                    tc_store_rows<C>(zimg, TC_AIMG, th, zb);       // z_bar_{h-1}
                    tc_publish(z_ready, th.lane);
                    TC_MARK(tr, 10 | (h << 5))
                }
            }

            // (3) Linear 0: W_0 gradient from z_bar_1, the coordinates and the direction vectors
            {
#pragma unroll
                for (int i = 0; i < PJ_MAX_COORDS; ++i)
                    if (i < net.n_in) {
                        const int ci = net.in_coord[i];
                        const float x = ti[(sp.n_yrows + NW + ci) * TP];
                        float v[UG];
#pragma unroll
                        for (int k = 0; k < UG; ++k) {
                            float s = zb[0][k] * x;
#pragma unroll
                            for (int f = 0; f < N1; ++f) s = fmaf(zb[1 + f][k], sp.dir[f][ci], s);
                            v[k] = s;
                        }
                        const float s = tc_reduce_points<C>(v, th.pt);
                        if (adder) sg[pl.g_w0[n] + uadd * net.n_in + i] += s;
                    }
            }
            TC_MARK(tr, 11)
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ti_empty[iter & 1]);           // the tile-info buffer may be refilled
    }
    TC_MARK(tr, 12)
    if (wg_pending) {
        mbar_wait(wg_done, ph_wg);
        ph_wg ^= 1u;
    }
    TC_MARK(tr, 14)

    // ---- this CTA's partial: small gradients (sum of the four quarter copies) from shared memory, hidden->hidden weight
    // gradients from TMEM.  Instances of one module (boundary instances, pinnjet.h) share w_off / b_off: a later instance
    // adds to what the first one stored. ----
    tc_fence_after();
    {
        const int SG = pl.sgrad_floats;
        int slot = 0;
#pragma unroll 1
        for (int n = 0; n < sp.n_nets; ++n) {
            bar_named(9, TC_NT);                                   // shared-memory sums complete / previous net's stores done
            TC_MARK(tr, 15)
            const PjNet& net = sp.net[n];
            bool shared_w = false;
            for (int m = 0; m < n; ++m) shared_w = shared_w || sp.net[m].w_off[0] == net.w_off[0];
            const int L = net.n_linear - 1;
            const int h1 = net.width[1], hL = net.width[L], n_out = net.width[net.n_linear];
            auto put = [&](long long off, int s_idx) {
                const float v = (sgrad[s_idx] + sgrad[SG + s_idx]) + (sgrad[2 * SG + s_idx] + sgrad[3 * SG + s_idx]);
                if (shared_w) gpart[off] += v; else gpart[off] = v;
            };
            // (cold code, executed once per CTA: kept small -- it runs at instruction-fetch speed)
#pragma unroll 1
            for (int e = tid; e < h1 * net.n_in; e += TC_NT) put(net.w_off[0] + e, pl.g_w0[n] + e);
#pragma unroll 1
            for (int hl = 0; hl < L; ++hl)
#pragma unroll 1
                for (int e = tid; e < net.width[hl + 1]; e += TC_NT) put(net.b_off[hl] + e, pl.g_b[n][hl] + e);
#pragma unroll 1
            for (int e = tid; e < n_out * hL; e += TC_NT) {
                const int o = e / hL, k = e - o * hL;
                put(net.w_off[L] + e, pl.g_wl[n] + o * TC_H + k);
            }
#pragma unroll 1
            for (int e = tid; e < n_out; e += TC_NT) put(net.b_off[L] + e, pl.g_bout[n] + e);
            TC_MARK(tr, 16)
#pragma unroll 1
            for (int l = 1; l < L; ++l, ++slot) {   // M = 64 accumulator: row j in lane (j % 16) + 32 * (j / 16)
                const int width_j = net.width[l + 1], width_k = net.width[l];
                uint32_t v[16];
                const uint32_t addr = tmem_base + 64u + (uint32_t)slot * 64u + (uint32_t)(th.j * 16) + ((uint32_t)(th.q * 32) << 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(addr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int jr = 16 * th.q + lane, k0 = th.j * 16;
                if (lane < 16 && jr < width_j) {
                    float* gw = gpart + net.w_off[l] + (size_t)jr * width_k + k0;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (k0 + i < width_k) {
                            if (shared_w) gw[i] += __uint_as_float(v[i]); else gw[i] = __uint_as_float(v[i]);
                        }
                }
            }
        }
    }
    TC_MARK(tr, 13)
    tc_fence_before();
    bar_named(9, TC_NT);
    if (warp == 0) {
        if (tmem_cols == 128)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base));
        else if (tmem_cols == 256)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base));
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

}  // namespace pj
