// pinnjet_k2.cuh -- K2: the single reverse pass  dL/dtheta  (replaces loss.backward(), reference solvers.py:393, i.e.
// the double/triple-backward sweep through TanhBackward/MmBackward nodes that dominates the reference's step).
//
// Inputs per tile: seeds dL/d(raw output jets) and the z-jets of every hidden layer, both written by K1.
// Per hidden layer h (from the last to the first) a CTA
//   1. pulls the adjoint of the layer's a-jets through W^T   -- register-tiled FFMA2 GEMM, out-major weights streamed by
//      the producer warp with bulk TMA (same ring as K1);
//   2. applies the reverse of the activation-jet rule (needs tanh''' / sin''') on the accumulator registers, re-creating
//      the a-jets of the layer below from its z-jets (bulk-TMA'd from the workspace) on the way;
//   3. accumulates the weight gradient  W_bar += z_bar (x) a_prev  over channels and points -- second FFMA2 GEMM whose
//      4x4 output tile per thread is added into a per-CTA partial buffer (thread-owned, no atomics);
// bias / first-layer / last-layer gradients are reduced over the point lanes with warp shuffles into shared memory.
// K2b sums the per-CTA partials into grad_theta (+=, like autograd accumulation, solvers.py:360-362).
#pragma once
#include "pinnjet_common.cuh"
#include "pinnjet_k1.cuh"   // weight_producer, RingCursor

namespace pj {

constexpr int WJ = 8, WK = 4;   // weight-gradient output tile per thread (8 rows of z_bar x 4 rows of a-jets)

// out[j][k] += sum_r G[j][r] * Z[k][r],  r over the C*T (channel, point) pairs; rows interleaved over lanes so that the
// float4 loads of 8 consecutive rows (stride RS = C*T+4 floats) hit 32 distinct banks.
__device__ __forceinline__ void wgrad_tile(f2 (&acc)[WJ][WK], const float* __restrict__ g_base, int j_step,
                                           const float* __restrict__ z_base, int k_step, int RS, int R) {
#pragma unroll 2
    for (int r = 0; r < R; r += 4) {
        ulonglong2 gv[WJ], zv[WK];
#pragma unroll
        for (int i = 0; i < WJ; ++i) gv[i] = *reinterpret_cast<const ulonglong2*>(g_base + (size_t)i * j_step * RS + r);
#pragma unroll
        for (int i = 0; i < WK; ++i) zv[i] = *reinterpret_cast<const ulonglong2*>(z_base + (size_t)i * k_step * RS + r);
#pragma unroll
        for (int i = 0; i < WJ; ++i)
#pragma unroll
            for (int j = 0; j < WK; ++j) {
                ffma2(acc[i][j], gv[i].x, zv[j].x);
                ffma2(acc[i][j], gv[i].y, zv[j].y);
            }
    }
}

template <int NTC, int MINB, int P, int Q, int N1, int N2, int WL>
__global__ void __launch_bounds__(NTC + 32, MINB) k2_backward_kernel(const __grid_constant__ K2Args A) {
    constexpr int C = 1 + N1 + N2;
    constexpr int NT_COMPUTE = NTC, NT_TOTAL = NTC + 32, N_CWARPS = NTC / 32;
    extern __shared__ __align__(128) unsigned char smem[];
    const PjSpec& sp = A.spec;
    const Plan& pl = A.plan;
    float* G = reinterpret_cast<float*>(smem + pl.k2_g0);    // adjoint of the current layer's z-jets
    float* G2 = reinterpret_cast<float*>(smem + pl.k2_g1);   // ... of the layer below (being produced)
    float* Zb = reinterpret_cast<float*>(smem + pl.k2_zb);   // z-jets -> a-jets of the layer below
    float* ring = reinterpret_cast<float*>(smem + pl.k2_ring);
    float* small = reinterpret_cast<float*>(smem + pl.k2_small);
    float* ybar = reinterpret_cast<float*>(smem + pl.k2_ybar);
    float* sgrad = reinterpret_cast<float*>(smem + pl.k2_sgrad);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + pl.k2_misc);
    uint64_t* empty = full + MAX_STAGES;
    uint64_t* zfull = empty + MAX_STAGES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = pl.T, RS = pl.RS;
    const int my_tiles = (pl.n_tiles > (int)blockIdx.x) ? (pl.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    float* gpart = A.gpart + (size_t)blockIdx.x * sp.n_theta;

    if (tid == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], N_CWARPS);
        }
        mbar_init(zfull, 1);
        fence_barrier_init();
    }
    pdl_launch_dependents();
    pdl_wait();                       // the forward kernel (records, seeds) and, before it, K0 have completed
    for (int i = tid; i < pl.small_floats; i += NT_TOTAL) small[i] = __ldg(A.pack + i);
    for (int i = tid; i < pl.sgrad_floats * pl.sgrad_copies; i += NT_TOTAL) sgrad[i] = 0.0f;
    for (long long i = tid; i < sp.n_theta; i += NT_TOTAL) gpart[i] = 0.0f;
    __syncthreads();

    if (warp == N_CWARPS) {
        if (lane == 0) weight_producer<false>(sp, pl, A.pack, ring, full, empty, my_tiles);
        return;
    }

    const JobMap jm(tid, T, P, Q);
    const int p0 = jm.p0, u0 = jm.u0;
    // every (point-group block, unit) pair is owned by exactly one lane -> private accumulation, no atomics
    float* sg = sgrad + (size_t)(warp % ((T / P) >> 3)) * pl.sgrad_floats;
    RingCursor cur{0, pl.n_stage_bwd, pl.resident_bwd != 0, full, empty, ring};
    uint32_t zphase = 0;
    // lane mapping of the weight-gradient GEMM: 8 k-lanes x 4 j-lanes per warp
    const int kl = lane & 7, jl = lane >> 3;
    PJ_T_DECL   // slots: 0 setup, 1 seeds+z wait, 2 last-linear stage, 3 adjoint gemm, 4 z wait, 5 reverse act, 6 wgrad, 7 layer0
    PJ_T_MARK(0)

    for (int iter = 0; iter < my_tiles; ++iter) {
        const long long tile = (long long)blockIdx.x + (long long)iter * gridDim.x;
        const long long base = tile * T;
        if (cur.resident) cur.it = 0;
        const float* zj_tile = A.zj + tile * pl.zj_tile_floats;
        const float* seed_tile = A.seeds + tile * ((long long)sp.n_yrows * T);

        for (int n = 0; n < sp.n_nets; ++n) {
            const PjNet& net = sp.net[n];
            const int L = net.n_linear - 1;
            const int act_kind = net.act;
            const int n_out = net.width[net.n_linear];
            const int hpL = pl.hp[n][L];
            float wq[P][WL > 0 ? WL : 1];   // weights of the combined second-order channel at this thread's points
#pragma unroll
            for (int p = 0; p < P; ++p)
#pragma unroll
                for (int dd = 0; dd < (WL > 0 ? WL : 1); ++dd)
                    wq[p][dd] = WL > 0 ? __ldg(A.wts + tile * ((long long)sp.n_nets * WL * T) + (n * WL + dd) * T + p0 + p) : 0.0f;

            // (0) seeds of this net + bulk load of the last hidden layer's z-jets
            if (tid == 0) {
                fence_proxy_async();
                const uint32_t bytes = (uint32_t)(hpL * RS) * 4u;
                mbar_arrive_expect_tx(zfull, bytes);
                tma_bulk_g2s(Zb, zj_tile + pl.zj_off[n][L], bytes, zfull);
            }
            for (int e = tid; e < n_out * C * T; e += NT_COMPUTE) ybar[e] = __ldg(seed_tile + net.yrow0 * T + e);
            bar_compute<NTC>();
            mbar_wait(zfull, zphase);
            zphase ^= 1u;
            PJ_T_MARK(1)

            // (1) last Linear: grads of W_out, b_out; adjoint of hidden-L a-jets; reverse activation -> G = z_bar_L
            {
                const float* wlo = small + pl.s_wlo[n];
                if (u0 < hpL) {
                    static_assert(Q == 4, "the gradient reductions below assume 4 units per thread");
                    float gbq[Q], gwq[PJ_MAX_NETS][Q];   // per-thread partials: bias of hidden L, W_out (n_out <= 4)
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        const int u = u0 + q;
                        float gw[PJ_MAX_NETS];
#pragma unroll
                        for (int o = 0; o < PJ_MAX_NETS; ++o) gw[o] = 0.0f;
                        float gb = 0.0f;
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            const int pt = p0 + p;
                            float z[C], ab[C], a[C], zb[C];
#pragma unroll
                            for (int c = 0; c < C; ++c) {
                                z[c] = Zb[u * RS + c * T + pt];
                                ab[c] = 0.0f;
                            }
#pragma unroll
                            for (int o = 0; o < PJ_MAX_NETS; ++o)
                                if (o < n_out) {
                                    const float w = wlo[o * hpL + u];
#pragma unroll
                                    for (int c = 0; c < C; ++c) ab[c] = fmaf(w, ybar[(o * C + c) * T + pt], ab[c]);
                                }
                            act_backward<N1, N2, WL>(act_kind, z, ab, a, zb, wq[p]);
#pragma unroll
                            for (int o = 0; o < PJ_MAX_NETS; ++o)
                                if (o < n_out) {
#pragma unroll
                                    for (int c = 0; c < C; ++c) gw[o] = fmaf(ybar[(o * C + c) * T + pt], a[c], gw[o]);
                                }
                            gb += zb[0];
#pragma unroll
                            for (int c = 0; c < C; ++c) G[u * RS + c * T + pt] = zb[c];
                        }
                        gbq[q] = gb;
#pragma unroll
                        for (int o = 0; o < PJ_MAX_NETS; ++o) gwq[o][q] = gw[o];
                    }
                    {   // reduce over the point lanes: [bias(4) | W_out row 0 (4)], then W_out rows 1.. two at a time
                        const int pl8 = jm.pg_lane, i4 = pl8 & 3;
                        const float v0[8] = {gbq[0], gbq[1], gbq[2], gbq[3], gwq[0][0], gwq[0][1], gwq[0][2], gwq[0][3]};
                        const float t0 = pg_reduce_scatter8(v0, pl8);
                        if (pl8 < 4) sg[pl.g_b[n][L - 1] + u0 + i4] += t0; else sg[pl.g_wl[n] + u0 + i4] += t0;
                        if (n_out > 1) {
                            const float v1[8] = {gwq[1][0], gwq[1][1], gwq[1][2], gwq[1][3],
                                                 gwq[2][0], gwq[2][1], gwq[2][2], gwq[2][3]};
                            const float t1 = pg_reduce_scatter8(v1, pl8);
                            if (pl8 < 4) sg[pl.g_wl[n] + hpL + u0 + i4] += t1;
                            else if (n_out > 2) sg[pl.g_wl[n] + 2 * hpL + u0 + i4] += t1;
                        }
                        if (n_out > 3) {
                            const float v2[4] = {gwq[3][0], gwq[3][1], gwq[3][2], gwq[3][3]};
                            const float t2 = pg_reduce_scatter4(v2, pl8);
                            if (!(pl8 & 1)) sg[pl.g_wl[n] + 3 * hpL + u0 + (pl8 >> 1)] += t2;
                        }
                    }
                }
                if (tid < n_out) {   // b_out gradient: sum over points of the value-channel seed
                    float s = 0.0f;
                    for (int pt = 0; pt < T; ++pt) s += ybar[(tid * C) * T + pt];
                    sgrad[pl.g_bout[n] + tid] += s;
                }
            }
            bar_compute<NTC>();
            PJ_T_MARK(2)

            // (2) hidden layers h = L .. 2: Linear l = h-1 maps hidden h-1 -> hidden h
            for (int h = L; h >= 2; --h) {
                const int l = h - 1;
                const int HJ = pl.hp[n][h], HK = pl.hp[n][h - 1];
                if (tid == 0) {   // z-jets of hidden h-1 (Zb is free: every reader passed the barrier above)
                    fence_proxy_async();
                    const uint32_t bytes = (uint32_t)(HK * RS) * 4u;
                    mbar_arrive_expect_tx(zfull, bytes);
                    tma_bulk_g2s(Zb, zj_tile + pl.zj_off[n][h - 1], bytes, zfull);
                }
                // (2a) a_bar_{h-1} = W_l^T z_bar_h
                const bool valid = u0 < HK;
                f2 acc[Q][C][P / 2];
#pragma unroll
                for (int q = 0; q < Q; ++q)
#pragma unroll
                    for (int c = 0; c < C; ++c)
#pragma unroll
                        for (int hh = 0; hh < P / 2; ++hh) acc[q][c][hh] = 0ull;
                const int rpc = CHUNK_FLOATS / HK;
                for (int r0 = 0; r0 < HJ; r0 += rpc) {
                    const float* chunk = cur.acquire();
                    if (valid) gemm_rows<P, Q, C>(acc, G + r0 * RS + p0, RS, T, chunk + u0, HK, min(rpc, HJ - r0));
                    cur.release(lane);
                }
                PJ_T_MARK(3)
                mbar_wait(zfull, zphase);
                zphase ^= 1u;
                PJ_T_MARK(4)
                // (2b) reverse activation of hidden h-1: Zb z-jets -> a-jets (in place), G2 <- z_bar_{h-1}
                if (valid) {
                    float gbq[Q];
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        const int u = u0 + q;
                        float gb = 0.0f;
                        float av[P][C], zv[P][C];
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            float z[C], ab[C];
#pragma unroll
                            for (int c = 0; c < C; ++c) {
                                z[c] = Zb[u * RS + c * T + p0 + p];
                                ab[c] = pick<P>(acc[q][c], p);
                            }
                            act_backward<N1, N2, WL>(act_kind, z, ab, av[p], zv[p], wq[p]);
                            gb += zv[p][0];
                        }
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            if constexpr (P == 4) {
                                *reinterpret_cast<float4*>(Zb + u * RS + c * T + p0) =
                                    make_float4(av[0][c], av[1][c], av[2][c], av[3][c]);
                                *reinterpret_cast<float4*>(G2 + u * RS + c * T + p0) =
                                    make_float4(zv[0][c], zv[1][c], zv[2][c], zv[3][c]);
                            } else {
                                *reinterpret_cast<float2*>(Zb + u * RS + c * T + p0) = make_float2(av[0][c], av[1][c]);
                                *reinterpret_cast<float2*>(G2 + u * RS + c * T + p0) = make_float2(zv[0][c], zv[1][c]);
                            }
                        }
                        gbq[q] = gb;
                    }
                    const float tb = pg_reduce_scatter4(gbq, jm.pg_lane);
                    if (!(jm.pg_lane & 1)) sg[pl.g_b[n][h - 2] + u0 + (jm.pg_lane >> 1)] += tb;
                }
                bar_compute<NTC>();
                PJ_T_MARK(5)
                // (2c) W_l gradient: out[j][k] += sum_{c,pt} G[j][c,pt] * Zb[k][c,pt]
                {
                    const int width_j = net.width[h], width_k = net.width[h - 1];   // unpadded
                    const int n_kb = HK / 32, n_jb = HJ / 32;   // warp tile = 32 rows j x 32 rows k
                    float* gw = gpart + net.w_off[l];
                    for (int wt = warp; wt < n_kb * n_jb; wt += N_CWARPS) {
                        const int jb = (wt / n_kb) * 32, kb = (wt % n_kb) * 32;
                        f2 wacc[WJ][WK];
#pragma unroll
                        for (int i = 0; i < WJ; ++i)
#pragma unroll
                            for (int jj = 0; jj < WK; ++jj) wacc[i][jj] = 0ull;
                        wgrad_tile(wacc, G + (size_t)(jb + jl) * RS, 4, Zb + (size_t)(kb + kl) * RS, 8, RS, C * T);
                        // every output element is owned by one thread of this CTA, so the fire-and-forget reduction
                        // (RED.ADD, no return value to wait for) into the CTA's private partial is race-free and ordered
#pragma unroll
                        for (int i = 0; i < WJ; ++i) {
                            const int j = jb + jl + 4 * i;
#pragma unroll
                            for (int jj = 0; jj < WK; ++jj) {
                                const int k = kb + kl + 8 * jj;
                                if (j < width_j && k < width_k) {
                                    const float2 v = unpack2(wacc[i][jj]);
                                    atomicAdd(&gw[(size_t)j * width_k + k], v.x + v.y);
                                }
                            }
                        }
                    }
                }
                bar_compute<NTC>();
                PJ_T_MARK(6)
                float* t = G;
                G = G2;
                G2 = t;
            }

            // (3) Linear 0: W_0 gradient from z_bar_1 (in G), the coordinates and the direction vectors
            {
                const int hp1 = pl.hp[n][1];
                if (u0 < hp1) {
                    float x[PJ_MAX_COORDS][P];
#pragma unroll
                    for (int i = 0; i < PJ_MAX_COORDS; ++i)
                        if (i < net.n_in) {
#pragma unroll
                            for (int p = 0; p < P; ++p) {
                                const long long g = min(base + p0 + p, A.N - 1);
                                x[i][p] = __ldg(A.coords[net.in_coord[i]] + g);
                            }
                        }
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        const int u = u0 + q;
                        float s0[P];
                        float sf[N1 > 0 ? N1 : 1];
#pragma unroll
                        for (int f = 0; f < N1; ++f) sf[f] = 0.0f;
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            s0[p] = G[u * RS + p0 + p];
#pragma unroll
                            for (int f = 0; f < N1; ++f) sf[f] += G[u * RS + (1 + f) * T + p0 + p];
                        }
                        float sv[PJ_MAX_COORDS];
#pragma unroll
                        for (int i = 0; i < PJ_MAX_COORDS; ++i) {
                            float s = 0.0f;
                            if (i < net.n_in) {
#pragma unroll
                                for (int p = 0; p < P; ++p) s = fmaf(s0[p], x[i][p], s);
#pragma unroll
                                for (int f = 0; f < N1; ++f) s = fmaf(sf[f], sp.dir[f][net.in_coord[i]], s);
                            }
                            sv[i] = s;
                        }
                        if (net.n_in <= 4) {
                            const float v4[4] = {sv[0], sv[1], sv[2], sv[3]};
                            const float t = pg_reduce_scatter4(v4, jm.pg_lane);
                            const int i = jm.pg_lane >> 1;
                            if (!(jm.pg_lane & 1) && i < net.n_in) sg[pl.g_w0[n] + u * net.n_in + i] += t;
                        } else {
                            const float t = pg_reduce_scatter8(sv, jm.pg_lane);
                            if (jm.pg_lane < net.n_in) sg[pl.g_w0[n] + u * net.n_in + jm.pg_lane] += t;
                        }
                    }
                }
            }
            bar_compute<NTC>();
            PJ_T_MARK(7)
        }
    }
    PJ_T_FLUSH(16)

    // flush the shared-memory gradient accumulators into this CTA's partial (padded units are dropped)
    bar_compute<NTC>();
    for (int n = 0; n < sp.n_nets; ++n) {
        const PjNet& net = sp.net[n];
        const int L = net.n_linear - 1;
        const int h1 = net.width[1], hL = net.width[L], hpL = pl.hp[n][L], n_out = net.width[net.n_linear];
        auto sgsum = [&](int idx) {
            float v = 0.0f;
            for (int c = 0; c < pl.sgrad_copies; ++c) v += sgrad[(size_t)c * pl.sgrad_floats + idx];
            return v;
        };
        for (int e = tid; e < h1 * net.n_in; e += NT_COMPUTE) gpart[net.w_off[0] + e] += sgsum(pl.g_w0[n] + e);
        for (int hl = 0; hl < L; ++hl)
            for (int e = tid; e < net.width[hl + 1]; e += NT_COMPUTE) gpart[net.b_off[hl] + e] += sgsum(pl.g_b[n][hl] + e);
        for (int e = tid; e < n_out * hL; e += NT_COMPUTE) {
            const int o = e / hL, k = e - o * hL;
            gpart[net.w_off[L] + e] += sgsum(pl.g_wl[n] + o * hpL + k);
        }
        for (int e = tid; e < n_out; e += NT_COMPUTE) gpart[net.b_off[L] + e] += sgsum(pl.g_bout[n] + e);
    }
}

}  // namespace pj
