// pinnjet_common.cuh -- shared definitions of the sm_100a kernels (plan, PTX helpers, jet algebra, FFMA2 microkernels).
//
// Kernel family (DESIGN.md has the full picture):
//   K0  pack      theta (torch layout) -> K-major / out-major padded copies the tiles stream with bulk TMA
//   K1  forward   coords -> FCNN forward in Taylor (jet) mode -> re-parameterisation + residual program
//                 (+ seeds dL/d(jet) and z-jets for K2 when training)
//   K2  backward  one reverse sweep per tile: adjoint GEMMs + weight-gradient GEMMs, per-CTA partials
//   K2b reduce    partials -> grad_theta (+=)
// One persistent CTA per SM: 8 compute warps + 1 producer warp (bulk-TMA weight stream through an mbarrier ring).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "../../include/pinnjet.h"

#ifndef PJ_USE_FFMA2
#define PJ_USE_FFMA2 1
#endif

namespace pj {

// PINNJET_PDL=1 switches programmatic dependent launch between the kernels of a step on.  Off by default: measured on B200
// (profiles/r02/bench_c2_v6.json vs bench_c2_nopdl_v6.json, same for C5) it does not pay -- inside a CUDA graph the kernel-to-
// kernel gaps are already ~1 us and the early-resident dependents cost more than they hide (C2 0.1121 vs 0.1099 ms/step,
// C5 0.2266 vs 0.2204).
inline bool pdl_enabled() {
    static const int on = [] {
        const char* e = getenv("PINNJET_PDL");
        return (e && e[0] == '1') ? 1 : 0;
    }();
    return on != 0;
}

// kern<<<grid, block, smem, s>>>(args...), optionally as a programmatic dependent of the kernel launched before it on `s`
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool dependent,
                                 Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (dependent && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// CTA shape: NTC compute threads (128 or 256: template parameter of the kernels) + one producer warp.  Narrow networks
// (hidden width <= 64) use 128-thread CTAs so that several CTAs share an SM and their GEMM / activation / program phases
// overlap; the residual program is batched over NTC points (one per compute thread).
constexpr int CHUNK_FLOATS = 4096;       // weight chunk = 16 KB
constexpr int MAX_STAGES = 8;
constexpr int ROW_PAD = 4;               // jet rows are C*T + 4 floats: conflict-free row-strided float4 loads

// opcodes of the residual program (mirror of neurodiffeq_b200/symbolic.py)
enum : int {
    OP_CONST = 0, OP_COORD, OP_NET, OP_RBAR, OP_PARAM, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_SIN, OP_COS, OP_EXP,
    OP_LOG, OP_TANH, OP_SQRT, OP_ABS, OP_SIGN, OP_POWC, OP_RCP, OP_ST_U, OP_ST_R, OP_ST_SEED, OP_TAN, OP_SINH, OP_COSH,
    OP_ATAN, OP_ERF, OP_ST_W
};

// Everything derived from (spec, N): identical on host and device, computed by make_plan() in pinnjet_api.cu.
struct Plan {
    int T, P, Q, C, RS;                  // K2 tile: points, thread tile, channels, jet row stride (floats); also the
                                         // layout of the z-jet records and seeds K1 leaves in the workspace
    int epi_batch;                       // points per residual-program batch (jet table double-buffered in smem)
    int T1, P1, Q1, RS1, ntc1, n_tiles1; // K1 tile (a multiple of T): wider thread tile (Q1 = 8) -> fewer smem wavefronts
    int n_tiles, grid, grid_bwd, hmax, ntc;   // grid: K1 CTAs (= loss partials), grid_bwd: K2 CTAs (= gradient partials)
    int n_stage, n_stage_bwd, resident_fwd, resident_bwd, chunks_fwd, chunks_bwd;   // n_stage: forward ring
    int hp[PJ_MAX_NETS][PJ_MAX_LINEAR + 1];   // padded widths (hidden -> multiple of 32; input/output unpadded)
    // ---- packed parameter copy (float offsets) ----
    int small_floats;
    int s_wt0[PJ_MAX_NETS];              // [n_in][hp1]       first Linear, K-major
    int s_dz[PJ_MAX_NETS];               // [PJ_MAX_DIRS][hp1] first-order seeds  W0 . dir_f  (point independent)
    int s_b[PJ_MAX_NETS][PJ_MAX_LINEAR]; // hidden biases, padded
    int s_wlt[PJ_MAX_NETS];              // [hpL][n_out]      last Linear, K-major        (forward)
    int s_wlo[PJ_MAX_NETS];              // [n_out][hpL]      last Linear, out-major      (backward)
    int s_bout[PJ_MAX_NETS];
    long long b_wt[PJ_MAX_NETS][PJ_MAX_LINEAR];   // hidden->hidden Linear l: [in_p][out_p]  (forward B operand)
    long long b_wo[PJ_MAX_NETS][PJ_MAX_LINEAR];   //                          [out_p][in_p]  (adjoint B operand)
    long long b_wimg[PJ_MAX_NETS][PJ_MAX_LINEAR];   // tensor-core path: 3 bf16 split images of W_l, K-major SWIZZLE_128B (float offset)
    long long b_woutimg[PJ_MAX_NETS];    // tensor-core path: 3 bf16 split images [16 x 64] of the output Linear (rows >= n_out zero)
    int tc;                              // 1: K1 runs the hidden-layer and output GEMMs on tcgen05 (pinnjet_k1tc3.cuh)
    int tc_bwd;                          // 1: K2 too (pinnjet_k2tc2.cuh); it reads K1-TC's records in place
    int tp;                              // tensor-core tile: points per 128 GEMM rows (pinnjet_tc.cuh: TcGeo::TP)
    int seed_T;                          // tile size of the seed / combined-weight layouts K1 writes ( = tp when tc_bwd, else T)
    long long tc_rec_layer_floats, tc_rec_tile_floats;   // tensor-core record layout [tile][hidden layer][thread][C*UG]
    long long ws_tcrec;                  // workspace offset (bytes) of those records
    long long pack_floats;
    // ---- small-gradient accumulators in shared memory (float offsets) ----
    int g_w0[PJ_MAX_NETS], g_b[PJ_MAX_NETS][PJ_MAX_LINEAR], g_wl[PJ_MAX_NETS], g_bout[PJ_MAX_NETS], sgrad_floats;
    int sgrad_copies;                    // one private copy per point-group block of warps (no atomics)
    // ---- workspace (byte offsets) ----
    int zj_off[PJ_MAX_NETS][PJ_MAX_LINEAR];       // float offset of hidden layer h (1..L) z-jets inside a tile block
    long long zj_tile_floats;
    long long ws_zj, ws_seed, ws_gpart, ws_loss, ws_bytes;
    // ---- shared memory (byte offsets) ----
    int k1_act, k1_ring, k1_small, k1_ycache, k1_slots, k1_prog, k1_misc, k1_bytes, k1_stage;
    int k1_wbuf, k1_wslots, k1_progw;    // combined second-order channel: per-point weights, their interpreter state
    long long ws_wts;                    // workspace: weights [tile][n_nets*wl][T] for K2
    int k2_g0, k2_g1, k2_zb, k2_ring, k2_small, k2_ybar, k2_sgrad, k2_misc, k2_bytes;
};

struct K1Args {
    PjSpec spec;
    Plan plan;
    const float* coords[PJ_MAX_COORDS];
    const float* pack;
    const int4* prog;
    int prog_len;
    const int4* prog_w;
    int prog_w_len;
    int mode;                            // 0 = eval (u, residual), 1 = train (residual, seeds, z-jets)
    long long N;
    float loss_scale;
    const float* rbar;
    float* u_out;
    float* r_out;
    float* zj;                           // z-jet records (tensor-core layout when plan.tc)
    float* zj_ffma;                      // plan.tc && !plan.tc_bwd: where the re-laid-out copy for the FFMA reverse kernel goes
    float* seeds;
    float* wts;
    float* loss_part;
    float* dbg;                          // diagnostic builds only (PJ_TIMING): phase cycle counters
    float* sumsq_out;                    // non-null: the LAST warp to deliver its partial folds them all into *sumsq_out (+=)
    unsigned* ticket;                    // ... found by this counter (zero between launches; lives in the loss-partial block)
};

struct K2Args {
    PjSpec spec;
    Plan plan;
    const float* coords[PJ_MAX_COORDS];
    const float* pack;
    long long N;
    const float* zj;
    const float* seeds;
    const float* wts;
    float* gpart;
    float* dbg;
};

// ---------------------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier, bulk TMA (cp.async.bulk -> SASS UBLKCP), named barriers, packed FP32 FMA (FFMA2)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Programmatic dependent launch (K0 -> K1 -> finalize -> K2 -> K2b are launched with programmatic stream serialization, see
// launch_kernel below): a kernel lets its successor's CTAs become resident right away (they take the SMs as this grid's
// CTAs retire), and the successor blocks in pdl_wait() -- until this grid has COMPLETED and its memory is visible -- before it
// touches anything this grid writes.  Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Loss finalisation inside the forward kernel (was a launch of its own): a program warp has stored its partial sum of
// squared residuals; the warp that draws the last ticket adds ALL partials in the fixed order of the former
// loss_finalize_kernel (lane-strided, then an xor-shuffle tree: run-to-run reproducible, identical numbers) to *sumsq_out and
// re-arms the counter.  Called by whole warps; `lane0_stored` = lane 0 has written part[my index].
__device__ __forceinline__ void fold_loss_partials(const float* part, unsigned n_parts, float* sumsq_out, unsigned* ticket, int lane) {
    if (sumsq_out == nullptr) return;
    unsigned last = 0;
    if (lane == 0) {
        __threadfence();                                   // my partial is visible before my ticket
        last = atomicAdd(ticket, 1u) == n_parts - 1u ? 1u : 0u;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    float s = 0.0f;
    for (unsigned p = lane; p < n_parts; p += 32) s += __ldcg(part + p);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        *sumsq_out += s;
        *ticket = 0u;
    }
}

// K2b arithmetic, shared by k2_reduce_kernel and the fused reduce + all-reduce kernel (pinnjet_comm.cu) so that both give the
// same bits: parameter i, partial group g of RED_GROUPS (contiguous ranges of the per-CTA partials, two accumulators each),
// combined as ((g0+g1)+(g2+g3)) + ((g4+g5)+(g6+g7)).  A block handles RED_PARAMS parameters with one warp per group.
constexpr int RED_PARAMS = 32, RED_GROUPS = 8;
__device__ __forceinline__ float red_group_sum(const float* __restrict__ gpart, int n_parts, long long n_theta, long long i, int g) {
    const int per = (n_parts + RED_GROUPS - 1) / RED_GROUPS, p_lo = g * per, p_hi = min(n_parts, p_lo + per);
    float s0 = 0.0f, s1 = 0.0f;
    int p = p_lo;
    for (; p + 1 < p_hi; p += 2) {
        s0 += gpart[(size_t)p * n_theta + i];
        s1 += gpart[(size_t)(p + 1) * n_theta + i];
    }
    if (p < p_hi) s0 += gpart[(size_t)p * n_theta + i];
    return s0 + s1;
}
__device__ __forceinline__ float red_combine(const float (*red)[RED_PARAMS], int il) {
    return ((red[0][il] + red[1][il]) + (red[2][il] + red[3][il])) + ((red[4][il] + red[5][il]) + (red[6][il] + red[7][il]));
}

__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk TMA global -> shared, completion signalled on an mbarrier (bytes multiple of 16, 16B-aligned both sides)
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// barrier among the compute threads only (the producer warp never joins)
template <int NTC>
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(NTC) : "memory"); }

// packed pair of fp32 in one 64-bit register pair: the operand type of fma.rn.f32x2 (SASS FFMA2)
typedef unsigned long long f2;
__device__ __forceinline__ f2 pack2(float x, float y) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ float2 unpack2(f2 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ void ffma2(f2& d, const f2 a, const f2 b) {
#if PJ_USE_FFMA2
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
#else
    const float2 x = unpack2(a), y = unpack2(b), z = unpack2(d);
    d = pack2(fmaf(x.x, y.x, z.x), fmaf(x.y, y.y, z.y));
#endif
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Reduce-scatter over the 8 point-group lanes (recursive halving): every thread contributes 8 (or 4) partial sums, lane l
// of each 8-lane group returns the total of value l (value l >> 1 for the 4-value form, in both lanes of a pair):
// 7 (4) shuffles instead of the 24 (12) of one butterfly all-reduce per value.
__device__ __forceinline__ float pg_reduce_scatter8(const float (&v)[8], int pg_lane) {
    const bool b2 = pg_lane & 4, b1 = pg_lane & 2, b0 = pg_lane & 1;
    float w[4], x[2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        w[i] = (b2 ? v[i + 4] : v[i]) + __shfl_xor_sync(0xffffffffu, b2 ? v[i] : v[i + 4], 4);
#pragma unroll
    for (int i = 0; i < 2; ++i)
        x[i] = (b1 ? w[i + 2] : w[i]) + __shfl_xor_sync(0xffffffffu, b1 ? w[i] : w[i + 2], 2);
    return (b0 ? x[1] : x[0]) + __shfl_xor_sync(0xffffffffu, b0 ? x[0] : x[1], 1);
}
__device__ __forceinline__ float pg_reduce_scatter4(const float (&v)[4], int pg_lane) {
    const bool b2 = pg_lane & 4, b1 = pg_lane & 2;
    float w[2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
        w[i] = (b2 ? v[i + 2] : v[i]) + __shfl_xor_sync(0xffffffffu, b2 ? v[i] : v[i + 2], 4);
    float x = (b1 ? w[1] : w[0]) + __shfl_xor_sync(0xffffffffu, b1 ? w[0] : w[1], 2);
    return x + __shfl_xor_sync(0xffffffffu, x, 1);   // value index = pg_lane >> 1
}

// ---------------------------------------------------------------------------------------------------------------------
// activation jets.  Channels: 0 value | 1..N1 first order | N1+1..N1+N2 pure second order.
// ---------------------------------------------------------------------------------------------------------------------
// Branch-free tanh, ~1-4 ulp: |x| < 0.6: x + x^3 P(x^2) (degree-4 minimax fit, 1.4 ulp); else 1 - 2/(exp(2|x|)+1) with
// ex2.approx / rcp.approx.  Straight-line code: the 8-16 independent calls of an activation epilogue pipeline instead of
// diverging like libdevice's tanhf.
__device__ __forceinline__ float tanh_fast(float x) {
    const float ax = fabsf(x), t = x * x;
    float p = fmaf(-0.006276387721300125f, t, 0.021116213873028755f);
    p = fmaf(p, t, -0.053875137120485306f);
    p = fmaf(p, t, 0.13332924246788025f);
    p = fmaf(p, t, -0.3333333134651184f);
    const float small = fmaf(x * t, p, x);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(ax, 15.0f) * 2.885390081777927f));
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    const float big = copysignf(fmaf(-2.0f, r, 1.0f), x);
    return ax < 0.6f ? small : big;
}

__device__ __forceinline__ void act_d2(int act, float z0, float& a0, float& s1, float& s2) {
    if (act == PJ_ACT_TANH) {
        a0 = tanh_fast(z0);
        s1 = fmaf(-a0, a0, 1.0f);
        s2 = -2.0f * a0 * s1;
    } else {
        sincosf(z0, &a0, &s1);
        s2 = -a0;
    }
}

// z-jet -> a-jet, in place.  WL > 0: the single second-order channel is the weighted combination L = sum_d w[d] D_d^2.
template <int N1, int N2, int WL>
__device__ __forceinline__ void act_forward(int act, float (&z)[1 + N1 + N2], const float* w) {
    float a0, s1, s2;
    act_d2(act, z[0], a0, s1, s2);
    if constexpr (WL > 0) {
        static_assert(N2 == 1, "combined mode carries one second-order channel");
        float q = 0.0f;
#pragma unroll
        for (int d = 0; d < WL; ++d) q = fmaf(w[d] * z[1 + d], z[1 + d], q);
        z[1 + N1] = fmaf(s2, q, s1 * z[1 + N1]);
    } else {
#pragma unroll
        for (int s = 0; s < N2; ++s) z[1 + N1 + s] = fmaf(s2 * z[1 + s], z[1 + s], s1 * z[1 + N1 + s]);
    }
#pragma unroll
    for (int f = 0; f < N1; ++f) z[1 + f] *= s1;
    z[0] = a0;
}

// reverse of the activation jet: given the stored record (channel 0 = tanh(z0) for tanh nets, z0 for sin nets; other
// channels z-jets) and the adjoint of the a-jet, produce the a-jet (for the weight-gradient GEMM) and the adjoint of the
// z-jet.
template <int N1, int N2, int WL>
__device__ __forceinline__ void act_backward(int act, const float (&z)[1 + N1 + N2], const float (&ab)[1 + N1 + N2],
                                             float (&a)[1 + N1 + N2], float (&zb)[1 + N1 + N2], const float* w) {
    float a0, s1, s2, s3;
    if (act == PJ_ACT_TANH) {   // record channel 0 = tanh(z0), stored by K1: no transcendental in the reverse pass
        a0 = z[0];
        s1 = fmaf(-a0, a0, 1.0f);
        s2 = -2.0f * a0 * s1;
        s3 = -2.0f * s1 * s1 - 2.0f * a0 * s2;
    } else {
        sincosf(z[0], &a0, &s1);
        s2 = -a0;
        s3 = -s1;
    }
    float zb0 = s1 * ab[0];
#pragma unroll
    for (int f = 0; f < N1; ++f) {
        zb[1 + f] = s1 * ab[1 + f];
        zb0 = fmaf(s2 * z[1 + f], ab[1 + f], zb0);
        a[1 + f] = s1 * z[1 + f];
    }
    if constexpr (WL > 0) {
        const float abL = ab[1 + N1], zL = z[1 + N1];
        float q = 0.0f;
#pragma unroll
        for (int d = 0; d < WL; ++d) {
            const float wz = w[d] * z[1 + d];
            q = fmaf(wz, z[1 + d], q);
            zb[1 + d] = fmaf(2.0f * s2 * wz, abL, zb[1 + d]);
        }
        zb[1 + N1] = s1 * abL;
        zb0 = fmaf(fmaf(s3, q, s2 * zL), abL, zb0);
        a[1 + N1] = fmaf(s2, q, s1 * zL);
    } else {
#pragma unroll
        for (int s = 0; s < N2; ++s) {
            const float zf = z[1 + s], zs = z[1 + N1 + s], abs_ = ab[1 + N1 + s];
            zb[1 + N1 + s] = s1 * abs_;
            zb[1 + s] = fmaf(2.0f * s2 * zf, abs_, zb[1 + s]);
            zb0 = fmaf(fmaf(s3 * zf, zf, s2 * zs), abs_, zb0);
            a[1 + N1 + s] = fmaf(s2 * zf, zf, s1 * zs);
        }
    }
    zb[0] = zb0;
    a[0] = a0;
}

// ---------------------------------------------------------------------------------------------------------------------
// register-tile GEMM: acc[q][c][p] += sum_k A[k][c][p0+p] * B[k][u0+q]
//   A: jet buffer rows (stride RS floats), channel c at +c*T, points contiguous       (shared memory)
//   B: weight chunk rows (stride ldb floats), output units contiguous                  (shared memory)
// Thread tile P points x Q units x C channels; point pairs are packed for FFMA2.
// ---------------------------------------------------------------------------------------------------------------------
template <int P, int Q, int C>
__device__ __forceinline__ void gemm_rows(f2 (&acc)[Q][C][P / 2], const float* __restrict__ a_ptr, int RS, int T,
                                          const float* __restrict__ b_ptr, int ldb, int nrows) {
#pragma unroll 2
    for (int k = 0; k < nrows; ++k) {
        f2 a[C][P / 2];
        const float* ar = a_ptr + k * RS;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if constexpr (P == 4) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(ar + c * T);
                a[c][0] = v.x;
                a[c][1] = v.y;
            } else {
                a[c][0] = *reinterpret_cast<const f2*>(ar + c * T);
            }
        }
        float b[Q];
        const float* br = b_ptr + k * ldb;
#pragma unroll
        for (int q4 = 0; q4 < Q / 4; ++q4) {
            const float4 v = *reinterpret_cast<const float4*>(br + 4 * q4);
            b[4 * q4 + 0] = v.x;
            b[4 * q4 + 1] = v.y;
            b[4 * q4 + 2] = v.z;
            b[4 * q4 + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const f2 bb = pack2(b[q], b[q]);
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int h = 0; h < P / 2; ++h) ffma2(acc[q][c][h], a[c][h], bb);
        }
    }
}

// scalar view of a packed accumulator tile (p is a compile-time constant after unrolling)
template <int P>
__device__ __forceinline__ float pick(const f2 (&v)[P / 2], int p) {
    const float2 t = unpack2(v[p >> 1]);
    return (p & 1) ? t.y : t.x;
}

// ---- optional phase timing (diagnostic build only: -DPJ_TIMING=1 -> libpinnjet_timing.so; never in the product) --------
#ifdef PJ_TIMING
#define PJ_T_DECL unsigned long long pj_t_last = clock64(), pj_t_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define PJ_T_MARK(slot)                              \
    {                                                \
        const unsigned long long now_ = clock64();   \
        pj_t_acc[slot] += now_ - pj_t_last;          \
        pj_t_last = now_;                            \
    }
#define PJ_T_FLUSH(base)                                                           \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                                     \
        for (int i_ = 0; i_ < 12; ++i_) A.dbg[(base) + i_] = (float)pj_t_acc[i_];     \
    }
#else
#define PJ_T_DECL
#define PJ_T_MARK(slot)
#define PJ_T_FLUSH(base)
#endif

// thread -> (point group, unit group) mapping shared by K1 and K2: a warp covers 8 point groups x 4 unit groups
struct JobMap {
    int p0, u0, pg_lane;
    __device__ __forceinline__ JobMap(int tid, int T, int P, int Q) {
        const int warp = tid >> 5, lane = tid & 31;
        const int n_pgb = (T / P) >> 3;
        pg_lane = lane & 7;
        p0 = P * ((warp % n_pgb) * 8 + pg_lane);
        u0 = Q * ((warp / n_pgb) * 4 + (lane >> 3));
    }
};

}  // namespace pj
