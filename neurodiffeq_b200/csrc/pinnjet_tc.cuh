// pinnjet_tc.cuh -- shared pieces of the tcgen05 kernels (K1-TC forward, K2-TC reverse): tile geometry, bf16x3 split
// images, the TMEM -> owner-layout transposition, MMA issue helpers.
//
// GEMM formulation (hidden width 64, C jet channels padded to CP in {2, 4, 8}):
//   * a tile is 128 GEMM rows r = CP*p + c  (point p < TP = 128/CP, channel c < C; rows with c >= C stay zero);
//   * operands are split into THREE bf16 terms (x = x1 + x2 + x3) stored as K-major SWIZZLE_128B shared-memory images
//     [rows x 64 units]; the six products whose weight is >= 2^-24 reproduce the fp32 contraction to ~5e-7;
//   * accumulators are fp32 in TMEM; row r of an M = 128 accumulator is TMEM lane r.
// Thread geometry of BOTH kernels (16 compute warps = 512 threads per tile):
//   warp w:  q = w & 3 (TMEM lane quarter = rows 32q..32q+31), j = w >> 2 (units 16j..16j+15);
//   lane l:  pt = l / NUG, ug = l % NUG   (NUG = 16 / UG unit groups per 16-unit block);
//   the thread OWNS point p = q*PW + pt and the UG adjacent units ubase = 16*j + UG*ug .. of it, all channels.
// A warp reads its 32 x 16 accumulator block from TMEM (one tcgen05.ld.32x32b.x16, lane = row), parks it in its private
// staging block and reads it back in owner layout (only __syncwarp in between): tanh once per (point, unit), no
// shuffles in the jet rules.  Because both kernels use the same map, the z-jet records K1 leaves for K2 are simply
// indexed by thread:  record(tile, layer)[tid][c][k], C*UG contiguous floats per thread (64 B for C = 4): coalesced
// 16-byte accesses.
#pragma once
#include "pinnjet_common.cuh"

namespace pj {

constexpr int TC_ROWS = 128;             // GEMM rows per tile
constexpr int TC_H = 64;                 // hidden width
constexpr int TC_AIMG = TC_ROWS * 128;   // bytes of one split image of a tile (128 rows x 64 bf16)
constexpr int TC_WIMG = TC_H * 128;      // bytes of one split image of a hidden->hidden weight matrix
constexpr int TC_WOUT = 16 * 128;        // bytes of one split image of an output layer (16 rows: outputs, zero padded)
constexpr int TC_NCW = 16;               // compute warps
constexpr int TC_NT = TC_NCW * 32;       // compute threads
constexpr int TC_STAGE_STRIDE = 20;      // floats per staged TMEM row (16 + 4: conflict-free 16-byte accesses)
constexpr int TC_STAGE_BYTES = TC_NCW * 32 * TC_STAGE_STRIDE * 4;   // one private 32 x 16 block per compute warp

template <int C>
struct TcGeo {
    static constexpr int CP = C <= 2 ? 2 : (C <= 4 ? 4 : 8);   // channels padded to a divisor of 32
    static constexpr int TP = TC_ROWS / CP;                    // points per tile
    static constexpr int PW = 32 / CP;                         // points per 32-row warp block
    static constexpr int NUG = 32 / PW;                        // unit groups per 16-unit block: 2 / 4 / 8
    static constexpr int UG = 16 / NUG;                        // adjacent units owned by a thread: 8 / 4 / 2
    static constexpr int REC = C * UG;                         // record floats per thread and hidden layer
};

__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    // 128-byte swizzle, 8-row groups 1024 B apart (SBO), descriptor version 1 (validated by experiments/tcgen05_probe)
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

// x = t1 + t2 + t3 (three bf16 terms) for a pair of values; the lower half of each word is the first value
__device__ __forceinline__ void split3_bf16(float x0, float x1, uint32_t& t1, uint32_t& t2, uint32_t& t3) {
    t1 = pack_bf16x2(x0, x1);
    const float r0 = x0 - bf16_lo_f32(t1), r1 = x1 - bf16_hi_f32(t1);
    t2 = pack_bf16x2(r0, r1);
    t3 = pack_bf16x2(r0 - bf16_lo_f32(t2), r1 - bf16_hi_f32(t2));
}

__device__ __forceinline__ void bar_named(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// instruction descriptor of tcgen05.mma.kind::f16: fp32 accumulator, bf16 A and B; optional MN-major operands
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the six split products over KSTEPS K = 16 steps, SMALLEST FIRST (a1 b3, a3 b1, a2 b2, a1 b2, a2 b1, a1 b1): the tensor core
// truncates when it adds into the fp32 accumulator, an error proportional to the accumulator's magnitude at that moment --
// only the last KSTEPS MMAs run at full magnitude (measured on C4: rms error 5.5e-7 -> see DESIGN.md).  `da0` / `db0` are the
// descriptors of term 0, K-step 0; the other 23 differ only in the start-address field (bytes >> 4), so every MMA costs
// two 64-bit adds with immediates.
// FIRST = 3 issues only the three largest products (a1 b2, a2 b1, a1 b1): ~2^-17 relative instead of ~2^-24.
template <int KSTEPS, uint32_t A_IMG, uint32_t A_KSTEP, uint32_t B_IMG, uint32_t B_KSTEP, int FIRST = 0>
__device__ __forceinline__ void tc_mma_split6(uint32_t d_tmem, uint64_t da0, uint64_t db0, uint32_t idesc, bool accumulate_first) {
#pragma unroll
    for (int pr = FIRST; pr < 6; ++pr)
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k) {
            constexpr int TA[6] = {0, 2, 1, 0, 1, 0}, TB[6] = {2, 0, 1, 1, 0, 0};
            const uint64_t da = da0 + (uint64_t)((TA[pr] * A_IMG + k * A_KSTEP) >> 4);
            const uint64_t db = db0 + (uint64_t)((TB[pr] * B_IMG + k * B_KSTEP) >> 4);
            tc_mma(d_tmem, da, db, idesc, (accumulate_first || pr > FIRST || k) ? 1u : 0u);
        }
}

// ---- optional event trace (diagnostic build only: -DPJ_TIMING=1 -> libpinnjet_timing.so; never in the product): lane 0 of
// selected warps of CTA 0 logs (tag << 24 | cycles since kernel start) words into the diagnostics area of the workspace ----
#ifdef PJ_TIMING
struct TcTrace {
    uint32_t* buf;
    int n, cap;
    unsigned long long t0;
    bool on;
    __device__ __forceinline__ void mark(int tag) {
        if (on && n < cap) buf[n++] = ((uint32_t)tag << 24) | (uint32_t)((clock64() - t0) & 0xFFFFFFull);
    }
};
#define TC_TRACE(name, dbg, base, cap_, t0_, on_) TcTrace name{reinterpret_cast<uint32_t*>(dbg) + (base), 0, cap_, t0_, on_};
#define TC_MARK(name, tag) name.mark(tag);
#else
#define TC_TRACE(name, dbg, base, cap_, t0_, on_)
#define TC_MARK(name, tag)
#endif

// per-thread constants of the owner layout
template <int C>
struct TcThread {
    using G = TcGeo<C>;
    int warp, lane, q, j, pt, ug;
    int p;              // tile-local point owned
    int ubase;          // first owned unit
    int R0;             // first GEMM row of the point: rows R0 + c
    uint32_t row_off;   // byte offset of row R0 inside an image (rows R0 + c are 128 B apart: R0 is a multiple of CP)
    uint32_t chunk_x;   // 16-byte chunk of the owned units, XORed per row with ((R0 + c) & 7)
    uint32_t chunk_b;   // byte inside that chunk
    __device__ __forceinline__ TcThread(int tid) {
        warp = tid >> 5;
        lane = tid & 31;
        q = warp & 3;
        j = (warp >> 2) & 3;
        pt = lane / G::NUG;
        ug = lane % G::NUG;
        p = q * G::PW + pt;
        ubase = j * 16 + ug * G::UG;
        R0 = G::CP * p;
        row_off = (uint32_t)((R0 >> 3) * 1024 + (R0 & 7) * 128);
        chunk_x = (uint32_t)(ubase >> 3);
        chunk_b = (uint32_t)(ubase & 7) * 2u;
    }
    // byte offset (inside one split image) of this thread's UG units of row R0 + c
    __device__ __forceinline__ uint32_t img_off(int c) const {
        const int r7 = (R0 + c) & 7;
        return row_off + (uint32_t)c * 128u + (((chunk_x ^ (uint32_t)r7) << 4) + chunk_b);
    }
};

// three bf16 terms of v[c][0..UG) into rows (point, channel c) of a split image set (images `img_bytes` apart)
template <int C>
__device__ __forceinline__ void tc_store_rows(unsigned char* img, uint32_t img_bytes, const TcThread<C>& t,
                                              const float (&v)[C][TcGeo<C>::UG]) {
    constexpr int UG = TcGeo<C>::UG;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        uint32_t t1[UG / 2], t2[UG / 2], t3[UG / 2];
#pragma unroll
        for (int e = 0; e < UG / 2; ++e) split3_bf16(v[c][2 * e], v[c][2 * e + 1], t1[e], t2[e], t3[e]);
        unsigned char* dst = img + t.img_off(c);
        if constexpr (UG == 2) {
            *reinterpret_cast<uint32_t*>(dst) = t1[0];
            *reinterpret_cast<uint32_t*>(dst + img_bytes) = t2[0];
            *reinterpret_cast<uint32_t*>(dst + 2 * img_bytes) = t3[0];
        } else if constexpr (UG == 4) {
            *reinterpret_cast<uint2*>(dst) = make_uint2(t1[0], t1[1]);
            *reinterpret_cast<uint2*>(dst + img_bytes) = make_uint2(t2[0], t2[1]);
            *reinterpret_cast<uint2*>(dst + 2 * img_bytes) = make_uint2(t3[0], t3[1]);
        } else {
            *reinterpret_cast<uint4*>(dst) = make_uint4(t1[0], t1[1], t1[2], t1[3]);
            *reinterpret_cast<uint4*>(dst + img_bytes) = make_uint4(t2[0], t2[1], t2[2], t2[3]);
            *reinterpret_cast<uint4*>(dst + 2 * img_bytes) = make_uint4(t3[0], t3[1], t3[2], t3[3]);
        }
    }
}

// TMEM accumulator block (rows 32q.., units 16j..) -> owner layout through the warp's private staging block.  `tmem_acc`:
// TMEM address of column 0 / lane 0 of the accumulator.  The caller has waited for the MMA (mbarrier) and issued
// tcgen05.fence::after_thread_sync.
template <int C>
__device__ __forceinline__ void tc_load_owner(uint32_t tmem_acc, float* stage, const TcThread<C>& t,
                                              float (&v)[C][TcGeo<C>::UG]) {
    using G = TcGeo<C>;
    constexpr int UG = G::UG;
    float* my_stage = stage + (size_t)t.warp * 32 * TC_STAGE_STRIDE;
    {
        uint32_t r[16];
        const uint32_t addr = tmem_acc + (uint32_t)(t.j * 16) + ((uint32_t)(t.q * 32) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float* dst = my_stage + t.lane * TC_STAGE_STRIDE;
#pragma unroll
        for (int s = 0; s < 4; ++s)
            *reinterpret_cast<uint4*>(dst + 4 * s) = make_uint4(r[4 * s], r[4 * s + 1], r[4 * s + 2], r[4 * s + 3]);
    }
    __syncwarp();
    const float* src = my_stage + (size_t)(G::CP * t.pt) * TC_STAGE_STRIDE + t.ug * UG;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        if constexpr (UG == 2) {
            const float2 x = *reinterpret_cast<const float2*>(src + c * TC_STAGE_STRIDE);
            v[c][0] = x.x;
            v[c][1] = x.y;
        } else {
#pragma unroll
            for (int s4 = 0; s4 < UG / 4; ++s4) {
                const float4 x = *reinterpret_cast<const float4*>(src + c * TC_STAGE_STRIDE + 4 * s4);
                v[c][4 * s4 + 0] = x.x;
                v[c][4 * s4 + 1] = x.y;
                v[c][4 * s4 + 2] = x.z;
                v[c][4 * s4 + 3] = x.w;
            }
        }
    }
    __syncwarp();   // every lane has its values: the block may be overwritten by the next call
}

// "my part of the operand is written": make the generic-proxy stores visible to the tensor core (async proxy), then one
// arrival per warp on an mbarrier the MMA warp waits on (count = TC_NCW)
__device__ __forceinline__ void tc_publish(uint64_t* bar, int lane) {
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// z-jet record of one hidden layer: C*UG contiguous floats per thread
template <int C>
__device__ __forceinline__ void tc_store_record(float* __restrict__ dst, const float (&z)[C][TcGeo<C>::UG]) {
    constexpr int UG = TcGeo<C>::UG;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        if constexpr (UG == 2) {
            *reinterpret_cast<float2*>(dst + c * UG) = make_float2(z[c][0], z[c][1]);
        } else {
#pragma unroll
            for (int s4 = 0; s4 < UG / 4; ++s4)
                *reinterpret_cast<float4*>(dst + c * UG + 4 * s4) =
                    make_float4(z[c][4 * s4], z[c][4 * s4 + 1], z[c][4 * s4 + 2], z[c][4 * s4 + 3]);
        }
    }
}
template <int C>
__device__ __forceinline__ void tc_load_record(const float* __restrict__ src, float (&z)[C][TcGeo<C>::UG]) {
    constexpr int UG = TcGeo<C>::UG;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        if constexpr (UG == 2) {
            const float2 x = __ldcg(reinterpret_cast<const float2*>(src + c * UG));
            z[c][0] = x.x;
            z[c][1] = x.y;
        } else {
#pragma unroll
            for (int s4 = 0; s4 < UG / 4; ++s4) {
                const float4 x = __ldcg(reinterpret_cast<const float4*>(src + c * UG + 4 * s4));
                z[c][4 * s4 + 0] = x.x;
                z[c][4 * s4 + 1] = x.y;
                z[c][4 * s4 + 2] = x.z;
                z[c][4 * s4 + 3] = x.w;
            }
        }
    }
}

// the same record from a shared-memory copy (K2-TC stages whole record blocks with bulk TMA)
template <int C>
__device__ __forceinline__ void tc_load_record_smem(const float* src, float (&z)[C][TcGeo<C>::UG]) {
    constexpr int UG = TcGeo<C>::UG;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        if constexpr (UG == 2) {
            const float2 x = *reinterpret_cast<const float2*>(src + c * UG);
            z[c][0] = x.x;
            z[c][1] = x.y;
        } else {
#pragma unroll
            for (int s4 = 0; s4 < UG / 4; ++s4) {
                const float4 x = *reinterpret_cast<const float4*>(src + c * UG + 4 * s4);
                z[c][4 * s4 + 0] = x.x;
                z[c][4 * s4 + 1] = x.y;
                z[c][4 * s4 + 2] = x.z;
                z[c][4 * s4 + 3] = x.w;
            }
        }
    }
}

// Sum of v[k] over the PW point lanes of a warp (lanes with equal ug) by recursive halving: UG = PW / 2 values cost UG
// shuffles instead of UG * log2(PW).  Afterwards the lane with point index pt holds the total of value pt >> 1 (both lanes
// of a pair hold the same value).
template <int C>
__device__ __forceinline__ float tc_reduce_points(const float (&v)[TcGeo<C>::UG], int pt) {
    using G = TcGeo<C>;
    constexpr int UG = G::UG;
    static_assert(UG * 2 == G::PW, "one value per pair of point lanes");
#ifdef PJ_DBG_BUTTERFLY
    {
        float r = 0.0f;
#pragma unroll
        for (int k = 0; k < UG; ++k) {
            float x = v[k];
#pragma unroll
            for (int m = G::NUG; m < 32; m <<= 1) x += __shfl_xor_sync(0xffffffffu, x, m);
            if (k == (pt >> 1)) r = x;
        }
        return r;
    }
#endif
    float w[UG];
#pragma unroll
    for (int k = 0; k < UG; ++k) w[k] = v[k];
#pragma unroll
    for (int cnt = UG, bit = G::PW / 2; cnt > 1; cnt >>= 1, bit >>= 1) {
        const bool up = pt & bit;
#pragma unroll
        for (int i = 0; i < cnt / 2; ++i) {
            const float keep = up ? w[i + cnt / 2] : w[i], send = up ? w[i] : w[i + cnt / 2];
            w[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit * G::NUG);
        }
    }
    return w[0] + __shfl_xor_sync(0xffffffffu, w[0], G::NUG);
}

}  // namespace pj
