// Tensor-core phase-screen synthesis for sm_100a: tcgen05.mma (kind::f16, fp32 accumulators in TMEM) fed by
// cp.async.bulk (TMA engine) through an mbarrier pipeline, with a float64 epilogue.
//
//   phi_hi[i][j] = sum_k P[k][i] Q[k][j]      (rings >= m_split; K2 = 2 (M - m_split) rows, see screen.cu)
//
// Precision: every operand is split into two fp16 numbers, v * s = hi + lo (s a power of two: 2^10 for P,
// 2^4 for Q; lo may be subnormal), and  D += Ph Qh + Ph Ql + Pl Qh  is accumulated in ONE fp32 TMEM accumulator.
// The dropped Pl Ql term is 2^-22 relative.  Rings are fed in DESCENDING order of radius so that the running sum
// stays small until the last chunks (the variance per ring falls like f^-5/3): with round-toward-zero
// accumulation this keeps the error at ~2e-6 rad rms for the README channel (measured: tests/test_gpu_screen_tc.py;
// CPU model: DESIGN.md "screen precision").
//
// Tiling: one CTA = 128 (i) x 256 (j) output tile, K chunks of 32, 3 smem stages of 48 KB, two TMEM accumulator
// buffers of 256 columns so that the float64 epilogue of tile t overlaps the MMAs of tile t+1.
// Warp roles: 0..7 = epilogue (TMEM lane = row; the two warps of a lane quarter split the 256 columns), 8 = bulk-copy
// producer, 9 = MMA issuer, 10 = TMEM allocator.  The single-thread roles get the HIGHEST warp ids on purpose: the
// warp scheduler favours high ids, and an MMA issuer starved by eight busy epilogue warps stalls the tensor pipe
// (measured: tile time 24.8 -> see profiles/).  The epilogue is float32-only ON PURPOSE: on B200 the FP64 pipe does not
// issue at all while tcgen05.mma instructions stream (tools/micro/fp64_vs_mma.cu: a DFMA loop takes exactly its own time
// PLUS the MMA time), so any float64 in the epilogue serialises it with the MMAs of the next tile.  The low-ring
// polynomial is therefore evaluated beforehand, in float64, at every 16th column (k_poly_rows -> k_poly_nodes, one
// launch each per batch of screens) and stored as float32 triples (hi, lo, node value in turns).
// Operands are pre-tiled in global memory by k_factors_tc in the canonical K-major / no-swizzle UMMA layout
// (8 rows x 16 bytes core matrices), so one bulk copy per operand and stage fills shared memory.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "internal_screen.h"

namespace pa {
namespace tc {

constexpr int TM = 128, TN = 256, BK = 32;
constexpr int P_HALF = TM * BK * 2;            // bytes of one fp16 operand block (hi or lo)
constexpr int Q_HALF = TN * BK * 2;
constexpr int P_STAGE = 2 * P_HALF;            // 16 KiB
constexpr int Q_STAGE = 2 * Q_HALF;            // 32 KiB
constexpr int STAGE_BYTES = P_STAGE + Q_STAGE;
constexpr int STAGES = 3;
constexpr int EPI_WARPS = 8;                  // two per TMEM lane quarter, each takes half of the 256 columns
constexpr int THREADS = 128 + 32 * EPI_WARPS;
constexpr int NODE_SP = 16;                   // exact polynomial evaluation every 16th column
constexpr float Q_SCALE = 16.0f;     // P is scaled by a.p_scale (power of two chosen on the host from the coefficient bound)
constexpr uint32_t SPIN_LIMIT = 1u << 28;

// ---- PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err) {
    uint32_t ok = 0;
#pragma unroll 1
    for (uint32_t spin = 0; spin < SPIN_LIMIT; ++spin) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
    }
    atomicExch(err, 1);     // never spin for ever on the GPU box: flag and abort the kernel
    __trap();
}
// pure polling variant (mbarrier.test_wait never suspends the thread): for barriers completed from the other SM of a pair
__device__ __forceinline__ void mbar_poll(uint32_t bar, uint32_t parity, int* err) {
    uint32_t ok = 0;
#pragma unroll 1
    for (uint32_t spin = 0; spin < SPIN_LIMIT; ++spin) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
    }
    atomicExch(err, 1);
    __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;      // descriptor version (Blackwell); layout_type = 0 (no swizzle)
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- CTA-pair (cta_group::2) helpers: one tcgen05.mma drives the tensor cores of both SMs of a 2-CTA cluster --------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n .reg .b32 ra;\n mapa.shared::cluster.u32 ra, %0, %1;\n mbarrier.arrive.shared::cluster.b64 _, [ra];\n}" ::"r"(bar),
        "r"(rank)
        : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Operand tile through a tensor map, issued by either CTA of a pair: the bytes land in the issuing CTA's shared memory, the
// transaction count is credited to the LEADER's mbarrier (peer bit of the shared::cluster address cleared), so the leader's
// "full" barrier sees the operands of both CTAs without any relay thread.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma2d_pair(uint32_t dst, const void* tmap, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(tmap), "r"(bar & kPeerBitMask), "r"(x), "r"(y)
                 : "memory");
}
// completion of all earlier MMAs of the pair -> one arrival on the mbarrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 consecutive accumulator columns of this thread's TMEM lane; asynchronous: pair with tmem_wait_ld()
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// instruction descriptor: fp32 accumulate, fp16 A and B, both K-major, M = 128, N = 256
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
// CTA pair: M = 256 (128 rows per CTA), N = 256 (each CTA stages 128 of the columns)
constexpr uint32_t IDESC_PAIR = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((2 * TM) >> 4) << 24);
// shared-memory ring: single CTA 3 stages of P (16 KiB) + Q (32 KiB); pair 3 stages of 2 x [P (16 KiB) + half of Q (16 KiB)]
template <bool PAIR> struct Ring {
    static constexpr int Q_BYTES = PAIR ? Q_STAGE / 2 : Q_STAGE;
    static constexpr int Q_HALF_BYTES = Q_BYTES / 2;               // hi or lo block
    static constexpr int SUB = PAIR ? 2 : 1;                       // K blocks of 32 per stage: the pair halves its cross-CTA handshakes
    static constexpr int KB_BYTES = P_STAGE + Q_BYTES;             // one K block of this CTA's operands
    static constexpr int BYTES = SUB * KB_BYTES;
    static constexpr int N = PAIR ? 3 : STAGES;
    static constexpr int NBARS = 2 * N + 6 + (PAIR ? N : 0);
};

// ---- operand generation -----------------------------------------------------------------------------------
// Global layout (fp16): P: [screen][row block i/128][k block k/32]{hi,lo}[k chunk (k%32)/8][row group (i%128)/8][i%8][k%8]
//                       Q: [screen][col block j/256][k block k/32]{hi,lo}[k chunk][col group (j%256)/8][j%8][k%8]
// k = 2 r + {0,1} for ring rank r, where rank r is ring  m - 1 - r  (descending radius); ranks beyond the last
// high ring are zero padding up to a multiple of 16 ranks (32 k).
// grid: (n/128, 2*nscreens, KF_SPLIT), block 128; one thread = one row or column, 4 rings (one k chunk of 8) per step.
// The phase argument coord * f is reduced mod 1 with an error-free float32 product (hi + lo), then the trigonometry
// (MUFU) and the scaling run in float32: the operands only carry 22 bits (hi + lo).
// One CTA = 128 rows (P) or columns (Q) of one screen x a range of K blocks: the ring frequencies and (pre-scaled)
// coefficients of that range are staged in shared memory once, then every thread walks the K blocks with constant address
// increments -- 4 rings (8 operand values, one 16-byte hi and one 16-byte lo store) per step.  (The first version spent 60 % of
// its instructions on per-thread index arithmetic and ran at 2.6 TB/s of stores: 186 us for the 40 screens of a step.)
constexpr int KF_SPLIT = 2;          // K range split over blockIdx.z: 2 x 16 x 80 = 2560 CTAs for a step of 8 realizations

__global__ void __launch_bounds__(128) k_factors_tc(ScreenLaunch a, __half* P, __half* Q, int kpad, int pair) {
    extern __shared__ __align__(16) unsigned char kf_smem[];
    const int nhigh = a.m - a.m_split;
    const int nranks = kpad / 2;                               // ring ranks incl. zero padding
    const int kblocks = kpad / BK;
    const int kb_per = (kblocks + KF_SPLIT - 1) / KF_SPLIT;
    const int kb0 = blockIdx.z * kb_per, kb1 = min(kblocks, kb0 + kb_per);
    if (kb0 >= kb1) return;
    const bool is_q = (blockIdx.y & 1) != 0;
    const int s = blockIdx.y >> 1;
    const int r0 = kb0 * (BK / 2), r1 = kb1 * (BK / 2);        // ranks [r0, r1)
    float* sF = reinterpret_cast<float*>(kf_smem);             // [r1 - r0] frequency of rank r (0 for padding)
    float2* sC = reinterpret_cast<float2*>(sF + (nranks + 3) / 4 * 4);   // [r1 - r0] coefficient * p_scale (P only)
    const float pscale = (float)a.p_scale;
    for (int r = r0 + threadIdx.x; r < r1; r += 128) {
        float f = 0.f;
        float2 c = make_float2(0.f, 0.f);
        if (r < nhigh) {
            const size_t o = (size_t)s * a.m + (a.m - 1 - r);  // rank r is ring m - 1 - r (descending radius)
            f = is_q ? a.fx[o] : a.fy[o];
            if (!is_q) {
                c = a.coef[o];
                c.x *= pscale;
                c.y *= pscale;
            }
        }
        sF[r - r0] = f;
        if (!is_q) sC[r - r0] = c;
    }
    __syncthreads();
    const int idx = blockIdx.x * 128 + threadIdx.x;            // row i (P) or column j (Q)
    const float coord = is_q ? __fadd_rn(a.x[idx], a.shift_x) : __fadd_rn(a.y[idx], a.shift_y);
    const int tile = is_q ? TN : TM;
    const int half_bytes = is_q ? Q_HALF : P_HALF;
    const int blk = idx / tile, within = idx % tile;
    // within one hi / lo block: [k chunk][col or row group][8][8]; for CTA pairs the Q block is stored as two contiguous
    // halves of 128 columns (one per CTA of the pair), each [k chunk][16 col groups][8][8]
    size_t inner = (size_t)(within / 8) * 128 + (within % 8) * 16;
    int kc_stride = (tile / 8) * 128;
    if (is_q && pair) {
        const int cg = within / 8;
        inner = (size_t)(cg / 16) * (half_bytes / 2) + (size_t)(cg % 16) * 128 + (within % 8) * 16;
        kc_stride = 16 * 128;
    }
    char* dst = (char*)(is_q ? Q : P) + (((size_t)s * (a.n / tile) + blk) * kblocks + kb0) * (size_t)(2 * half_bytes) + inner;
    for (int kb = kb0; kb < kb1; ++kb, dst += 2 * half_bytes) {
#pragma unroll
        for (int kc = 0; kc < BK / 8; ++kc) {
            __half2 hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = (kb - kb0) * (BK / 2) + kc * 4 + q;
                // coord * f mod 1 without float64: the product is hi + lo exactly (lo from one FMA), hi - rint(hi) is exact,
                // and |lo| <= ulp(hi)/2 ~ 1.5e-5 turns, so the sum is good to 3e-8 turns like a rounded float64 result
                const float f = sF[r];
                const float hi_t = __fmul_rn(coord, f);
                const float lo_t = __fmaf_rn(coord, f, -hi_t);
                const float turns = (hi_t - rintf(hi_t)) + lo_t;
                // |angle| <= pi (+ 1e-4): the hardware approximations (abs. error ~4e-7 there) are as good as the 22-bit
                // hi+lo operands
                const float ang = 6.283185307179586f * turns;
                const float sn = __sinf(ang), cs = __cosf(ang);
                float v0, v1;
                if (is_q) {
                    v0 = cs * Q_SCALE;
                    v1 = sn * Q_SCALE;
                } else {
                    const float2 c = sC[r];
                    v0 = c.x * cs - c.y * sn;
                    v1 = -(c.x * sn + c.y * cs);
                }
                hi[q] = __floats2half2_rn(v0, v1);
                const float2 back = __half22float2(hi[q]);
                lo[q] = __floats2half2_rn(v0 - back.x, v1 - back.y);
            }
            *reinterpret_cast<uint4*>(dst + kc * kc_stride) = *reinterpret_cast<const uint4*>(hi);
            *reinterpret_cast<uint4*>(dst + kc * kc_stride + half_bytes) = *reinterpret_cast<const uint4*>(lo);
        }
    }
}

// ---- low-ring polynomial at the interpolation nodes --------------------------------------------------------------
// Nodes sit every NODE_SP pixels of the UNIFORM axes xu_j = x_first + j dxu (the reference's float32 axes are those plus a
// rounding jitter of ~1e-7 m): node q (0 <= q < nq = n/16 + 5) is pixel 16 (q - 2), in x and in y alike.
//   k_poly_rows   : Uc[s][p][qy] = sum_q' T_pq' yh(qy)^q'          grid (ceil(nq/128), D+1, nscreens)
//   k_poly_coarse : E[s][qy][qx] = sum_p Uc[s][p][qy] xh(qx)^p      grid (ceil(nq/160), nq, nscreens)   (float64 Horner)
//   k_poly_nodes  : e(i, qx) = sum_t w_t(s_i) E[s][b_i + t][qx]     6-point Lagrange in y at the ACTUAL position of row i,
//                   stored for the float32 epilogue as nodes[screen][row block i/128][qx]{hi, lo, turns}[i % 128]:
//                   e = hi + lo, turns = frac(e / 2 pi) reduced in float64; also writes the column jitter table.
// Interpolating the exact values of every 16th row instead of evaluating every row exactly costs < 1.5e-7 rad (the
// polynomial is as smooth in y as in x) and 5x fewer float64 operations, which B200 is short of (32 DFMA/clk/SM).
// Per-screen scratch inside the workspace slab of u_stride doubles: Uc at 0, E at (kMaxPolyDegree + 1) * nq.
constexpr int NODES_PT = 19;
constexpr int NODE_FLOATS = 3 * TM;          // floats per (row block, node)
constexpr int TILE_NODES = TN / NODE_SP + 5; // nodes one output tile needs

__global__ void __launch_bounds__(128) k_poly_rows(ScreenLaunch a, double* W, size_t u_stride, int nq) {
    const int D = a.degree;
    const int i = blockIdx.x * 128 + threadIdx.x;      // row node
    const int p = blockIdx.y;
    const int s = blockIdx.z;
    if (i >= nq) return;
    const double* tcf = a.polyc + (size_t)s * (D + 1) * (D + 1) + (size_t)p * (D + 1);
    const double yh = (a.y_first + a.dyu * (double)((i - 2) * NODE_SP)) * a.inv_y0;
    double u0 = 0.0, u1 = 0.0;                 // even / odd powers separately: two independent Horner chains in yh^2
    const double y2 = yh * yh;
    const int top = D - p;
    for (int q = top; q >= 0; --q) {
        if ((q & 1) == 0) u0 = fma(u0, y2, __ldg(tcf + q));
        else u1 = fma(u1, y2, __ldg(tcf + q));
    }
    W[(size_t)s * u_stride + (size_t)p * nq + i] = fma(u1, yh, u0);
}

__global__ void __launch_bounds__(160) k_poly_coarse(ScreenLaunch a, double* W, size_t u_stride, int nq) {
    const int D = a.degree;
    const int qx = blockIdx.x * 160 + threadIdx.x, qy = blockIdx.y, s = blockIdx.z;
    __shared__ double su[kMaxPolyDegree + 4];            // coefficients of this row node: one load latency, then broadcasts
    if (threadIdx.x < kMaxPolyDegree + 4)
        su[threadIdx.x] = threadIdx.x <= D ? __ldg(W + (size_t)s * u_stride + (size_t)threadIdx.x * nq + qy) : 0.0;
    __syncthreads();
    if (qx >= nq) return;
    const double xh = (a.x_first + a.dxu * (double)((qx - 2) * NODE_SP)) * a.inv_x0;
    double e[4] = {0.0, 0.0, 0.0, 0.0};        // four Horner chains in xh^4
    const double x2 = xh * xh, x4 = x2 * x2;
    for (int p4 = D / 4; p4 >= 0; --p4) {
#pragma unroll
        for (int c = 0; c < 4; ++c) e[c] = fma(e[c], x4, su[4 * p4 + c]);
    }
    W[(size_t)s * u_stride + (size_t)(kMaxPolyDegree + 1) * nq + (size_t)qy * nq + qx] = fma(fma(e[3], xh, e[2]), x2, fma(e[1], xh, e[0]));
}

__global__ void __launch_bounds__(128) k_poly_nodes(ScreenLaunch a, const double* W, size_t u_stride, float* nodes, float* jit, int nq) {
    const int n = a.n, D = a.degree;
    const int i = blockIdx.x * 128 + threadIdx.x;
    const int q0 = blockIdx.y * NODES_PT;
    const int s = blockIdx.z;
    if (blockIdx.y == 0 && s == 0)          // column jitter of the float32 axis, in units of x0 (i is a column index here)
        jit[i] = (float)(((double)__fadd_rn(a.x[i], a.shift_x) - (a.x_first + a.dxu * (double)i)) * a.inv_x0);
    if (D < 0) return;
    // Lagrange weights of row nodes b-2 .. b+3 (b = the node at or below row i) at the actual float32 row coordinate
    const int r = i % NODE_SP;
    const double sy = ((double)__fadd_rn(a.y[i], a.shift_y) - (a.y_first + a.dyu * (double)(i - r))) / ((double)NODE_SP * a.dyu);
    const double m2 = sy + 2.0, m1 = sy + 1.0, p1 = sy - 1.0, p2 = sy - 2.0, p3 = sy - 3.0;
    double w[6];
    w[0] = m1 * sy * p1 * p2 * p3 * (-1.0 / 120.0);
    w[1] = m2 * sy * p1 * p2 * p3 * (1.0 / 24.0);
    w[2] = m2 * m1 * p1 * p2 * p3 * (-1.0 / 12.0);
    w[3] = m2 * m1 * sy * p2 * p3 * (1.0 / 12.0);
    w[4] = m2 * m1 * sy * p1 * p3 * (-1.0 / 24.0);
    w[5] = m2 * m1 * sy * p1 * p2 * (1.0 / 120.0);
    // row nodes b-2 .. b+3 are rows (i / 16) .. (i / 16) + 5 of E
    const double* E = W + (size_t)s * u_stride + (size_t)(kMaxPolyDegree + 1) * nq + (size_t)(i / NODE_SP) * nq;
    float* dst = nodes + (((size_t)s * (n / TM) + i / TM) * nq) * NODE_FLOATS + (i % TM);
#pragma unroll
    for (int t = 0; t < NODES_PT; ++t) {
        const int q = q0 + t;
        if (q >= nq) break;
        double e = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) e = fma(w[k], __ldg(E + (size_t)k * nq + q), e);
        const float hi = (float)e;
        double tz = e * 0.15915494309189533576888376;
        tz -= rint(tz);
        dst[(size_t)q * NODE_FLOATS] = hi;
        dst[(size_t)q * NODE_FLOATS + TM] = (float)(e - (double)hi);
        dst[(size_t)q * NODE_FLOATS + 2 * TM] = (float)tz;
    }
}

// 6-point Lagrange weights for nodes -2..3 at s = r/16 (sum to one; applied to differences from node 0)
__device__ constexpr float kLag[16][6] = {
    {0.0000000000e+00f, 0.0000000000e+00f, 1.0000000000e+00f, 0.0000000000e+00f, 0.0000000000e+00f, 0.0000000000e+00f},
    {2.9526948929e-03f, -2.8658509254e-02f, 9.7438931465e-01f, 6.4959287643e-02f, -1.5715956688e-02f, 2.0731687546e-03f},
    {5.5274963379e-03f, -5.2204132080e-02f, 9.3967437744e-01f, 1.3423919678e-01f, -3.1322479248e-02f, 4.0855407715e-03f},
    {7.6850652695e-03f, -7.0783495903e-02f, 8.9659094810e-01f, 2.0690560341e-01f, -4.6375393867e-02f, 5.9772729874e-03f},
    {9.3994140625e-03f, -8.4594726562e-02f, 8.4594726562e-01f, 2.8198242188e-01f, -6.0424804688e-02f, 7.6904296875e-03f},
    {1.0656952858e-02f, -9.3882679939e-02f, 7.8861451149e-01f, 3.5846114159e-01f, -7.3019862175e-02f, 9.1699361801e-03f},
    {1.1455535889e-02f, -9.8934173584e-02f, 7.2551727295e-01f, 4.3531036377e-01f, -8.3713531494e-02f, 1.0364532471e-02f},
    {1.1803507805e-02f, -1.0007321835e-01f, 6.5762400627e-01f, 5.1148533821e-01f, -9.2067360878e-02f, 1.1227726936e-02f},
    {1.1718750000e-02f, -9.7656250000e-02f, 5.8593750000e-01f, 5.8593750000e-01f, -9.7656250000e-02f, 1.1718750000e-02f},
    {1.1227726936e-02f, -9.2067360878e-02f, 5.1148533821e-01f, 6.5762400627e-01f, -1.0007321835e-01f, 1.1803507805e-02f},
    {1.0364532471e-02f, -8.3713531494e-02f, 4.3531036377e-01f, 7.2551727295e-01f, -9.8934173584e-02f, 1.1455535889e-02f},
    {9.1699361801e-03f, -7.3019862175e-02f, 3.5846114159e-01f, 7.8861451149e-01f, -9.3882679939e-02f, 1.0656952858e-02f},
    {7.6904296875e-03f, -6.0424804688e-02f, 2.8198242188e-01f, 8.4594726562e-01f, -8.4594726562e-02f, 9.3994140625e-03f},
    {5.9772729874e-03f, -4.6375393867e-02f, 2.0690560341e-01f, 8.9659094810e-01f, -7.0783495903e-02f, 7.6850652695e-03f},
    {4.0855407715e-03f, -3.1322479248e-02f, 1.3423919678e-01f, 9.3967437744e-01f, -5.2204132080e-02f, 5.5274963379e-03f},
    {2.0731687546e-03f, -1.5715956688e-02f, 6.4959287643e-02f, 9.7438931465e-01f, -2.8658509254e-02f, 2.9526948929e-03f}};

// ---- contraction + epilogue ---------------------------------------------------------------------------------
struct TcArgs {
    // operand buffers as 2-D tensors of fp16: rows of 256 halves (512 bytes); one box = 16 rows = 8 KiB (pair kernel, TMAP mode)
    alignas(64) CUtensorMap mapP;
    alignas(64) CUtensorMap mapQ;
    int use_tmap;
    ScreenLaunch a;
    const __half* P;
    const __half* Q;
    int kblocks;        // kpad / 32
    int total_tiles;    // nscreens * (n/128) * (n/256)
    int swap_lbo_sbo;   // debug bits: 1 = swap LBO/SBO, 2 = no bulk copies, 4 = no MMAs, 8 = no epilogue math (timing experiments)
    int* err;
    const float* nodes; // [nscreens][n/128][nq]{hi, lo, turns}[128]: polynomial at the interpolation nodes (k_poly_nodes)
    const float* jit;   // [n] jitter of the float32 column axis around the uniform one, in units of x0
    int nq;             // nodes per row = n/16 + 5
    float inv_h;        // 1 / node spacing in normalised units
    float out_scale;    // 1 / (p_scale * Q_SCALE): accumulator -> radians
    long long* trace;   // debug bit 128: per-tile clock64 stamps of CTA 0, [tile][16]
};
#define PA_TRACE(slot) do { if (g.trace && blockIdx.x == 0 && it < 16) g.trace[it * 16 + (slot)] = clock64(); } while (0)

// PAIR = true (default; PYATM_TC_PAIR=0 selects the single-CTA variant): clusters of two CTAs share one 256 x 256 output tile
// (cta_group::2).  Each CTA stages its own 128 rows of P and 128 of the 256 columns of Q, so an MMA reads 8 KiB of shared
// memory per SM instead of 12 KiB and a K block fills 32 KiB instead of 48 KiB: the single-CTA kernel is bound by shared-
// memory bandwidth (operand reads + TMA fills = 150 B/clk against 128 B/clk per SM; measured per 8 screens: MMA floor 75 us,
// bulk copies alone 67 us, both together 105 us).  The leader CTA (cluster rank 0) issues the MMAs for both; the peer's
// MMA warp only relays "my operands have landed".
template <bool PAIR>
__global__ void __launch_bounds__(THREADS, 1) k_screen_tc(const __grid_constant__ TcArgs g) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const ScreenLaunch& a = g.a;
    using R = Ring<PAIR>;
    constexpr int STAGES = R::N, STAGE_BYTES = R::BYTES;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    unsigned char* stage_base = smem;                                        // STAGES * STAGE_BYTES
    float* sN = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);       // [TILE_NODES]{hi, lo, turns}[128]: nodes of this tile
    float* sX = sN + TILE_NODES * NODE_FLOATS;                               // [2][256] column jitter, per tile parity
    const int D = a.degree;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 2 * TN);
    // bars[0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty, [2S+4] nodes_full, [2S+5] nodes_free,
    // pair only: [2S+6..3S+6) peer_full (in the leader: the peer's operands of that stage have landed)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + R::NBARS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + b); };
    auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 2 + b); };
    const uint32_t ufull_bar = bar0 + 8u * (2 * STAGES + 4), ufree_bar = bar0 + 8u * (2 * STAGES + 5);
    auto peerfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 6 + s); };

    constexpr int W_PROD = EPI_WARPS, W_MMA = EPI_WARPS + 1, W_ALLOC = EPI_WARPS + 2;   // high warp ids: the scheduler favours them
    if (warp == W_MMA && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), EPI_WARPS * (PAIR ? 2 : 1));   // one arrival per epilogue warp (of both CTAs)
        }
        mbar_init(ufull_bar, 1);
        mbar_init(ufree_bar, EPI_WARPS);
        if constexpr (PAIR)
            for (int s = 0; s < STAGES; ++s) mbar_init(peerfull_bar(s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == W_ALLOC) {
        if constexpr (PAIR) {      // the same warp of both CTAs allocates the same columns in both tensor memories
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();        // the peer's barriers exist before anything arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n = a.n;
    const int rblocks = n / TM, cblocks = n / TN;
    // work items: single CTA = one 128 x 256 tile; pair = one 256 x 256 tile, this CTA owning row block 2 * pair_row + rank
    const int tiles_per_screen = (PAIR ? rblocks / 2 : rblocks) * cblocks;
    const int first_tile = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    auto row_block = [&](int rem) { return PAIR ? 2 * (rem / cblocks) + (int)rank : rem / cblocks; };

    if (warp == W_PROD) {
        // ===== producer: one bulk copy per operand and stage =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = first_tile; tile < g.total_tiles; tile += tile_step, ++it) {
                const int s = tile / tiles_per_screen, rem = tile % tiles_per_screen;
                const int rb = row_block(rem), cb = rem % cblocks;
                PA_TRACE(0);
                const char* psrc = (const char*)g.P + ((size_t)s * rblocks + rb) * g.kblocks * (size_t)P_STAGE;
                const char* qsrc = (const char*)g.Q + ((size_t)s * cblocks + cb) * g.kblocks * (size_t)Q_STAGE;
                for (int kb0 = 0; kb0 < g.kblocks; kb0 += R::SUB) {
                    const int nsub = g.kblocks - kb0 < R::SUB ? g.kblocks - kb0 : R::SUB;      // K blocks in this stage
                    if constexpr (PAIR) mbar_poll(empty_bar(stage), phase ^ 1, g.err);
                    else mbar_wait(empty_bar(stage), phase ^ 1, g.err);
                    if (g.swap_lbo_sbo & 2) {
                        mbar_arrive(full_bar(stage));
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    const bool tmap = PAIR && g.use_tmap;
                    // tensor-map mode: the leader's barrier collects the bytes of both CTAs; the peer only issues its copies
                    if (!tmap) mbar_expect_tx(full_bar(stage), (uint32_t)nsub * R::KB_BYTES);
                    else if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * (uint32_t)nsub * R::KB_BYTES);
                    constexpr int CH = 8192;      // several 8 KiB copies in flight instead of two large ones
#pragma unroll
                    for (int sub = 0; sub < R::SUB; ++sub) {
                        if (R::SUB > 1 && sub >= nsub) break;
                        const int kb = kb0 + sub;
                        const uint32_t dst = smem_u32(stage_base + (size_t)stage * STAGE_BYTES + (size_t)sub * R::KB_BYTES);
                        if (tmap) {
                            // the same 8 KiB pieces as below, addressed as 16-row boxes of the operand tensors
                            const size_t prow = (size_t)(psrc - (const char*)g.P + (size_t)kb * P_STAGE) / 512;
                            const size_t qrow = (size_t)(qsrc - (const char*)g.Q + (size_t)kb * Q_STAGE + rank * R::Q_HALF_BYTES) / 512;
#pragma unroll
                            for (int o = 0; o < P_STAGE; o += CH) tma2d_pair(dst + o, &g.mapP, 0, (int)(prow + o / 512), full_bar(stage));
#pragma unroll
                            for (int hl = 0; hl < 2; ++hl)
                                tma2d_pair(dst + P_STAGE + hl * R::Q_HALF_BYTES, &g.mapQ, 0, (int)(qrow + (size_t)hl * Q_HALF / 512), full_bar(stage));
                            continue;
                        }
#pragma unroll
                        for (int o = 0; o < P_STAGE; o += CH) bulk_g2s(dst + o, psrc + (size_t)kb * P_STAGE + o, CH, full_bar(stage));
                        if constexpr (PAIR) {
#pragma unroll
                            for (int hl = 0; hl < 2; ++hl)      // this CTA's 128 columns are one contiguous 8 KiB piece of the hi / lo block
                                bulk_g2s(dst + P_STAGE + hl * R::Q_HALF_BYTES,
                                         qsrc + (size_t)kb * Q_STAGE + (size_t)hl * Q_HALF + rank * R::Q_HALF_BYTES, R::Q_HALF_BYTES, full_bar(stage));
                        } else {
#pragma unroll
                            for (int o = 0; o < Q_STAGE; o += CH) bulk_g2s(dst + P_STAGE + o, qsrc + (size_t)kb * Q_STAGE + o, CH, full_bar(stage));
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                PA_TRACE(1);
            }
        }
    } else if (warp == W_MMA) {
        // ===== MMA issuer =====
        if (PAIR && rank != 0) {
            // ===== peer CTA: no MMAs of its own -- tell the leader when this CTA's operands of a stage have landed =====
            if (lane == 0 && !g.use_tmap) {
                int stage = 0;
                uint32_t phase = 0;
                for (int tile = first_tile; tile < g.total_tiles; tile += tile_step) {
                    for (int kb0 = 0; kb0 < g.kblocks; kb0 += R::SUB) {
                        mbar_poll(full_bar(stage), phase, g.err);
                        if (!(g.swap_lbo_sbo & 16)) mbar_arrive_remote(peerfull_bar(stage), 0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        } else if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            constexpr int QCOLS = PAIR ? TN / 2 : TN;          // columns of Q staged in this CTA
            const uint32_t p_lbo = (g.swap_lbo_sbo & 1) ? 128u : (uint32_t)(TM / 8) * 128u, p_sbo = (g.swap_lbo_sbo & 1) ? (uint32_t)(TM / 8) * 128u : 128u;
            const uint32_t q_lbo = (g.swap_lbo_sbo & 1) ? 128u : (uint32_t)(QCOLS / 8) * 128u, q_sbo = (g.swap_lbo_sbo & 1) ? (uint32_t)(QCOLS / 8) * 128u : 128u;
            for (int tile = first_tile; tile < g.total_tiles; tile += tile_step, ++it) {
                const int buf = it & 1;
                PA_TRACE(2);
                if constexpr (PAIR) mbar_poll(tempty_bar(buf), ((it >> 1) & 1) ^ 1, g.err);
                else mbar_wait(tempty_bar(buf), ((it >> 1) & 1) ^ 1, g.err);     // epilogue has drained this accumulator
                tc_fence_after();
                PA_TRACE(3);
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * TN;
                for (int kb0 = 0; kb0 < g.kblocks; kb0 += R::SUB) {
                    const int nsub = g.kblocks - kb0 < R::SUB ? g.kblocks - kb0 : R::SUB;
                    mbar_wait(full_bar(stage), phase, g.err);
                    if constexpr (PAIR)
                        if (!(g.swap_lbo_sbo & 16) && !g.use_tmap) mbar_poll(peerfull_bar(stage), phase, g.err);   // bit 16: timing experiment without the relay
                    tc_fence_after();
#pragma unroll
                  for (int sub = 0; sub < R::SUB; ++sub) {
                    if (R::SUB > 1 && sub >= nsub) break;
                    const int kb = kb0 + sub;
                    const uint32_t sp = smem_u32(stage_base + (size_t)stage * STAGE_BYTES + (size_t)sub * R::KB_BYTES);
                    const uint32_t sq = sp + P_STAGE;
#pragma unroll
                    for (int k16 = 0; k16 < BK / 16; ++k16) {
                        if (g.swap_lbo_sbo & 4) break;
                        const uint32_t poff = (uint32_t)k16 * 2u * (TM / 8) * 128u, qoff = (uint32_t)k16 * 2u * (QCOLS / 8) * 128u;
                        const uint64_t ah = umma_desc(sp + poff, p_lbo, p_sbo), al = umma_desc(sp + P_HALF + poff, p_lbo, p_sbo);
                        const uint64_t bh = umma_desc(sq + qoff, q_lbo, q_sbo), bl = umma_desc(sq + R::Q_HALF_BYTES + qoff, q_lbo, q_sbo);
                        if constexpr (PAIR) {
                            umma_f16_pair(d_tmem, al, bh, IDESC_PAIR, (kb | k16) != 0);
                            umma_f16_pair(d_tmem, ah, bl, IDESC_PAIR, 1);
                            umma_f16_pair(d_tmem, ah, bh, IDESC_PAIR, 1);
                        } else {
                            umma_f16(d_tmem, al, bh, IDESC, (kb | k16) != 0);
                            umma_f16(d_tmem, ah, bl, IDESC, 1);
                            umma_f16(d_tmem, ah, bh, IDESC, 1);
                        }
                    }
                  }
                    // smem slot free (in both CTAs of a pair) once these MMAs have read it
                    if constexpr (PAIR) umma_commit_pair(empty_bar(stage));
                    else umma_commit(empty_bar(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if constexpr (PAIR) umma_commit_pair(tfull_bar(buf));     // accumulator complete, in both CTAs
                else umma_commit(tfull_bar(buf));
                PA_TRACE(4);
            }
        }
    } else if (warp == W_ALLOC) {
        // ===== node producer: the polynomial values of the next tile travel into sN as soon as the epilogue warps
        // have read those of the current one =====
        if (lane == 0 && D >= 0) {
            int it = 0;
            constexpr uint32_t BYTES = TILE_NODES * NODE_FLOATS * 4u, PIECE = BYTES / 3u;
            static_assert(BYTES % 48 == 0, "three 16-byte aligned pieces");
            for (int tile = first_tile; tile < g.total_tiles; tile += tile_step, ++it) {
                const int s = tile / tiles_per_screen, rem = tile % tiles_per_screen;
                const int rb = row_block(rem), cb = rem % cblocks;
                if (it > 0) mbar_wait(ufree_bar, (it - 1) & 1, g.err);
                const char* src = reinterpret_cast<const char*>(g.nodes + (((size_t)s * rblocks + rb) * g.nq + (size_t)cb * (TN / NODE_SP)) * NODE_FLOATS);
                mbar_expect_tx(ufull_bar, BYTES);
#pragma unroll
                for (uint32_t o = 0; o < BYTES; o += PIECE) bulk_g2s(smem_u32(sN) + o, src + o, PIECE, ufull_bar);
            }
        }
    } else if (warp < EPI_WARPS) {
        // ===== epilogue: TMEM lane = output row; low-ring polynomial + accumulator -> turns, float32 only =====
        // The polynomial is smooth on the scale of tens of pixels (its highest harmonic has a wavelength of ~1000
        // pixels): between two nodes the DIFFERENCE from the nearer-left node (a few radians at most) is interpolated
        // with 6-point Lagrange weights, the node value itself arrives already reduced to turns, and the rounding jitter
        // of the reference's float32 axis is put back to first order:  phi(x_j) = phi(xu_j) + (x_j - xu_j) dphi/dx.
        // Measured against the float64 path: < 2e-7 rad.
        const int ew = warp & 3;                       // TMEM lane quarter this warp may read
        const int ch = warp >> 2;                      // which half of the 256 columns
        const int et = threadIdx.x;                    // index among the epilogue threads (warps 0..7)
        const int row_in_tile = ew * 32 + lane;
        const float out_scale_f = g.out_scale;
        int it = 0;
        for (int tile = first_tile; tile < g.total_tiles; tile += tile_step, ++it) {
            const int s = tile / tiles_per_screen, rem = tile % tiles_per_screen;
            const int rb = row_block(rem), cb = rem % cblocks;
            const int i = rb * TM + row_in_tile;
            float* sXt = sX + (it & 1) * TN;           // the other parity may still be read by a slower warp
            if (et < TN) sXt[et] = __ldg(g.jit + cb * TN + et);
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
            const int buf = it & 1;
            if (threadIdx.x == 0) PA_TRACE(5);
            mbar_wait(tfull_bar(buf), (it >> 1) & 1, g.err);
            tc_fence_after();
            if (D >= 0) mbar_wait(ufull_bar, it & 1, g.err);          // nodes of this tile have landed in sN
            if (threadIdx.x == 0) PA_TRACE(6);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                if (g.swap_lbo_sbo & 8) {          // timing experiment without the epilogue math: only hand sN back
                    __syncwarp();
                    if (lane == 0 && D >= 0) mbar_arrive(ufree_bar);
                    break;
                }
                const int cbase = ch * (TN / 2) + half * 64;            // first of the 64 columns of this round
                const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * TN + cbase);
                uint32_t acc0[32], acc1[32];
                tmem_ld32_async(t_row, acc0);
                // the 9 nodes cbase + 16 (t - 2), t = 0..8, of this row
                float nh[9], nl[9], nt[4];
                if (D >= 0) {
                    const float* sn = sN + (size_t)(cbase / NODE_SP) * NODE_FLOATS + row_in_tile;
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        nh[t] = sn[t * NODE_FLOATS];
                        nl[t] = sn[t * NODE_FLOATS + TM];
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) nt[k] = sn[(k + 2) * NODE_FLOATS + 2 * TM];
                } else {
#pragma unroll
                    for (int t = 0; t < 9; ++t) nh[t] = nl[t] = 0.f;
#pragma unroll
                    for (int k = 0; k < 4; ++k) nt[k] = 0.f;
                }
                if (threadIdx.x == 0) PA_TRACE(8 + 3 * half);
                if (half == 1 && D >= 0) {        // sN is no longer needed by this warp: let the next tile's nodes come in
                    __syncwarp();
                    if (lane == 0) mbar_arrive(ufree_bar);
                }
                const int j0 = cb * TN + cbase;
                float* turns = (a.turns && !(g.swap_lbo_sbo & 64)) ? (float*)a.turns + ((size_t)s * n + i) * n + j0 : nullptr;
                tmem_wait_ld();                                          // first 32 accumulator columns have arrived
                if (threadIdx.x == 0) PA_TRACE(9 + 3 * half);
                tmem_ld32_async(t_row + 32u, acc1);                      // next 32 travel while these are finished
#pragma unroll
                for (int k = 0; k < 4; ++k) {                            // interval between nodes k+2 and k+3: 16 columns
                    if (k == 2) tmem_wait_ld();
                    const float zh = nh[k + 2], zl = nl[k + 2];
                    const float t1f = nt[k];                             // node value in turns, reduced in float64 beforehand
                    // hi parts are a few radians apart at magnitude <= 1e4: their difference is exact in float32
                    const float d0 = (nh[k] - zh) + (nl[k] - zl), d1 = (nh[k + 1] - zh) + (nl[k + 1] - zl);
                    const float d3 = (nh[k + 3] - zh) + (nl[k + 3] - zl), d4 = (nh[k + 4] - zh) + (nl[k + 4] - zl);
                    const float d5 = (nh[k + 5] - zh) + (nl[k + 5] - zl);
                    const float slope = d3 * g.inv_h;
                    float tv[16], dv[16];
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const int c = 16 * k + r;
                        const float accv = __uint_as_float(c < 32 ? acc0[c & 31] : acc1[c & 31]);
                        float delta = kLag[r][0] * d0;
                        delta = fmaf(kLag[r][1], d1, delta);
                        delta = fmaf(kLag[r][3], d3, delta);
                        delta = fmaf(kLag[r][4], d4, delta);
                        delta = fmaf(kLag[r][5], d5, delta);
                        delta = fmaf(sXt[cbase + c], slope, delta);
                        dv[r] = fmaf(accv, out_scale_f, delta);          // phase minus the node value (a few radians)
                        float tt = fmaf(dv[r], 0.15915494309189533577f, t1f);
                        tv[r] = tt - rintf(tt);
                    }
                    if (turns) {        // 256-bit stores: every lane writes whole 32-byte sectors of its own row
#pragma unroll
                        for (int q = 0; q < 2; ++q)
                            st_global_v8(turns + 16 * k + 8 * q, tv + 8 * q);
                    }
                    if (a.phi) {                                         // full phase requested (generator / inspection path)
                        const size_t o = ((size_t)s * n + i) * n + j0 + 16 * k;
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            const double ph = (double)zh + (double)zl + (double)dv[r];
                            if (a.phi_f64) ((double*)a.phi)[o + r] = ph;
                            else ((float*)a.phi)[o + r] = (float)ph;
                        }
                    }
                }
                if (threadIdx.x == 0) PA_TRACE(10 + 3 * half);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {        // the leader's MMA warp waits for the epilogue warps of both CTAs
                if (PAIR && rank != 0) mbar_arrive_remote(tempty_bar(buf), 0);
                else mbar_arrive(tempty_bar(buf));
            }
            if (threadIdx.x == 0) PA_TRACE(7);
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();       // neither CTA leaves (or frees tensor memory) while its peer may still touch it
    if (warp == W_ALLOC) {
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

template <bool PAIR> constexpr int smem_bytes() {
    return Ring<PAIR>::N * Ring<PAIR>::BYTES + (TILE_NODES * NODE_FLOATS + 2 * TN) * (int)sizeof(float) + Ring<PAIR>::NBARS * 8 + 16;
}
constexpr int SMEM_BYTES = smem_bytes<false>();
constexpr int SMEM_BYTES_PAIR = smem_bytes<true>();

}  // namespace tc

// workspace requirement of the tensor-core path, in bytes, for `nscreens` screens
size_t screen_tc_workspace(int n, int m, int m_split, int nscreens) {
    const int nhigh = m - m_split;
    const int kpad = ((2 * nhigh + tc::BK - 1) / tc::BK) * tc::BK;
    const size_t nq = (size_t)n / tc::NODE_SP + 5;
    return (size_t)nscreens * 2 /*P,Q*/ * 2 /*hi,lo*/ * n * (size_t)(kpad > 0 ? kpad : tc::BK) * sizeof(__half) +
           (size_t)nscreens * (kMaxPolyDegree + 1) * n * sizeof(double) +                   // U table
           (size_t)nscreens * (n / tc::TM) * nq * tc::NODE_FLOATS * sizeof(float) +         // node table
           (size_t)n * sizeof(float);                                                        // column jitter
}

// phase: 0 = prepare only (operands + row coefficients for the a.nscreens screens of `a`, stored as screens
// [first_screen, first_screen + a.nscreens) of a workspace laid out for total_screens); 1 = contraction only for those
// screens (their operands must have been prepared); 2 = both.
int launch_screen_tc(const ScreenLaunch& a, void* workspace, int* err_flag, int num_sms, int swap, cudaStream_t st, int phase,
                     int first_screen, int total_screens, cudaStream_t factors_stream, cudaEvent_t factors_done) {
    using namespace tc;
    if (a.n % TN != 0) return (int)cudaErrorInvalidValue;
    const int nhigh = a.m - a.m_split;
    int kpad = ((2 * nhigh + BK - 1) / BK) * BK;
    if (kpad == 0) kpad = BK;
    if (total_screens < 0) total_screens = a.nscreens;
    const size_t pq_stride = (size_t)2 * a.n * kpad;                    // halves per screen (hi + lo)
    const size_t u_stride = (size_t)(kMaxPolyDegree + 1) * a.n;        // doubles reserved per screen
    __half* Pall = (__half*)workspace;
    __half* Qall = Pall + (size_t)total_screens * pq_stride;
    double* Uall = reinterpret_cast<double*>(Qall + (size_t)total_screens * pq_stride);
    const int nq = a.n / NODE_SP + 5;
    const size_t node_stride = (size_t)(a.n / TM) * nq * NODE_FLOATS;   // floats per screen
    float* Nall = reinterpret_cast<float*>(Uall + (size_t)total_screens * u_stride);
    float* jit = Nall + (size_t)total_screens * node_stride;
    float* nodes = Nall + (size_t)first_screen * node_stride;
    __half* P = Pall + (size_t)first_screen * pq_stride;
    __half* Q = Qall + (size_t)first_screen * pq_stride;
    // row coefficients are packed with the actual degree: [(D+1)][n] per screen inside the reserved slab
    double* U = Uall + (size_t)first_screen * u_stride;
    // CTA pairs (cta_group::2) by default; PYATM_TC_PAIR=0 selects the single-CTA kernel and its operand layout.
    // Per 8 screens of 2048^2: single 120 us, pair 113 us (two K blocks per stage; 127 us with one, 165 us with a
    // release.cluster fence on the remote arrives).  Without the peer -> leader "operands landed" relay (debug bit 16, wrong
    // results) the pair runs in 100 us: the next step is to let the peer's TMA signal the leader's barrier directly
    // (cp.async.bulk.tensor ... cta_group::2), which needs the operands behind tensor maps.
    static const bool pair = !(getenv("PYATM_TC_PAIR") && atoi(getenv("PYATM_TC_PAIR")) == 0) && !(swap & 128);
    if (phase == 0 || phase == 2) {
        dim3 gf(a.n / 128, 2 * a.nscreens, KF_SPLIT);
        const size_t kf_smem = ((size_t)(kpad / 2 + 3) / 4 * 4) * sizeof(float) + (size_t)(kpad / 2) * sizeof(float2);
        // store-bound operand generation next to the float64 / latency-bound polynomial kernels below, on a stream of its own
        k_factors_tc<<<gf, 128, kf_smem, factors_stream ? factors_stream : st>>>(a, P, Q, kpad, pair ? 1 : 0);
        if (factors_stream && factors_done) cudaEventRecord(factors_done, factors_stream);
        if (a.degree >= 0) {
            dim3 gu((nq + 127) / 128, a.degree + 1, a.nscreens);
            k_poly_rows<<<gu, 128, 0, st>>>(a, U, u_stride, nq);
            dim3 gc((nq + 159) / 160, nq, a.nscreens);
            k_poly_coarse<<<gc, 160, 0, st>>>(a, U, u_stride, nq);
        }
        dim3 gn(a.n / 128, (nq + NODES_PT - 1) / NODES_PT, a.nscreens);
        k_poly_nodes<<<gn, 128, 0, st>>>(a, U, u_stride, nodes, jit, nq);
        if (factors_stream && factors_done) cudaStreamWaitEvent(st, factors_done, 0);      // join: operands ready for `st`
        if (phase == 0) return (int)cudaGetLastError();
    }
    static SmemOptIn attr_done, attr_done_pair;
    if (cudaError_t e = attr_done.raise(k_screen_tc<false>, SMEM_BYTES); e != cudaSuccess) return (int)e;
    if (pair)
        if (cudaError_t e = attr_done_pair.raise(k_screen_tc<true>, SMEM_BYTES_PAIR); e != cudaSuccess) return (int)e;
    TcArgs g;
    memset(&g.mapP, 0, sizeof(g.mapP));
    memset(&g.mapQ, 0, sizeof(g.mapQ));
    // The operands of the pair kernel go through tensor maps with cta_group::2 so that the peer's bytes are credited to the
    // leader's barrier (no relay thread): 110.9 -> 108.8 us per 8 screens, tests/test_gpu_screen_tc.py and
    // tests/test_gpu_c3_parity.py green with it (round 2).  PYATM_TC_PAIR_TMAP=0 brings the relay back.
    static const bool want_tmap = !(getenv("PYATM_TC_PAIR_TMAP") && atoi(getenv("PYATM_TC_PAIR_TMAP")) == 0);
    g.use_tmap = 0;
    if (pair && want_tmap) {
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static EncodeFn encode = nullptr;
        if (!encode) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
                return (int)cudaErrorNotSupported;
            encode = (EncodeFn)fn;
        }
        const size_t bytes = (size_t)a.nscreens * pq_stride * sizeof(__half);       // per operand, this launch's screens
        const cuuint64_t dims[2] = {256, (cuuint64_t)(bytes / 512)};
        const cuuint64_t strides[1] = {512};
        const cuuint32_t box[2] = {256, 16};
        const cuuint32_t estr[2] = {1, 1};
        for (int w = 0; w < 2; ++w) {
            const CUresult r = encode(w ? &g.mapQ : &g.mapP, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w ? (void*)Q : (void*)P, dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
        }
        g.use_tmap = 1;
    }
    g.a = a;
    g.P = P;
    g.Q = Q;
    g.kblocks = kpad / BK;
    g.total_tiles = a.nscreens * (a.n / TM) * (a.n / TN);
    g.swap_lbo_sbo = swap;
    g.err = err_flag;
    g.nodes = nodes;
    g.jit = jit;
    g.nq = nq;
    g.inv_h = (float)(1.0 / ((double)NODE_SP * a.dxu * a.inv_x0));
    g.out_scale = (float)(1.0 / (a.p_scale * (double)Q_SCALE));
    g.trace = nullptr;
    if (pair) {
        g.total_tiles = a.nscreens * (a.n / (2 * TM)) * (a.n / TN);      // 256 x 256 tiles, one per CTA pair
        const int pairs = g.total_tiles < num_sms / 2 ? g.total_tiles : num_sms / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = SMEM_BYTES_PAIR;
        cfg.stream = st;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = 2;
        at.val.clusterDim.y = 1;
        at.val.clusterDim.z = 1;
        cfg.attrs = &at;
        cfg.numAttrs = 1;
        return (int)cudaLaunchKernelEx(&cfg, k_screen_tc<true>, g);
    }
    const int grid = g.total_tiles < num_sms ? g.total_tiles : num_sms;
    if (swap & 128) {           // debug: clock64 stamps of CTA 0 printed to stderr (synchronises)
        static long long* trace_dev = nullptr;
        if (!trace_dev) cudaMalloc(&trace_dev, 16 * 16 * sizeof(long long));
        cudaMemsetAsync(trace_dev, 0, 16 * 16 * sizeof(long long), st);
        g.trace = trace_dev;
        k_screen_tc<false><<<grid, THREADS, SMEM_BYTES, st>>>(g);
        long long h[16 * 16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost);
        long long t0 = h[0];
        fprintf(stderr, "k_screen_tc trace (CTA 0, cycles since first stamp): prod_start prod_end mma_wait mma_go mma_commit epi_wait epi_go epi_done | half0: horner tmem interp  half1: horner tmem interp\n");
        for (int it = 0; it < 16 && h[it * 16] != 0; ++it) {
            fprintf(stderr, "  tile %2d:", it);
            for (int k = 0; k < 14; ++k) fprintf(stderr, " %9lld", h[it * 16 + k] ? h[it * 16 + k] - t0 : -1);
            fprintf(stderr, "\n");
        }
        return (int)cudaGetLastError();
    }
    k_screen_tc<false><<<grid, THREADS, SMEM_BYTES, st>>>(g);
    return (int)cudaGetLastError();
}

}  // namespace pa
