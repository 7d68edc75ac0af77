// Sparse-spectrum phase-screen synthesis (replaces SSPhaseScreen.generate_phase_screen,
// /root/reference/pyatmosphere/phase_screens.py:108-136, contraction at :125-126).
//
//   phi[i][j] = Re sum_m c_m exp(2 pi i ys_i fy_m) exp(2 pi i fx_m xs_j),  xs = fl32(x + sx), ys = fl32(y + sy)
//
// Rings are ordered by radius (RandLogPolarGrid annuli are nested), so the sum is split at m_split:
//   * m <  m_split ("low" rings: huge |c_m|, arguments <= theta_cut over the whole grid): the sum of harmonics
//     is a bivariate polynomial  sum_{p+q<=D} T_pq xh^p yh^q  (Taylor series of exp), evaluated in float64.
//   * m >= m_split ("high" rings): real contraction  phi_hi = P^T Q  with K2 = 2 (M - m_split) rows,
//       P[2k][i] = Re(c a_i), P[2k+1][i] = -Im(c a_i), Q[2k][j] = cos(2 pi fx xs_j), Q[2k+1][j] = sin(...).
// This file holds the float64 CUDA-core contraction (exact path, used for complex128 runs and as the on-device
// check of the tensor-core path in screen_tc.cu).  The result is written as "turns" = frac(phi / 2 pi) in
// [-0.5, 0.5] (what the row pass consumes) and, optionally, as the full phase.
#include "common.cuh"
#include "internal_screen.h"

namespace pa {

__constant__ double c_invfact[kMaxPolyDegree + 2];

int screen_init_constants() {
    double f[kMaxPolyDegree + 2];
    double acc = 1.0;
    f[0] = 1.0;
    for (int i = 1; i < kMaxPolyDegree + 2; ++i) {
        acc *= (double)i;
        f[i] = 1.0 / acc;
    }
    return (int)cudaMemcpyToSymbol(c_invfact, f, sizeof(f));
}

// ---- factor rows for the high rings, float64 ------------------------------------------------------------
// grid: (ceil(n/256), m - m_split, nscreens)
__global__ void k_screen_factors64(ScreenLaunch a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    const int s = blockIdx.z;
    if (i >= a.n) return;
    const int m = a.m_split + k;
    const float fx = a.fx[(size_t)s * a.m + m];
    const float fy = a.fy[(size_t)s * a.m + m];
    const float2 c = a.coef[(size_t)s * a.m + m];
    const int k2 = 2 * (a.m - a.m_split);
    double* P = a.P + ((size_t)s * k2 + 2 * k) * a.n;
    double* Q = a.Q + ((size_t)s * k2 + 2 * k) * a.n;
    const double ys = (double)__fadd_rn(a.y[i], a.shift_y);
    const double xs = (double)__fadd_rn(a.x[i], a.shift_x);
    double sa, ca, sb, cb;
    sincospi(2.0 * (ys * (double)fy), &sa, &ca);
    sincospi(2.0 * (xs * (double)fx), &sb, &cb);
    P[i] = (double)c.x * ca - (double)c.y * sa;
    P[a.n + i] = -((double)c.x * sa + (double)c.y * ca);
    Q[i] = cb;
    Q[a.n + i] = sb;
}

// ---- polynomial coefficients of the low rings -----------------------------------------------------------
// T^_pq = Re( i^(p+q) sum_{m<m_split} c_m (2 pi fx_m X0)^p (2 pi fy_m Y0)^q ) / (p! q!),   p + q <= D
// grid: (nscreens, kPolySplit), block 256.  Rings are processed in chunks of 32: their power tables a^p, b^q
// (p, q <= D) are built once per chunk in shared memory, then every thread accumulates its (p,q) pairs with one
// DMUL + two DFMA per ring (the float64 pipe is the bound: 32 lanes/clk/SM on B200).  Entries with p+q > D are 0.
constexpr int kPolySplit = 8;
constexpr int kPolyChunk = 32;
// pairs of block y: rows p = y, y + kPolySplit, ... ; a thread's pairs have consecutive q within a warp, so the a^p
// reads are broadcasts and the b^q reads are conflict-free.  320 blocks for 40 screens: one wave, 2-3 blocks per SM.
__global__ void __launch_bounds__(256) k_screen_poly_coef(ScreenLaunch a) {
    const int D = a.degree, D1 = D + 1;
    const int s = blockIdx.x;
    constexpr int kPad = kMaxPolyDegree + 5;      // pitch of 68 doubles: the interleaved table build below is conflict-free
    __shared__ double apow[kPolyChunk][kPad];
    __shared__ double bpow[kPolyChunk][kPad];
    __shared__ double cre[kPolyChunk], cim[kPolyChunk];
    constexpr int kRows = (kMaxPolyDegree + kPolySplit) / kPolySplit;                 // rows of p per block
    constexpr int kPer = (kRows * (kMaxPolyDegree + 1) + 255) / 256;
    double sr[kPer], si[kPer];
    int pp[kPer], qq[kPer];
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
        sr[u] = 0.0;
        si[u] = 0.0;
        const int l = threadIdx.x + 256 * u;
        const int p = blockIdx.y + kPolySplit * (l / D1), q = l % D1;
        pp[u] = (p <= D && q <= D - p) ? p : -1;
        qq[u] = q;
    }
    for (int m0 = 0; m0 < a.m_split; m0 += kPolyChunk) {
        __syncthreads();
        {   // four threads per (ring, a-or-b) build a power table: thread k fills the powers p = k (mod 4)
            const int r = threadIdx.x >> 3, which = (threadIdx.x >> 2) & 1, part = threadIdx.x & 3, m = m0 + r;
            double base = 0.0;
            if (m < a.m_split)
                base = 6.283185307179586476925287 * (which ? (double)a.fy[(size_t)s * a.m + m] * a.y0 : (double)a.fx[(size_t)s * a.m + m] * a.x0);
            const double b2 = base * base, b4 = b2 * b2;
            double w = part == 0 ? 1.0 : (part == 1 ? base : (part == 2 ? b2 : b2 * base));
            double* row = which ? bpow[r] : apow[r];
            for (int p = part; p <= D; p += 4) {
                row[p] = w;
                w *= b4;
            }
            if (!which && part == 0) {
                const float2 c = m < a.m_split ? a.coef[(size_t)s * a.m + m] : make_float2(0.f, 0.f);
                cre[r] = (double)c.x;
                cim[r] = (double)c.y;
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            if (pp[u] < 0) continue;
            const int p = pp[u], q = qq[u];
            double tr0 = sr[u], ti0 = si[u], tr1 = 0.0, ti1 = 0.0;       // two independent chains per pair
#pragma unroll 4
            for (int r = 0; r < kPolyChunk; r += 2) {
                const double w0 = apow[r][p] * bpow[r][q], w1 = apow[r + 1][p] * bpow[r + 1][q];
                tr0 = fma(cre[r], w0, tr0);
                ti0 = fma(cim[r], w0, ti0);
                tr1 = fma(cre[r + 1], w1, tr1);
                ti1 = fma(cim[r + 1], w1, ti1);
            }
            sr[u] = tr0 + tr1;
            si[u] = ti0 + ti1;
        }
    }
    double* out = a.polyc + (size_t)s * D1 * D1;
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
        const int l = threadIdx.x + 256 * u;
        const int p = blockIdx.y + kPolySplit * (l / D1), q = l % D1;
        if (p > D) continue;
        double v = 0.0;
        if (pp[u] >= 0) {
            switch ((p + q) & 3) {   // Re(i^k (sr + i si))
                case 0: v = sr[u]; break;
                case 1: v = -si[u]; break;
                case 2: v = -sr[u]; break;
                default: v = si[u]; break;
            }
            v *= c_invfact[p] * c_invfact[q];
        }
        out[p * D1 + q] = v;
    }
}

// ---- float64 contraction + epilogue -----------------------------------------------------------------------
// tile 64 (i) x 64 (j), 256 threads, 4x4 outputs per thread, K chunk 16.  grid: (n/64, n/64, nscreens)
constexpr int kTS = 64;
constexpr int kKC = 16;

template <typename TOUT>
__global__ void __launch_bounds__(256) k_screen_gemm64(ScreenLaunch a) {
    extern __shared__ __align__(16) double smd[];
    double* sP = smd;                    // [2][kKC][kTS]
    double* sQ = smd + 2 * kKC * kTS;    // [2][kKC][kTS]
    const int s = blockIdx.z;
    const int i0 = blockIdx.y * kTS;
    const int j0 = blockIdx.x * kTS;
    const int tx = threadIdx.x % 16;     // j direction
    const int ty = threadIdx.x / 16;     // i direction
    const int k2 = 2 * (a.m - a.m_split);
    const double* P = a.P + (size_t)s * k2 * a.n;
    const double* Q = a.Q + (size_t)s * k2 * a.n;
    double acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int w = 0; w < 4; ++w) acc[u][w] = 0.0;

    // each thread loads 4 doubles of P and 4 of Q per chunk: chunk has 16*64 = 1024 per matrix
    const int lk = threadIdx.x / 16;           // 0..15
    const int li = (threadIdx.x % 16) * 4;     // 0..60
    auto load = [&](int buf, int kc) {
        const int k = kc * kKC + lk;
        double2 p0 = make_double2(0, 0), p1 = p0, q0 = p0, q1 = p0;
        if (k < k2) {
            const double2* pp = reinterpret_cast<const double2*>(P + (size_t)k * a.n + i0 + li);
            const double2* qq = reinterpret_cast<const double2*>(Q + (size_t)k * a.n + j0 + li);
            p0 = pp[0]; p1 = pp[1]; q0 = qq[0]; q1 = qq[1];
        }
        double2* dp = reinterpret_cast<double2*>(sP + ((size_t)buf * kKC + lk) * kTS + li);
        double2* dq = reinterpret_cast<double2*>(sQ + ((size_t)buf * kKC + lk) * kTS + li);
        dp[0] = p0; dp[1] = p1; dq[0] = q0; dq[1] = q1;
    };
    const int nchunk = (k2 + kKC - 1) / kKC;
    if (nchunk > 0) load(0, 0);
    __syncthreads();
    for (int kc = 0; kc < nchunk; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < nchunk) load(buf ^ 1, kc + 1);
#pragma unroll
        for (int kk = 0; kk < kKC; ++kk) {
            const double2* pr = reinterpret_cast<const double2*>(sP + ((size_t)buf * kKC + kk) * kTS + ty * 4);
            const double2* qr = reinterpret_cast<const double2*>(sQ + ((size_t)buf * kKC + kk) * kTS + tx * 4);
            const double2 pa = pr[0], pb = pr[1], qa = qr[0], qb = qr[1];
            const double pv[4] = {pa.x, pa.y, pb.x, pb.y};
            const double qv[4] = {qa.x, qa.y, qb.x, qb.y};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int w = 0; w < 4; ++w) acc[u][w] = fma(pv[u], qv[w], acc[u][w]);
        }
        __syncthreads();
    }

    // ---- epilogue: low-ring polynomial + reduction to turns --------------------------------------------
    const int D = a.degree;
    double* sU = smd;      // [(D+1)][kTS] : U_p(yh_i) = sum_q T^_pq yh^q   (reuses the GEMM buffers)
    if (D >= 0) {
        const double* tc = a.polyc + (size_t)s * (D + 1) * (D + 1);
        for (int e = threadIdx.x; e < (D + 1) * kTS; e += blockDim.x) {
            const int p = e / kTS, ii = e % kTS;
            const double yh = (double)__fadd_rn(a.y[i0 + ii], a.shift_y) * a.inv_y0;
            double u = 0.0;
            for (int q = D - p; q >= 0; --q) u = fma(u, yh, tc[p * (D + 1) + q]);
            sU[p * kTS + ii] = u;
        }
        __syncthreads();
    }
    TOUT* turns = (TOUT*)a.turns + (size_t)s * a.n * a.n;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const int j = j0 + tx * 4 + w;
        const double xh = (double)__fadd_rn(a.x[j], a.shift_x) * a.inv_x0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int ii = ty * 4 + u;
            double phi = acc[u][w];
            if (D >= 0) {
                double pl = 0.0;
                for (int p = D; p >= 0; --p) pl = fma(pl, xh, sU[p * kTS + ii]);
                phi += pl;
            }
            const size_t o = (size_t)(i0 + ii) * a.n + j;
            if (a.turns) {
                const double tt = phi * 0.15915494309189533576888376;
                turns[o] = (TOUT)(tt - rint(tt));
            }
            if (a.phi) {
                if (a.phi_f64) ((double*)a.phi)[(size_t)s * a.n * a.n + o] = phi;
                else ((float*)a.phi)[(size_t)s * a.n * a.n + o] = (float)phi;
            }
        }
    }
}

int launch_screen_poly(const ScreenLaunch& a, cudaStream_t st) {
    if (a.degree >= 0) {
        dim3 g(a.nscreens, kPolySplit);
        k_screen_poly_coef<<<g, 256, 0, st>>>(a);
    }
    return (int)cudaGetLastError();
}

int launch_screen_exact(const ScreenLaunch& a, cudaStream_t st) {
    const int khalf = a.m - a.m_split;
    if (khalf > 0) {
        dim3 g((a.n + 255) / 256, khalf, a.nscreens);
        k_screen_factors64<<<g, 256, 0, st>>>(a);
    }
    if (a.degree >= 0) {
        dim3 g(a.nscreens, kPolySplit);
        k_screen_poly_coef<<<g, 256, 0, st>>>(a);
    }
    const int smem_gemm = 4 * kKC * kTS * (int)sizeof(double);
    const int smem_poly = (a.degree + 1) * kTS * (int)sizeof(double);
    const int smem = smem_gemm > smem_poly ? smem_gemm : smem_poly;
    dim3 g(a.n / kTS, a.n / kTS, a.nscreens);
    if (a.turns_f64) {
        k_screen_gemm64<double><<<g, 256, smem, st>>>(a);
    } else {
        k_screen_gemm64<float><<<g, 256, smem, st>>>(a);
    }
    return (int)cudaGetLastError();
}

}  // namespace pa
