// In-register / shared-memory FFT building blocks for the split-step passes.
//
// Scheme ("in-place positions"): a length-N transform is factored N = R0*R1*...*R(L-1).  Element positions
// p in [0,N) never move: stage s works on the digit of p with stride SIGMA_s = N/(R0..Rs).  The forward
// transform is decimation-in-frequency (butterfly, then twiddle), takes natural-order input and leaves
// frequency k = sum_s d_s * prod_{u<s} R_u at position p = sum_s d_s * SIGMA_s (mixed-radix digit
// reversal).  The inverse runs the same stages backwards (conjugate twiddle, then inverse butterfly) and so
// maps the permuted spectrum back to natural order.  The permuted order is never undone on the device: the
// split-step path only multiplies spectra element-wise, by tables that the host stores in the same order.
//
// Each thread owns E elements; a transform uses N/E threads.  At stage s a thread owns G = E/R butterflies;
// butterfly b = t + g*(N/E) touches positions base(b) + j*SIGMA, base(b) = (b/SIGMA)*SIGMA*R + b%SIGMA.
// Between stages the E registers go through shared memory (write at stage-s positions, one barrier, read at
// stage-(s+1) positions).  Because every stage reads exactly the positions it later writes, one barrier per
// exchange is enough.
//
// LAST STAGE (SIGMA = 1): its butterflies are handed out differently (plan_last_butterfly) -- thread t takes the ones that
// lie inside the blocks of SIGMA_(L-2) * R_(L-2) positions it already works on in stage L-2.  Such a block is shared by
// R_(L-1) CONSECUTIVE threads, so the exchange between the last two stages never leaves a group of R_(L-1) adjacent
// threads: when that group sits inside one warp the exchange needs __syncwarp() only (Addr::kLocalLast), which removes
// two of the four barriers of a 3-stage transform pair (forward + inverse).
#pragma once
#include "common.cuh"

#ifndef PA_E32
#define PA_E32 16       // complex64 elements per thread (internal.h); 8 is an experiment switch
#endif

namespace pa {

// ---- radix plan (compile time) ------------------------------------------------------------------------
__host__ __device__ constexpr int plan_pick(int rem, int e) { return rem >= e ? (rem == 2 * e && e >= 8 ? e / 2 : e) : rem; }
__host__ __device__ constexpr int plan_rem(int n, int e, int s) {  // size still to be factored before stage s
    int rem = n;
    for (int u = 0; u < s; ++u) rem /= plan_pick(rem, e);
    return rem;
}
__host__ __device__ constexpr int plan_radix(int n, int e, int s) { return plan_pick(plan_rem(n, e, s), e); }
__host__ __device__ constexpr int plan_sigma(int n, int e, int s) { return plan_rem(n, e, s) / plan_radix(n, e, s); }
__host__ __device__ constexpr int plan_len(int n, int e) {
    int s = 0;
    while (plan_rem(n, e, s) > 1) ++s;
    return s;
}
// offset (in complex elements) of stage s inside the concatenated twiddle table; stage L-1 has none.  Stage s holds
// (R-1) * SIGMA entries: the factor of output j of butterfly b is W^(j * (b % SIGMA)), stored at (j-1) * SIGMA + b % SIGMA
// (compact: threads that share b % SIGMA read the same 8 bytes, one wavefront instead of two per warp-wide load).
__host__ __device__ constexpr int plan_tw_off(int n, int e, int s) {
    int off = 0;
    for (int u = 0; u < s; ++u) off += (plan_radix(n, e, u) - 1) * plan_sigma(n, e, u);
    return off;
}
__host__ __device__ constexpr int plan_tw_size(int n, int e) { return plan_tw_off(n, e, plan_len(n, e) - 1); }

// Butterfly of the LAST stage (SIGMA = 1, radix RL) that thread t works on as its g-th (g < E / RL):
// with R' = radix of stage L-2 (whose SIGMA is RL), G' = E / R' blocks per thread there and H = R' / RL butterflies per
// thread and block:  g = g' H + h,  bt = t + g' (N/E),  b = R' (bt / RL) + bt % RL + RL h.   (L = 1: b = t + g N/E.)
__host__ __device__ constexpr int plan_last_butterfly(int n, int e, int t, int g) {
    const int L = plan_len(n, e);
    if (L < 2) return t + g * (n / e);
    const int rl = plan_radix(n, e, L - 1), rp = plan_radix(n, e, L - 2);
    const int gp_count = e / rp, h_count = (e / rl) / gp_count;
    const int gp = g / h_count, h = g % h_count;
    const int bt = t + gp * (n / e);
    return rp * (bt / rl) + (bt % rl) + rl * h;
}

template <int N, int E, int S> struct Stage {
    static constexpr int R = plan_radix(N, E, S);
    static constexpr int SIGMA = plan_sigma(N, E, S);
    static constexpr int G = E / R;
    static constexpr int TPF = N / E;  // threads per transform
    static constexpr int NB = N / R;   // butterflies per transform
    static constexpr int TW = plan_tw_off(N, E, S);
    static constexpr bool LAST = S == plan_len(N, E) - 1;
    __device__ static __forceinline__ int base(int t, int g) {
        if constexpr (LAST && plan_len(N, E) >= 2) {
            // plan_last_butterfly with every plan quantity folded at compile time (the generic function is for the host)
            constexpr int RP = plan_radix(N, E, plan_len(N, E) - 2);
            constexpr int H = (E / R) / (E / RP);
            const int bt = t + (g / H) * TPF;
            return (RP * (bt / R) + (bt % R) + R * (g % H)) * R;
        } else {
            const int b = t + g * TPF;
            return (b / SIGMA) * (SIGMA * R) + (b % SIGMA);
        }
    }
};

// ---- small DFTs on registers v[OFF .. OFF+R), forward sign exp(-2 pi i nk/R), natural-order output -----
template <typename C, int OFF, int STR, int E> __device__ __forceinline__ void dft2(C (&v)[E]) {
    C a = v[OFF], b = v[OFF + STR];
    v[OFF] = cadd(a, b);
    v[OFF + STR] = csub(a, b);
}

// radix-4 over v[OFF + STR*{0,1,2,3}], in place, natural order
template <typename C, int OFF, int STR, int E> __device__ __forceinline__ void dft4(C (&v)[E]) {
    C a = v[OFF], b = v[OFF + STR], c = v[OFF + 2 * STR], d = v[OFF + 3 * STR];
    C t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = cmul_mi(csub(b, d));
    v[OFF] = cadd(t0, t2);
    v[OFF + STR] = cadd(t1, t3);
    v[OFF + 2 * STR] = csub(t0, t2);
    v[OFF + 3 * STR] = csub(t1, t3);
}

template <typename T> struct Consts {
    static constexpr T RSQRT2 = (T)0.70710678118654752440084436210485;
    static constexpr T C1 = (T)0.92387953251128675612818318939679;  // cos(pi/8)
    static constexpr T S1 = (T)0.38268343236508977172845998403040;  // sin(pi/8)
};

template <typename T, int OFF, int E> __device__ __forceinline__ void dft8(cplx<T> (&v)[E]) {
    using C = cplx<T>;
    // evens -> v[OFF+0,2,4,6], odds -> v[OFF+1,3,5,7]
    dft4<C, OFF, 2, E>(v);
    dft4<C, OFF + 1, 2, E>(v);
    // E[k] at OFF+2k, O[k] at OFF+2k+1
    const T h = Consts<T>::RSQRT2;
    C o0 = v[OFF + 1];
    C o1 = v[OFF + 3];
    C o2 = v[OFF + 5];
    C o3 = v[OFF + 7];
    o1 = mkc<T>((o1.x + o1.y) * h, (o1.y - o1.x) * h);    // * (1-i)/sqrt2
    o2 = cmul_mi(o2);                                     // * (-i)
    o3 = mkc<T>((o3.y - o3.x) * h, -(o3.x + o3.y) * h);   // * (-1-i)/sqrt2
    C e0 = v[OFF], e1 = v[OFF + 2], e2 = v[OFF + 4], e3 = v[OFF + 6];
    v[OFF + 0] = cadd(e0, o0);
    v[OFF + 1] = cadd(e1, o1);
    v[OFF + 2] = cadd(e2, o2);
    v[OFF + 3] = cadd(e3, o3);
    v[OFF + 4] = csub(e0, o0);
    v[OFF + 5] = csub(e1, o1);
    v[OFF + 6] = csub(e2, o2);
    v[OFF + 7] = csub(e3, o3);
}

template <typename T, int OFF, int E> __device__ __forceinline__ void dft16(cplx<T> (&v)[E]) {
    using C = cplx<T>;
    // n = 4*n1 + n2, k = k1 + 4*k2.  Step 1: radix-4 over n1 for each n2 (stride 4) -> y[n2][k1] at OFF+4*k1+n2
    dft4<C, OFF + 0, 4, E>(v);
    dft4<C, OFF + 1, 4, E>(v);
    dft4<C, OFF + 2, 4, E>(v);
    dft4<C, OFF + 3, 4, E>(v);
    // twiddle y[n2][k1] *= w16^(n2*k1)
    const T h = Consts<T>::RSQRT2, c1 = Consts<T>::C1, s1 = Consts<T>::S1;
    // k1 = 1: n2 = 1,2,3 -> w^1, w^2, w^3
    v[OFF + 5] = cmul(v[OFF + 5], mkc<T>(c1, -s1));
    v[OFF + 6] = mkc<T>((v[OFF + 6].x + v[OFF + 6].y) * h, (v[OFF + 6].y - v[OFF + 6].x) * h);
    v[OFF + 7] = cmul(v[OFF + 7], mkc<T>(s1, -c1));
    // k1 = 2: w^2, w^4, w^6
    v[OFF + 9] = mkc<T>((v[OFF + 9].x + v[OFF + 9].y) * h, (v[OFF + 9].y - v[OFF + 9].x) * h);
    v[OFF + 10] = cmul_mi(v[OFF + 10]);
    v[OFF + 11] = mkc<T>((v[OFF + 11].y - v[OFF + 11].x) * h, -(v[OFF + 11].x + v[OFF + 11].y) * h);
    // k1 = 3: w^3, w^6, w^9
    v[OFF + 13] = cmul(v[OFF + 13], mkc<T>(s1, -c1));
    v[OFF + 14] = mkc<T>((v[OFF + 14].y - v[OFF + 14].x) * h, -(v[OFF + 14].x + v[OFF + 14].y) * h);
    v[OFF + 15] = cmul(v[OFF + 15], mkc<T>(-c1, s1));
    // Step 2: radix-4 over n2 for each k1 (contiguous groups of 4) -> X[k1 + 4*k2] at OFF + 4*k1 + k2
    dft4<C, OFF + 0, 1, E>(v);
    dft4<C, OFF + 4, 1, E>(v);
    dft4<C, OFF + 8, 1, E>(v);
    dft4<C, OFF + 12, 1, E>(v);
    // transpose 4x4 so that X[k] sits at OFF + k
    C tmp;
#define PA_SWAP(a, b) tmp = v[OFF + a]; v[OFF + a] = v[OFF + b]; v[OFF + b] = tmp;
    PA_SWAP(1, 4) PA_SWAP(2, 8) PA_SWAP(3, 12) PA_SWAP(6, 9) PA_SWAP(7, 13) PA_SWAP(11, 14)
#undef PA_SWAP
}

// forward/inverse DFT of size R over v[OFF..OFF+R).  Inverse via IDFT(x) = swap(DFT(swap(x))).
template <typename T, int R, bool INV, int OFF, int E> __device__ __forceinline__ void dftR(cplx<T> (&v)[E]) {
    using C = cplx<T>;
    if constexpr (INV) {
#pragma unroll
        for (int j = 0; j < R; ++j) v[OFF + j] = cswap(v[OFF + j]);
    }
    if constexpr (R == 2) dft2<C, OFF, 1, E>(v);
    else if constexpr (R == 4) dft4<C, OFF, 1, E>(v);
    else if constexpr (R == 8) dft8<T, OFF, E>(v);
    else if constexpr (R == 16) dft16<T, OFF, E>(v);
    else static_assert(R == 2 || R == 4 || R == 8 || R == 16, "unsupported radix");
    if constexpr (INV) {
#pragma unroll
        for (int j = 0; j < R; ++j) v[OFF + j] = cswap(v[OFF + j]);
    }
}

template <typename T, int R, bool INV, int G, int E, int GI = 0> __device__ __forceinline__ void dft_groups(cplx<T> (&v)[E]) {
    if constexpr (GI < G) {
        dftR<T, R, INV, GI * R, E>(v);
        dft_groups<T, R, INV, G, E, GI + 1>(v);
    }
}

template <typename T> __device__ __forceinline__ cplx<T> ldg_c(const cplx<T>* p) { return __ldg(p); }

// ---- one stage ------------------------------------------------------------------------------------------
// forward (DIF): butterflies, then multiply output j of butterfly b by W^(j * (b % SIGMA)), W = exp(-2 pi i/(SIGMA R))
template <typename T, int N, int E, int S> __device__ __forceinline__ void stage_fwd(cplx<T> (&v)[E], int t, const cplx<T>* __restrict__ tw) {
    using St = Stage<N, E, S>;
    dft_groups<T, St::R, false, St::G, E>(v);
    if constexpr (St::SIGMA > 1) {
#pragma unroll
        for (int g = 0; g < St::G; ++g) {
            const int b = t + g * St::TPF;
#pragma unroll
            for (int j = 1; j < St::R; ++j) v[g * St::R + j] = cmul(v[g * St::R + j], ldg_c<T>(tw + St::TW + (j - 1) * St::SIGMA + b % St::SIGMA));
        }
    }
}
// inverse (DIT): conjugate twiddle first, then inverse butterflies (unnormalised)
template <typename T, int N, int E, int S> __device__ __forceinline__ void stage_inv(cplx<T> (&v)[E], int t, const cplx<T>* __restrict__ tw) {
    using St = Stage<N, E, S>;
    if constexpr (St::SIGMA > 1) {
#pragma unroll
        for (int g = 0; g < St::G; ++g) {
            const int b = t + g * St::TPF;
#pragma unroll
            for (int j = 1; j < St::R; ++j) v[g * St::R + j] = cmulc(v[g * St::R + j], ldg_c<T>(tw + St::TW + (j - 1) * St::SIGMA + b % St::SIGMA));
        }
    }
    dft_groups<T, St::R, true, St::G, E>(v);
}

// STORAGE ORDER of a spectrum: the register idx of thread t after the last forward stage (which holds position
// reg_pos<L-1>(t, idx)) is stored at index  t + idx * (N/E).  All global / TMA-staged I/O is therefore a unit-
// stride access over the threads of a transform, for natural data (reg_pos<0> has the same form) and for
// spectra alike; the resulting permutation (storage index -> frequency) is exported by pa_ctx_permutation.
template <int N, int E> __device__ __forceinline__ int io_pos(int t, int idx) { return t + idx * (N / E); }

// position of register idx of thread t in the distribution of stage S
template <int N, int E, int S> __device__ __forceinline__ int reg_pos(int t, int idx) {
    using St = Stage<N, E, S>;
    return St::base(t, idx / St::R) + (idx % St::R) * St::SIGMA;
}

// ---- shared-memory exchange from the distribution of stage SA to that of stage SB --------------------------
// A stage with SIGMA == 1 owns runs of R consecutive positions: with a contiguous address map (rows) and
// complex64 data these are moved as 128-bit accesses (two elements), which is what the row swizzle is
// conflict-free for (tools/bank_sim2.py).
// Addresses.  Every address map is  position -> (XOR swizzle that is GF(2)-linear in the bits of the position), and the
// position of register idx of thread t splits into a thread part reg_pos(t, 0) and a register part reg_pos(0, idx) on
// DISJOINT bits, so  addr(t, idx) = swz(thread part) ^ swz(register part)  with the second factor a compile-time constant
// K.  Addr::at<S>(t, idx) returns it as  (A_t ^ (K & low)) + (K & ~low)  where `low` are the bits a swizzle may write: the
// few distinct low parts cost one LOP3 each per exchange and everything else folds into the immediate offset of the
// LDS / STS, instead of 16 separately computed (and register-resident) addresses per exchange.
// at<S, LOCAL>: LOCAL marks the warp-local exchange between the last two stages; maps with kDualLayout use a layout of its
// own there (ColAddrDual, fft_tma.cuh), all others ignore it.
template <typename T, int N, int E, int S, bool LOCAL, typename Addr>
__device__ __forceinline__ void smem_put(const cplx<T> (&v)[E], int t, cplx<T>* sm, const Addr& addr) {
    using St = Stage<N, E, S>;
    if constexpr (Addr::kContiguous && St::SIGMA == 1 && sizeof(cplx<T>) == 8 && St::R % 2 == 0) {
#pragma unroll
        for (int i = 0; i < E; i += 2)
            *reinterpret_cast<float4*>(sm + addr.template at<S, LOCAL>(t, i)) = make_float4(v[i].x, v[i].y, v[i + 1].x, v[i + 1].y);
    } else {
#pragma unroll
        for (int i = 0; i < E; ++i) sm[addr.template at<S, LOCAL>(t, i)] = v[i];
    }
}
template <typename T, int N, int E, int S, bool LOCAL, typename Addr>
__device__ __forceinline__ void smem_get(cplx<T> (&v)[E], int t, const cplx<T>* sm, const Addr& addr) {
    using St = Stage<N, E, S>;
    if constexpr (Addr::kContiguous && St::SIGMA == 1 && sizeof(cplx<T>) == 8 && St::R % 2 == 0) {
#pragma unroll
        for (int i = 0; i < E; i += 2) {
            const float4 q = *reinterpret_cast<const float4*>(sm + addr.template at<S, LOCAL>(t, i));
            v[i] = mkc<T>(q.x, q.y);
            v[i + 1] = mkc<T>(q.z, q.w);
        }
    } else {
#pragma unroll
        for (int i = 0; i < E; ++i) v[i] = sm[addr.template at<S, LOCAL>(t, i)];
    }
}
template <typename T, int N, int E, int SA, int SB, typename Addr>
__device__ __forceinline__ void exchange(cplx<T> (&v)[E], int t, cplx<T>* sm, const Addr& addr) {
    constexpr int L = plan_len(N, E);
    // LOCAL: the exchange between the last two stages when its R_(L-1) adjacent threads share a warp
    constexpr bool LOCAL = Addr::kLocalLast && (SA == L - 1 || SB == L - 1);
    if constexpr (Addr::kDualLayout && L >= 3 && SA == L - 2) {
        // dual-layout address maps keep the local exchange in a layout of its own: this thread's previous read of its
        // stage-(L-2) positions used the other layout, whose addresses (inside the same block of the same R_(L-1) threads)
        // the writes below may hit -- let every lane of the group finish that read first
        __syncwarp();
    }
    smem_put<T, N, E, SA, LOCAL>(v, t, sm, addr);
    if constexpr (LOCAL) {
        __syncwarp();           // between the last two stages data stays inside groups of R_(L-1) adjacent threads of one warp
    } else {
        addr.sync();
    }
    smem_get<T, N, E, SB, LOCAL>(v, t, sm, addr);
}

// ---- whole transforms on registers ------------------------------------------------------------------------
// forward: input in stage-0 distribution (natural order), output in stage-(L-1) distribution (permuted spectrum)
template <typename T, int N, int E, typename Addr, int S = 0>
__device__ __forceinline__ void fft_fwd(cplx<T> (&v)[E], int t, cplx<T>* sm, const Addr& addr, const cplx<T>* __restrict__ tw) {
    constexpr int L = plan_len(N, E);
    stage_fwd<T, N, E, S>(v, t, tw);
    if constexpr (S + 1 < L) {
        exchange<T, N, E, S, S + 1>(v, t, sm, addr);
        fft_fwd<T, N, E, Addr, S + 1>(v, t, sm, addr, tw);
    }
}
// inverse (unnormalised): input in stage-(L-1) distribution (permuted spectrum), output in stage-0 distribution
template <typename T, int N, int E, typename Addr, int S = plan_len(N, E) - 1>
__device__ __forceinline__ void fft_inv(cplx<T> (&v)[E], int t, cplx<T>* sm, const Addr& addr, const cplx<T>* __restrict__ tw) {
    stage_inv<T, N, E, S>(v, t, tw);
    if constexpr (S > 0) {
        exchange<T, N, E, S, S - 1>(v, t, sm, addr);
        fft_inv<T, N, E, Addr, S - 1>(v, t, sm, addr, tw);
    }
}

// inverse split in two so that a caller can issue independent loads between the last exchange and the last stage:
// head = stages L-1 .. 1 plus the exchange into the stage-0 distribution, tail = stage 0
template <typename T, int N, int E, typename Addr, int S = plan_len(N, E) - 1>
__device__ __forceinline__ void fft_inv_head(cplx<T> (&v)[E], int t, cplx<T>* sm, const Addr& addr, const cplx<T>* __restrict__ tw) {
    if constexpr (S > 0) {
        stage_inv<T, N, E, S>(v, t, tw);
        exchange<T, N, E, S, S - 1>(v, t, sm, addr);
        fft_inv_head<T, N, E, Addr, S - 1>(v, t, sm, addr, tw);
    }
}
template <typename T, int N, int E> __device__ __forceinline__ void fft_inv_tail(cplx<T> (&v)[E], int t, const cplx<T>* __restrict__ tw) {
    stage_inv<T, N, E, 0>(v, t, tw);
}

// host-side mirrors of the plan, used to build twiddle/permutation tables
inline int host_plan_len(int n, int e) { return plan_len(n, e); }

}  // namespace pa
