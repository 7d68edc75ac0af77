// TMA-fed persistent variants of the two FFT passes.
//
// One CTA per SM loops over tiles (FPB rows, or TC adjacent columns, = 64 KiB of field).  Tiles travel through a
// ring of three shared-memory slots: while the 512 compute threads transform the tile in slot k%3 (the exchanges
// of the register FFT run inside that same slot), the TMA engine loads tile k+1 / k+2 into the other slots and
// drains the result of tile k-1 back to HBM.  The compute threads never issue a global load or store for the
// field, so the `lg_throttle` / `long_scoreboard` stalls of the direct-access kernels (profiles/r1_prof_*.txt)
// are gone and loads, math and stores overlap.
//
//   rows   : tiles are contiguous in memory -> 1-D bulk copies (cp.async.bulk), no tensor map
//   columns: TC columns x N rows -> 2-D tiled tensor copies (cp.async.bulk.tensor.2d), boxes of 256 rows
//
// Data is staged in plain linear order; thanks to the unit-stride storage order (fft_core.cuh: io_pos) the
// register <-> slot moves at both ends are bank-conflict-free without any hardware swizzle.
#pragma once
#include <cuda.h>

#include "async_ptx.cuh"
#include "fft_passes.cuh"

namespace pa {

constexpr int kSlots = 3;
// Column pass: three slots, or two (the refill of the idle slot is then issued in the middle of the next tile, when the
// store of the previous one has long left shared memory).  Two slots leave ~96 KiB of L1 for the twiddle / transfer-
// function tables; measured per size (tools/prof_sizes.py, us per pass, 3 vs 2 slots): 1024: 46.5 / 50.4, 2048: 202.7 /
// 202.1, 4096: 283.7 / 260.9, 8192 (split pass): 633 / 656 -> two slots at 4096 only.
#ifdef PA_COL_SLOTS
template <int N> struct ColSlots { static constexpr int value = PA_COL_SLOTS; };     // experiment switch
#else
template <int N> struct ColSlots { static constexpr int value = N == 4096 ? 2 : 3; };
#endif
#ifndef PA_COL_MINBLOCKS
#define PA_COL_MINBLOCKS 2          // CTAs per SM the column kernel is compiled for when a tile needs <= 256 threads
#endif
#ifndef PA_TMA_ROW_THREADS
#define PA_TMA_ROW_THREADS 128      // row pass: small CTAs (one 2048-point row each), several per SM
#endif
#ifndef PA_TMA_ROW_TURNS
#define PA_TMA_ROW_TURNS 1
#endif
#ifndef PA_TMA_ROW_SLOTS_8192
#define PA_TMA_ROW_SLOTS_8192 2
#endif
// Row pass: three slots, or two + a staging region for the tile's screen rows (TmaRowGeo::TURNS_STAGED).  Two where the
// direct-access kernel cannot hide its loads: the 64 KiB rows of 8192^2 (one 512-thread CTA per SM; 192 KiB of slots would
// also leave the L1 too small for the 61 KiB twiddle table of the first stage) and complex128 at every size (us per pass,
// direct -> two slots + staging: 1024^2 x 8: 113.7 -> 97.1, 2048^2 x 4: 238 -> 205, 4096^2: 248 -> 224).  complex64 below
// 8192 keeps the direct kernel (131.5 against 146 at 2048^2).  With two slots the refill of the idle slot is issued in the
// middle of the next tile (as in the two-slot column ring).
template <typename T, int N> struct RowSlots { static constexpr int value = (N == 8192 || sizeof(T) == 8) ? PA_TMA_ROW_SLOTS_8192 : kSlots; };
#ifndef PA_TMA_COL_BYTES
#define PA_TMA_COL_BYTES 65536      // column pass: 64 KiB tiles (4 columns of 2048 complex64)
#endif
#ifndef PA_TMA_COL_BYTES_256
#define PA_TMA_COL_BYTES_256 32768  // 256-row tiles (256^2 grids, inner transform of the split pass): 32 KiB, two CTAs per SM (8192^2 complex128: 1326 -> 1250 us)
#endif

// geometry of the row pass: FPB rows per tile so that a CTA has about PA_TMA_ROW_THREADS threads
template <typename T, int N, int E> struct TmaRowGeo {
    using C = cplx<T>;
    static constexpr int TPF = N / E;
    static constexpr int FPB = (PA_TMA_ROW_THREADS / TPF) > 0 ? (PA_TMA_ROW_THREADS / TPF) : 1;
    static constexpr int THREADS = FPB * TPF;
    static constexpr int SLOT = FPB * N * (int)sizeof(C);
    static constexpr int CHUNK = SLOT < 16384 ? SLOT : 16384;               // bytes per bulk copy
    static constexpr int SLOTS = RowSlots<T, N>::value;
    // PA_TMA_ROW_TURNS: with a two-slot ring the screen values of a tile (FPB rows of N reals) travel through one more,
    // single-buffered region instead of per-thread global loads
    static constexpr int TSLOT = FPB * N * (int)sizeof(T);
    static constexpr bool TURNS_STAGED = PA_TMA_ROW_TURNS && SLOTS == 2 && TSLOT % 16 == 0;
    static constexpr int SMEM = SLOTS * SLOT + (TURNS_STAGED ? TSLOT : 0) + 64;
    static constexpr bool OK = THREADS <= 1024 && THREADS >= 32 && SMEM <= 227 * 1024 && SLOT % 16 == 0;
};
// geometry of the column pass: TC adjacent columns per tile
template <typename T, int N, int E> struct TmaGeo {
    using C = cplx<T>;
    static constexpr int TPF = N / E;
    static constexpr int TC = (N == 256 ? PA_TMA_COL_BYTES_256 : PA_TMA_COL_BYTES) / (N * (int)sizeof(C));
    static constexpr int THREADS = TC * TPF;
    static constexpr int SLOT = TC * N * (int)sizeof(C);
    static constexpr int BOXR = N < 256 ? N : 256;                          // rows per tensor box
    static constexpr int SLOTS = ColSlots<N>::value;
    static constexpr bool OK = TC >= 1 && THREADS <= 1024 && THREADS >= 64 && TC * (int)sizeof(C) >= 16 && TC <= 32 &&
                               TmaRowGeo<T, N, E>::OK;
    static constexpr int SMEM = SLOTS * SLOT + 64;
};

// Address map of the TMA-fed column kernel.  The tile arrives (and leaves) in the natural order [row][TC columns], and that
// order is conflict-free for every exchange except the one between the last two stages (tools/bank_sim2.py).  So the
// transform runs IN the delivered tile: exchanges between earlier stages use the natural order -- a thread's first exchange
// writes exactly the addresses it read the tile from, its last exchange reads exactly the addresses it leaves its results
// at -- and only the warp-local exchange between the last two stages moves to the swizzled order (swz_col), which stays
// inside the block of 16 R_last rows that the same R_last adjacent threads own (kDualLayout; exchange() separates the two
// layouts with a __syncwarp).  No copy between a "tile layout" and an "exchange layout", and none of their barriers.
template <int N, int E, int TC> struct ColAddrDual {
    static constexpr bool kContiguous = false;
    static constexpr int L = plan_len(N, E);
    static constexpr bool kLocalLast = L >= 2 && plan_radix(N, E, L - 1) * TC <= 32;
    static constexpr bool kDualLayout = L >= 3 && kLocalLast;
    static constexpr int kLow = 0xF;
    // without the local layout the natural order must do for every exchange: true when a tile row fills a 128-byte bank row
    static_assert(kDualLayout || TC * (E == 16 ? 8 : 16) >= 128 || PA_E32 == 8, "TMA column tile: no conflict-free single layout");
    int c;
    template <int S, bool LOCAL> __device__ __forceinline__ int at(int t, int idx) const {
        if constexpr (LOCAL && kDualLayout) {
            const int vt = swz_col<N, E, TC>(reg_pos<N, E, S>(t, 0));
            const int k = swz_col<N, E, TC>(reg_pos<N, E, S>(0, idx));
            return ((vt ^ (k & kLow)) * TC + c) + (k & ~kLow) * TC;
        } else {
            return reg_pos<N, E, S>(t, 0) * TC + c + reg_pos<N, E, S>(0, idx) * TC;
        }
    }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
// the same map with a hook that runs in front of every block-wide exchange barrier (the column kernel hands its deferred TMA
// store to it: see k_cols_tma)
template <int N, int E, int TC, typename Hook> struct ColAddrDualHook : ColAddrDual<N, E, TC> {
    Hook* hook;
    __device__ __forceinline__ ColAddrDualHook(int c_, Hook* h) : ColAddrDual<N, E, TC>{c_}, hook(h) {}
    __device__ __forceinline__ void sync() const {
        (*hook)();
        __syncthreads();
    }
};


template <typename T, int N, int E, bool IN_PERM, bool OUT_PERM>
__global__ void __launch_bounds__(TmaRowGeo<T, N, E>::THREADS) k_rows_tma(RowArgs<T> a, int ntiles) {
    using C = cplx<T>;
    using G = TmaRowGeo<T, N, E>;
    constexpr int TPF = G::TPF, FPB = G::FPB;
    constexpr int kSlotBytes = G::SLOT, RS = G::SLOTS;
    extern __shared__ __align__(1024) unsigned char smem_tma[];
    C* slots = reinterpret_cast<C*>(smem_tma);
    const uint32_t slot0 = ptx::smem_u32(smem_tma);
    constexpr bool TS = G::TURNS_STAGED;
    const uint32_t tslot = slot0 + RS * kSlotBytes;                           // screen rows of the current tile (TS)
    const uint32_t bar0 = tslot + (TS ? G::TSLOT : 0);
    const uint32_t tbar = bar0 + 8 * RS;
    const T* tsm = reinterpret_cast<const T*>(smem_tma + RS * kSlotBytes);
    const int f = threadIdx.x / TPF, t = threadIdx.x % TPF;
    if (threadIdx.x == 0) {
        for (int s = 0; s < RS; ++s) ptx::mbar_init(bar0 + 8 * s, 1);
        ptx::mbar_init(tbar, 1);
        ptx::fence_mbar_init();
    }
    __syncthreads();
    auto issue_turns = [&](int k) {          // thread 0, after every thread has finished reading the previous tile's values
        const int tile = blockIdx.x + k * gridDim.x;
        if (!TS || a.turns == nullptr || tile >= ntiles) return;
        ptx::mbar_expect_tx(tbar, G::TSLOT);
        const char* src = reinterpret_cast<const char*>(a.turns) + (size_t)tile * G::TSLOT;
        constexpr int TCH = G::TSLOT < 16384 ? G::TSLOT : 16384;
#pragma unroll
        for (int o = 0; o < G::TSLOT; o += TCH) ptx::bulk_g2s(tslot + o, src + o, TCH, tbar);
    };
    auto issue_load = [&](int k) {
        const int tile = blockIdx.x + k * gridDim.x;
        if (tile >= ntiles) return;
        const int slot = k % RS;
        ptx::mbar_expect_tx(bar0 + 8 * slot, kSlotBytes);
        const char* src = reinterpret_cast<const char*>(a.field) + (size_t)tile * kSlotBytes;
#pragma unroll
        for (int o = 0; o < kSlotBytes; o += G::CHUNK) ptx::bulk_g2s(slot0 + slot * kSlotBytes + o, src + o, G::CHUNK, bar0 + 8 * slot);
        if (!TS && a.turns != nullptr)      // pull the tile's screen rows towards L2 ahead of the compute threads
            ptx::bulk_prefetch_l2(reinterpret_cast<const char*>(a.turns) + (size_t)tile * (kSlotBytes / 2), kSlotBytes / 2);
    };
    if (threadIdx.x == 0) {
        issue_load(0);
        issue_turns(0);
        issue_load(1);
    }
    const RowAddr<N, E> addr{f * N};
    for (int k = 0;; ++k) {
        const int tile = blockIdx.x + k * gridDim.x;
        if (tile >= ntiles) break;
        const int slot = k % RS;
        C* sm = slots + (size_t)slot * (kSlotBytes / sizeof(C));
        const int row = tile * FPB + f;
        ptx::mbar_wait(bar0 + 8 * slot, (k / RS) & 1);
        C v[E];
#pragma unroll
        for (int i = 0; i < E; ++i) v[i] = sm[f * N + (IN_PERM ? io_pos<N, E>(t, i) : reg_pos<N, E, 0>(t, i))];
        __syncthreads();                                   // slot is now scratch for the exchanges
        if constexpr (RS == 2) {
            // two-slot ring: the other slot held tile k-1, whose store was committed at the end of the previous iteration
            if (threadIdx.x == 0 && k >= 1) {
                ptx::bulk_wait_read<0>();
                issue_load(k + 1);
            }
        }
        if constexpr (IN_PERM) fft_inv<T, N, E>(v, t, sm, addr, a.tw);
        if (a.turns != nullptr) {
            const T* tr = a.turns + (size_t)row * N;
            if constexpr (TS) {
                ptx::mbar_wait(tbar, k & 1);
                tr = tsm + (size_t)f * N;
            }
#pragma unroll
            for (int i = 0; i < E; ++i) {
                C e = expm2pi(tr[reg_pos<N, E, 0>(t, i)]);
                e.x *= a.scale;
                e.y *= a.scale;
                v[i] = cmul(v[i], e);
            }
        } else if (a.scale != (T)1) {
#pragma unroll
            for (int i = 0; i < E; ++i) {
                v[i].x *= a.scale;
                v[i].y *= a.scale;
            }
        }
        if constexpr (OUT_PERM) fft_fwd<T, N, E>(v, t, sm, addr, a.tw);
        __syncthreads();                                   // every exchange read is done
#pragma unroll
        for (int i = 0; i < E; ++i) sm[f * N + (OUT_PERM ? io_pos<N, E>(t, i) : reg_pos<N, E, 0>(t, i))] = v[i];
        ptx::fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            ptx::bulk_s2g(reinterpret_cast<char*>(a.field) + (size_t)tile * kSlotBytes, slot0 + slot * kSlotBytes, kSlotBytes);
            ptx::bulk_commit();
            if constexpr (RS >= 3) {
                ptx::bulk_wait_read<1>();                  // the store of tile k-1 has left its slot ...
                issue_load(k + 2);                         // ... which is the slot of tile k+2
            }
            issue_turns(k + 1);                            // the barriers above: nobody reads this tile's screen values any more
        }
    }
    if (threadIdx.x == 0) ptx::bulk_wait_read<0>();
}

// Column pass: persistent, one CTA per SM, tiles of TC adjacent columns travel through a ring of shared-memory slots
// (TMA tensor loads in, TMA tensor stores out).  Per tile the compute threads meet at TWO block-wide barriers -- the two
// exchanges that cross warps (first <-> second stage, forward and inverse); everything else is warp-local or an mbarrier the
// warps arrive on without waiting: see ColAddrDual above.
template <typename T, int N, int E>
__global__ void __launch_bounds__(TmaGeo<T, N, E>::THREADS, TmaGeo<T, N, E>::THREADS <= 256 ? PA_COL_MINBLOCKS : 1) k_cols_tma(const __grid_constant__ CUtensorMap tmap, ColArgs<T> a, int ntiles) {
    using C = cplx<T>;
    using G = TmaGeo<T, N, E>;
    constexpr int TC = G::TC, BOXR = G::BOXR, kColSlots = G::SLOTS;
    constexpr int kSlotBytes = G::SLOT;
    const int tiles_shift = 31 - __clz(a.ncols / TC);       // tiles per block of N rows (a power of two), as a shift
    constexpr int BOX_BYTES = BOXR * TC * (int)sizeof(C);
    extern __shared__ __align__(1024) unsigned char smem_tma[];
    C* slots = reinterpret_cast<C*>(smem_tma);
    const uint32_t slot0 = ptx::smem_u32(smem_tma);
    const uint32_t full0 = slot0 + kColSlots * kSlotBytes;          // "tile has landed" (TMA transaction bytes)
    const uint32_t done0 = full0 + 8 * kColSlots;                   // "every warp has left its results in the slot"
    const int c = threadIdx.x % TC, t = threadIdx.x / TC;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kColSlots; ++s) {
            ptx::mbar_init(full0 + 8 * s, 1);
            ptx::mbar_init(done0 + 8 * s, G::THREADS / 32);
        }
        ptx::fence_mbar_init();
    }
    __syncthreads();
    auto issue_load = [&](int k) {
        const int tile = blockIdx.x + k * gridDim.x;
        if (tile >= ntiles) return;
        const int slot = k % kColSlots;
        const int b = tile >> tiles_shift, col0 = (tile & ((1 << tiles_shift) - 1)) * TC;
        ptx::mbar_expect_tx(full0 + 8 * slot, kSlotBytes);
#pragma unroll
        for (int r = 0; r < N / BOXR; ++r)
            ptx::tma_load_2d(slot0 + slot * kSlotBytes + r * BOX_BYTES, &tmap, col0 * 2, b * N + r * BOXR, full0 + 8 * slot);
    };
    auto issue_store = [&](int k) {
        const int tile = blockIdx.x + k * gridDim.x;
        const int slot = k % kColSlots;
        const int b = tile >> tiles_shift, col0 = (tile & ((1 << tiles_shift) - 1)) * TC;
        ptx::mbar_wait(done0 + 8 * slot, (k / kColSlots) & 1);
#pragma unroll
        for (int r = 0; r < N / BOXR; ++r) ptx::tma_store_2d(&tmap, col0 * 2, b * N + r * BOXR, slot0 + slot * kSlotBytes + r * BOX_BYTES);
        ptx::bulk_commit();
    };
    if (threadIdx.x == 0) {
        issue_load(0);
        issue_load(1);
    }
    // Deferred store (2048- and 4096-row tiles; measured 158.0 -> 154.7 / 225.6 -> 219.0 us, but 622.9 -> 649.3 for the
    // 256-row tiles of the split pass and 39.0 -> 39.6 at 1024): the TMA store of tile k is issued by thread 0 in front of
    // the first block-wide barrier of tile k+1, when every warp has certainly left tile k -- so warp 0 never waits for the
    // others at the end of a tile (which made the hand-over a third block-wide meeting point).
    constexpr bool kLateStore = N >= 2048;
    int pending_store = -1;
    auto hook = [&]() {
        if (kLateStore && threadIdx.x == 0 && pending_store >= 0) {
            issue_store(pending_store);
            pending_store = -1;
        }
    };
    const ColAddrDualHook<N, E, TC, decltype(hook)> addr(c, &hook);
    for (int k = 0;; ++k) {
        const int tile = blockIdx.x + k * gridDim.x;
        if (tile >= ntiles) break;
        const int slot = k % kColSlots;
        C* sm = slots + (size_t)slot * (kSlotBytes / sizeof(C));
        const int col = (tile & ((1 << tiles_shift) - 1)) * TC + c;
        ptx::mbar_wait(full0 + 8 * slot, (k / kColSlots) & 1);
        C v[E];
#pragma unroll
        for (int i = 0; i < E; ++i) v[i] = sm[addr.template at<0, false>(t, i)];
        fft_fwd<T, N, E>(v, t, sm, addr, a.tw);
        {
            const C hx = cmul(ldg_c<T>(a.hp + col), mkc<T>(a.alpha_re, a.alpha_im));
            const C* hy = a.hpy + ((tile >> tiles_shift) & (a.nsub - 1)) * N;
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const C h = cmul(ldg_c<T>(hy + io_pos<N, E>(t, i)), hx);
                v[i] = cmul(v[i], h);
            }
        }
        if constexpr (kColSlots == 2) {
            // two-slot ring: the other slot holds tile k-1, whose store was committed at the end of the previous iteration and
            // has long left shared memory by now; refill it with tile k+1 (half a tile time ahead of its use)
            if (threadIdx.x == 0 && k >= 1) {
                ptx::bulk_wait_read<0>();
                issue_load(k + 1);
            }
        }
        fft_inv<T, N, E>(v, t, sm, addr, a.tw);
#pragma unroll
        for (int i = 0; i < E; ++i) sm[addr.template at<0, false>(t, i)] = v[i];
        ptx::fence_proxy_async();                      // generic-proxy writes -> visible to the TMA store
        __syncwarp();
        if ((threadIdx.x & 31) == 0) ptx::mbar_arrive(done0 + 8 * slot);
        if (threadIdx.x == 0) {
            if constexpr (kLateStore) {
                if constexpr (kColSlots >= 3) {
                    ptx::bulk_wait_read<0>();              // the store of tile k-1 (issued early in this tile) has left its slot ...
                    issue_load(k + 2);                     // ... which is the slot of tile k+2
                }
                pending_store = k;
            } else {
                issue_store(k);
                if constexpr (kColSlots >= 3) {
                    ptx::bulk_wait_read<1>();              // the store of tile k-1 has left its slot ...
                    issue_load(k + 2);                     // ... which is the slot of tile k+2
                }
            }
        }
    }
    if (threadIdx.x == 0) {
        if (pending_store >= 0) issue_store(pending_store);
        ptx::bulk_wait_read<0>();
    }
}

}  // namespace pa
