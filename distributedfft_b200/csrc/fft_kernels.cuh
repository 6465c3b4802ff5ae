// fft_kernels.cuh — sm_100a batched 1D FFT kernels over "segmented strided views".
//
// One launch = one axis pass of the distributed 3D transform, i.e. one cufftExec* of the reference
// (/root/reference/src/slab/default/mpicufft_slab.cpp:788,806,847,856;
//  /root/reference/src/pencil/mpicufft_pencil.cpp:1662,1695-1698,1723) fused with the pack / unpack
// copies around it (mpicufft_slab.cpp:646-655,691-695; mpicufft_pencil.cpp:873-927,1507-1538): the
// store side of a pass can scatter along the transformed axis into up to MAXSEG destination segments
// (local send slots or peer-mapped receive buffers on other GPUs), the load side can gather likewise.
//
// Two thread mappings:
//   CONTIG — the transformed axis is contiguous in memory (z passes, R2C/C2R).  TB lines per CTA;
//            consecutive threads walk along the line, so global accesses are coalesced over n.
//   TILED  — the transformed axis is strided (y and x passes).  A CTA owns an N x TB tile with TB
//            consecutive elements of the contiguous dimension b; consecutive threads walk along b, so
//            every global access is a TB*sizeof(complex) byte run.
// Both stage through shared memory between radix stages only; first-stage loads and last-stage stores
// go registers <-> HBM directly (one read + one write of the array per pass = the algorithmic minimum).
//
// Instruction diet (the kernels are close to issue-bound in fp64): direction is a template parameter,
// single-segment views use one base pointer plus immediate / incremental offsets, multi-segment views
// read a per-line table of segment bases from shared memory, and shared-memory indices are a
// per-thread base plus compile-time constants (SmemLayout in fft_core.cuh).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <type_traits>

#include "fft_core.cuh"

namespace dfft {

constexpr int MAXSEG = 16;

struct Seg {
    void* base;     // element pointer (complex elements of the pass' precision)
    long long sA0;  // stride of batch index a0 (elements)
    long long sA1;  // stride of batch index a1
    int n0;         // first n covered by this segment
    int pad_;
};

// element (a0, a1, n, b) of segment s lives at  seg[s].base + a0*sA0 + a1*sA1 + (n - n0)*sN + b
struct View {
    const unsigned char* seg_of_n;  // n -> segment id (device memory); ignored when nseg == 1
    long long sN;                   // stride along the transformed axis (same for every segment)
    int nseg;
    int pad_;
    Seg seg[MAXSEG];
};

struct FftParams {
    View in, out;
    int A0, A1;       // batch extents: line id = a0*A1 + a1
    int B;            // TILED: extent of the contiguous dimension; CONTIG: 1
    int inverse;      // 0 forward (e^-), 1 inverse (e^+), both unnormalised like cuFFT
    const void* tw;   // exp(-2*pi*i*m/N), m < N, N = pass length (complex length for R2C/C2R)
    const void* tw2;  // R2C/C2R only: exp(-2*pi*i*k/(2N)), k <= N/2
    int max_ctas;     // > 0: run as a persistent kernel on at most this many CTAs (SM partitioning for the
                      // overlapped schedule: the exchange pass keeps a few SMs, the local passes the rest)
    int tile_pref;    // TILED: 0 automatic, 1 narrow tiles, 2 wide tiles
    int bulk_out;     // TILED, TMA-fed kernel, experiment (DFFT_BULK_STORE=1): store each destination's rows with one
                      // cp.async.bulk from shared memory instead of 16-byte stores from registers
    int tile_swz;     // TILED: log2 G of the tile-order blocking: G x G tiles of (a0, a1) are numbered consecutively, so
                      // that CTAs running at the same time touch neighbouring rows on BOTH sides of a transposing pass
                      // (0 = plain order: b tiles fastest, then a1, then a0)
};

enum PassKind { PASS_C2C_CONTIG = 0, PASS_C2C_TILED = 1, PASS_R2C = 2, PASS_C2R = 3 };

// ---- tile shape choices (compile time) -----------------------------------------------------------
template <typename T, int LOG2N>
struct Shape {
    // points per thread: 2^LOG2E_MAX (experiment knobs DFFT_LOG2E_F64 / DFFT_LOG2E_F32; 16 points = radix-16
    // butterflies and ~120 registers in f64, 8 points = radix-8 and <= 64 registers, i.e. twice the resident warps)
#ifndef DFFT_LOG2E_F64
#define DFFT_LOG2E_F64 4
#endif
#ifndef DFFT_LOG2E_F32
#define DFFT_LOG2E_F32 4
#endif
    static constexpr int LOG2E_MAX = sizeof(T) == 8 ? DFFT_LOG2E_F64 : DFFT_LOG2E_F32;
    static constexpr int LOG2E = LOG2N < LOG2E_MAX ? LOG2N : LOG2E_MAX;
    static constexpr int N = 1 << LOG2N;
    static constexpr int TPL = N >> LOG2E;
    // CONTIG: lines per CTA.  Measured on B200 (tools/axis_bench.py): f64 lines of >= 512 points run 6-9 % faster
    // with 128-thread CTAs (four CTAs per SM interleave their load / compute / store phases), short lines
    // and f32 prefer 256 threads.  -DDFFT_CONTIG_THREADS=<n> overrides for experiments.
#ifdef DFFT_CONTIG_THREADS
    static constexpr int CT = DFFT_CONTIG_THREADS;
#else
    static constexpr int CT = (sizeof(T) == 8 && LOG2N >= 9) ? 128 : 256;
#endif
    static constexpr int TBC = (CT / TPL) < 1 ? 1 : ((CT / TPL) > 64 ? 64 : (CT / TPL));
    // TILED: tile width; rows of >= DFFT_MINROW_BYTES where shared memory allows, DFFT_TILE_POINTS_F64 / _F32
    // points per tile (experiment knobs; defaults measured best on B200)
#ifndef DFFT_MINROW_BYTES
#define DFFT_MINROW_BYTES 64
#endif
#ifndef DFFT_TILE_POINTS_F64
#define DFFT_TILE_POINTS_F64 4096
#endif
#ifndef DFFT_TILE_POINTS_F32
#define DFFT_TILE_POINTS_F32 4096
#endif
    static constexpr int MINROW = DFFT_MINROW_BYTES / int(2 * sizeof(T));
    static constexpr int WANT = (sizeof(T) == 8 ? DFFT_TILE_POINTS_F64 : DFFT_TILE_POINTS_F32) / N;
    static constexpr int TBT_ = WANT < MINROW ? MINROW : (WANT > 32 ? 32 : WANT);
    // keep the tile within 128 KB of shared memory and 1024 threads
    static constexpr int CAP1 = (128 * 1024) / (N * int(2 * sizeof(T)));
    static constexpr int CAP2 = 1024 / TPL;
    static constexpr int CAP = CAP1 < CAP2 ? CAP1 : CAP2;
    static constexpr int TBT = TBT_ > CAP ? (CAP < 1 ? 1 : CAP) : TBT_;
    // experimental wider tile (DFFT_WIDE_TILES=1): double width if it still fits 128 KB / 1024 threads
    static constexpr int TBT_WIDE = (2 * TBT <= CAP && 2 * TBT <= 32) ? 2 * TBT : TBT;
};

// Resident CTAs per SM the register allocator must leave room for: 128 registers per thread for f64 and 64 for
// f32 (16 points per thread) — without it the grid-stride version of the f32 kernels drifted to 90+ registers and
// lost its second CTA per SM (tiled y pass 5300 -> 3850 GB/s).
template <typename T, int LOG2E = 4>
constexpr int min_ctas_per_sm(int threads) {
    constexpr int regs = (sizeof(T) == 8 ? 128 : 64) >> (LOG2E >= 4 ? 0 : 1);  // 8 points per thread: half the budget
    return (65536 / (threads * regs)) < 1 ? 1 : (65536 / (threads * regs));
}

template <typename T>
__device__ __forceinline__ cx<T> ld_elem(const cx<T>* p) {
    if constexpr (sizeof(T) == 8) {
        double2 r = *reinterpret_cast<const double2*>(p);
        return cx<T>{r.x, r.y};
    } else {
        float2 r = *reinterpret_cast<const float2*>(p);
        return cx<T>{r.x, r.y};
    }
}
template <typename T>
__device__ __forceinline__ void st_elem(cx<T>* p, cx<T> v) {
    if constexpr (sizeof(T) == 8) *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y);
    else *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y);
}

// ---- per-thread access to one line of a view ----------------------------------------------------------
// LINES = lines per CTA that need their own segment table (TB for CONTIG, 1 for TILED).
template <typename T, int LINES>
struct LineAccess {
    static constexpr int TABLE_ELEMS = LINES * MAXSEG;  // unsigned long long entries in shared memory
    const View& v;
    cx<T>* p0;                  // single segment: element n = 0 of this thread's line (b included)
    const unsigned long long* tab;  // multi segment: tab[s] = address of element n = 0 of segment s (b excluded)
    long long boff;             // multi segment: + b (elements)
    bool multi;

    // fills the shared-memory table (all threads of the CTA call this; a __syncthreads follows outside)
    static __device__ __forceinline__ void fill_table(const View& v, unsigned long long* table, int line_in_cta, int lane_in_line,
                                                      int lanes_per_line, int a0, int a1) {
        if (v.nseg <= 1) return;
        for (int s = lane_in_line; s < v.nseg; s += lanes_per_line) {
            const Seg& g = v.seg[s];
            cx<T>* p = reinterpret_cast<cx<T>*>(g.base) + (a0 * g.sA0 + a1 * g.sA1 - (long long)g.n0 * v.sN);
            table[line_in_cta * MAXSEG + s] = reinterpret_cast<unsigned long long>(p);
        }
    }
    __device__ __forceinline__ LineAccess(const View& v_, const unsigned long long* table, int line_in_cta, int a0, int a1, int b)
        : v(v_) {
        multi = v.nseg > 1;
        const Seg& g = v.seg[0];
        p0 = reinterpret_cast<cx<T>*>(g.base) + (a0 * g.sA0 + a1 * g.sA1 - (long long)g.n0 * v.sN + b);
        tab = table + line_in_cta * MAXSEG;
        boff = b;
    }
    __device__ __forceinline__ cx<T>* at(int n) const {
        if (!multi) return p0 + (long long)n * v.sN;
        const int s = v.seg_of_n[n];
        return reinterpret_cast<cx<T>*>(tab[s]) + ((long long)n * v.sN + boff);
    }
    // unit stride along n (every CONTIG view): no multiply
    __device__ __forceinline__ cx<T>* at1(int n) const {
        if (!multi) return p0 + n;
        const int s = v.seg_of_n[n];
        return reinterpret_cast<cx<T>*>(tab[s]) + (n + boff);
    }
};

// ---- the in-CTA transform ---------------------------------------------------------------------------
template <typename T, int LOG2N, int LOG2E, int TB, bool TILED>
struct CtaFft {
    using Core = FftCore<T, LOG2N, LOG2E>;
    using L = SmemLayout<LOG2N, LOG2E, TB, TILED, int(sizeof(cx<T>))>;
    static constexpr int N = Core::N, E = Core::E, TPL = Core::TPL, NST = Core::NST;
    static constexpr int THREADS = TPL * TB;
    static constexpr int LINES = TILED ? 1 : TB;
    static constexpr size_t TILE_BYTES = size_t(L::ELEMS) * sizeof(cx<T>);
    static constexpr size_t TABLE_BYTES = size_t(2) * LINES * MAXSEG * sizeof(unsigned long long) + 16;  // + one mbarrier (TMA-fed kernel)
    static constexpr size_t SMEM_BYTES = TILE_BYTES + TABLE_BYTES;
    static_assert(THREADS >= 1 && THREADS <= 1024, "CTA size out of range");
    static_assert(SMEM_BYTES <= 227 * 1024, "tile does not fit the 227 KB of shared memory a CTA can use on sm_100a");

    static __device__ __forceinline__ int sidx(int n, int t) { return L::idx(n, t); }

    // synchronise the threads that cooperate on one line (CONTIG) or the whole tile (TILED)
    static __device__ __forceinline__ void sync(int t) {
        if constexpr (TILED || TB == 1) {
            __syncthreads();
        } else if constexpr (TPL <= 32) {
            __syncwarp();
        } else if constexpr (TB <= 15) {
            asm volatile("bar.sync %0, %1;" ::"r"(t + 1), "n"(TPL) : "memory");
        } else {
            __syncthreads();
        }
    }

    // gather: v[e] = tile[j + e*TPL]
    static __device__ __forceinline__ void gather(cx<T> (&v)[E], const cx<T>* sm, int j, int t) {
        if constexpr (L::GATHER_SPLIT) {
            const cx<T>* p = sm + L::idx(j, t);
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = p[L::off(e * TPL)];
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = sm[L::idx(j + e * TPL, t)];
        }
    }

    template <int ST>
    static __device__ __forceinline__ void stages(cx<T> (&v)[E], int j, int t, cx<T>* sm, const cx<T>* tw) {
        Core::template stage_compute<ST>(v, j, tw);
        if constexpr (ST + 1 < NST) {
            if constexpr (ST > 0) sync(t);
            cx<T>* p = sm + L::idx(Core::template scatter_base<ST>(j), t);
#pragma unroll
            for (int e = 0; e < E; ++e) p[L::off(Core::template scatter_off<ST>(e))] = v[e];
            sync(t);
            gather(v, sm, j, t);
            stages<ST + 1>(v, j, t, sm, tw);
        }
    }
};

// linear tile number -> (b tile, a1, a0) of a TILED pass
template <int TB>
__device__ __forceinline__ void tiled_decode(const FftParams& p, long long tile, int& bt, int& a1, int& a0) {
    const int nbt = (p.B + TB - 1) / TB;
    bt = int(tile % nbt);
    const int a = int(tile / nbt);
    if (p.tile_swz > 0) {
        const int lg = p.tile_swz, G = 1 << lg;
        const int lo = a & (G * G - 1), hi = a >> (2 * lg);
        const int w1 = p.A1 >> lg;  // A0, A1 are multiples of G (checked by the launcher)
        a1 = ((hi % w1) << lg) + (lo & (G - 1));
        a0 = ((hi / w1) << lg) + (lo >> lg);
    } else {
        a1 = a % p.A1;
        a0 = a / p.A1;
    }
}

// common prologue: thread -> (j, t), tile -> (a0, a1, b), segment tables
template <typename C, bool TILED, int TB>
struct TileCoord {
    int j, t, a0, a1, b;
    bool valid;
    __device__ __forceinline__ TileCoord(const FftParams& p, long long tile = blockIdx.x) {
        const int tid = threadIdx.x;
        if constexpr (TILED) { t = tid % TB; j = tid / TB; }
        else { j = tid % C::TPL; t = tid / C::TPL; }
        if constexpr (TILED) {
            int bt;
            tiled_decode<TB>(p, tile, bt, a1, a0);
            b = bt * TB + t;
            valid = b < p.B;
        } else {
            const long long line = tile * TB + t;
            valid = line < (long long)p.A0 * p.A1;
            a1 = int(line % p.A1);
            a0 = int(line / p.A1);
            b = 0;
        }
    }
    static __device__ __forceinline__ long long num_tiles(const FftParams& p) {
        if constexpr (TILED) return (long long)p.A0 * p.A1 * ((p.B + TB - 1) / TB);
        else return ((long long)p.A0 * p.A1 + TB - 1) / TB;
    }
};

// ---- C2C pass ------------------------------------------------------------------------------------------
// Grid-stride over tiles: launched with one CTA per tile it is the plain kernel; launched with fewer CTAs
// (FftParams::max_ctas) it is persistent and occupies only that many CTA slots, which is how the overlapped
// schedule leaves the rest of the GPU to the other passes.
template <typename T, int LOG2N, int LOG2E, int TB, bool TILED, bool INV>
__global__ void __launch_bounds__((1 << (LOG2N - LOG2E)) * TB, min_ctas_per_sm<T, LOG2E>((1 << (LOG2N - LOG2E)) * TB))
fft_c2c_kernel(const __grid_constant__ FftParams p) {
    using C = CtaFft<T, LOG2N, LOG2E, TB, TILED>;
    using LA = LineAccess<T, C::LINES>;
    using TC = TileCoord<C, TILED, TB>;
    constexpr int E = C::E, TPL = C::TPL;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    cx<T>* sm = reinterpret_cast<cx<T>*>(smem_raw);
    unsigned long long* tab_in = reinterpret_cast<unsigned long long*>(smem_raw + C::TILE_BYTES);
    unsigned long long* tab_out = tab_in + C::LINES * MAXSEG;
    const bool multi = p.in.nseg > 1 || p.out.nseg > 1;
    const long long ntiles = TC::num_tiles(p);

#pragma unroll 1
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TC tc(p, tile);
        const int j = tc.j, t = tc.t;
        if (multi) {
            // TILED: one table per CTA (a0, a1 are CTA-uniform), filled by the first threads;
            // CONTIG: one table per line, filled by that line's threads.
            const int line = TILED ? 0 : t;
            const int lane = TILED ? int(threadIdx.x) : j;
            const int lanes = TILED ? C::THREADS : TPL;
            LA::fill_table(p.in, tab_in, line, lane, lanes, tc.a0, tc.a1);
            LA::fill_table(p.out, tab_out, line, lane, lanes, tc.a0, tc.a1);
            __syncthreads();
        }
        const LA in(p.in, tab_in, TILED ? 0 : t, tc.a0, tc.a1, tc.b);
        const LA out(p.out, tab_out, TILED ? 0 : t, tc.a0, tc.a1, tc.b);

        cx<T> v[E];
        if (tc.valid) {
            if (!in.multi) {
                if constexpr (!TILED) {
                    // contiguous line: sN == 1, immediate offsets
                    const cx<T>* q = in.p0 + j;
#pragma unroll
                    for (int e = 0; e < E; ++e) v[e] = ld_elem<T>(q + e * TPL);
                } else {
                    const long long step = (long long)TPL * p.in.sN;
                    const cx<T>* q = in.p0 + (long long)j * p.in.sN;
#pragma unroll
                    for (int e = 0; e < E; ++e) { v[e] = ld_elem<T>(q); q += step; }
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = ld_elem<T>(TILED ? in.at(j + e * TPL) : in.at1(j + e * TPL));
            }
            if constexpr (INV) {
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = cswap(v[e]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = cx<T>{T(0), T(0)};
        }

        C::template stages<0>(v, j, t, sm, reinterpret_cast<const cx<T>*>(p.tw));

        if (tc.valid) {
            if (!out.multi) {
                if constexpr (!TILED) {
                    cx<T>* q = out.p0 + j;
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const cx<T> x = v[C::Core::final_slot(e)];
                        st_elem<T>(q + e * TPL, INV ? cswap(x) : x);
                    }
                } else {
                    const long long step = (long long)TPL * p.out.sN;
                    cx<T>* q = out.p0 + (long long)j * p.out.sN;
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const cx<T> x = v[C::Core::final_slot(e)];
                        st_elem<T>(q, INV ? cswap(x) : x);
                        q += step;
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const cx<T> x = v[C::Core::final_slot(e)];
                    st_elem<T>(TILED ? out.at(j + e * TPL) : out.at1(j + e * TPL), INV ? cswap(x) : x);
                }
            }
        }
        if (tile + gridDim.x < ntiles) {
            // persistent use: the next tile's first scatter must not overtake this tile's last gather, and
            // the segment tables are rewritten
            if (multi) __syncthreads();
            else if constexpr (C::NST > 1) C::sync(t);
        }
    }
}

// ---- C2C TILED pass fed by TMA ----------------------------------------------------------------------------
// Same tile, same stages, but the N x TB input tile is fetched by the TMA engine (cp.async.bulk.tensor, SASS UTMALDG)
// straight into the tile buffer and announced through an mbarrier, instead of 16 LDG.128 per thread:
//   * the loads no longer pass through the LSU / L1TEX data pipe — the busiest unit of the register-fed kernel
//     (ncu: 63-74 % of peak, about half of its wavefronts are the 64-byte-row global accesses);
//   * no registers are tied to bytes in flight, so the kernel is persistent and the NEXT tile is requested as soon
//     as the last gather of the current tile has left the buffer: it lands while the last radix stage and the stores
//     run (and while the second resident CTA computes).
// Input view: single segment; described by a rank-4 tensor map (b, n, a1, a0) built by the launcher, box TB x 256.
// Output: either 16-byte stores from registers through any (segmented) view, or — `bulk_out`, blocked hand-over
// layouts whose rows per destination are adjacent — one cp.async.bulk (UBLKCP) per destination from the buffer.
struct TmaBar {
    static __device__ __forceinline__ void init(unsigned long long* bar, unsigned count) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(unsigned(__cvta_generic_to_shared(bar))), "r"(count) : "memory");
    }
    static __device__ __forceinline__ void expect(unsigned long long* bar, unsigned bytes) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(unsigned(__cvta_generic_to_shared(bar))), "r"(bytes) : "memory");
    }
    // bounded wait: a lost transaction must not hang the GPU (returns false after ~2 s)
    static __device__ __forceinline__ bool wait(unsigned long long* bar, unsigned parity) {
        const unsigned a = unsigned(__cvta_generic_to_shared(bar));
        for (long long spin = 0; spin < (1ll << 26); ++spin) {
            unsigned ok;
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok) : "r"(a), "r"(parity) : "memory");
            if (ok) return true;
        }
        return false;
    }
};

template <typename T, int LOG2N, int LOG2E, int TB, bool INV>
__global__ void __launch_bounds__((1 << (LOG2N - LOG2E)) * TB, min_ctas_per_sm<T, LOG2E>((1 << (LOG2N - LOG2E)) * TB))
fft_c2c_tma_kernel(const __grid_constant__ FftParams p, const __grid_constant__ CUtensorMap tm_in) {
    using C = CtaFft<T, LOG2N, LOG2E, TB, true>;
    using Core = typename C::Core;
    using L = typename C::L;
    using LA = LineAccess<T, 1>;
    using TC = TileCoord<C, true, TB>;
    constexpr int E = C::E, TPL = C::TPL, N = C::N, NST = C::NST;
    static_assert(NST >= 2, "the TMA-fed kernel is for lines that exchange through shared memory");
    constexpr unsigned TILE_TX = unsigned(N) * TB * unsigned(sizeof(cx<T>));
    constexpr int ROWS_PER_OP = N < 256 ? N : 256;  // box height (a TMA box dimension is at most 256)
    constexpr int D0 = sizeof(T) == 8 ? 2 : 1;      // inner-dimension units per complex element (8-byte units)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    cx<T>* sm = reinterpret_cast<cx<T>*>(smem_raw);
    unsigned long long* tab_in = reinterpret_cast<unsigned long long*>(smem_raw + C::TILE_BYTES);
    unsigned long long* tab_out = tab_in + MAXSEG;
    unsigned long long* bar = tab_out + MAXSEG;
    const long long ntiles = TC::num_tiles(p);
    const bool out_multi = p.out.nseg > 1;
    const bool bulk = p.bulk_out != 0;
    const int tid = threadIdx.x;

    auto request = [&](long long tile) {  // one thread: arm the barrier and ask the TMA engine for the tile
        const TC tc(p, tile);
        const int b0 = (tc.b - tc.t) * D0;
        TmaBar::expect(bar, TILE_TX);
        const unsigned dst = unsigned(__cvta_generic_to_shared(sm));
        const unsigned mb = unsigned(__cvta_generic_to_shared(bar));
#pragma unroll
        for (int r = 0; r < N; r += ROWS_PER_OP)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(dst + unsigned(r) * TB * unsigned(sizeof(cx<T>))), "l"(&tm_in), "r"(mb), "r"(b0), "r"(r), "r"(tc.a1), "r"(tc.a0)
                         : "memory");
    };

    if (tid == 0) {
        TmaBar::init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && (long long)blockIdx.x < ntiles) request(blockIdx.x);
    unsigned parity = 0;

#pragma unroll 1
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TC tc(p, tile);
        const int j = tc.j, t = tc.t;
        const long long next = tile + gridDim.x;
        if (out_multi && !bulk) LA::fill_table(p.out, tab_out, 0, tid, C::THREADS, tc.a0, tc.a1);
        if (bulk)
            for (int s = tid; s < p.out.nseg; s += C::THREADS) {
                const Seg& g = p.out.seg[s];
                cx<T>* q = reinterpret_cast<cx<T>*>(g.base) + (tc.a0 * g.sA0 + tc.a1 * g.sA1 - (long long)g.n0 * p.out.sN);
                tab_out[s] = reinterpret_cast<unsigned long long>(q);
            }
        if (!TmaBar::wait(bar, parity)) __trap();
        parity ^= 1;

        cx<T> v[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const cx<T> x = sm[(j + e * TPL) * TB + t];  // dense [n][TB], as the TMA engine wrote it
            v[e] = INV ? cswap(x) : x;
        }
        __syncthreads();  // everyone has its inputs: the buffer turns into the (padded) exchange space; tables visible

        // the stages of CtaFft::stages, with a hook where the buffer is free again (after the last gather)
        auto run = [&](auto self, auto st_tag) -> void {
            constexpr int ST = decltype(st_tag)::value;
            Core::template stage_compute<ST>(v, j, reinterpret_cast<const cx<T>*>(p.tw));
            if constexpr (ST + 1 < NST) {
                if constexpr (ST > 0) __syncthreads();
                cx<T>* q = sm + L::idx(Core::template scatter_base<ST>(j), t);
#pragma unroll
                for (int e = 0; e < E; ++e) q[L::off(Core::template scatter_off<ST>(e))] = v[e];
                __syncthreads();
                C::gather(v, sm, j, t);
                if constexpr (ST + 2 == NST) {
                    if (!bulk) {
                        // generic-proxy reads of the buffer are ordered before the async-proxy write of the next tile
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncthreads();
                        if (tid == 0 && next < ntiles) request(next);
                    }
                }
                self(self, std::integral_constant<int, ST + 1>{});
            }
        };
        run(run, std::integral_constant<int, 0>{});

        if (!bulk) {
            if (tc.valid) {
                const LA out(p.out, tab_out, 0, tc.a0, tc.a1, tc.b);
                if (!out.multi) {
                    const long long step = (long long)TPL * p.out.sN;
                    cx<T>* q = out.p0 + (long long)j * p.out.sN;
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const cx<T> x = v[Core::final_slot(e)];
                        st_elem<T>(q, INV ? cswap(x) : x);
                        q += step;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const cx<T> x = v[Core::final_slot(e)];
                        st_elem<T>(out.at(j + e * TPL), INV ? cswap(x) : x);
                    }
                }
            }
            if (out_multi) __syncthreads();  // the segment table is rewritten by the next tile
        } else {
            __syncthreads();  // last gather finished everywhere: stage the finished tile, natural order [n][TB]
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const cx<T> x = v[Core::final_slot(e)];
                sm[(j + e * TPL) * TB + t] = INV ? cswap(x) : x;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                for (int s = 0; s < p.out.nseg; ++s) {
                    const int n0 = p.out.nseg > 1 ? p.out.seg[s].n0 : 0;
                    const int n1 = (s + 1 < p.out.nseg) ? p.out.seg[s + 1].n0 : N;
                    const unsigned bytes = unsigned(n1 - n0) * TB * unsigned(sizeof(cx<T>));
                    if (!bytes) continue;
                    const unsigned src = unsigned(__cvta_generic_to_shared(sm + n0 * TB));
                    const unsigned long long dst = tab_out[s] + (unsigned long long)n0 * TB * sizeof(cx<T>);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the buffer has been read: it may be refilled
                if (next < ntiles) request(next);
            }
            __syncthreads();  // tab_out is rewritten by the next tile
        }
    }
    if (bulk && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all bulk stores performed before exit
}


// complex value from another lane of the warp
template <typename T>
__device__ __forceinline__ cx<T> shfl_cx(cx<T> v, int src_lane) {
    return cx<T>{__shfl_sync(0xffffffffu, v.x, src_lane), __shfl_sync(0xffffffffu, v.y, src_lane)};
}

// ---- R2C pass (CONTIG): real line of 2M points -> M+1 complex points ------------------------------------
// The real line is read as M complex points z[m] = x[2m] + i x[2m+1], transformed with the length-M
// core and split into even/odd spectra in shared memory:  X[k] = Xe[k] + W_2M^k Xo[k].
template <typename T, int LOG2M, int LOG2E, int TB>
__global__ void __launch_bounds__((1 << (LOG2M - LOG2E)) * TB, min_ctas_per_sm<T, LOG2E>((1 << (LOG2M - LOG2E)) * TB))
fft_r2c_kernel(const __grid_constant__ FftParams p) {
    using C = CtaFft<T, LOG2M, LOG2E, TB, false>;
    using LA = LineAccess<T, C::LINES>;
    constexpr int E = C::E, TPL = C::TPL, M = C::N, NST = C::NST;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    cx<T>* sm = reinterpret_cast<cx<T>*>(smem_raw);
    unsigned long long* tab_in = reinterpret_cast<unsigned long long*>(smem_raw + C::TILE_BYTES);
    unsigned long long* tab_out = tab_in + C::LINES * MAXSEG;
    const TileCoord<C, false, TB> tc(p);
    const int j = tc.j, t = tc.t;
    const bool valid = tc.valid;
    if (p.out.nseg > 1) {
        LA::fill_table(p.out, tab_out, t, j, TPL, tc.a0, tc.a1);
        __syncthreads();
    }
    const LA in(p.in, tab_in, t, tc.a0, tc.a1, 0);
    const LA out(p.out, tab_out, t, tc.a0, tc.a1, 0);

    cx<T> v[E];
    if (valid) {
        const cx<T>* q = in.p0 + j;
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = ld_elem<T>(q + e * TPL);
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = cx<T>{T(0), T(0)};
    }
    C::template stages<0>(v, j, t, sm, reinterpret_cast<const cx<T>*>(p.tw));

    const cx<T>* tw2 = reinterpret_cast<const cx<T>*>(p.tw2);
    // (A shuffle-based split like fft_c2r_kernel's was measured for this direction too: 5066 vs 5138 GB/s at 1024 real
    // points — the 16 extra twiddle multiplies cost what the shared-memory round trip saves — so the forward kernel
    // keeps the shared-memory split.)
    // Z in natural order -> shared memory
    if constexpr (NST > 1) C::sync(t);
#pragma unroll
    for (int e = 0; e < E; ++e) sm[C::sidx(j + e * TPL, t)] = v[C::Core::final_slot(e)];
    C::sync(t);

    auto emit = [&](int k) {
        const int kk = (M - k) & (M - 1);
        const cx<T> zk = sm[C::sidx(k, t)], zp = cconj(sm[C::sidx(kk, t)]);
        const cx<T> xe = cx<T>{T(0.5) * (zk.x + zp.x), T(0.5) * (zk.y + zp.y)};
        const cx<T> d = csub(zk, zp);
        const cx<T> xo = cx<T>{T(0.5) * d.y, T(-0.5) * d.x};  // -i/2 * d
        const cx<T> tt = cmul(ld_tw(tw2, k), xo);
        if (valid) {
            st_elem<T>(out.at1(k), cadd(xe, tt));
            st_elem<T>(out.at1(M - k), cconj(csub(xe, tt)));
        }
    };
#pragma unroll
    for (int e = 0; e < E / 2; ++e) emit(j + e * TPL);
    if (j == 0) emit(M / 2);
}

// ---- C2R pass (CONTIG): M+1 complex points -> real line of 2M points, unnormalised -----------------------
template <typename T, int LOG2M, int LOG2E, int TB>
__global__ void __launch_bounds__((1 << (LOG2M - LOG2E)) * TB, min_ctas_per_sm<T, LOG2E>((1 << (LOG2M - LOG2E)) * TB))
fft_c2r_kernel(const __grid_constant__ FftParams p) {
    using C = CtaFft<T, LOG2M, LOG2E, TB, false>;
    using LA = LineAccess<T, C::LINES>;
    constexpr int E = C::E, TPL = C::TPL, M = C::N;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    cx<T>* sm = reinterpret_cast<cx<T>*>(smem_raw);
    unsigned long long* tab_in = reinterpret_cast<unsigned long long*>(smem_raw + C::TILE_BYTES);
    unsigned long long* tab_out = tab_in + C::LINES * MAXSEG;
    const TileCoord<C, false, TB> tc(p);
    const int j = tc.j, t = tc.t;
    const bool valid = tc.valid;
    if (p.in.nseg > 1) {
        LA::fill_table(p.in, tab_in, t, j, TPL, tc.a0, tc.a1);
        __syncthreads();
    }
    const LA in(p.in, tab_in, t, tc.a0, tc.a1, 0);
    const LA out(p.out, tab_out, t, tc.a0, tc.a1, 0);

    const cx<T>* tw2 = reinterpret_cast<const cx<T>*>(p.tw2);
    cx<T> v[E];
    if constexpr (TPL <= 32) {
        // The line lives in one warp (TPL lanes): thread j loads X[j + e*TPL] (coalesced), fetches the partners X[M - k] from lane TPL - j
        // (slot E-1-e; lane 0: its own slot E-e, and X[M] for k = 0) with one shuffle each, and builds
        //   Z[k] = (X[k] + conj X[M-k]) + i conj(W_2M^k) (X[k] - conj X[M-k])
        // directly in the register layout the first stage expects — no staging through shared memory.
        cx<T> xk[E];
        cx<T> xM{T(0), T(0)};
        if (valid) {
#pragma unroll
            for (int e = 0; e < E; ++e) xk[e] = ld_elem<T>(in.at1(j + e * TPL));
            if (j == 0) xM = ld_elem<T>(in.at1(M));
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) xk[e] = cx<T>{T(0), T(0)};
        }
        const int lane = threadIdx.x & 31;
        const int src = (lane & ~(TPL - 1)) + ((TPL - j) & (TPL - 1));
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const cx<T> other = shfl_cx<T>(xk[E - 1 - e], src);
            const cx<T> self = e == 0 ? xM : xk[(E - e) & (E - 1)];
            const cx<T> xm = cconj(j == 0 ? self : other);
            const int k = j + e * TPL;
            cx<T> wc;  // conj(W_2M^k)
            if (2 * k <= M) wc = cconj(ld_tw(tw2, k));
            else { const cx<T> u = ld_tw(tw2, M - k); wc = cx<T>{-u.x, -u.y}; }
            const cx<T> xe = cadd(xk[e], xm);
            const cx<T> xo = cmul(wc, csub(xk[e], xm));
            // stored with re/im swapped: the inverse transform is run as swap(fwd(swap(.)))
            v[e] = cswap(cx<T>{xe.x - xo.y, xe.y + xo.x});
        }
    } else {
    // Z[k] = (X[k] + conj X[M-k]) + i * conj(W^k) * (X[k] - conj X[M-k]);  Z[M-k] = conj(Xe' - i Xo')
    auto build = [&](int k) {
        cx<T> xk{T(0), T(0)}, xm{T(0), T(0)};
        if (valid) {
            xk = ld_elem<T>(in.at1(k));
            xm = cconj(ld_elem<T>(in.at1(M - k)));
        }
        const cx<T> xe = cadd(xk, xm);
        const cx<T> xo = cmul(cconj(ld_tw(tw2, k)), csub(xk, xm));
        // stored with re/im swapped: the inverse transform is run as swap(fwd(swap(.)))
        sm[C::sidx(k, t)] = cswap(cx<T>{xe.x - xo.y, xe.y + xo.x});
        if (k != 0) sm[C::sidx(M - k, t)] = cswap(cx<T>{xe.x + xo.y, -(xe.y - xo.x)});
    };
#pragma unroll
    for (int e = 0; e < E / 2; ++e) build(j + e * TPL);
    if (j == 0) build(M / 2);
    C::sync(t);
    C::gather(v, sm, j, t);
    if constexpr (C::NST > 1) C::sync(t);
    }
    C::template stages<0>(v, j, t, sm, reinterpret_cast<const cx<T>*>(p.tw));
    if (valid) {
        cx<T>* q = out.p0 + j;
#pragma unroll
        for (int e = 0; e < E; ++e) st_elem<T>(q + e * TPL, cswap(v[C::Core::final_slot(e)]));
    }
}

// ---- host launchers (one explicit instantiation per precision and size, see fft_inst.cu) ----------------
template <typename T, int LOG2N>
cudaError_t launch_pass(PassKind kind, const FftParams& p, cudaStream_t stream);

// launch_pass for a runtime size; returns cudaErrorInvalidValue for unsupported lengths.
cudaError_t launch_pass_f64(int log2n, PassKind kind, const FftParams& p, cudaStream_t stream);
cudaError_t launch_pass_f32(int log2n, PassKind kind, const FftParams& p, cudaStream_t stream);

constexpr int MAX_LOG2N = 13;

}  // namespace dfft
