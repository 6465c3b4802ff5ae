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
#pragma once
#include <cuda_runtime.h>

#include "fft_core.cuh"

namespace dfft {

constexpr int MAXSEG = 16;

struct Seg {
    void* base;        // element pointer (complex elements of the pass' precision)
    long long sA0;     // stride of batch index a0 (elements)
    long long sA1;     // stride of batch index a1
    long long sN;      // stride along the transformed axis
    int n0;            // first n covered by this segment
    int pad_;
};

struct View {
    const unsigned char* seg_of_n;  // n -> segment id (device memory); ignored when nseg == 1
    int nseg;
    int pad_;
    Seg seg[MAXSEG];
};

struct FftParams {
    View in, out;
    int A0, A1;        // batch extents: line id = a0*A1 + a1
    int B;             // TILED: extent of the contiguous dimension; CONTIG: 1
    int inverse;       // 0 forward (e^-), 1 inverse (e^+), both unnormalised like cuFFT
    const void* tw;    // exp(-2*pi*i*m/N), m < N, N = pass length (complex length for R2C/C2R)
    const void* tw2;   // R2C/C2R only: exp(-2*pi*i*k/(2N)), k <= N/2
};

enum PassKind { PASS_C2C_CONTIG = 0, PASS_C2C_TILED = 1, PASS_R2C = 2, PASS_C2R = 3 };

// ---- tile shape choices (compile time) -----------------------------------------------------------
template <typename T, int LOG2N>
struct Shape {
    static constexpr int LOG2E = LOG2N < 4 ? LOG2N : 4;
    static constexpr int N = 1 << LOG2N;
    static constexpr int TPL = N >> LOG2E;
    // CONTIG: lines per CTA, aim at 256 threads
    static constexpr int TBC = (256 / TPL) < 1 ? 1 : ((256 / TPL) > 64 ? 64 : (256 / TPL));
    // TILED: tile width; rows of >= 64 bytes where shared memory allows, 4096 (f64) / 8192 (f32) points
    static constexpr int MINROW = 64 / int(2 * sizeof(T));
    static constexpr int WANT = (sizeof(T) == 8 ? 4096 : 8192) / N;
    static constexpr int TBT_ = WANT < MINROW ? MINROW : (WANT > 32 ? 32 : WANT);
    // keep the tile within 128 KB of shared memory and 1024 threads
    static constexpr int CAP1 = (128 * 1024) / (N * int(2 * sizeof(T)));
    static constexpr int CAP2 = 1024 / TPL;
    static constexpr int CAP = CAP1 < CAP2 ? CAP1 : CAP2;
    static constexpr int TBT = TBT_ > CAP ? (CAP < 1 ? 1 : CAP) : TBT_;
};

// ---- addressing ------------------------------------------------------------------------------------
template <typename T>
struct Addr {
    // pointer to element (a0,a1,n,b) of a view
    static __device__ __forceinline__ cx<T>* at(const View& v, int a0, int a1, int n, int b) {
        int s = 0;
        if (v.nseg > 1) s = v.seg_of_n[n];
        const Seg& g = v.seg[s];
        return reinterpret_cast<cx<T>*>(g.base) + (a0 * g.sA0 + a1 * g.sA1 + (long long)(n - g.n0) * g.sN + b);
    }
};

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

// ---- the in-CTA transform ---------------------------------------------------------------------------
template <typename T, int LOG2N, int LOG2E, int TB, bool TILED>
struct CtaFft {
    using Core = FftCore<T, LOG2N, LOG2E>;
    static constexpr int N = Core::N, E = Core::E, TPL = Core::TPL, NST = Core::NST;
    static constexpr int THREADS = TPL * TB;
    static constexpr int ROWB = TB * int(sizeof(cx<T>));
    static constexpr int PADSH = StagePlan<LOG2N, LOG2E>::bits(0) < 3 ? 3 : StagePlan<LOG2N, LOG2E>::bits(0);
    static constexpr bool PAD = TILED ? (ROWB < 128) : true;
    static constexpr int PADSH_T = StagePlan<LOG2N, LOG2E>::bits(0);
    static constexpr int NPAD = TILED ? (PAD ? N + (N >> PADSH_T) : N) : (N + (N >> PADSH));
    static constexpr size_t SMEM_BYTES = (NST > 1 || true) ? size_t(NPAD) * TB * sizeof(cx<T>) : 0;

    static __device__ __forceinline__ int sidx(int n, int t) {
        if constexpr (TILED) {
            if constexpr (PAD) return (n + (n >> PADSH_T)) * TB + t;
            else return n * TB + t;
        } else {
            return t * NPAD + n + (n >> PADSH);
        }
    }

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

    template <int ST>
    static __device__ __forceinline__ void stages(cx<T> (&v)[E], int j, int t, cx<T>* sm, const cx<T>* tw) {
        Core::template stage_compute<ST>(v, j, tw);
        if constexpr (ST + 1 < NST) {
            if constexpr (ST > 0) sync(t);
#pragma unroll
            for (int e = 0; e < E; ++e) sm[sidx(Core::template scatter_pos<ST>(j, e), t)] = v[e];
            sync(t);
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = sm[sidx(j + e * TPL, t)];
            stages<ST + 1>(v, j, t, sm, tw);
        }
    }
};

// ---- C2C pass ------------------------------------------------------------------------------------------
template <typename T, int LOG2N, int LOG2E, int TB, bool TILED>
__global__ void __launch_bounds__((1 << (LOG2N - LOG2E)) * TB)
fft_c2c_kernel(const __grid_constant__ FftParams p) {
    using C = CtaFft<T, LOG2N, LOG2E, TB, TILED>;
    constexpr int E = C::E, TPL = C::TPL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T>* sm = reinterpret_cast<cx<T>*>(smem_raw);

    const int tid = threadIdx.x;
    int j, t;
    if constexpr (TILED) { t = tid % TB; j = tid / TB; }
    else { j = tid % TPL; t = tid / TPL; }

    int a0, a1, b;
    bool valid;
    if constexpr (TILED) {
        const int nbt = (p.B + TB - 1) / TB;
        const int bt = blockIdx.x % nbt;
        const int a = blockIdx.x / nbt;
        b = bt * TB + t;
        valid = b < p.B;
        a1 = a % p.A1;
        a0 = a / p.A1;
    } else {
        const long long line = (long long)blockIdx.x * TB + t;
        valid = line < (long long)p.A0 * p.A1;
        a1 = int(line % p.A1);
        a0 = int(line / p.A1);
        b = 0;
    }

    cx<T> v[E];
    if (valid) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            cx<T> x = ld_elem<T>(Addr<T>::at(p.in, a0, a1, j + e * TPL, b));
            v[e] = p.inverse ? cswap(x) : x;
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = cx<T>{T(0), T(0)};
    }

    C::template stages<0>(v, j, t, sm, reinterpret_cast<const cx<T>*>(p.tw));

    if (valid) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            cx<T> x = v[C::Core::final_slot(e)];
            st_elem<T>(Addr<T>::at(p.out, a0, a1, j + e * TPL, b), p.inverse ? cswap(x) : x);
        }
    }
}

// ---- R2C pass (CONTIG): real line of 2M points -> M+1 complex points ------------------------------------
// The real line is read as M complex points z[m] = x[2m] + i x[2m+1], transformed with the length-M
// core and split into even/odd spectra in shared memory:  X[k] = Xe[k] + W_2M^k Xo[k].
template <typename T, int LOG2M, int LOG2E, int TB>
__global__ void __launch_bounds__((1 << (LOG2M - LOG2E)) * TB)
fft_r2c_kernel(const __grid_constant__ FftParams p) {
    using C = CtaFft<T, LOG2M, LOG2E, TB, false>;
    constexpr int E = C::E, TPL = C::TPL, M = C::N, NST = C::NST;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T>* sm = reinterpret_cast<cx<T>*>(smem_raw);
    const int tid = threadIdx.x;
    const int j = tid % TPL, t = tid / TPL;
    const long long line = (long long)blockIdx.x * TB + t;
    const bool valid = line < (long long)p.A0 * p.A1;
    const int a1 = int(line % p.A1), a0 = int(line / p.A1);

    cx<T> v[E];
    if (valid) {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = ld_elem<T>(Addr<T>::at(p.in, a0, a1, j + e * TPL, 0));
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = cx<T>{T(0), T(0)};
    }
    C::template stages<0>(v, j, t, sm, reinterpret_cast<const cx<T>*>(p.tw));

    // Z in natural order -> shared memory
    if constexpr (NST > 1) C::sync(t);
#pragma unroll
    for (int e = 0; e < E; ++e) sm[C::sidx(j + e * TPL, t)] = v[C::Core::final_slot(e)];
    C::sync(t);

    const cx<T>* tw2 = reinterpret_cast<const cx<T>*>(p.tw2);
    auto emit = [&](int k) {
        const int kk = (M - k) & (M - 1);
        const cx<T> zk = sm[C::sidx(k, t)], zp = cconj(sm[C::sidx(kk, t)]);
        const cx<T> xe = cx<T>{T(0.5) * (zk.x + zp.x), T(0.5) * (zk.y + zp.y)};
        const cx<T> d = csub(zk, zp);
        const cx<T> xo = cx<T>{T(0.5) * d.y, T(-0.5) * d.x};  // -i/2 * d
        const cx<T> tt = cmul(ld_tw(tw2, k), xo);
        if (valid) {
            st_elem<T>(Addr<T>::at(p.out, a0, a1, k, 0), cadd(xe, tt));
            st_elem<T>(Addr<T>::at(p.out, a0, a1, M - k, 0), cconj(csub(xe, tt)));
        }
    };
    if constexpr (E >= 2) {
#pragma unroll
        for (int e = 0; e < E / 2; ++e) emit(j + e * TPL);
        if (j == 0) emit(M / 2);
    } else {
        emit(0);
    }
}

// ---- C2R pass (CONTIG): M+1 complex points -> real line of 2M points, unnormalised -----------------------
template <typename T, int LOG2M, int LOG2E, int TB>
__global__ void __launch_bounds__((1 << (LOG2M - LOG2E)) * TB)
fft_c2r_kernel(const __grid_constant__ FftParams p) {
    using C = CtaFft<T, LOG2M, LOG2E, TB, false>;
    constexpr int E = C::E, TPL = C::TPL, M = C::N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T>* sm = reinterpret_cast<cx<T>*>(smem_raw);
    const int tid = threadIdx.x;
    const int j = tid % TPL, t = tid / TPL;
    const long long line = (long long)blockIdx.x * TB + t;
    const bool valid = line < (long long)p.A0 * p.A1;
    const int a1 = int(line % p.A1), a0 = int(line / p.A1);

    const cx<T>* tw2 = reinterpret_cast<const cx<T>*>(p.tw2);
    // Z[k] = (X[k] + conj X[M-k]) + i * conj(W^k) * (X[k] - conj X[M-k]);  Z[M-k] = conj(Xe' - i Xo')
    auto build = [&](int k) {
        cx<T> xk{T(0), T(0)}, xm{T(0), T(0)};
        if (valid) {
            xk = ld_elem<T>(Addr<T>::at(p.in, a0, a1, k, 0));
            xm = cconj(ld_elem<T>(Addr<T>::at(p.in, a0, a1, M - k, 0)));
        }
        const cx<T> xe = cadd(xk, xm);
        const cx<T> xo = cmul(cconj(ld_tw(tw2, k)), csub(xk, xm));
        // stored with re/im swapped: the inverse transform is run as swap(fwd(swap(.)))
        sm[C::sidx(k, t)] = cswap(cx<T>{xe.x - xo.y, xe.y + xo.x});
        if (k != 0) sm[C::sidx(M - k, t)] = cswap(cx<T>{xe.x + xo.y, -(xe.y - xo.x)});
    };
    if constexpr (E >= 2) {
#pragma unroll
        for (int e = 0; e < E / 2; ++e) build(j + e * TPL);
        if (j == 0) build(M / 2);
    } else {
        build(0);
    }
    C::sync(t);
    cx<T> v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = sm[C::sidx(j + e * TPL, t)];
    if constexpr (C::NST > 1) C::sync(t);
    C::template stages<0>(v, j, t, sm, reinterpret_cast<const cx<T>*>(p.tw));
    if (valid) {
#pragma unroll
        for (int e = 0; e < E; ++e)
            st_elem<T>(Addr<T>::at(p.out, a0, a1, j + e * TPL, 0), cswap(v[C::Core::final_slot(e)]));
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
