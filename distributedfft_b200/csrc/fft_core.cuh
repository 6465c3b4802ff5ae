// fft_core.cuh — register-resident Stockham radix-2^k FFT core, shared by every sm_100a kernel.
//
// Replaces the closed cuFFT kernels behind the reference's plans
// (/root/reference/src/slab/default/mpicufft_slab.cpp:159-165,
//  /root/reference/src/pencil/mpicufft_pencil.cpp:156-200): every cufftExec* there becomes one
// launch of a kernel built from this core.
//
// Everything here is __host__ __device__ so that tests/emu_fft_core.cpp can run the very same index
// arithmetic on the CPU (one "thread" at a time, stage by stage) and compare it with a naive DFT.
//
// Algorithm (per line of N = 2^LOG2N points, E = 2^LOG2E points held per thread, TPL = N/E threads):
//   * the line is transformed in NST = ceil(LOG2N/LOG2E) Stockham stages of radix r_s = 2^bits(s);
//   * before every stage thread j holds x[j + e*TPL], e = 0..E-1 (slot e) — this is the global-load
//     pattern of the first stage (coalesced over j) and the shared-memory read pattern of later ones;
//   * a stage of radix r < E performs S = E/r independent butterflies per thread: butterfly b works on
//     slots b + q*S, q = 0..r-1, i.e. on Stockham butterfly index jb = j + b*TPL;
//   * butterflies are radix-2 DIF networks in registers, leaving natural output q in slot
//     b + bitrev_r(q)*S;
//   * stage s scatters output q of butterfly jb to n' = expand(jb) + q*Ns (Ns = prod of earlier
//     radices) through shared memory; the last stage's outputs land at jb + q*N/r, which is again slot
//     pattern j + e*TPL, so the global store is coalesced as well.
#pragma once

#if defined(__CUDACC__)
#define DFFT_HD __host__ __device__ __forceinline__
#define DFFT_HDC __host__ __device__ constexpr
#else
#define DFFT_HD inline
#define DFFT_HDC constexpr
#endif

namespace dfft {

template <typename T>
struct alignas(2 * sizeof(T)) cx {
    T x, y;
};

template <typename T>
DFFT_HD cx<T> cadd(cx<T> a, cx<T> b) { return cx<T>{a.x + b.x, a.y + b.y}; }
template <typename T>
DFFT_HD cx<T> csub(cx<T> a, cx<T> b) { return cx<T>{a.x - b.x, a.y - b.y}; }
template <typename T>
DFFT_HD cx<T> cmul(cx<T> a, cx<T> b) { return cx<T>{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
template <typename T>
DFFT_HD cx<T> cconj(cx<T> a) { return cx<T>{a.x, -a.y}; }
template <typename T>
DFFT_HD cx<T> cswap(cx<T> a) { return cx<T>{a.y, a.x}; }

DFFT_HDC int ilog2c(unsigned v) { return v <= 1 ? 0 : 1 + ilog2c(v >> 1); }
DFFT_HDC int bitrev_c(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}

// ---------------------------------------------------------------------------------------------
// Stage plan: how LOG2N bits are split over stages of at most LOG2E bits (smaller radices first, so
// that the twiddle-free first stage is the cheap one and later stages are full radix-E butterflies).
// ---------------------------------------------------------------------------------------------
template <int LOG2N, int LOG2E>
struct StagePlan {
    static_assert(LOG2N >= 1 && LOG2E >= 1 && LOG2E <= LOG2N, "bad stage plan");
    static constexpr int N = 1 << LOG2N;
    static constexpr int E = 1 << LOG2E;
    static constexpr int TPL = N / E;  // threads per line
    static constexpr int NST = (LOG2N + LOG2E - 1) / LOG2E;
    static DFFT_HDC int bits(int s) { return LOG2N / NST + ((s >= NST - LOG2N % NST) ? 1 : 0); }
    static DFFT_HDC int log2ns(int s) {
        int a = 0;
        for (int i = 0; i < s; ++i) a += bits(i);
        return a;
    }
};

// Constant twiddle W_len^q = exp(-2*pi*i*q/len) for len <= 16 (forward sign).
template <typename T>
DFFT_HD cx<T> const_twiddle16(int idx16) {
    // idx16 = q * (16/len) in [0,8)
    constexpr double c1 = 0.92387953251128675613;  // cos(pi/8)
    constexpr double s1 = 0.38268343236508977173;  // sin(pi/8)
    constexpr double h = 0.70710678118654752440;   // sqrt(1/2)
    switch (idx16) {
        case 0: return cx<T>{T(1), T(0)};
        case 1: return cx<T>{T(c1), T(-s1)};
        case 2: return cx<T>{T(h), T(-h)};
        case 3: return cx<T>{T(s1), T(-c1)};
        case 4: return cx<T>{T(0), T(-1)};
        case 5: return cx<T>{T(-s1), T(-c1)};
        case 6: return cx<T>{T(-h), T(-h)};
        default: return cx<T>{T(-c1), T(-s1)};
    }
}

// d * W_len^q with the cheap special cases spelled out (all indices are compile-time after unrolling).
template <int IDX16, typename T>
DFFT_HD cx<T> mul_const_tw(cx<T> d) {
    constexpr double h = 0.70710678118654752440;
    if (IDX16 == 0) return d;
    if (IDX16 == 4) return cx<T>{d.y, -d.x};                                   // * (-i)
    if (IDX16 == 2) return cx<T>{T(h) * (d.x + d.y), T(h) * (d.y - d.x)};      // * (1-i)/sqrt2
    if (IDX16 == 6) return cx<T>{T(h) * (d.y - d.x), T(-h) * (d.x + d.y)};     // * (-1-i)/sqrt2
    return cmul(d, const_twiddle16<T>(IDX16));
}

// One radix-2 DIF level over a block: template recursion keeps every index a compile-time constant so
// the slot array stays in registers.
template <int R, int S, int LEN, int BLK, int Q, typename T>
struct DifPair {
    static DFFT_HD void run(cx<T>* v) {
        constexpr int i0 = (BLK + Q) * S;
        constexpr int i1 = (BLK + Q + LEN / 2) * S;
        cx<T> a = v[i0], c = v[i1];
        v[i0] = cadd(a, c);
        v[i1] = mul_const_tw<Q*(16 / LEN)>(csub(a, c));
        if constexpr (Q + 1 < LEN / 2) DifPair<R, S, LEN, BLK, Q + 1, T>::run(v);
    }
};
template <int R, int S, int LEN, int BLK, typename T>
struct DifBlocks {
    static DFFT_HD void run(cx<T>* v) {
        DifPair<R, S, LEN, BLK, 0, T>::run(v);
        if constexpr (BLK + LEN < R) DifBlocks<R, S, LEN, BLK + LEN, T>::run(v);
    }
};
template <int R, int S, int LEN, typename T>
struct DifLevels {
    static DFFT_HD void run(cx<T>* v) {
        DifBlocks<R, S, LEN, 0, T>::run(v);
        if constexpr (LEN > 2) DifLevels<R, S, LEN / 2, T>::run(v);
    }
};
// In-register radix-R DFT over v[q*S], q = 0..R-1; natural output q ends in v[bitrev_R(q)*S].
template <int R, int S, typename T>
DFFT_HD void dif_radix(cx<T>* v) {
    static_assert(R >= 2 && R <= 16, "radix must be 2..16");
    DifLevels<R, S, R, T>::run(v);
}

// ---------------------------------------------------------------------------------------------
// Inter-stage twiddles. tw[m] = exp(-2*pi*i*m/N), m = 0..N-1 (forward sign; the inverse transform is
// obtained by swapping re/im on load and store, so only forward twiddles exist).
// Powers w^q are composed from the binary powers w, w^2, w^4, w^8 which are read from the table
// (log2 r loads per butterfly instead of r-1).
// ---------------------------------------------------------------------------------------------
template <typename T>
DFFT_HD cx<T> ld_tw(const cx<T>* tw, int idx) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(T) == 8) {
        double2 r = __ldg(reinterpret_cast<const double2*>(tw) + idx);
        return cx<T>{T(r.x), T(r.y)};
    } else {
        float2 r = __ldg(reinterpret_cast<const float2*>(tw) + idx);
        return cx<T>{T(r.x), T(r.y)};
    }
#else
    return tw[idx];
#endif
}

template <int R, int S, int Q, typename T>
struct TwApply {
    static DFFT_HD void run(cx<T>* v, const cx<T>* wp /* wp[i] = w^(2^i) */) {
        // compose w^Q from set bits of Q
        cx<T> w{T(1), T(0)};
        bool first = true;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (Q & (1 << i)) {
                if (first) { w = wp[i]; first = false; }
                else w = cmul(w, wp[i]);
            }
        }
        v[Q * S] = cmul(v[Q * S], w);
        if constexpr (Q + 1 < R) TwApply<R, S, Q + 1, T>::run(v, wp);
    }
};

template <typename T, int LOG2N, int LOG2E>
struct FftCore {
    using SP = StagePlan<LOG2N, LOG2E>;
    static constexpr int N = SP::N, E = SP::E, TPL = SP::TPL, NST = SP::NST;

    // Stage ST on the E slots of thread j (0 <= j < TPL): twiddle (if Ns > 1) then butterflies.
    template <int ST, int B = 0>
    static DFFT_HD void stage_compute(cx<T> (&v)[E], int j, const cx<T>* tw) {
        constexpr int BITS = SP::bits(ST);
        constexpr int R = 1 << BITS;
        constexpr int S = E / R;
        constexpr int L2NS = SP::log2ns(ST);
        constexpr int NS = 1 << L2NS;
        if constexpr (NS > 1) {
            const int jb = j + B * TPL;
            const int k = jb & (NS - 1);
            const int base = k << (LOG2N - L2NS - BITS);  // index of w = W_{Ns*R}^k in the N-table
            cx<T> wp[4];
#pragma unroll
            for (int i = 0; i < BITS; ++i) wp[i] = ld_tw(tw, base << i);
            TwApply<R, S, 1, T>::run(&v[B], wp);
        }
        dif_radix<R, S, T>(&v[B]);
        if constexpr (B + 1 < S) stage_compute<ST, B + 1>(v, j, tw);
    }

    // Where slot `slot` of thread j goes after stage ST (ST < NST-1): position n' inside the line.
    template <int ST>
    static DFFT_HD int scatter_pos(int j, int slot) {
        constexpr int BITS = SP::bits(ST);
        constexpr int R = 1 << BITS;
        constexpr int S = E / R;
        constexpr int L2NS = SP::log2ns(ST);
        constexpr int NS = 1 << L2NS;
        const int b = slot % S;
        const int q = bitrev_c(slot / S, BITS);  // natural output index held by this slot
        const int jb = j + b * TPL;
        return ((jb >> L2NS) << (L2NS + BITS)) + (jb & (NS - 1)) + q * NS;
    }

    // scatter_pos(j, slot) == scatter_base(j) + scatter_off(slot): TPL is a multiple of Ns for every
    // scattering stage, so the butterfly index b and output index q only add compile-time constants.
    template <int ST>
    static DFFT_HD int scatter_base(int j) {
        constexpr int BITS = SP::bits(ST);
        constexpr int L2NS = SP::log2ns(ST);
        constexpr int NS = 1 << L2NS;
        static_assert(TPL % NS == 0, "threads per line must be a multiple of Ns");
        return ((j >> L2NS) << (L2NS + BITS)) + (j & (NS - 1));
    }
    template <int ST>
    static DFFT_HDC int scatter_off(int slot) {
        constexpr int BITS = SP::bits(ST);
        constexpr int R = 1 << BITS;
        constexpr int S = E / R;
        constexpr int NS = 1 << SP::log2ns(ST);
        return (slot % S) * TPL * R + bitrev_c(slot / S, BITS) * NS;
    }

    // After the LAST stage, the value that belongs to output position j + e*TPL sits in slot
    // final_slot(e).
    static DFFT_HDC int final_slot(int e) {
        constexpr int BITS = SP::bits(NST - 1);
        constexpr int R = 1 << BITS;
        constexpr int S = E / R;
        // e = b + q*S  ->  slot b + bitrev(q)*S
        return (e % S) + bitrev_c(e / S, BITS) * S;
    }
};

// ---------------------------------------------------------------------------------------------
// Shared-memory layout of one CTA tile (N points x TB lines) between stages.
//   CONTIG: idx(n, t) = t*NPAD + pad(n)      (lines side by side, threads walk along n)
//   TILED:  idx(n, t) = pad(n)*TB + t        (threads walk along t; rows of TB elements)
// pad(n) = n + (n >> SH) inserts one element (CONTIG) / one row (TILED) every 2^SH so that the strided
// first-stage scatter is bank-conflict free; TILED rows of >= 128 bytes need no padding.
// Both the scatter and the gather index split into a per-thread base plus a compile-time constant
// (idx(base + off) = idx(base) + off_const) — see split_ok().
// ---------------------------------------------------------------------------------------------
template <int LOG2N, int LOG2E, int TB, bool TILED, int ELEM_BYTES>
struct SmemLayout {
    using SP = StagePlan<LOG2N, LOG2E>;
    static constexpr int N = SP::N, TPL = SP::TPL;
    static constexpr bool PAD = TILED ? (TB * ELEM_BYTES < 128) : true;
    static constexpr int SH = TILED ? SP::bits(0) : (SP::bits(0) < 3 ? 3 : SP::bits(0));
    static constexpr int NPAD = PAD ? N + (N >> SH) : N;
    static constexpr int ELEMS = NPAD * TB;
    static DFFT_HDC int pad(int n) { return PAD ? n + (n >> SH) : n; }
    static DFFT_HDC int idx(int n, int t) { return TILED ? pad(n) * TB + t : t * NPAD + pad(n); }
    // constant part contributed by an offset `off` added to n (valid when no carry crosses bit SH)
    static DFFT_HDC int off(int o) { return TILED ? pad(o) * TB : pad(o); }
    // gather positions j + e*TPL split when TPL is a multiple of the pad period
    static constexpr bool GATHER_SPLIT = !PAD || (TPL % (1 << SH) == 0);
};

DFFT_HDC int pad_idx(int i) { return i + (i >> 3); }
DFFT_HDC int padded_len(int n) { return n + (n >> 3); }

}  // namespace dfft
