// CPU emulation of the CUDA FFT core: runs fft_core.cuh's stage arithmetic "one thread at a time" with
// explicit exchange buffers standing in for shared memory, and checks it against a naive O(N^2) DFT.
// Built and run by tests/test_core_emulation.py (no GPU needed). Exit code 0 = all sizes pass.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../distributedfft_b200/csrc/fft_core.cuh"

using namespace dfft;

static int split_errors = 0;

// Runs TB lines through the stages with the same shared-memory indexing the kernels use
// (per-thread base + compile-time offset), checking it against the direct index.
template <typename T, int LOG2N, int LOG2E, int TB, bool TILED, int ST>
struct RunStages {
    using Core = FftCore<T, LOG2N, LOG2E>;
    using L = SmemLayout<LOG2N, LOG2E, TB, TILED, int(sizeof(cx<T>))>;
    static void run(std::vector<cx<T>>& regs, std::vector<cx<T>>& smem, const cx<T>* tw) {
        constexpr int E = Core::E, TPL = Core::TPL, NST = Core::NST;
        for (int t = 0; t < TB; ++t)
            for (int j = 0; j < TPL; ++j) {
                cx<T>(&v)[E] = *reinterpret_cast<cx<T>(*)[E]>(&regs[(size_t(t) * TPL + j) * E]);
                Core::template stage_compute<ST>(v, j, tw);
            }
        if constexpr (ST + 1 < NST) {
            for (auto& c : smem) c = cx<T>{T(1e30), T(1e30)};
            for (int t = 0; t < TB; ++t)
                for (int j = 0; j < TPL; ++j) {
                    const int sb = L::idx(Core::template scatter_base<ST>(j), t);
                    for (int e = 0; e < E; ++e) {
                        const int direct = L::idx(Core::template scatter_pos<ST>(j, e), t);
                        const int split = sb + L::off(Core::template scatter_off<ST>(e));
                        if (direct != split) ++split_errors;
                        smem[split] = regs[(size_t(t) * TPL + j) * E + e];
                    }
                }
            for (int t = 0; t < TB; ++t)
                for (int j = 0; j < TPL; ++j) {
                    const int gb = L::idx(j, t);
                    for (int e = 0; e < E; ++e) {
                        const int direct = L::idx(j + e * TPL, t);
                        const int split = L::GATHER_SPLIT ? gb + L::off(e * TPL) : direct;
                        if (direct != split) ++split_errors;
                        regs[(size_t(t) * TPL + j) * E + e] = smem[split];
                    }
                }
            RunStages<T, LOG2N, LOG2E, TB, TILED, ST + 1>::run(regs, smem, tw);
        }
    }
};

template <typename T, int LOG2N, int LOG2E, int TB = 2, bool TILED = false>
double check(bool inverse) {
    using Core = FftCore<T, LOG2N, LOG2E>;
    using L = SmemLayout<LOG2N, LOG2E, TB, TILED, int(sizeof(cx<T>))>;
    constexpr int N = Core::N, E = Core::E, TPL = Core::TPL;
    std::vector<cx<T>> x(size_t(N) * TB), out(size_t(N) * TB), regs(size_t(TPL) * E * TB), smem(L::ELEMS), tw(N);
    for (int m = 0; m < N; ++m) {
        long double a = -2.0L * M_PIl * m / N;
        tw[m] = cx<T>{T(cosl(a)), T(sinl(a))};
    }
    srand(1234 + LOG2N);
    for (auto& c : x) c = cx<T>{T(rand() / double(RAND_MAX) - 0.5), T(rand() / double(RAND_MAX) - 0.5)};
    for (int t = 0; t < TB; ++t)
        for (int j = 0; j < TPL; ++j)
            for (int e = 0; e < E; ++e) {
                cx<T> v = x[size_t(t) * N + j + e * TPL];
                regs[(size_t(t) * TPL + j) * E + e] = inverse ? cswap(v) : v;
            }
    RunStages<T, LOG2N, LOG2E, TB, TILED, 0>::run(regs, smem, tw.data());
    for (int t = 0; t < TB; ++t)
        for (int j = 0; j < TPL; ++j)
            for (int e = 0; e < E; ++e) {
                cx<T> r = regs[(size_t(t) * TPL + j) * E + Core::final_slot(e)];
                out[size_t(t) * N + j + e * TPL] = inverse ? cswap(r) : r;
            }
    // naive DFT in long double (cos/sin tabulated once per N)
    double num = 0, den = 0;
    std::vector<long double> ct(N), st(N);
    for (int m = 0; m < N; ++m) {
        long double a = 2.0L * M_PIl * m / N;
        ct[m] = cosl(a);
        st[m] = (inverse ? 1.0L : -1.0L) * sinl(a);
    }
    for (int t = 0; t < TB; ++t)
        for (int k = 0; k < N; ++k) {
            long double sr = 0, si = 0;
            const cx<T>* xl = &x[size_t(t) * N];
            for (int n = 0; n < N; ++n) {
                const int idx = int((long long)k * n % N);
                long double c = ct[idx], s = st[idx];
                sr += xl[n].x * c - xl[n].y * s;
                si += xl[n].x * s + xl[n].y * c;
            }
            const cx<T> o = out[size_t(t) * N + k];
            num += double((o.x - sr) * (o.x - sr) + (o.y - si) * (o.y - si));
            den += double(sr * sr + si * si);
        }
    return std::sqrt(num / den);
}

static int fails = 0;
template <typename T, int LOG2N, int LOG2E>
void one(double tol) {
    for (int inv = 0; inv < 2; ++inv) {
        double e = check<T, LOG2N, LOG2E, 2, false>(inv);
        e = std::max(e, check<T, LOG2N, LOG2E, 4, true>(inv));   // tiled, padded rows (< 128 B for f64/f32)
        e = std::max(e, check<T, LOG2N, LOG2E, 16, true>(inv));  // tiled, unpadded rows
        bool ok = e < tol && split_errors == 0;
        printf("%s N=2^%d E=2^%d %s relL2=%.3e %s\n", sizeof(T) == 8 ? "f64" : "f32", LOG2N, LOG2E, inv ? "inv" : "fwd", e,
               ok ? "ok" : "FAIL");
        if (!ok) ++fails;
    }
}

template <typename T, int LOG2N>
void sizes(double tol) {
    one<T, LOG2N, (LOG2N < 4 ? LOG2N : 4)>(tol);
    if constexpr (LOG2N >= 3) one<T, LOG2N, 3>(tol);
    if constexpr (LOG2N >= 2) one<T, LOG2N, 2>(tol);
    if constexpr (LOG2N > 1) sizes<T, LOG2N - 1>(tol);
}

int main(int argc, char** argv) {
    const int maxlog = 12;
    (void)argc; (void)argv; (void)maxlog;
    sizes<double, 12>(1e-14);
    sizes<float, 11>(2e-6);
    printf("split_errors=%d\n", split_errors);
    if (split_errors) ++fails;
    printf("fails=%d\n", fails);
    return fails ? 1 : 0;
}
