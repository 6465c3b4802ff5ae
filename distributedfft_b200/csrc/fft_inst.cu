// fft_inst.cu — compiled once per (precision, log2 length): -DDFFT_T=double -DDFFT_LOG2N=10.
// Splitting the instantiations over translation units lets build.py compile them in parallel.
#include <cstdlib>

#include "fft_kernels.cuh"

#ifndef DFFT_T
#error "compile with -DDFFT_T=<float|double> -DDFFT_LOG2N=<1..13>"
#endif

namespace dfft {

template <typename K>
static cudaError_t set_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
}

static int wide_tiles() {
    const char* e = getenv("DFFT_WIDE_TILES");
    return e ? atoi(e) : 0;
}

template <typename T, int LOG2N, int TB>
static cudaError_t launch_tiled(const FftParams& p, cudaStream_t stream, long long lines) {
    using S = Shape<T, LOG2N>;
    constexpr int LOG2E = S::LOG2E;
    using C = CtaFft<T, LOG2N, LOG2E, TB, true>;
    auto kf = fft_c2c_kernel<T, LOG2N, LOG2E, TB, true, false>;
    auto ki = fft_c2c_kernel<T, LOG2N, LOG2E, TB, true, true>;
    static cudaError_t once = set_smem(kf, C::SMEM_BYTES) != cudaSuccess ? cudaErrorInvalidValue : set_smem(ki, C::SMEM_BYTES);
    if (once != cudaSuccess) return once;
    if (p.B <= 0) return cudaSuccess;
    const long long grid = lines * ((p.B + TB - 1) / TB);
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    {
        // TMA bulk-store variant (see fft_c2c_bulk_kernel): only when the launcher's contract holds
        if (p.bulk_out && p.in.nseg == 1 && p.out.sN == TB && p.B == TB && (TB * sizeof(cx<T>)) % 16 == 0) {
            auto bf = fft_c2c_bulk_kernel<T, LOG2N, LOG2E, TB, false>;
            auto bi = fft_c2c_bulk_kernel<T, LOG2N, LOG2E, TB, true>;
            static cudaError_t onceb = set_smem(bf, C::SMEM_BYTES) != cudaSuccess ? cudaErrorInvalidValue : set_smem(bi, C::SMEM_BYTES);
            if (onceb != cudaSuccess) return onceb;
            const unsigned gb = unsigned((p.max_ctas > 0 && grid > p.max_ctas) ? p.max_ctas : grid);
            if (p.inverse) bi<<<gb, C::THREADS, C::SMEM_BYTES, stream>>>(p);
            else bf<<<gb, C::THREADS, C::SMEM_BYTES, stream>>>(p);
            return cudaGetLastError();
        }
    }
    const unsigned g = unsigned((p.max_ctas > 0 && grid > p.max_ctas) ? p.max_ctas : grid);
    if (p.inverse) ki<<<g, C::THREADS, C::SMEM_BYTES, stream>>>(p);
    else kf<<<g, C::THREADS, C::SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

template <typename T, int LOG2N>
cudaError_t launch_pass(PassKind kind, const FftParams& p, cudaStream_t stream) {
    using S = Shape<T, LOG2N>;
    constexpr int LOG2E = S::LOG2E;
    const long long lines = (long long)p.A0 * p.A1;
    if (lines <= 0) return cudaSuccess;
    cudaError_t err = cudaSuccess;
    switch (kind) {
        case PASS_C2C_CONTIG: {
            using C = CtaFft<T, LOG2N, LOG2E, S::TBC, false>;
            auto kf = fft_c2c_kernel<T, LOG2N, LOG2E, S::TBC, false, false>;
            auto ki = fft_c2c_kernel<T, LOG2N, LOG2E, S::TBC, false, true>;
            static cudaError_t once = set_smem(kf, C::SMEM_BYTES) != cudaSuccess ? cudaErrorInvalidValue : set_smem(ki, C::SMEM_BYTES);
            if (once != cudaSuccess) return once;
            if (p.in.sN != 1 || p.out.sN != 1) return cudaErrorInvalidValue;
            const long long grid = (lines + S::TBC - 1) / S::TBC;
            if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
            if (p.inverse) ki<<<unsigned(grid), C::THREADS, C::SMEM_BYTES, stream>>>(p);
            else kf<<<unsigned(grid), C::THREADS, C::SMEM_BYTES, stream>>>(p);
            break;
        }
        case PASS_C2C_TILED: {
            // Wide tiles (128-byte rows, one CTA per SM) win when the rows of a tile sit in different 2 MB pages
            // (x passes with a row pitch >= 1 MiB: 3012 -> 3863 GB/s at N=1024 f64) and lose otherwise
            // (y passes: 4611 -> 3876 GB/s).  DFFT_WIDE_TILES=1 / -1 forces them on / off.
            if constexpr (S::TBT_WIDE != S::TBT) {
                const int w = wide_tiles();
                const bool far_rows = (unsigned long long)p.in.sN * sizeof(cx<T>) >= (1ull << 20) ||
                                      (unsigned long long)p.out.sN * sizeof(cx<T>) >= (1ull << 20);
                const bool wide = w > 0 || p.tile_pref == 2 || (w == 0 && p.tile_pref == 0 && LOG2N >= 10 && far_rows);
                if (wide && w >= 0 && p.tile_pref != 1) return launch_tiled<T, LOG2N, S::TBT_WIDE>(p, stream, lines);
            }
            return launch_tiled<T, LOG2N, S::TBT>(p, stream, lines);
        }
        case PASS_R2C: {
            using C = CtaFft<T, LOG2N, LOG2E, S::TBC, false>;
            auto k = fft_r2c_kernel<T, LOG2N, LOG2E, S::TBC>;
            static cudaError_t once = set_smem(k, C::SMEM_BYTES);
            if (once != cudaSuccess) return once;
            if (p.in.nseg != 1 || p.in.sN != 1 || p.out.sN != 1) return cudaErrorInvalidValue;
            const long long grid = (lines + S::TBC - 1) / S::TBC;
            if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
            k<<<unsigned(grid), C::THREADS, C::SMEM_BYTES, stream>>>(p);
            break;
        }
        case PASS_C2R: {
            using C = CtaFft<T, LOG2N, LOG2E, S::TBC, false>;
            auto k = fft_c2r_kernel<T, LOG2N, LOG2E, S::TBC>;
            static cudaError_t once = set_smem(k, C::SMEM_BYTES);
            if (once != cudaSuccess) return once;
            if (p.out.nseg != 1 || p.in.sN != 1 || p.out.sN != 1) return cudaErrorInvalidValue;
            const long long grid = (lines + S::TBC - 1) / S::TBC;
            if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
            k<<<unsigned(grid), C::THREADS, C::SMEM_BYTES, stream>>>(p);
            break;
        }
        default: return cudaErrorInvalidValue;
    }
    err = cudaGetLastError();
    return err;
}

template cudaError_t launch_pass<DFFT_T, DFFT_LOG2N>(PassKind, const FftParams&, cudaStream_t);

}  // namespace dfft
