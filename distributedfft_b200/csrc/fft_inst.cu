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

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);  // read per launch (cheap) so that one process can compare variants
    return e ? atoi(e) : dflt;
}

// cuTensorMapEncodeTiled through the runtime (libdfft.so does not link libcuda)
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn encode_tiled() {
    static encode_tiled_fn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) { cudaGetLastError(); f = nullptr; }
        return (encode_tiled_fn)f;
    }();
    return fn;
}

// rank-4 tensor map (b, n, a1, a0) of a single-segment view; box TB x min(N, 256) x 1 x 1.  false: the view does not
// meet the TMA constraints (16-byte aligned base and strides) and the register-fed kernel is used instead.
template <typename T>
static bool make_view_map(const View& v, int A0, int A1, int N, int B, int TB, CUtensorMap* tm) {
    encode_tiled_fn enc = encode_tiled();
    if (!enc || v.nseg != 1) return false;
    const unsigned long long es = sizeof(cx<T>);
    const int d0 = sizeof(T) == 8 ? 2 : 1;  // 8-byte units per complex element
    char* base = reinterpret_cast<char*>(v.seg[0].base) - (long long)v.seg[0].n0 * v.sN * (long long)es;
    if (reinterpret_cast<unsigned long long>(base) % 16) return false;
    if (v.sN <= 0 || v.seg[0].sA0 < 0 || v.seg[0].sA1 < 0) return false;
    cuuint64_t dims[4] = {cuuint64_t(B) * d0, cuuint64_t(N), cuuint64_t(A1 > 0 ? A1 : 1), cuuint64_t(A0 > 0 ? A0 : 1)};
    unsigned long long sn = (unsigned long long)v.sN * es;
    unsigned long long s1 = A1 > 1 ? (unsigned long long)v.seg[0].sA1 * es : sn * N;
    unsigned long long s0 = A0 > 1 ? (unsigned long long)v.seg[0].sA0 * es : sn * N;
    if (s1 == 0) s1 = sn * N;
    if (s0 == 0) s0 = sn * N;
    cuuint64_t strides[3] = {sn, s1, s0};
    for (int i = 0; i < 3; ++i)
        if (strides[i] % 16 || strides[i] >= (1ull << 40)) return false;
    if ((cuuint64_t(TB) * d0 * 8) % 16) return false;
    cuuint32_t box[4] = {cuuint32_t(TB * d0), cuuint32_t(N < 256 ? N : 256), 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    // (L2 promotion to 128 / 256 bytes was measured: no effect at 128 bytes, -5 ... -12 % at 256: profiles/r02/axis_f64_s1b_tma_p*.log)
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int wide_tiles() {
    const char* e = getenv("DFFT_WIDE_TILES");
    return e ? atoi(e) : 0;
}

template <typename T, int LOG2N, int TB>
static cudaError_t launch_tiled(const FftParams& p_in, cudaStream_t stream, long long lines) {
    FftParams p = p_in;
    {
        const int e = env_int("DFFT_TILE_SWZ", -1);  // experiment override of the tile-order blocking
        if (e >= 0) p.tile_swz = e;
        if (p.tile_swz > 0) {
            const int G = 1 << p.tile_swz;
            if (p.tile_swz > 4 || p.A0 % G || p.A1 % G) p.tile_swz = 0;
        } else p.tile_swz = 0;
    }
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
    if constexpr (C::NST >= 2 && LOG2N >= 7) {
        // TMA-fed persistent kernel (fft_c2c_tma_kernel).  Measured on B200 (tools/axis_bench.py, profiles/r02/axis_*):
        // f64 strided passes of 1024 points +7 % (64-byte rows, 16 KB pitch) ... +22 % (128-byte rows) ... +50 % (128-byte
        // rows, 16 KB pitch: 3876 -> 5838 GB/s), 2048 points +15 ... +40 %, 4096 points +10 ... +23 %; lines of <= 512
        // points and all f32 tiles lose (their register-fed kernels already run at 0.9+ of the copy bandwidth, or the
        // tile leaves room for one CTA per SM only), and so do 64-byte rows of the blocked hand-over layout (3009 vs 4044).
        // Default (-1): f64 lines of >= 1024 points, except 64-byte blocked rows; DFFT_TMA=1 forces it on, 0 off.
        const int mode = env_int("DFFT_TMA", -1);
        const bool narrow_blocked = p.B == TB && TB * sizeof(cx<T>) < 128;
        const bool want = mode > 0 || (mode < 0 && sizeof(T) == 8 && LOG2N >= 10 && !narrow_blocked);
        const bool bulk_ok = p.bulk_out && p.out.sN == TB && p.B == TB;
        alignas(64) CUtensorMap tm;
        if (want && p.in.nseg == 1 && (!p.bulk_out || bulk_ok) && make_view_map<T>(p.in, p.A0, p.A1, C::N, p.B, TB, &tm)) {
            auto tf = fft_c2c_tma_kernel<T, LOG2N, LOG2E, TB, false>;
            auto ti = fft_c2c_tma_kernel<T, LOG2N, LOG2E, TB, true>;
            static cudaError_t oncet = set_smem(tf, C::SMEM_BYTES) != cudaSuccess ? cudaErrorInvalidValue : set_smem(ti, C::SMEM_BYTES);
            if (oncet != cudaSuccess) return oncet;
            static int resident = [&] {
                int dev = 0, sms = 0, per = 0;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, tf, C::THREADS, C::SMEM_BYTES) != cudaSuccess || per < 1) per = 1;
                return sms * per;
            }();
            long long g = grid < resident ? grid : resident;
            if (p.max_ctas > 0 && g > p.max_ctas) g = p.max_ctas;
            FftParams q = p;
            q.bulk_out = bulk_ok ? 1 : 0;
            if (p.inverse) ti<<<unsigned(g), C::THREADS, C::SMEM_BYTES, stream>>>(q, tm);
            else tf<<<unsigned(g), C::THREADS, C::SMEM_BYTES, stream>>>(q, tm);
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
                // with the TMA-fed kernel (f64, >= 1024 points) the wide tile wins for near rows as well (5838 vs 5089 GB/s)
                const bool tma_auto = env_int("DFFT_TMA", -1) != 0 && sizeof(T) == 8 && LOG2N >= 10 && p.in.nseg == 1;
                const bool wide = w > 0 || p.tile_pref == 2 || (w == 0 && p.tile_pref == 0 && LOG2N >= 10 && (far_rows || tma_auto));
                // never a wide tile for a view that is only one narrow tile wide (blocked hand-over with CH = TBT): half of
                // its columns would be empty
                if (wide && w >= 0 && p.tile_pref != 1 && p.B > S::TBT) return launch_tiled<T, LOG2N, S::TBT_WIDE>(p, stream, lines);
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
