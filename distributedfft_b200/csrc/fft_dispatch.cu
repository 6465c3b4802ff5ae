// fft_dispatch.cu — runtime length -> compile-time instantiation (see fft_inst.cu).
#include "fft_kernels.cuh"

namespace dfft {

#define DFFT_DECL(T, L) extern template cudaError_t launch_pass<T, L>(PassKind, const FftParams&, cudaStream_t);
#define DFFT_ALL(M, T) M(T, 1) M(T, 2) M(T, 3) M(T, 4) M(T, 5) M(T, 6) M(T, 7) M(T, 8) M(T, 9) M(T, 10) M(T, 11) M(T, 12) M(T, 13)
DFFT_ALL(DFFT_DECL, double)
DFFT_ALL(DFFT_DECL, float)

#define DFFT_CASE(T, L) \
    case L: return launch_pass<T, L>(kind, p, stream);

cudaError_t launch_pass_f64(int log2n, PassKind kind, const FftParams& p, cudaStream_t stream) {
    switch (log2n) {
        DFFT_ALL(DFFT_CASE, double)
        default: return cudaErrorInvalidValue;
    }
}
cudaError_t launch_pass_f32(int log2n, PassKind kind, const FftParams& p, cudaStream_t stream) {
    switch (log2n) {
        DFFT_ALL(DFFT_CASE, float)
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace dfft
