// cufft_ref.cu — single-GPU cuFFT 3D transform: the oracle of the reference's own testcase 1
// (cufftMakePlan3d + exec on the coordinator rank, /root/reference/tests/src/slab/random_dist_default.cu:
// 300-303,337) and the compute engine of its P=1 path (src/slab/default/mpicufft_slab.cpp:142-145).
// TEST INFRASTRUCTURE ONLY: built into oracle/_ref/libcufft_ref.so by oracle/Makefile, loaded only by
// tests/ and bench.py.  kind: 0 Z2Z/C2C forward, 1 inverse, 2 D2Z/R2C, 3 Z2D/C2R.  prec: 0 f32, 1 f64.
#include <cuda_runtime.h>
#include <cufft.h>

extern "C" int cufft_ref_3d(int prec, int kind, int nx, int ny, int nz, void* out, void* in, float* ms, int reps) {
    cufftHandle plan;
    cufftType type;
    if (kind <= 1) type = prec ? CUFFT_Z2Z : CUFFT_C2C;
    else if (kind == 2) type = prec ? CUFFT_D2Z : CUFFT_R2C;
    else type = prec ? CUFFT_Z2D : CUFFT_C2R;
    if (cufftPlan3d(&plan, nx, ny, nz, type) != CUFFT_SUCCESS) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int rc = 0;
    if (reps < 1) reps = 1;
    for (int r = 0; r < reps + 1; ++r) {
        if (r == 1) cudaEventRecord(e0);
        cufftResult res;
        if (kind <= 1) {
            int dir = kind == 0 ? CUFFT_FORWARD : CUFFT_INVERSE;
            res = prec ? cufftExecZ2Z(plan, (cufftDoubleComplex*)in, (cufftDoubleComplex*)out, dir)
                       : cufftExecC2C(plan, (cufftComplex*)in, (cufftComplex*)out, dir);
        } else if (kind == 2) {
            res = prec ? cufftExecD2Z(plan, (double*)in, (cufftDoubleComplex*)out) : cufftExecR2C(plan, (float*)in, (cufftComplex*)out);
        } else {
            res = prec ? cufftExecZ2D(plan, (cufftDoubleComplex*)in, (double*)out) : cufftExecC2R(plan, (cufftComplex*)in, (float*)out);
        }
        if (res != CUFFT_SUCCESS) rc = -2;
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1);
    if (ms) *ms = (reps > 0) ? t / reps : 0.f;
    if (reps == 1 && ms) { /* single timed rep after one warm-up */ }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cufftDestroy(plan);
    if (cudaDeviceSynchronize() != cudaSuccess) rc = -3;
    return rc;
}
