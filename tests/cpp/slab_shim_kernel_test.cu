// slab_shim_kernel_test.cu — the reference's testcase 4 call pattern, verbatim in shape
// (/root/reference/tests/src/slab/random_dist_default.cu:704-724): pageable/pinned cudaMemcpyAsync on the default
// stream, execR2C, a DEFAULT-STREAM KERNEL that scales the spectrum in place, then execC2R directly behind it with
// no synchronisation in between.  The reference relies on cuFFT running on the legacy default stream for the
// ordering (mpicufft_slab.cpp:788-807); the drop-in must give the same guarantee.  Also repeats the sequence with
// large buffers so that the scaling kernel is still running when execC2R is entered.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -Iinclude slab_shim_kernel_test.cu -Ldistributedfft_b200 -ldfft
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <vector>

#include "dfft.hpp"

#define CUDA_CALL(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("Error %d at %s:%d\n", int(e_), __FILE__, __LINE__); return 2; } } while (0)

// same job as Difference_Slab_Default::derivativeCoefficients (tests/src/slab/random_dist_default.cu:40-66):
// out[x][y][z] *= -(k1^2 + k2^2 + k3^2) / sqrt(N); written from the formula, grid-stride, with a deliberately slow
// inner loop (`spin`) so that the kernel is certainly still in flight when the host reaches execC2R
__global__ void derivative_coefficients(double2* out, size_t Nx, size_t Ny, size_t Nz, size_t y0, size_t ny_loc, int spin) {
    const size_t nzo = Nz / 2 + 1, total = Nx * ny_loc * nzo;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const size_t z = i % nzo, y = (i / nzo) % ny_loc + y0, x = i / (nzo * ny_loc);
        const double k1 = x <= Nx / 2 ? double(x) : double(x) - double(Nx);
        const double k2 = y <= Ny / 2 ? double(y) : double(y) - double(Ny);
        const double k3 = double(z);
        double c = -(k1 * k1 + k2 * k2 + k3 * k3) / sqrt(double(Nx) * double(Ny) * double(Nz));
        double w = 1.0;
        for (int s = 0; s < spin; ++s) w = w * 1.0000001 - (w - 1.0);  // stays 1.0 up to rounding noise removed below
        if (w < 0.5) c = 0;  // never true; keeps the loop alive
        double2 v = out[i];
        v.x *= c; v.y *= c;
        out[i] = v;
    }
}

template <typename Plan>
static int laplacian(const char* name, Plan& fft, size_t Nx, size_t Ny, size_t Nz, int spin) {
    GlobalSize gs(Nx, Ny, Nz);
    fft.initFFT(&gs, nullptr, true);
    size_t isz[3], osz[3], ost[3];
    fft.getInSize(isz); fft.getOutSize(osz); fft.getOutStart(ost);
    const size_t nin = isz[0] * isz[1] * isz[2];
    const double N = double(Nx) * Ny * Nz, pi = 3.14159265358979323846;
    std::vector<double> f(nin), back(nin);
    for (size_t x = 0; x < isz[0]; ++x)
        for (size_t y = 0; y < isz[1]; ++y)
            for (size_t z = 0; z < isz[2]; ++z)
                f[(x * isz[1] + y) * isz[2] + z] = std::sin(2 * pi * x / Nx) * std::sin(2 * pi * y / Ny) * std::sin(2 * pi * z / Nz);
    double *in_d, *back_d;
    double2* out_d;
    CUDA_CALL(cudaMalloc(&in_d, nin * sizeof(double)));
    CUDA_CALL(cudaMalloc(&back_d, nin * sizeof(double)));
    CUDA_CALL(cudaMalloc(&out_d, fft.getDomainSize()));
    double worst = 0;
    for (int run = 0; run < 3; ++run) {
        // pageable source, default stream, NO synchronisation before the exec (the staging DMA may still be in flight)
        CUDA_CALL(cudaMemcpyAsync(in_d, f.data(), nin * sizeof(double), cudaMemcpyHostToDevice));
        fft.execR2C(out_d, in_d);
        derivative_coefficients<<<296, 256>>>(out_d, Nx, Ny, Nz, ost[1], osz[1], spin);
        fft.execC2R(back_d, out_d);
        CUDA_CALL(cudaMemcpy(back.data(), back_d, nin * sizeof(double), cudaMemcpyDeviceToHost));
        const double amp = 3.0 * std::sqrt(N);
        double e = 0;
        for (size_t i = 0; i < nin; ++i) e = std::fmax(e, std::fabs(back[i] - (-amp * f[i])));
        worst = std::fmax(worst, e / amp);
    }
    std::printf("%s %zux%zux%zu spin=%d  Result (laplacian max / 3sqrt(N)): %.3e\n", name, Nx, Ny, Nz, spin, worst);
    cudaFree(in_d); cudaFree(back_d); cudaFree(out_d);
    // rounding grows with the grid (the multiplication by k^2 up to 3 (N/2)^2 amplifies it; the reference records
    // 7.5e-12 for this check at 1024^3): 1e-10 is BASELINE's bound; an ordering bug shows up as an error of O(1)
    return worst < 1e-10 ? 0 : 1;
}

int main() {
    dfft_comm_t comm;
    DFFT_CALL(dfft_comm_create(0, 1, nullptr, 0, &comm));
    Configurations config{true, 0, Peer2Peer, Sync, "", Peer2Peer, Sync};
    int rc = 0;
    {
        MPIcuFFT_Slab<double> fft(config, comm);
        rc |= laplacian("MPIcuFFT_Slab<double>", fft, 64, 32, 128, 0);
    }
    {
        MPIcuFFT_Slab<double> fft(config, comm);
        rc |= laplacian("MPIcuFFT_Slab<double>", fft, 256, 256, 256, 2000);  // scaling kernel runs for milliseconds
    }
    {
        MPIcuFFT_Slab_Z_Then_YX<double> fft(config, comm);
        rc |= laplacian("MPIcuFFT_Slab_Z_Then_YX<double>", fft, 128, 64, 64, 500);
    }
    {
        MPIcuFFT_Pencil<double> fft(config, comm);
        Pencil_Partition part(1, 1);
        GlobalSize gs(64, 64, 64);
        rc |= 0;
        fft.initFFT(&gs, &part, true);
    }
    dfft_comm_destroy(comm);
    std::printf(rc ? "FAILED\n" : "PASSED\n");
    return rc;
}
