// nvlink_micro.cu — how fast can ONE GPU push data into a peer's memory over NVLink, as a function of the number of
// CTAs (= SMs) that do the pushing and of the instruction that moves the bytes:
//   mode 0: 16-byte st.global from registers (what an FFT pass' epilogue does), each warp writes 512 contiguous bytes
//   mode 1: cp.async.bulk shared -> peer global (TMA engine, SASS UBLKCP), one bulk copy of `chunk` bytes per step
//   mode 2: like 0 but to LOCAL memory (reference)          mode 3: like 1 but to LOCAL memory
// One process, two GPUs (peer access enabled).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 nvlink_micro.cu -o nvlink_micro
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); std::exit(1); } } while (0)

__global__ void __launch_bounds__(256) push_stg(double2* dst, size_t elems_per_cta, int iters) {
    // each CTA owns a contiguous region of elems_per_cta double2; 256 threads x 16 stores per step
    double2* base = dst + size_t(blockIdx.x) * elems_per_cta;
    const double2 v = make_double2(threadIdx.x, blockIdx.x);
    for (int it = 0; it < iters; ++it) {
        double2* p = base + (size_t(it) * 4096) % elems_per_cta + threadIdx.x;
#pragma unroll
        for (int e = 0; e < 16; ++e) p[e * 256] = v;
    }
}

__global__ void __launch_bounds__(256) push_bulk(char* dst, size_t bytes_per_cta, int iters, unsigned chunk, int depth) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (unsigned i = threadIdx.x; i < chunk * depth / 16; i += blockDim.x) reinterpret_cast<double2*>(sm)[i] = make_double2(i, blockIdx.x);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        char* base = dst + size_t(blockIdx.x) * bytes_per_cta;
        for (int it = 0; it < iters; ++it) {
            const unsigned src = unsigned(__cvta_generic_to_shared(sm + size_t(it % depth) * chunk));
            char* p = base + (size_t(it) * chunk) % bytes_per_cta;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p), "r"(src), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // keep at most `depth` bulk copies reading shared memory (a real kernel must not overwrite the tile earlier)
            if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else if (depth == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main(int argc, char** argv) {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { std::printf("needs 2 GPUs\n"); return 0; }
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    const size_t total = size_t(1) << 30;  // 1 GiB target buffer
    char *remote = nullptr, *local = nullptr;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&remote, total)); CK(cudaMemset(remote, 0, total));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&local, total));
    CK(cudaFuncSetAttribute(push_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grids[] = {8, 16, 32, 48, 64, 96, 148, 296};
    std::printf("mode,target,grid,chunk_bytes,depth,GBps,GBps_per_cta\n");
    for (int target = 0; target < 2; ++target) {
        char* dst = target == 0 ? remote : local;
        for (int g : grids) {
            const size_t per = (total / g) & ~size_t(65535);
            // st.global from registers: 64 KiB per step per CTA
            {
                const int iters = 256;
                for (int rep = 0; rep < 2; ++rep) {
                    CK(cudaEventRecord(e0));
                    push_stg<<<g, 256>>>(reinterpret_cast<double2*>(dst), per / 16, iters);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                }
                float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
                const double gb = double(g) * iters * 65536.0 / 1e9;
                std::printf("stg128,%s,%d,65536,0,%.1f,%.2f\n", target ? "local" : "peer", g, gb / (ms * 1e-3), gb / (ms * 1e-3) / g);
            }
            const unsigned chunks[] = {2048, 8192, 16384, 65536};
            for (unsigned chunk : chunks)
                for (int depth : {1, 2, 4}) {
                    if (size_t(chunk) * depth > 196608) continue;
                    if (g == 296 && size_t(chunk) * depth > 98304) continue;
                    const int iters = int((size_t(16) << 20) / chunk);  // 16 MiB per CTA
                    for (int rep = 0; rep < 2; ++rep) {
                        CK(cudaEventRecord(e0));
                        push_bulk<<<g, 256, size_t(chunk) * depth>>>(dst, per, iters, chunk, depth);
                        CK(cudaEventRecord(e1));
                        CK(cudaEventSynchronize(e1));
                    }
                    CK(cudaGetLastError());
                    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
                    const double gb = double(g) * iters * double(chunk) / 1e9;
                    std::printf("bulk,%s,%d,%u,%d,%.1f,%.2f\n", target ? "local" : "peer", g, chunk, depth, gb / (ms * 1e-3), gb / (ms * 1e-3) / g);
                }
        }
    }
    return 0;
}
