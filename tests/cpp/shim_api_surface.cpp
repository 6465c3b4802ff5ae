// Compile-time check of the drop-in surface: every public member of the reference's classes and parameter structs
// (include/mpicufft.hpp:55-105, mpicufft_slab.hpp:88-125, mpicufft_slab_z_then_yx.hpp:34-35, mpicufft_pencil.hpp:71-122,
// params.hpp:24-93) must exist in include/dfft.hpp with a signature the reference's own test drivers compile against.
// tests/test_abi.py compiles it with -fsyntax-only (all classes, float and double) and builds + runs the parameter
// part with -DPARAMS_ONLY (no library needed).
#include "dfft.hpp"

#ifndef PARAMS_ONLY
template <typename P> void use_base(P* f, void* o, const void* i) {
    GlobalSize g(8, 8, 8);
    f->initFFT(&g, true);
    f->setWorkArea(nullptr, nullptr);
    f->execR2C(o, i);
    f->execC2R(o, i);
    size_t v[3];
    f->getInSize(v); f->getInStart(v); f->getOutSize(v); f->getOutStart(v);
    size_t s = f->getDomainSize() + f->getWorkSizeDevice() + f->getWorkSizeHost();
    void* w = f->getWorkAreaDevice(); void* h = f->getWorkAreaHost();
    int r = f->getRank() + f->getWorldSize();
    (void)s; (void)w; (void)h; (void)r;
}
template <typename T> void use_all(dfft_comm_t c, void* o, const void* i) {
    Configurations cfg;
    MPIcuFFT_Slab<T> a(cfg, c, 8); use_base(&a, o, i);
    MPIcuFFT_Slab_Opt1<T> a1(cfg, c, 8); use_base(&a1, o, i);
    MPIcuFFT_Slab_Z_Then_YX<T> b(cfg, c, 8); use_base(&b, o, i);
    MPIcuFFT_Slab_Z_Then_YX_Opt1<T> b1(cfg, c, 8); use_base(&b1, o, i);
    MPIcuFFT_Pencil<T> p(cfg, c, 8);
    GlobalSize g(8, 8, 8); Pencil_Partition part(2, 4);
    p.initFFT(&g, &part, true);
    p.execR2C(o, i); p.execC2R(o, i); p.execR2C(o, i, 2); p.execC2R(o, i, 1);
    size_t v[3]; p.getInSize(v); p.getOutStart(v);
    Partition_Dimensions x, y, z;
    p.getPartitionDimensions(x, y, z);
    MPIcuFFT_Pencil_Opt1<T> p1(cfg, c, 8); p1.setWorkArea(nullptr, nullptr);
}
// how the reference's test drivers hold the plans (tests/src/slab/random_dist_default.cu:703-710,
// tests/src/pencil/random_dist_3D.cu:717-721): a base-class pointer that receives either variant
template <typename T> void driver_pattern(dfft_comm_t c, int opt, int world_size) {
    Configurations config;
    MPIcuFFT_Slab<T>* slab = opt == 1 ? new MPIcuFFT_Slab_Opt1<T>(config, c, world_size) : new MPIcuFFT_Slab<T>(config, c, world_size);
    MPIcuFFT_Slab_Z_Then_YX<T>* zyx = opt == 1 ? new MPIcuFFT_Slab_Z_Then_YX_Opt1<T>(config, c, world_size) : new MPIcuFFT_Slab_Z_Then_YX<T>(config, c, world_size);
    MPIcuFFT_Pencil<T>* pen = opt == 1 ? new MPIcuFFT_Pencil_Opt1<T>(config, c, world_size) : new MPIcuFFT_Pencil<T>(config, c, world_size);
    MPIcuFFT<T>* any = slab;
    GlobalSize global_size(8, 8, 8);
    any->initFFT(&global_size, nullptr, true);
    zyx->initFFT(&global_size, true);
    delete slab; delete zyx; delete pen;
}
template void driver_pattern<float>(dfft_comm_t, int, int);
template void driver_pattern<double>(dfft_comm_t, int, int);
template void use_all<float>(dfft_comm_t, void*, const void*);
template void use_all<double>(dfft_comm_t, void*, const void*);

#endif  // PARAMS_ONLY

int params_api() {
    Slab_Partition sp = 4;  // implicit, like params.hpp:44-49
    Partition_Dimensions d; d.size_x = {3, 2}; d.size_y = {1}; d.size_z = {4, 4, 1}; d.computeOffsets();
    Configurations c; c.cuda_aware = false; c.warmup_rounds = 0; c.comm_method = Peer2Peer; c.send_method = MPI_Type; c.benchmark_dir = "";
    c.comm_method2 = All2All; c.send_method2 = Streams;
    return int(sp.P1 + sp.P2 + d.start_x[1] + d.start_z[2]) == 4 + 1 + 3 + 8 ? 0 : 1;
}
int main() { return params_api(); }
