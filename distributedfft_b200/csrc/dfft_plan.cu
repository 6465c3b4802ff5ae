// dfft_plan.cu — plan construction, step schedule and execution behind include/dfft.h.
//
// A plan is a short list of steps — FFT passes (fft_kernels.cuh), device rendezvous kernels and NCCL
// all-to-all-v exchanges — built once from the partition geometry, replayed by every exec:
//
//   reference call                                              here
//   ------------------------------------------------------------------------------------------------
//   cufftExecR2C/D2Z 2D (y,z)   mpicufft_slab.cpp:788           R2C z pass + C2C y pass (scattering)
//   pack memcpy2D + Alltoallv   mpicufft_slab.cpp:646-661       fused into the y pass' stores (Peer2Peer)
//                                                               or send slots + ncclSend/Recv (All2All)
//   cufftExecZ2Z 1D x           mpicufft_slab.cpp:806           C2C x pass (gathering)
//   pencil transposes 1 and 2   mpicufft_pencil.cpp:858-930,    same mechanism, row / column groups
//                               1490-1576
//
// Peer2Peer: every rank maps the other ranks' work slots with CUDA IPC; an FFT pass writes its output
// rows straight into the layout the receiver's next pass reads (NVLink stores), and ranks meet at tiny
// flag kernels (st.release.sys / ld.acquire.sys on peer-mapped flags) instead of MPI_Waitall.
#include <cuda_runtime.h>
#include <nccl.h>

#include <sys/stat.h>

#include <chrono>
#include <cmath>
#include <fstream>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/dfft.h"
#include "fft_kernels.cuh"
#include "geometry.hpp"

namespace dfft {

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CK_CUDA(x)                                                                                     \
    do {                                                                                               \
        cudaError_t e_ = (x);                                                                          \
        if (e_ != cudaSuccess)                                                                         \
            return fail(DFFT_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                                           std::to_string(__LINE__) + ")");                            \
    } while (0)
#define CK_NCCL(x)                                                                                     \
    do {                                                                                               \
        ncclResult_t r_ = (x);                                                                         \
        if (r_ != ncclSuccess)                                                                         \
            return fail(DFFT_ERR_NCCL, std::string(#x) + ": " + ncclGetErrorString(r_) + " (" + __FILE__ + ":" + \
                                           std::to_string(__LINE__) + ")");                            \
    } while (0)

static int ilog2_exact(size_t n) {
    if (n == 0 || (n & (n - 1))) return -1;
    int l = 0;
    while ((size_t(1) << l) < n) ++l;
    return l;
}

// ---- twiddle tables ---------------------------------------------------------------------------------
template <typename T>
static cudaError_t make_table(size_t count, size_t denom, void** dev) {
    std::vector<cx<T>> h(count);
    for (size_t m = 0; m < count; ++m) {
        // exact symmetric evaluation in long double
        long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)denom;
        h[m] = cx<T>{T(cosl(a)), T(sinl(a))};
    }
    cudaError_t e = cudaMalloc(dev, count * sizeof(cx<T>));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dev, h.data(), count * sizeof(cx<T>), cudaMemcpyHostToDevice);
}

struct Tables {
    int prec = DFFT_F64;
    bool dry = false;  // geometry-only plans: no device allocations, segment tables live on the host
    std::map<const void*, std::vector<unsigned char>> host_tabs;  // segment table pointer -> host copy
    std::vector<std::vector<unsigned char>*> dry_store;
    std::map<int, void*> tw;   // log2n -> exp(-2 pi i m / n), m < n
    std::map<int, void*> tw2;  // log2m -> exp(-2 pi i k / 2m), k <= m/2
    std::vector<void*> bytes;  // seg_of_n tables
    cudaError_t get_tw(int log2n, void** out) {
        if (dry) { *out = nullptr; return cudaSuccess; }
        auto it = tw.find(log2n);
        if (it == tw.end()) {
            void* d = nullptr;
            size_t n = size_t(1) << log2n;
            cudaError_t e = prec == DFFT_F64 ? make_table<double>(n, n, &d) : make_table<float>(n, n, &d);
            if (e != cudaSuccess) return e;
            it = tw.emplace(log2n, d).first;
        }
        *out = it->second;
        return cudaSuccess;
    }
    cudaError_t get_tw2(int log2m, void** out) {
        if (dry) { *out = nullptr; return cudaSuccess; }
        auto it = tw2.find(log2m);
        if (it == tw2.end()) {
            void* d = nullptr;
            size_t m = size_t(1) << log2m;
            cudaError_t e = prec == DFFT_F64 ? make_table<double>(m / 2 + 1, 2 * m, &d) : make_table<float>(m / 2 + 1, 2 * m, &d);
            if (e != cudaSuccess) return e;
            it = tw2.emplace(log2m, d).first;
        }
        *out = it->second;
        return cudaSuccess;
    }
    cudaError_t seg_table(const Split& s, const unsigned char** out) {
        size_t n = s.start.back() + s.size.back();
        std::vector<unsigned char> h(n);
        for (size_t p = 0; p < s.size.size(); ++p)
            for (size_t k = 0; k < s.size[p]; ++k) h[s.start[p] + k] = (unsigned char)p;
        if (dry) {
            auto* keep = new std::vector<unsigned char>(h);
            dry_store.push_back(keep);
            host_tabs[keep->data()] = h;
            *out = keep->data();
            return cudaSuccess;
        }
        void* d = nullptr;
        cudaError_t e = cudaMalloc(&d, n ? n : 1);
        if (e != cudaSuccess) return e;
        e = cudaMemcpy(d, h.data(), n, cudaMemcpyHostToDevice);
        bytes.push_back(d);
        host_tabs[d] = h;
        *out = (const unsigned char*)d;
        return e;
    }
    // segment tables belong to the schedules: dropped whenever the schedules are rebuilt (dfft_set_work_area)
    void release_seg_tables() {
        for (void* p : bytes) cudaFree(p);
        for (auto* v : dry_store) delete v;
        bytes.clear(); dry_store.clear(); host_tabs.clear();
    }
    void release() {
        for (auto& kv : tw) cudaFree(kv.second);
        for (auto& kv : tw2) cudaFree(kv.second);
        for (void* p : bytes) cudaFree(p);
        for (auto* v : dry_store) delete v;
        tw.clear(); tw2.clear(); bytes.clear(); dry_store.clear(); host_tabs.clear();
    }
};

// ---- device rendezvous ---------------------------------------------------------------------------------
// flags layout (uint64): [phase][rank], NPHASE phases.  Thread i handles group member i: publishes my
// epoch in the peer's flag row, then waits for the peer's epoch in mine.
constexpr int NPHASE = 4;
constexpr unsigned long long FLAG_POISON = ~0ull;  // written into every peer's flag rows by a rank that gave up waiting
// A missing peer: the reference would block in MPI_Waitall forever (mpicufft_slab.cpp:802-803).  Here the wait
// gives up after `timeout_cycles` (DFFT_RENDEZVOUS_TIMEOUT_S, default 300 s, 0 = wait forever), records the error
// and POISONS the flag rows of every group member, so that no rank silently continues on half-delivered slots:
// every later rendezvous of every member sees the poison, records an error as well, and dfft_plan_wait returns
// DFFT_ERR_TIMEOUT on all of them (the plan is dead after that and must be destroyed).
__global__ void rendezvous_kernel(unsigned long long* const* peer_flags, unsigned long long* my_flags, const int* group,
                                  int gsize, int me, int nranks, int phase, unsigned long long epoch, int* err,
                                  long long timeout_cycles) {
    const int i = threadIdx.x;
    if (i >= gsize) return;
    const int q = group[i];
    if (q == me) return;
    __threadfence_system();
    unsigned long long* dst = peer_flags[q] + size_t(phase) * nranks + me;
    const unsigned long long* src = my_flags + size_t(phase) * nranks + q;
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
    if (v == FLAG_POISON) {  // a peer gave up earlier: do not overwrite its poison with my epoch
        atomicExch(err, 100 + phase);
        return;
    }
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(epoch) : "memory");
    const long long t0 = clock64();
    while (true) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
        if (v == FLAG_POISON) {
            atomicExch(err, 100 + phase);
            break;
        }
        if (v >= epoch) break;
        if (timeout_cycles > 0 && clock64() - t0 > timeout_cycles) {
            atomicExch(err, 1 + phase);
            for (int ph = 0; ph < NPHASE; ++ph)
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flags[q] + size_t(ph) * nranks + me), "l"(FLAG_POISON) : "memory");
            break;
        }
        __nanosleep(64);
    }
}


}  // namespace dfft

using namespace dfft;

struct dfft_comm_s {
    int rank = 0, nranks = 1, device = 0;
    ncclComm_t nccl = nullptr;
    bool dry = false;  // geometry-only: plans never touch CUDA (dfft_comm_create_dry)
};

namespace dfft {

enum StepType { STEP_PASS = 0, STEP_RENDEZVOUS = 1, STEP_A2A = 2 };

struct Step {
    StepType type = STEP_PASS;
    const char* phase = nullptr;  // timer section recorded after this step
    const char* label = "";       // short name of the step ("z pass", "y pass", "rendezvous 2", ...)
    // PASS
    PassKind kind = PASS_C2C_CONTIG;
    int log2n = 0;
    FftParams prm{};
    int in_user = 0, out_user = 0;  // 1: patch seg[0].base with the caller's in, 2: with the caller's out
    // RENDEZVOUS
    int group = 0, phase_id = 0;  // group: 0 all, 1 first transposition group, 2 second
    // A2A
    std::vector<size_t> scount, soff, rcount, roff;  // elements, indexed by group member
    int send_slot = 0, recv_slot = 0;
    // overlapped schedules: which plan stream runs the step (0 = caller's stream, 1 = exchange stream,
    // 2 = follow-up stream), events to wait for before it and the event to record after it
    int stream = 0;
    std::vector<int> waits;
    int record = -1;
};

struct Schedule {
    std::vector<Step> steps;
    bool built = false;
    bool overlapped = false;  // uses the plan's auxiliary streams
    int nevents = 0;
};

}  // namespace dfft

struct dfft_plan_s {
    dfft_comm_t comm = nullptr;
    dfft_config cfg{};
    std::string bench_dir;
    Geometry g;
    int prec = DFFT_F64;
    size_t esize = 16;  // bytes per complex element
    int rank = 0, P = 1;
    size_t domain_bytes = 0;  // this rank (getDomainSize)
    size_t slot_bytes = 0;    // uniform over ranks
    int nslots = 2;
    size_t work_bytes = 0;
    void* work = nullptr;
    bool work_owned = false;
    bool any_direct = false;  // some transposition uses peer stores across >1 ranks
    bool direct1 = true, direct2 = true;
    // peer mapping
    std::vector<std::vector<void*>> slot_ptr;  // [slot][rank]
    std::vector<void*> opened;                 // IPC mappings to close
    unsigned long long* flags = nullptr;       // my flags (NPHASE * P)
    unsigned long long** peer_flags_d = nullptr;
    std::vector<void*> opened_flags;
    int* groups_d = nullptr;  // [3][P] group member lists on device
    std::vector<int> grp[3];
    int* err_d = nullptr;
    unsigned long long epoch = 0;
    unsigned long long ticket[4] = {0, 0, 0, 0};  // per-phase rendezvous counters (same sequence on every rank)
    cudaStream_t aux[2] = {nullptr, nullptr};     // exchange stream (high priority), follow-up stream
    std::vector<cudaEvent_t> sync_events;
    cudaEvent_t fork_ev = nullptr, join_ev[2] = {nullptr, nullptr};
    int xchg_ctas = 0;                            // SMs given to the exchange pass in overlapped schedules
    int tuned_seq[2] = {0, 0};                    // dfft_plan_tune: [fwd/inv] 1 = the sequential schedule won
    int tuned_chunks[2] = {0, 0};                 // dfft_plan_tune: [fwd/inv] z chunks of the winning overlapped schedule (0 = ovl_chunks)
    int tuned_groups_inv = 0;                     // dfft_plan_tune: plane groups of the inverse overlapped pencil schedule (0 = ovl_groups)
    int tuned_ctas[2] = {-2, -2};                 // dfft_plan_tune: [fwd/inv] exchange CTAs of the winning overlapped schedule (-2 = not tuned)
    std::string tune_report;
    int ovl_groups = 4, ovl_chunks = 4;           // overlapped schedules: plane groups of the z pass, z chunks of the y / x passes
    int blocked_ch = 0;                           // > 0: slab forward keeps the y->x intermediate as [b/CH][Nx][CH]
    int single_rank_layout = 1;                   // one rank: hand-over [ny][nzm/CH][nx][CH] (x innermost) instead of [nzm/CH][nx][ny][CH] (DFFT_N1_LAYOUT=0)
    int x_swz = 1;                                // tile-order blocking (log2 G) of passes that read the blocked hand-over layout
    int blocked_inv = 1;                          // the inverse x -> y hand-over is blocked as well (DFFT_BLOCKED_INV=0: plain)
    int xchg_tile_pref = 2;                       // tile preference of passes that store into other GPUs (2 = wide rows)
    long long rendezvous_timeout_cycles = 0;      // device clock cycles a rendezvous waits for a peer (0 = forever)
    int bulk_store = 0;                           // experimental: exchanging y pass stores with cp.async.bulk (DFFT_BULK_STORE=1)
    cudaStream_t own_stream = nullptr, last_stream = nullptr;  // last_stream: stream of the last exec
    cudaEvent_t entry_ev = nullptr;  // synchronous execs: "everything the caller queued on the legacy default stream"
    Tables tabs;
    // schedules: [fwd/inv][d-1]
    Schedule sched[2][3];
    // timing
    bool timing = false;
    std::vector<cudaEvent_t> events;
    std::vector<const char*> ev_names;
    std::vector<int> ev_is_fft;  // 1 fft pass, 0 exchange
    std::vector<const char*> ev_labels;
    int n_events_used = 0;
    // timeline of the last timed exec: one (begin, end) event pair per step, recorded on the step's own stream
    std::vector<cudaEvent_t> tl_events;
    std::vector<const char*> tl_labels;
    std::vector<int> tl_streams;
    int tl_used = 0;
    int last_launches = 0;
    int execs = 0;
    double init_ms = 0;          // "init" section of the reference's CSV
    int csv_warmup_left = 0;     // execs still to be skipped (Configurations::warmup_rounds)
    std::string csv_path;        // empty: no CSV
};

namespace dfft {

// element offset helper
static inline void* eptr(void* base, size_t elems, size_t esize) { return (char*)base + elems * esize; }

static View single_view(void* base, long long sA0, long long sA1, long long sN) {
    View v{};
    v.seg_of_n = nullptr;
    v.nseg = 1;
    v.sN = sN;
    v.seg[0].base = base;
    v.seg[0].sA0 = sA0;
    v.seg[0].sA1 = sA1;
    v.seg[0].n0 = 0;
    return v;
}

}  // namespace dfft

// =====================================================================================================
// Schedule construction
// =====================================================================================================
namespace dfft {

static int build_schedule(dfft_plan_s* p, int inverse, int d, Schedule& sc);
static int build_overlapped_slab(dfft_plan_s* p, int inverse, Schedule& sc);
static int build_overlapped_pencil(dfft_plan_s* p, int inverse, Schedule& sc);
static bool pencil_overlap_enabled();

}  // namespace dfft

static int plan_setup_memory(dfft_plan_s* p, void* user_device);
static void plan_release_memory(dfft_plan_s* p);

// -----------------------------------------------------------------------------------------------------
namespace dfft {

// Fill a segmented view over group members. f(q_pos, rank) returns the segment and the (view-wide) stride
// along the transformed axis.
struct SegN {
    Seg s;
    long long sN;
};
static bool g_view_error = false;
template <typename F>
static void seg_view(View& v, const unsigned char* table, const std::vector<int>& ranks, F f) {
    v.seg_of_n = table;
    v.nseg = int(ranks.size());
    for (size_t q = 0; q < ranks.size(); ++q) {
        SegN sn = f(int(q), ranks[q]);
        v.seg[q] = sn.s;
        if (q == 0) v.sN = sn.sN;
        else if (v.sN != sn.sN) g_view_error = true;
    }
    if (v.nseg == 1) v.seg_of_n = nullptr;
}

static SegN mkseg(void* base, long long sA0, long long sA1, long long sN, size_t n0) {
    SegN r{};
    r.s.base = base; r.s.sA0 = sA0; r.s.sA1 = sA1; r.s.n0 = int(n0);
    r.sN = sN;
    return r;
}

static int build_schedule(dfft_plan_s* p, int inverse, int d, Schedule& sc) {
    const Geometry& g = p->g;
    const int me = p->rank, P = p->P;
    const int pi = g.pi(me), pj = g.pj(me);
    const size_t es = p->esize;
    const bool c2c = g.transform == DFFT_C2C;
    Tables& T = p->tabs;
    sc.steps.clear();

    auto new_pass = [&](PassKind kind, size_t n, const char* phase, Step& s) -> int {
        s = Step();
        s.type = STEP_PASS;
        s.label = (kind == PASS_C2C_TILED) ? "strided pass" : (kind == PASS_R2C ? "z pass (R2C)" : (kind == PASS_C2R ? "z pass (C2R)" : "z pass"));
        s.kind = kind;
        s.phase = phase;
        s.log2n = ilog2_exact(n);
        if (s.log2n < 1 || s.log2n > MAX_LOG2N)
            return fail(DFFT_ERR_UNSUPPORTED, "axis length " + std::to_string(n) + " is not a supported power of two (2..8192)");
        s.prm.A0 = 1; s.prm.A1 = 1; s.prm.B = 1;
        s.prm.inverse = inverse;
        void* tw = nullptr;
        if (T.get_tw(s.log2n, &tw) != cudaSuccess) return fail(DFFT_ERR_CUDA, "twiddle table allocation failed");
        s.prm.tw = tw;
        s.prm.tw2 = nullptr;
        if (kind == PASS_R2C || kind == PASS_C2R) {
            void* tw2 = nullptr;
            if (T.get_tw2(s.log2n, &tw2) != cudaSuccess) return fail(DFFT_ERR_CUDA, "twiddle table allocation failed");
            s.prm.tw2 = tw2;
        }
        return DFFT_SUCCESS;
    };
    auto rendezvous = [&](int group, int phase_id, const char* phase) {
        Step s;
        s.type = STEP_RENDEZVOUS;
        s.label = group == 0 ? "entry rendezvous" : (group == 1 ? "rendezvous 1" : "rendezvous 2");
        s.group = group;
        s.phase_id = phase_id;
        s.phase = phase;
        sc.steps.push_back(s);
    };

    // groups: transposition 1 / 2 member lists (positions along the split axis)
    const std::vector<int>& G1 = p->grp[1];
    const std::vector<int>& G2 = p->grp[2];
    const bool dir1 = p->direct1 || G1.size() == 1;
    const bool dir2 = p->direct2 || G2.size() == 1;
    // slot ids
    int nxt = 0;
    int D1 = -1, D2 = -1, SS = -1, SR = -1;
    if (dir1) D1 = nxt++;
    if (dir2) D2 = nxt++;
    if (!dir1 || !dir2) { SS = nxt++; SR = nxt++; }
    auto slotp = [&](int s, int r) -> void* { return p->slot_ptr[s][r]; };

    if (p->any_direct) rendezvous(0, 0, nullptr);  // everyone has left the previous exec: slots are free

    const size_t nzc = g.nzc;
    const bool zyx = g.decomp == DFFT_SLAB_Z_THEN_YX;
    int rc;

    // sizes of this rank
    const size_t nx_i = g.sx.size[pi], x0_i = g.sx.start[pi];
    const size_t ny_j = zyx ? g.ny : g.sy.size[pj], y0_j = zyx ? 0 : g.sy.start[pj];
    const size_t nz_j = zyx ? g.sz.size[me] : g.sz.size[pj], z0_j = zyx ? g.sz.start[me] : g.sz.start[pj];
    const size_t oy_i = zyx ? g.ny : g.oy.size[pi], oy0_i = zyx ? 0 : g.oy.start[pi];
    (void)x0_i; (void)y0_j; (void)z0_j; (void)oy0_i;

    const unsigned char *tab_z = nullptr, *tab_y_in = nullptr, *tab_y_out = nullptr, *tab_x = nullptr;
    if (G1.size() > 1) {
        if (T.seg_table(g.sz, &tab_z) != cudaSuccess) return fail(DFFT_ERR_CUDA, "segment table");
        if (!zyx && T.seg_table(g.sy, &tab_y_in) != cudaSuccess) return fail(DFFT_ERR_CUDA, "segment table");
    }
    if (G2.size() > 1 || (zyx && G1.size() > 1)) {
        if (!zyx && T.seg_table(g.oy, &tab_y_out) != cudaSuccess) return fail(DFFT_ERR_CUDA, "segment table");
        if (T.seg_table(g.sx, &tab_x) != cudaSuccess) return fail(DFFT_ERR_CUDA, "segment table");
    }

    auto a2a_counts = [&](Step& s, const std::vector<int>& G, auto send_elems, auto recv_elems) {
        s.type = STEP_A2A;
        s.label = "nccl all-to-all";
        size_t so = 0, ro = 0;
        s.scount.resize(G.size()); s.soff.resize(G.size()); s.rcount.resize(G.size()); s.roff.resize(G.size());
        for (size_t q = 0; q < G.size(); ++q) {
            s.scount[q] = send_elems(int(q));
            s.rcount[q] = recv_elems(int(q));
            s.soff[q] = so; s.roff[q] = ro;
            so += s.scount[q]; ro += s.rcount[q];
        }
        s.send_slot = SS; s.recv_slot = SR;
    };

    if (!inverse) {
        // ------------------------------------------------------------------ forward
        // pass 1: z (contiguous), lines (x in nx_i, y in ny_j)
        Step s1;
        rc = new_pass(c2c ? PASS_C2C_CONTIG : PASS_R2C, c2c ? g.nz : g.nz / 2, zyx || g.decomp == DFFT_PENCIL ? "1D FFT Z-Direction" : nullptr, s1);
        if (rc) return rc;
        s1.prm.A0 = int(nx_i); s1.prm.A1 = int(ny_j);
        {   // input: caller's buffer; real lines of nz reals = nz/2 complex
            const long long pitch = c2c ? (long long)g.nz : (long long)(g.nz / 2);
            s1.prm.in = single_view(nullptr, pitch * (long long)ny_j, pitch, 1);
            s1.in_user = 1;
        }
        if (d == 1) {
            s1.prm.out = single_view(nullptr, (long long)(ny_j * nzc), (long long)nzc, 1);
            s1.out_user = 2;
            sc.steps.push_back(s1);
            sc.built = true;
            return DFFT_SUCCESS;
        }
        // transposition 1: scatter along z over G1
        Step x1;  // optional a2a
        const bool t1_a2a = !dir1;
        if (zyx) {
            // x split -> z split: dest q holds [nx][ny][nz_q]
            if (dir1) {
                seg_view(s1.prm.out, tab_z, G1, [&](int q, int r) {
                    const size_t nzq = g.sz.size[q];
                    return mkseg(eptr(slotp(D1, r), g.sx.start[me] * g.ny * nzq, es), (long long)(g.ny * nzq), (long long)nzq, 1, g.sz.start[q]);
                });
            } else {
                a2a_counts(x1, G1, [&](int q) { return nx_i * g.ny * g.sz.size[q]; }, [&](int q) { return g.sx.size[q] * g.ny * nz_j; });
                seg_view(s1.prm.out, tab_z, G1, [&](int q, int) {
                    const size_t nzq = g.sz.size[q];
                    return mkseg(eptr(slotp(SS, me), x1.soff[q], es), (long long)(g.ny * nzq), (long long)nzq, 1, g.sz.start[q]);
                });
            }
        } else {
            // pencil row: dest (i,q) holds [nx_i][ny][nz_q]
            if (dir1) {
                seg_view(s1.prm.out, tab_z, G1, [&](int q, int r) {
                    const size_t nzq = g.sz.size[q];
                    return mkseg(eptr(slotp(D1, r), y0_j * nzq, es), (long long)(g.ny * nzq), (long long)nzq, 1, g.sz.start[q]);
                });
            } else {
                a2a_counts(x1, G1, [&](int q) { return nx_i * ny_j * g.sz.size[q]; }, [&](int q) { return nx_i * g.sy.size[q] * nz_j; });
                seg_view(s1.prm.out, tab_z, G1, [&](int q, int) {
                    const size_t nzq = g.sz.size[q];
                    return mkseg(eptr(slotp(SS, me), x1.soff[q], es), (long long)(ny_j * nzq), (long long)nzq, 1, g.sz.start[q]);
                });
            }
        }
        if (g.decomp == DFFT_SLAB_ZY_THEN_X) s1.phase = nullptr;
        sc.steps.push_back(s1);
        if (t1_a2a) { x1.phase = zyx ? "Transpose (Finished All2All)" : "First Transpose (Finished All2All)"; x1.group = 1; sc.steps.push_back(x1); }
        else if (G1.size() > 1) rendezvous(1, 1, zyx ? "Transpose (Finished Receive)" : "First Transpose (Finished Receive)");

        // pass 2: y (tiled): a0 = x, n = y, b = z in nz_j
        Step s2;
        const size_t nx2 = zyx ? g.nx : nx_i;  // x extent held after transposition 1
        rc = new_pass(PASS_C2C_TILED, g.ny, g.decomp == DFFT_SLAB_ZY_THEN_X ? "2D FFT Y-Z-Direction" : (zyx ? nullptr : "1D FFT Y-Direction"), s2);
        if (rc) return rc;
        s2.label = "y pass";
        s2.prm.A0 = int(nx2); s2.prm.A1 = 1; s2.prm.B = int(nz_j);
        if (t1_a2a && !zyx) {
            seg_view(s2.prm.in, tab_y_in, G1, [&](int q, int) {
                const size_t nyq = g.sy.size[q];
                return mkseg(eptr(slotp(SR, me), x1.roff[q], es), (long long)(nyq * nz_j), 0, (long long)nz_j, g.sy.start[q]);
            });
        } else {
            void* base = t1_a2a ? slotp(SR, me) : slotp(D1, me);
            s2.prm.in = single_view(base, (long long)(g.ny * nz_j), 0, (long long)nz_j);
        }
        if (d == 2 || zyx) {
            if (d == 2) {
                s2.prm.out = single_view(nullptr, (long long)(g.ny * nz_j), 0, (long long)nz_j);
                s2.out_user = 2;
                sc.steps.push_back(s2);
                sc.built = true;
                return DFFT_SUCCESS;
            }
            // z_then_yx: y pass in place, then x pass into the caller's out
            s2.prm.out = s2.prm.in;
            if (s2.prm.in.nseg != 1) return fail(DFFT_ERR_STATE, "internal: z_then_yx gather view");
            sc.steps.push_back(s2);
            Step s3;
            rc = new_pass(PASS_C2C_TILED, g.nx, "2D FFT Y-X-Direction", s3);
            if (rc) return rc;
            s3.label = "x pass";
            s3.prm.A0 = 1; s3.prm.A1 = 1; s3.prm.B = int(g.ny * nz_j);
            s3.prm.in = single_view(s2.prm.in.seg[0].base, 0, 0, (long long)(g.ny * nz_j));
            s3.prm.out = single_view(nullptr, 0, 0, (long long)(g.ny * nz_j));
            s3.out_user = 2;
            sc.steps.push_back(s3);
            sc.built = true;
            return DFFT_SUCCESS;
        }
        // transposition 2: scatter along y over G2 (column of the grid, or all ranks for the slab)
        Step x2;
        const bool t2_a2a = !dir2;
        const size_t CH = (dir2 && !t1_a2a) ? size_t(p->blocked_ch) : 0;  // slab and pencil (second transposition)
        const size_t rem = CH ? nz_j % CH : 0, nzm = nz_j - rem;
        // every rank emits the same step list: when any rank of the grid has leftover columns the others run the
        // tail steps as empty passes (B = 0)
        bool any_rem = false;
        if (CH) {
            if (g.decomp == DFFT_PENCIL) { for (size_t v : g.sz.size) any_rem = any_rem || (v % CH != 0); }
            else any_rem = rem != 0;
        }
        Step s2t;
        bool have_tail = false;
        if (dir2 && CH) {
            // blocked hand-over: receiver q holds [nzc/CH][nx][ny_q][CH] — the rows a y-pass tile sends to one
            // destination (consecutive y) are adjacent, so every warp store is 512 contiguous bytes
            s2.prm.A1 = int(nzm / CH); s2.prm.B = int(CH);
            s2.prm.in.seg[0].sA1 = (long long)CH;
            // rows are only adjacent when the tile is CH wide: with remote peers take the wide tile (TB = CH = 8 for
            // the 1024-point f64 pass; measured 449 -> 700 GB/s per direction), locally the faster narrow one
            s2.prm.tile_pref = G2.size() > 1 ? p->xchg_tile_pref : 1;
            s2.prm.bulk_out = G2.size() > 1 ? p->bulk_store : 0;
            if (G2.size() == 1 && p->single_rank_layout) {
                // one rank: keep x innermost, [ny][nzm/CH][nx][CH] — every x-pass tile is one contiguous block
                // (measured 6254 GB/s for the 512-point x pass, profiles/r01_bench_n1.json)
                s2.prm.out = single_view(slotp(D2, me), (long long)CH, (long long)(g.nx * CH), (long long)((nzm / CH) * g.nx * CH));
            } else {
                seg_view(s2.prm.out, tab_y_out, G2, [&](int q, int r) {
                    const size_t nyq = g.oy.size[q];
                    return mkseg(eptr(slotp(D2, r), x0_i * nyq * CH, es), (long long)(nyq * CH), (long long)(g.nx * nyq * CH), (long long)CH, g.oy.start[q]);
                });
            }
            if (any_rem) {  // leftover columns z in [nzm, nzc): plain layout [nx][ny_q][rem] behind the blocked part
                s2t = s2;
                have_tail = true;
                s2t.label = "y pass (tail)";
                s2t.prm.bulk_out = 0;
                s2t.phase = nullptr;
                s2t.prm.A1 = 1; s2t.prm.B = int(rem);
                s2t.prm.in.seg[0].base = eptr(s2.prm.in.seg[0].base, nzm, es);
                s2t.prm.in.seg[0].sA1 = 0;
                seg_view(s2t.prm.out, tab_y_out, G2, [&](int q, int r) {
                    const size_t nyq = g.oy.size[q];
                    return mkseg(eptr(slotp(D2, r), g.nx * nyq * nzm + x0_i * nyq * rem, es), (long long)(nyq * rem), 0, (long long)rem, g.oy.start[q]);
                });
            }
        } else if (dir2) {
            if (G2.size() > 1) s2.prm.tile_pref = p->xchg_tile_pref;
            seg_view(s2.prm.out, tab_y_out, G2, [&](int q, int r) {
                const size_t nyq = g.oy.size[q];
                return mkseg(eptr(slotp(D2, r), x0_i * nyq * nz_j, es), (long long)(nyq * nz_j), 0, (long long)nz_j, g.oy.start[q]);
            });
        } else {
            a2a_counts(x2, G2, [&](int q) { return nx_i * g.oy.size[q] * nz_j; }, [&](int q) { return g.sx.size[q] * oy_i * nz_j; });
            seg_view(s2.prm.out, tab_y_out, G2, [&](int q, int) {
                const size_t nyq = g.oy.size[q];
                return mkseg(eptr(slotp(SS, me), x2.soff[q], es), (long long)(nyq * nz_j), 0, (long long)nz_j, g.oy.start[q]);
            });
        }
        if (have_tail) { const char* ph = s2.phase; s2.phase = nullptr; sc.steps.push_back(s2); s2t.phase = ph; sc.steps.push_back(s2t); }
        else sc.steps.push_back(s2);
        const bool slab = g.decomp == DFFT_SLAB_ZY_THEN_X;
        if (t2_a2a) { x2.phase = slab ? "Transpose (Finished All2All)" : "Second Transpose (Finished All2All)"; x2.group = 2; sc.steps.push_back(x2); }
        else if (G2.size() > 1) rendezvous(2, 2, slab ? "Transpose (Finished Receive)" : "Second Transpose (Finished Receive)");

        // pass 3: x (tiled): n = x, b = (y in oy_i, z in nz_j); received blocks are x-major => contiguous
        Step s3;
        rc = new_pass(PASS_C2C_TILED, g.nx, "1D FFT X-Direction", s3);
        if (rc) return rc;
        s3.label = "x pass";
        if (CH) {
            // in: [nzc/CH][nx][oy_i][CH]: a0 = y_loc, a1 = z chunk, n = x
            s3.prm.A0 = int(oy_i); s3.prm.A1 = int(nzm / CH); s3.prm.B = int(CH);
            if (G2.size() == 1 && p->single_rank_layout)
                s3.prm.in = single_view(slotp(D2, me), (long long)((nzm / CH) * g.nx * CH), (long long)(g.nx * CH), (long long)CH);
            else
                s3.prm.in = single_view(slotp(D2, me), (long long)CH, (long long)(g.nx * oy_i * CH), (long long)(oy_i * CH));
            s3.prm.out = single_view(nullptr, (long long)nz_j, (long long)CH, (long long)(oy_i * nz_j));
            s3.prm.tile_swz = p->x_swz;
            if (any_rem) {
                Step s3t = s3;
                s3t.label = "x pass (tail)";
                s3t.phase = nullptr;
                s3t.prm.A0 = int(oy_i); s3t.prm.A1 = 1; s3t.prm.B = int(rem);
                s3t.prm.in = single_view(eptr(slotp(D2, me), g.nx * oy_i * nzm, es), (long long)rem, 0, (long long)(oy_i * rem));
                s3t.prm.out = single_view((void*)(size_t)(nzm * es), (long long)nz_j, 0, (long long)(oy_i * nz_j));
                s3t.out_user = 2;
                sc.steps.push_back(s3t);
            }
        } else {
            s3.prm.A0 = 1; s3.prm.A1 = 1; s3.prm.B = int(oy_i * nz_j);
            s3.prm.in = single_view(t2_a2a ? slotp(SR, me) : slotp(D2, me), 0, 0, (long long)(oy_i * nz_j));
            s3.prm.out = single_view(nullptr, 0, 0, (long long)(oy_i * nz_j));
        }
        s3.out_user = 2;
        sc.steps.push_back(s3);
        sc.built = true;
        return DFFT_SUCCESS;
    }

    // ---------------------------------------------------------------------- inverse
    // Reverse order: x, transposition 2', y, transposition 1', z.  Partial d: only the last d of them.
    const PassKind zkind = c2c ? PASS_C2C_CONTIG : PASS_C2R;
    const size_t zlen = c2c ? g.nz : g.nz / 2;
    const long long zpitch_out = c2c ? (long long)g.nz : (long long)(g.nz / 2);  // output line pitch in complex units

    if (d == 1) {
        Step s;
        rc = new_pass(zkind, zlen, "1D FFT Z-Direction", s);
        if (rc) return rc;
        s.prm.A0 = int(nx_i); s.prm.A1 = int(ny_j);
        s.prm.in = single_view(nullptr, (long long)(ny_j * nzc), (long long)nzc, 1);
        s.in_user = 1;
        s.prm.out = single_view(nullptr, zpitch_out * (long long)ny_j, zpitch_out, 1);
        s.out_user = 2;
        sc.steps.push_back(s);
        sc.built = true;
        return DFFT_SUCCESS;
    }

    if (zyx) {
        // input [nx][ny][nz_me]: y pass (to a slot), x pass scattering along x to the owners of x,
        // then z inverse on [nx_me][ny][nzc].
        Step sy_;
        rc = new_pass(PASS_C2C_TILED, g.ny, nullptr, sy_);
        if (rc) return rc;
        sy_.label = "y pass";
        const int W = dir1 ? D2 : SR;  // a local scratch slot that is not the transposition target
        sy_.prm.A0 = int(g.nx); sy_.prm.B = int(nz_j);
        sy_.prm.in = single_view(nullptr, (long long)(g.ny * nz_j), 0, (long long)nz_j);
        sy_.in_user = 1;
        sy_.prm.out = single_view(slotp(W, me), (long long)(g.ny * nz_j), 0, (long long)nz_j);
        sc.steps.push_back(sy_);
        Step sx_;
        rc = new_pass(PASS_C2C_TILED, g.nx, "2D FFT Y-X-Direction", sx_);
        if (rc) return rc;
        sx_.label = "x pass";
        sx_.prm.A0 = 1; sx_.prm.A1 = int(g.ny); sx_.prm.B = int(nz_j);
        sx_.prm.in = single_view(slotp(W, me), 0, (long long)nz_j, (long long)(g.ny * nz_j));
        Step xa;
        if (dir1) {
            seg_view(sx_.prm.out, tab_x, G1, [&](int q, int r) {
                return mkseg(eptr(slotp(D1, r), z0_j, es), 0, (long long)nzc, (long long)(g.ny * nzc), g.sx.start[q]);
            });
        } else {
            a2a_counts(xa, G1, [&](int q) { return g.sx.size[q] * g.ny * nz_j; }, [&](int q) { return nx_i * g.ny * g.sz.size[q]; });
            seg_view(sx_.prm.out, tab_x, G1, [&](int q, int) {
                return mkseg(eptr(slotp(SS, me), xa.soff[q], es), 0, (long long)nz_j, (long long)(g.ny * nz_j), g.sx.start[q]);
            });
        }
        sc.steps.push_back(sx_);
        if (!dir1) { xa.phase = "Transpose (Finished All2All)"; xa.group = 1; sc.steps.push_back(xa); }
        else if (G1.size() > 1) rendezvous(1, 1, "Transpose (Finished Receive)");
        Step sz_;
        rc = new_pass(zkind, zlen, "1D FFT Z-Direction", sz_);
        if (rc) return rc;
        sz_.prm.A0 = int(nx_i); sz_.prm.A1 = int(g.ny);
        if (dir1) sz_.prm.in = single_view(slotp(D1, me), (long long)(g.ny * nzc), (long long)nzc, 1);
        else
            seg_view(sz_.prm.in, tab_z, G1, [&](int q, int) {
                const size_t nzq = g.sz.size[q];
                return mkseg(eptr(slotp(SR, me), xa.roff[q], es), (long long)(g.ny * nzq), (long long)nzq, 1, g.sz.start[q]);
            });
        sz_.prm.out = single_view(nullptr, zpitch_out * (long long)g.ny, zpitch_out, 1);
        sz_.out_user = 2;
        sc.steps.push_back(sz_);
        sc.built = true;
        return DFFT_SUCCESS;
    }

    // pencil / slab zy_x inverse
    const bool slab = g.decomp == DFFT_SLAB_ZY_THEN_X;
    // Blocked hand-over for the inverse as well (mirror of the forward one): the x pass stores into the receivers as
    // [nz_j/CH][ny][nx_q][CH] (+ a plain tail of nz_j % CH columns), so the rows a tile sends to one destination —
    // consecutive x — are adjacent (one contiguous run per destination), and the y pass reads rows nx_i*CH apart
    // instead of one row per (y, z) plane.  DFFT_BLOCKED_INV=0 keeps the plain layout.
    const size_t CHI = (d == 3 && dir2 && dir1 && p->blocked_inv) ? size_t(p->blocked_ch) : 0;
    if (CHI) {
        const size_t CH = CHI;
        const size_t rem = nz_j % CH, nzm = nz_j - rem;
        bool any_rem = false;
        if (g.decomp == DFFT_PENCIL) { for (size_t v : g.sz.size) any_rem = any_rem || (v % CH != 0); }
        else any_rem = rem != 0;
        // x pass: in = caller's [nx][oy_i][nz_j]; tile = (y_loc, z block); n = x
        Step s3;
        rc = new_pass(PASS_C2C_TILED, g.nx, "1D FFT X-Direction", s3);
        if (rc) return rc;
        s3.label = "x pass";
        s3.prm.A0 = int(oy_i); s3.prm.A1 = int(nzm / CH); s3.prm.B = int(CH);
        s3.prm.in = single_view(nullptr, (long long)nz_j, (long long)CH, (long long)(oy_i * nz_j));
        s3.in_user = 1;
        s3.prm.tile_pref = G2.size() > 1 ? p->xchg_tile_pref : 1;
        s3.prm.bulk_out = G2.size() > 1 ? p->bulk_store : 0;
        seg_view(s3.prm.out, tab_x, G2, [&](int q, int r) {
            const size_t nxq = g.sx.size[q];
            return mkseg(eptr(slotp(D2, r), oy0_i * nxq * CH, es), (long long)(nxq * CH), (long long)(g.ny * nxq * CH), (long long)CH, g.sx.start[q]);
        });
        Step s3t;
        if (any_rem) {  // leftover columns: plain [nx_q][ny][rem] behind the blocked part
            s3t = s3;
            s3t.label = "x pass (tail)";
            s3t.phase = nullptr;
            s3t.prm.bulk_out = 0;
            s3t.prm.A0 = int(oy_i); s3t.prm.A1 = 1; s3t.prm.B = int(rem);
            s3t.prm.in = single_view((void*)(size_t)(nzm * es), (long long)nz_j, 0, (long long)(oy_i * nz_j));
            seg_view(s3t.prm.out, tab_x, G2, [&](int q, int r) {
                const size_t nxq = g.sx.size[q];
                return mkseg(eptr(slotp(D2, r), g.ny * nxq * nzm + oy0_i * rem, es), (long long)rem, 0, (long long)(g.ny * rem), g.sx.start[q]);
            });
            const char* ph = s3.phase; s3.phase = nullptr; s3t.phase = ph;
            sc.steps.push_back(s3);
            sc.steps.push_back(s3t);
        } else {
            sc.steps.push_back(s3);
        }
        if (G2.size() > 1) rendezvous(2, 2, slab ? "Transpose (Finished Receive)" : "Second Transpose (Finished Receive)");
        // y pass: in = my slot, blocked [nz_j/CH][ny][nx_i][CH]; tile = (x_loc, z block); n = y
        Step s2;
        rc = new_pass(PASS_C2C_TILED, g.ny, slab ? nullptr : "1D FFT Y-Direction", s2);
        if (rc) return rc;
        s2.label = "y pass";
        s2.prm.A0 = int(nx_i); s2.prm.A1 = int(nzm / CH); s2.prm.B = int(CH);
        s2.prm.in = single_view(slotp(D2, me), (long long)CH, (long long)(g.ny * nx_i * CH), (long long)(nx_i * CH));
        if (G1.size() > 1) s2.prm.tile_pref = p->xchg_tile_pref;
        else s2.prm.tile_pref = 1;
        seg_view(s2.prm.out, tab_y_in, G1, [&](int q, int r) {
            const size_t nyq = g.sy.size[q];
            return mkseg(eptr(slotp(D1, r), z0_j, es), (long long)(nyq * nzc), (long long)CH, (long long)nzc, g.sy.start[q]);
        });
        if (any_rem) {
            Step s2t = s2;
            s2t.label = "y pass (tail)";
            s2t.phase = nullptr;
            s2t.prm.A0 = int(nx_i); s2t.prm.A1 = 1; s2t.prm.B = int(rem);
            s2t.prm.in = single_view(eptr(slotp(D2, me), g.ny * nx_i * nzm, es), (long long)(g.ny * rem), 0, (long long)rem);
            seg_view(s2t.prm.out, tab_y_in, G1, [&](int q, int r) {
                const size_t nyq = g.sy.size[q];
                return mkseg(eptr(slotp(D1, r), z0_j + nzm, es), (long long)(nyq * nzc), 0, (long long)nzc, g.sy.start[q]);
            });
            const char* ph = s2.phase; s2.phase = nullptr; s2t.phase = ph;
            sc.steps.push_back(s2);
            sc.steps.push_back(s2t);
        } else {
            sc.steps.push_back(s2);
        }
        if (G1.size() > 1) rendezvous(1, 1, "First Transpose (Finished Receive)");
        Step s1;
        rc = new_pass(zkind, zlen, slab ? "2D FFT Y-Z-Direction" : "1D FFT Z-Direction", s1);
        if (rc) return rc;
        s1.prm.A0 = int(nx_i); s1.prm.A1 = int(ny_j);
        s1.prm.in = single_view(slotp(D1, me), (long long)(ny_j * nzc), (long long)nzc, 1);
        s1.prm.out = single_view(nullptr, zpitch_out * (long long)ny_j, zpitch_out, 1);
        s1.out_user = 2;
        sc.steps.push_back(s1);
        sc.built = true;
        return DFFT_SUCCESS;
    }
    Step x2;
    bool have_in_slot = false;  // whether the y pass input is in a slot (d == 3) or the caller's buffer (d == 2)
    if (d == 3) {
        Step s3;
        rc = new_pass(PASS_C2C_TILED, g.nx, "1D FFT X-Direction", s3);
        if (rc) return rc;
        s3.label = "x pass";
        s3.prm.A0 = 1; s3.prm.A1 = 1; s3.prm.B = int(oy_i * nz_j);
        s3.prm.in = single_view(nullptr, 0, 0, (long long)(oy_i * nz_j));
        s3.in_user = 1;
        if (dir2 && G2.size() > 1) s3.prm.tile_pref = p->xchg_tile_pref;
        if (dir2) {
            // dest (q,j) holds [nx_q][ny][nz_j]; my rows y in [oy0_i, +oy_i)
            seg_view(s3.prm.out, tab_x, G2, [&](int q, int r) {
                return mkseg(eptr(slotp(D2, r), oy0_i * nz_j, es), 0, 0, (long long)(g.ny * nz_j), g.sx.start[q]);
            });
        } else {
            a2a_counts(x2, G2, [&](int q) { return g.sx.size[q] * oy_i * nz_j; }, [&](int q) { return nx_i * g.oy.size[q] * nz_j; });
            seg_view(s3.prm.out, tab_x, G2, [&](int q, int) {
                return mkseg(eptr(slotp(SS, me), x2.soff[q], es), 0, 0, (long long)(oy_i * nz_j), g.sx.start[q]);
            });
        }
        sc.steps.push_back(s3);
        if (!dir2) { x2.phase = slab ? "Transpose (Finished All2All)" : "Second Transpose (Finished All2All)"; x2.group = 2; sc.steps.push_back(x2); }
        else if (G2.size() > 1) rendezvous(2, 2, slab ? "Transpose (Finished Receive)" : "Second Transpose (Finished Receive)");
        have_in_slot = true;
    }
    // y pass on [nx_i][ny][nz_j]
    Step s2;
    rc = new_pass(PASS_C2C_TILED, g.ny, slab ? nullptr : "1D FFT Y-Direction", s2);
    if (rc) return rc;
    s2.label = "y pass";
    s2.prm.A0 = int(nx_i); s2.prm.A1 = 1; s2.prm.B = int(nz_j);
    if (!have_in_slot) {
        s2.prm.in = single_view(nullptr, (long long)(g.ny * nz_j), 0, (long long)nz_j);
        s2.in_user = 1;
    } else if (dir2) {
        s2.prm.in = single_view(slotp(D2, me), (long long)(g.ny * nz_j), 0, (long long)nz_j);
    } else {
        seg_view(s2.prm.in, tab_y_out, G2, [&](int q, int) {
            const size_t nyq = g.oy.size[q];
            return mkseg(eptr(slotp(SR, me), x2.roff[q], es), (long long)(nyq * nz_j), 0, (long long)nz_j, g.oy.start[q]);
        });
    }
    Step x1;
    if (dir1 && G1.size() > 1) s2.prm.tile_pref = p->xchg_tile_pref;
    if (dir1) {
        // dest (i,q) holds [nx_i][ny_q][nzc]; my columns z in [z0_j, +nz_j)
        seg_view(s2.prm.out, tab_y_in, G1, [&](int q, int r) {
            const size_t nyq = g.sy.size[q];
            return mkseg(eptr(slotp(D1, r), z0_j, es), (long long)(nyq * nzc), 0, (long long)nzc, g.sy.start[q]);
        });
    } else {
        a2a_counts(x1, G1, [&](int q) { return nx_i * g.sy.size[q] * nz_j; }, [&](int q) { return nx_i * ny_j * g.sz.size[q]; });
        seg_view(s2.prm.out, tab_y_in, G1, [&](int q, int) {
            const size_t nyq = g.sy.size[q];
            return mkseg(eptr(slotp(SS, me), x1.soff[q], es), (long long)(nyq * nz_j), 0, (long long)nz_j, g.sy.start[q]);
        });
    }
    sc.steps.push_back(s2);
    if (!dir1) { x1.phase = "First Transpose (Finished All2All)"; x1.group = 1; sc.steps.push_back(x1); }
    else if (G1.size() > 1) rendezvous(1, 1, "First Transpose (Finished Receive)");
    // z pass on [nx_i][ny_j][nzc]
    Step s1;
    rc = new_pass(zkind, zlen, slab ? "2D FFT Y-Z-Direction" : "1D FFT Z-Direction", s1);
    if (rc) return rc;
    s1.prm.A0 = int(nx_i); s1.prm.A1 = int(ny_j);
    if (dir1) s1.prm.in = single_view(slotp(D1, me), (long long)(ny_j * nzc), (long long)nzc, 1);
    else
        seg_view(s1.prm.in, tab_z, G1, [&](int q, int) {
            const size_t nzq = g.sz.size[q];
            return mkseg(eptr(slotp(SR, me), x1.roff[q], es), (long long)(ny_j * nzq), (long long)nzq, 1, g.sz.start[q]);
        });
    s1.prm.out = single_view(nullptr, zpitch_out * (long long)ny_j, zpitch_out, 1);
    s1.out_user = 2;
    sc.steps.push_back(s1);
    sc.built = true;
    (void)P;
    return DFFT_SUCCESS;
}

}  // namespace dfft

namespace dfft {

// Overlapped slab (ZY_Then_X) schedule, Peer2Peer, SendMethod Streams.
//
// The reference's Streams variant overlaps pack(q+1) with send(q) (mpicufft_slab.cpp:398-415).  Here the
// exchange is the y pass itself (NVLink-bound), so the overlap is between whole FFT passes:
//   stream 0 (caller) : z pass, plane group by plane group                       (HBM-bound)
//   stream 1 (exchange, high priority): y pass per (plane group, z chunk), persistent on `xchg_ctas`
//                       SMs, storing straight into the peers' slots              (NVLink-bound)
//   stream 2 (follow-up): per z chunk: rendezvous with the peers, then the x pass of that chunk (HBM-bound)
// The y pass of chunk c+1 runs while the x pass consumes chunk c, and while later plane groups are still in
// the z pass.  Inverse: x pass per z chunk scatters (stream 1), y pass per chunk follows (stream 2), the z
// pass runs last on the caller's stream.
static int exchange_ctas(const dfft_plan_s* p, int inverse) {
    const int t = p->tuned_ctas[inverse ? 1 : 0];
    return t != -2 ? t : p->xchg_ctas;
}

static int build_overlapped_slab(dfft_plan_s* p, int inverse, Schedule& sc) {
    const Geometry& g = p->g;
    const int me = p->rank;
    const size_t es = p->esize;
    const bool c2c = g.transform == DFFT_C2C;
    Tables& T = p->tabs;
    sc.steps.clear();
    sc.overlapped = true;
    const size_t nzc = g.nzc, ny = g.ny, nx = g.nx;
    const size_t nx_p = g.sx.size[me], x0 = g.sx.start[me];
    const size_t oy_me = g.oy.size[me], oy0_me = g.oy.start[me];
    const std::vector<int>& G2 = p->grp[2];
    const int D1 = 0, D2 = 1;
    auto slotp = [&](int s_, int r) -> void* { return p->slot_ptr[s_][r]; };
    int nev = 0;

    auto new_pass = [&](PassKind kind, size_t n, const char* label, Step& s) -> int {
        s = Step();
        s.type = STEP_PASS;
        s.kind = kind;
        s.label = label;
        s.log2n = ilog2_exact(n);
        if (s.log2n < 1 || s.log2n > MAX_LOG2N) return fail(DFFT_ERR_UNSUPPORTED, "unsupported axis length");
        s.prm.A0 = 1; s.prm.A1 = 1; s.prm.B = 1;
        s.prm.inverse = inverse;
        void* tw = nullptr;
        if (T.get_tw(s.log2n, &tw) != cudaSuccess) return fail(DFFT_ERR_CUDA, "twiddle table allocation failed");
        s.prm.tw = tw;
        if (kind == PASS_R2C || kind == PASS_C2R) {
            void* tw2 = nullptr;
            if (T.get_tw2(s.log2n, &tw2) != cudaSuccess) return fail(DFFT_ERR_CUDA, "twiddle table allocation failed");
            s.prm.tw2 = tw2;
        }
        return DFFT_SUCCESS;
    };
    auto rendezvous = [&](int group, int phase_id, int stream) {
        Step s;
        s.type = STEP_RENDEZVOUS;
        s.label = group == 0 ? "entry rendezvous" : "rendezvous 2";
        s.group = group;
        s.phase_id = phase_id;
        s.stream = stream;
        return s;
    };

    // plane groups (of my x planes) and z chunks
    Split groups, chunks;
    size_t min_nx = nx;  // the same number of plane groups (and steps) on every rank, also for uneven splits of x
    for (size_t v : g.sx.size) min_nx = std::min(min_nx, v);
    const size_t NG = std::min<size_t>(size_t(p->ovl_groups), std::max<size_t>(min_nx, 1));
    const size_t want_chunks = size_t(p->tuned_chunks[inverse ? 1 : 0] > 0 ? p->tuned_chunks[inverse ? 1 : 0] : p->ovl_chunks);
    const size_t NS = nzc >= 32 * want_chunks ? want_chunks : (nzc >= 32 ? 2 : 1);
    groups.make(nx_p, NG);
    const size_t CH = (inverse && !p->blocked_inv) ? 0 : size_t(p->blocked_ch);  // blocked hand-over, both directions
    const size_t rem = CH ? nzc % CH : 0, nzm = nzc - rem;
    if (CH) {  // z chunks are whole multiples of the block width; the Nzc % CH leftover columns ride with the last chunk
        Split u;
        u.make(nzm / CH, std::min<size_t>(NS, nzm / CH));
        chunks.size.clear(); chunks.start.clear();
        for (size_t c = 0; c < u.size.size(); ++c) { chunks.size.push_back(u.size[c] * CH); chunks.start.push_back(u.start[c] * CH); }
    } else {
        chunks.make(nzc, NS);
    }
    const size_t NSc = chunks.size.size();
    const unsigned char *tab_y = nullptr, *tab_x = nullptr;
    if (T.seg_table(g.oy, &tab_y) != cudaSuccess || T.seg_table(g.sx, &tab_x) != cudaSuccess) return fail(DFFT_ERR_CUDA, "segment table");

    sc.steps.push_back(rendezvous(0, 0, 0));  // everyone has left the previous exec
    // the auxiliary streams fork from the caller's stream at the start of run_schedule, i.e. BEFORE the
    // entry rendezvous: make them wait for it explicitly
    const int ev_entry = nev++;
    sc.steps.back().record = ev_entry;
    int rc;

    if (!inverse) {
        std::vector<int> ev_z(NG), ev_y(NSc);
        for (size_t gi = 0; gi < NG; ++gi) {
            Step s;
            rc = new_pass(c2c ? PASS_C2C_CONTIG : PASS_R2C, c2c ? g.nz : g.nz / 2, c2c ? "z pass" : "z pass (R2C)", s);
            if (rc) return rc;
            const size_t pl0 = groups.start[gi], npl = groups.size[gi];
            const long long pitch = c2c ? (long long)g.nz : (long long)(g.nz / 2);
            s.prm.A0 = int(npl); s.prm.A1 = int(ny);
            s.prm.in = single_view((void*)(size_t)(pl0 * ny * pitch * es), pitch * (long long)ny, pitch, 1);
            s.in_user = 1;
            s.prm.out = single_view(eptr(slotp(D1, me), pl0 * ny * nzc, es), (long long)(ny * nzc), (long long)nzc, 1);
            s.stream = 0;
            s.record = ev_z[gi] = nev++;
            sc.steps.push_back(s);
        }
        for (size_t c = 0; c < NSc; ++c) {
            const size_t z0 = chunks.start[c], zc = chunks.size[c];
            for (size_t gi = 0; gi < NG; ++gi) {
                Step s;
                rc = new_pass(PASS_C2C_TILED, ny, "y pass", s);
                if (rc) return rc;
                const size_t pl0 = groups.start[gi], npl = groups.size[gi];
                if (CH) {
                    s.prm.A0 = int(npl); s.prm.A1 = int(zc / CH); s.prm.B = int(CH);
                    s.prm.in = single_view(eptr(slotp(D1, me), pl0 * ny * nzc + z0, es), (long long)(ny * nzc), (long long)CH, (long long)nzc);
                    s.prm.tile_pref = p->xchg_tile_pref;
                    s.prm.bulk_out = p->bulk_store;
                    seg_view(s.prm.out, tab_y, G2, [&](int q, int r) {
                        const size_t nyq = g.oy.size[q];
                        return mkseg(eptr(slotp(D2, r), ((z0 / CH) * nx + x0 + pl0) * nyq * CH, es), (long long)(nyq * CH), (long long)(nx * nyq * CH),
                                     (long long)CH, g.oy.start[q]);
                    });
                } else {
                s.prm.A0 = int(npl); s.prm.A1 = 1; s.prm.B = int(zc);
                s.prm.tile_pref = p->xchg_tile_pref;
                s.prm.in = single_view(eptr(slotp(D1, me), pl0 * ny * nzc + z0, es), (long long)(ny * nzc), 0, (long long)nzc);
                seg_view(s.prm.out, tab_y, G2, [&](int q, int r) {
                    const size_t nyq = g.oy.size[q];
                    return mkseg(eptr(slotp(D2, r), (x0 + pl0) * nyq * nzc + z0, es), (long long)(nyq * nzc), 0, (long long)nzc, g.oy.start[q]);
                });
                }
                s.prm.max_ctas = exchange_ctas(p, inverse);
                s.stream = 1;
                if (c == 0) s.waits.push_back(ev_z[gi]);
                if (c == 0 && gi == 0) s.waits.push_back(ev_entry);
                const bool tail_here = CH && rem && c + 1 == NSc;
                if (gi + 1 == NG && !tail_here) s.record = ev_y[c] = nev++;
                sc.steps.push_back(s);
                if (tail_here) {
                    Step t = s;
                    t.label = "y pass (tail)";
                    t.prm.bulk_out = 0;
                    t.waits.clear();
                    t.record = -1;
                    t.prm.A0 = int(npl); t.prm.A1 = 1; t.prm.B = int(rem);
                    t.prm.in = single_view(eptr(slotp(D1, me), pl0 * ny * nzc + nzm, es), (long long)(ny * nzc), 0, (long long)nzc);
                    seg_view(t.prm.out, tab_y, G2, [&](int q, int r) {
                        const size_t nyq = g.oy.size[q];
                        return mkseg(eptr(slotp(D2, r), nx * nyq * nzm + (x0 + pl0) * nyq * rem, es), (long long)(nyq * rem), 0, (long long)rem, g.oy.start[q]);
                    });
                    if (gi + 1 == NG) t.record = ev_y[c] = nev++;
                    sc.steps.push_back(t);
                }
            }
        }
        for (size_t c = 0; c < NSc; ++c) {
            const size_t z0 = chunks.start[c], zc = chunks.size[c];
            Step r = rendezvous(2, 2, 2);
            r.waits.push_back(ev_y[c]);
            sc.steps.push_back(r);
            Step s;
            rc = new_pass(PASS_C2C_TILED, nx, "x pass", s);
            if (rc) return rc;
            if (CH) {
                // in: [nzc/CH][nx][oy_me][CH]; a0 = y_loc, a1 = z chunk within this z range, n = x
                s.prm.A0 = int(oy_me); s.prm.A1 = int(zc / CH); s.prm.B = int(CH);
                s.prm.in = single_view(eptr(slotp(D2, me), (z0 / CH) * nx * oy_me * CH, es), (long long)CH, (long long)(nx * oy_me * CH), (long long)(oy_me * CH));
                s.prm.out = single_view((void*)(size_t)(z0 * es), (long long)nzc, (long long)CH, (long long)(oy_me * nzc));
                s.prm.tile_swz = p->x_swz;
            } else {
            s.prm.A0 = 1; s.prm.A1 = int(oy_me); s.prm.B = int(zc);
            s.prm.in = single_view(eptr(slotp(D2, me), z0, es), 0, (long long)nzc, (long long)(oy_me * nzc));
            s.prm.out = single_view((void*)(size_t)(z0 * es), 0, (long long)nzc, (long long)(oy_me * nzc));
            }
            s.out_user = 2;
            s.stream = 2;
            sc.steps.push_back(s);
            if (CH && rem && c + 1 == NSc) {
                Step t = s;
                t.label = "x pass (tail)";
                t.prm.A0 = int(oy_me); t.prm.A1 = 1; t.prm.B = int(rem);
                t.prm.in = single_view(eptr(slotp(D2, me), nx * oy_me * nzm, es), (long long)rem, 0, (long long)(oy_me * rem));
                t.prm.out = single_view((void*)(size_t)(nzm * es), (long long)nzc, 0, (long long)(oy_me * nzc));
                sc.steps.push_back(t);
            }
        }
    } else {
        std::vector<int> ev_x(NSc), ev_yi(NSc);
        for (size_t c = 0; c < NSc; ++c) {
            const size_t z0 = chunks.start[c], zc = chunks.size[c];
            Step s;
            rc = new_pass(PASS_C2C_TILED, nx, "x pass", s);
            if (rc) return rc;
            s.in_user = 1;
            if (CH) {
                // dest q holds [nzc/CH][ny][nx_q][CH]: the rows a tile sends to one destination (consecutive x) are adjacent
                s.prm.A0 = int(oy_me); s.prm.A1 = int(zc / CH); s.prm.B = int(CH);
                s.prm.in = single_view((void*)(size_t)(z0 * es), (long long)nzc, (long long)CH, (long long)(oy_me * nzc));
                s.prm.bulk_out = p->bulk_store;
                seg_view(s.prm.out, tab_x, G2, [&](int q, int r) {
                    const size_t nxq = g.sx.size[q];
                    return mkseg(eptr(slotp(D2, r), ((z0 / CH) * ny + oy0_me) * nxq * CH, es), (long long)(nxq * CH), (long long)(ny * nxq * CH), (long long)CH,
                                 g.sx.start[q]);
                });
            } else {
                s.prm.A0 = 1; s.prm.A1 = int(oy_me); s.prm.B = int(zc);
                s.prm.in = single_view((void*)(size_t)(z0 * es), 0, (long long)nzc, (long long)(oy_me * nzc));
                // dest q holds [nx_q][ny][nzc]; my rows y in [oy0_me, +oy_me)
                seg_view(s.prm.out, tab_x, G2, [&](int q, int r) {
                    return mkseg(eptr(slotp(D2, r), oy0_me * nzc + z0, es), 0, (long long)nzc, (long long)(ny * nzc), g.sx.start[q]);
                });
            }
            s.prm.max_ctas = exchange_ctas(p, inverse);
            s.prm.tile_pref = p->xchg_tile_pref;
            s.stream = 1;
            if (c == 0) s.waits.push_back(ev_entry);
            const bool tail_here = CH && rem && c + 1 == NSc;
            if (!tail_here) s.record = ev_x[c] = nev++;
            sc.steps.push_back(s);
            if (tail_here) {
                Step t = s;
                t.label = "x pass (tail)";
                t.prm.bulk_out = 0;
                t.waits.clear();
                t.prm.A0 = int(oy_me); t.prm.A1 = 1; t.prm.B = int(rem);
                t.prm.in = single_view((void*)(size_t)(nzm * es), (long long)nzc, 0, (long long)(oy_me * nzc));
                seg_view(t.prm.out, tab_x, G2, [&](int q, int r) {
                    const size_t nxq = g.sx.size[q];
                    return mkseg(eptr(slotp(D2, r), ny * nxq * nzm + oy0_me * rem, es), (long long)rem, 0, (long long)(ny * rem), g.sx.start[q]);
                });
                t.record = ev_x[c] = nev++;
                sc.steps.push_back(t);
            }
        }
        for (size_t c = 0; c < NSc; ++c) {
            const size_t z0 = chunks.start[c], zc = chunks.size[c];
            Step r = rendezvous(2, 2, 2);
            r.waits.push_back(ev_x[c]);
            sc.steps.push_back(r);
            Step s;
            rc = new_pass(PASS_C2C_TILED, ny, "y pass", s);
            if (rc) return rc;
            s.stream = 2;
            if (CH) {
                s.prm.A0 = int(nx_p); s.prm.A1 = int(zc / CH); s.prm.B = int(CH);
                s.prm.tile_pref = 1;
                s.prm.in = single_view(eptr(slotp(D2, me), (z0 / CH) * ny * nx_p * CH, es), (long long)CH, (long long)(ny * nx_p * CH), (long long)(nx_p * CH));
                s.prm.out = single_view(eptr(slotp(D1, me), z0, es), (long long)(ny * nzc), (long long)CH, (long long)nzc);
                const bool tail_here = rem && c + 1 == NSc;
                if (!tail_here) s.record = ev_yi[c] = nev++;
                sc.steps.push_back(s);
                if (tail_here) {
                    Step t = s;
                    t.label = "y pass (tail)";
                    t.prm.A0 = int(nx_p); t.prm.A1 = 1; t.prm.B = int(rem);
                    t.prm.in = single_view(eptr(slotp(D2, me), ny * nx_p * nzm, es), (long long)(ny * rem), 0, (long long)rem);
                    t.prm.out = single_view(eptr(slotp(D1, me), nzm, es), (long long)(ny * nzc), 0, (long long)nzc);
                    t.record = ev_yi[c] = nev++;
                    sc.steps.push_back(t);
                }
            } else {
                s.prm.A0 = int(nx_p); s.prm.A1 = 1; s.prm.B = int(zc);
                s.prm.in = single_view(eptr(slotp(D2, me), z0, es), (long long)(ny * nzc), 0, (long long)nzc);
                s.prm.out = s.prm.in;  // in place
                s.record = ev_yi[c] = nev++;
                sc.steps.push_back(s);
            }
        }
        Step s;
        const PassKind zkind = c2c ? PASS_C2C_CONTIG : PASS_C2R;
        const long long zpitch = c2c ? (long long)g.nz : (long long)(g.nz / 2);
        rc = new_pass(zkind, c2c ? g.nz : g.nz / 2, c2c ? "z pass" : "z pass (C2R)", s);
        if (rc) return rc;
        s.prm.A0 = int(nx_p); s.prm.A1 = int(ny);
        s.prm.in = single_view(slotp(CH ? D1 : D2, me), (long long)(ny * nzc), (long long)nzc, 1);
        s.prm.out = single_view(nullptr, zpitch * (long long)ny, zpitch, 1);
        s.out_user = 2;
        s.stream = 0;
        for (size_t c = 0; c < NSc; ++c) s.waits.push_back(ev_yi[c]);
        sc.steps.push_back(s);
    }
    sc.nevents = nev;
    sc.built = true;
    return DFFT_SUCCESS;
}

}  // namespace dfft

namespace dfft {

// DFFT_PENCIL_OVERLAP: 1 (default) = Streams plans on a grid with two real transpositions run the forward overlapped
// schedule, and dfft_plan_tune may select an overlapped schedule for any grid and direction after measuring it;
// 0 = sequential schedules only; 2 = overlapped schedules wherever they can be built, untuned (tests, experiments).
static int pencil_overlap_mode() {
    const char* e = getenv("DFFT_PENCIL_OVERLAP");
    return e ? atoi(e) : 1;
}
static bool pencil_overlap_enabled() { return pencil_overlap_mode() != 0; }

// inverse half of build_overlapped_pencil (see the comment there); NG plane groups, the same on every rank
static int build_overlapped_pencil_inverse(dfft_plan_s* p, Schedule& sc, size_t NG) {
    const Geometry& g = p->g;
    const int me = p->rank;
    const int pi = g.pi(me), pj = g.pj(me);
    const size_t es = p->esize;
    const bool c2c = g.transform == DFFT_C2C;
    Tables& T = p->tabs;
    const size_t ny = g.ny, nx = g.nx, nzc = g.nzc;
    const size_t nx_i = g.sx.size[pi];
    const size_t ny_j = g.sy.size[pj];
    const size_t nz_j = g.sz.size[pj], z0_j = g.sz.start[pj];
    const size_t oy_i = g.oy.size[pi], oy0_i = g.oy.start[pi];
    const std::vector<int>& G1 = p->grp[1];
    const std::vector<int>& G2 = p->grp[2];
    const int D1 = 0, D2 = 1;
    auto slotp = [&](int s_, int r) -> void* { return p->slot_ptr[s_][r]; };
    int nev = 0;
    auto new_pass = [&](PassKind kind, size_t n, const char* label, Step& s) -> int {
        s = Step();
        s.type = STEP_PASS;
        s.kind = kind;
        s.label = label;
        s.log2n = ilog2_exact(n);
        if (s.log2n < 1 || s.log2n > MAX_LOG2N) return fail(DFFT_ERR_UNSUPPORTED, "unsupported axis length");
        s.prm.A0 = 1; s.prm.A1 = 1; s.prm.B = 1;
        s.prm.inverse = 1;
        void* tw = nullptr;
        if (T.get_tw(s.log2n, &tw) != cudaSuccess) return fail(DFFT_ERR_CUDA, "twiddle table allocation failed");
        s.prm.tw = tw;
        if (kind == PASS_C2R) {
            void* tw2 = nullptr;
            if (T.get_tw2(s.log2n, &tw2) != cudaSuccess) return fail(DFFT_ERR_CUDA, "twiddle table allocation failed");
            s.prm.tw2 = tw2;
        }
        return DFFT_SUCCESS;
    };
    auto rendezvous = [&](int group, int stream) {
        Step s;
        s.type = STEP_RENDEZVOUS;
        s.label = group == 0 ? "entry rendezvous" : (group == 1 ? "rendezvous 1" : "rendezvous 2");
        s.group = group;
        s.phase_id = group;
        s.stream = stream;
        return s;
    };
    Split groups;
    groups.make(nx_i, NG);
    // hand-over of the x -> y transposition: blocked [nz_j/CH][ny][nx_q][CH] + plain tail [nx_q][ny][rem] (build_schedule)
    const size_t CH = p->blocked_inv ? size_t(p->blocked_ch) : 0;
    const size_t rem = CH ? nz_j % CH : 0, nzm = nz_j - rem;
    bool any_rem = false;  // ranks without leftover columns run the tail steps as empty passes: same step list everywhere
    if (CH) for (size_t v : g.sz.size) any_rem = any_rem || (v % CH != 0);
    const unsigned char *tab_x = nullptr, *tab_y_in = nullptr;
    if (T.seg_table(g.sx, &tab_x) != cudaSuccess || T.seg_table(g.sy, &tab_y_in) != cudaSuccess) return fail(DFFT_ERR_CUDA, "segment table");
    int rc;

    sc.steps.push_back(rendezvous(0, 0));  // everyone has left the previous exec
    // x pass on the caller's [nx][oy_i][nz_j], whole, uncapped: nothing local can run beside it yet
    {
        Step s;
        rc = new_pass(PASS_C2C_TILED, nx, "x pass", s);
        if (rc) return rc;
        s.in_user = 1;
        s.prm.tile_pref = G2.size() > 1 ? p->xchg_tile_pref : 1;
        if (CH) {
            s.prm.A0 = int(oy_i); s.prm.A1 = int(nzm / CH); s.prm.B = int(CH);
            s.prm.in = single_view(nullptr, (long long)nz_j, (long long)CH, (long long)(oy_i * nz_j));
            s.prm.bulk_out = G2.size() > 1 ? p->bulk_store : 0;
            seg_view(s.prm.out, tab_x, G2, [&](int q, int r) {
                const size_t nxq = g.sx.size[q];
                return mkseg(eptr(slotp(D2, r), oy0_i * nxq * CH, es), (long long)(nxq * CH), (long long)(ny * nxq * CH), (long long)CH, g.sx.start[q]);
            });
            sc.steps.push_back(s);
            if (any_rem) {
                Step t = s;
                t.label = "x pass (tail)";
                t.prm.bulk_out = 0;
                t.prm.A0 = int(oy_i); t.prm.A1 = 1; t.prm.B = int(rem);
                t.prm.in = single_view((void*)(size_t)(nzm * es), (long long)nz_j, 0, (long long)(oy_i * nz_j));
                seg_view(t.prm.out, tab_x, G2, [&](int q, int r) {
                    const size_t nxq = g.sx.size[q];
                    return mkseg(eptr(slotp(D2, r), ny * nxq * nzm + oy0_i * rem, es), (long long)rem, 0, (long long)(ny * rem), g.sx.start[q]);
                });
                sc.steps.push_back(t);
            }
        } else {
            s.prm.A0 = 1; s.prm.A1 = 1; s.prm.B = int(oy_i * nz_j);
            s.prm.in = single_view(nullptr, 0, 0, (long long)(oy_i * nz_j));
            // dest (q, j) holds [nx_q][ny][nz_j]; my rows y in [oy0_i, +oy_i)
            seg_view(s.prm.out, tab_x, G2, [&](int q, int r) {
                return mkseg(eptr(slotp(D2, r), oy0_i * nz_j, es), 0, 0, (long long)(ny * nz_j), g.sx.start[q]);
            });
            sc.steps.push_back(s);
        }
    }
    sc.steps.push_back(rendezvous(2, 0));  // all column peers have delivered their rows
    const int ev_x = nev++;
    sc.steps.back().record = ev_x;
    std::vector<int> ev_y(NG);
    for (size_t gi = 0; gi < NG; ++gi) {
        const size_t pl0 = groups.start[gi], npl = groups.size[gi];
        Step s;
        rc = new_pass(PASS_C2C_TILED, ny, "y pass", s);
        if (rc) return rc;
        s.stream = 1;
        if (gi == 0) s.waits.push_back(ev_x);
        s.prm.max_ctas = exchange_ctas(p, 1);
        s.prm.tile_pref = G1.size() > 1 ? p->xchg_tile_pref : 1;
        if (CH) {
            // in = my slot, blocked [nz_j/CH][ny][nx_i][CH]; tile = (x_loc, z block); dest (i, q) holds [nx_i][ny_q][nzc]
            s.prm.A0 = int(npl); s.prm.A1 = int(nzm / CH); s.prm.B = int(CH);
            s.prm.in = single_view(eptr(slotp(D2, me), pl0 * CH, es), (long long)CH, (long long)(ny * nx_i * CH), (long long)(nx_i * CH));
            seg_view(s.prm.out, tab_y_in, G1, [&](int q, int r) {
                const size_t nyq = g.sy.size[q];
                return mkseg(eptr(slotp(D1, r), pl0 * nyq * nzc + z0_j, es), (long long)(nyq * nzc), (long long)CH, (long long)nzc, g.sy.start[q]);
            });
            if (!any_rem) s.record = ev_y[gi] = nev++;
            sc.steps.push_back(s);
            if (any_rem) {
                Step t = s;
                t.label = "y pass (tail)";
                t.waits.clear();
                t.prm.A0 = int(npl); t.prm.A1 = 1; t.prm.B = int(rem);
                t.prm.in = single_view(eptr(slotp(D2, me), ny * nx_i * nzm + pl0 * ny * rem, es), (long long)(ny * rem), 0, (long long)rem);
                seg_view(t.prm.out, tab_y_in, G1, [&](int q, int r) {
                    const size_t nyq = g.sy.size[q];
                    return mkseg(eptr(slotp(D1, r), pl0 * nyq * nzc + z0_j + nzm, es), (long long)(nyq * nzc), 0, (long long)nzc, g.sy.start[q]);
                });
                t.record = ev_y[gi] = nev++;
                sc.steps.push_back(t);
            }
        } else {
            s.prm.A0 = int(npl); s.prm.A1 = 1; s.prm.B = int(nz_j);
            s.prm.in = single_view(eptr(slotp(D2, me), pl0 * ny * nz_j, es), (long long)(ny * nz_j), 0, (long long)nz_j);
            seg_view(s.prm.out, tab_y_in, G1, [&](int q, int r) {
                const size_t nyq = g.sy.size[q];
                return mkseg(eptr(slotp(D1, r), pl0 * nyq * nzc + z0_j, es), (long long)(nyq * nzc), 0, (long long)nzc, g.sy.start[q]);
            });
            s.record = ev_y[gi] = nev++;
            sc.steps.push_back(s);
        }
    }
    const PassKind zkind = c2c ? PASS_C2C_CONTIG : PASS_C2R;
    const long long zpitch = c2c ? (long long)g.nz : (long long)(g.nz / 2);  // output line pitch in complex units
    for (size_t gi = 0; gi < NG; ++gi) {
        const size_t pl0 = groups.start[gi], npl = groups.size[gi];
        Step r = rendezvous(1, 2);  // all row peers have delivered plane group gi
        r.waits.push_back(ev_y[gi]);
        sc.steps.push_back(r);
        Step s;
        rc = new_pass(zkind, c2c ? g.nz : g.nz / 2, c2c ? "z pass" : "z pass (C2R)", s);
        if (rc) return rc;
        s.prm.A0 = int(npl); s.prm.A1 = int(ny_j);
        s.prm.in = single_view(eptr(slotp(D1, me), pl0 * ny_j * nzc, es), (long long)(ny_j * nzc), (long long)nzc, 1);
        s.prm.out = single_view((void*)(size_t)(pl0 * ny_j * size_t(zpitch) * es), zpitch * (long long)ny_j, zpitch, 1);
        s.out_user = 2;
        s.stream = 2;
        sc.steps.push_back(s);
    }
    sc.nevents = nev;
    sc.built = true;
    return DFFT_SUCCESS;
}

// Overlapped pencil schedules, Peer2Peer on both transpositions (send_method Streams; forward measured at 8 GPUs:
// 2048x2048x1024 complex-float 8.67 ms on a 4x2 grid vs 10.7 ms sequential, profiles/r02/8gpu_b; parity at full size
// against cuFFT).  Forward:
//   stream 0: z pass per plane group, scattering along z into the row peers' slot A            (NVLink-bound)
//   stream 1: per plane group: meet the row peers; then the y pass per (plane group, z chunk), persistent on
//             `xchg_ctas` CTAs, scattering along y into the column peers' slot B                 (NVLink-bound)
//   stream 2: per z chunk: meet the column peers, x pass of that chunk                           (HBM-bound)
// Only the x pass can hide behind an exchange (both scatters share the NVLink ports).  The inverse mirrors it — again
// only the local pass, now the z pass, can hide:
//   stream 0: x pass, scattering along x into the column peers' slot B; meet the column peers   (NVLink-bound)
//   stream 1: y pass per plane group on a capped grid, scattering along y into the row peers' slot A  (NVLink-bound)
//   stream 2: per plane group: meet the row peers, z pass of that group into the caller's buffer  (HBM-bound)
static int build_overlapped_pencil(dfft_plan_s* p, int inverse, Schedule& sc) {
    const Geometry& g = p->g;
    const int me = p->rank;
    const int pi = g.pi(me), pj = g.pj(me);
    const size_t es = p->esize;
    const bool c2c = g.transform == DFFT_C2C;
    Tables& T = p->tabs;
    sc.steps.clear();
    sc.overlapped = true;
    const size_t ny = g.ny, nx = g.nx;
    const size_t nx_i = g.sx.size[pi], x0_i = g.sx.start[pi];
    const size_t ny_j = g.sy.size[pj], y0_j = g.sy.start[pj];
    const size_t nz_j = g.sz.size[pj];
    const size_t oy_i = g.oy.size[pi];
    const std::vector<int>& G1 = p->grp[1];
    const std::vector<int>& G2 = p->grp[2];
    const int D1 = 0, D2 = 1;
    auto slotp = [&](int s_, int r) -> void* { return p->slot_ptr[s_][r]; };
    int nev = 0;

    auto new_pass = [&](PassKind kind, size_t n, const char* label, Step& s) -> int {
        s = Step();
        s.type = STEP_PASS;
        s.kind = kind;
        s.label = label;
        s.log2n = ilog2_exact(n);
        if (s.log2n < 1 || s.log2n > MAX_LOG2N) return fail(DFFT_ERR_UNSUPPORTED, "unsupported axis length");
        s.prm.A0 = 1; s.prm.A1 = 1; s.prm.B = 1;
        s.prm.inverse = inverse;
        void* tw = nullptr;
        if (T.get_tw(s.log2n, &tw) != cudaSuccess) return fail(DFFT_ERR_CUDA, "twiddle table allocation failed");
        s.prm.tw = tw;
        if (kind == PASS_R2C || kind == PASS_C2R) {
            void* tw2 = nullptr;
            if (T.get_tw2(s.log2n, &tw2) != cudaSuccess) return fail(DFFT_ERR_CUDA, "twiddle table allocation failed");
            s.prm.tw2 = tw2;
        }
        return DFFT_SUCCESS;
    };
    auto rendezvous = [&](int group, int stream) {
        Step s;
        s.type = STEP_RENDEZVOUS;
        s.label = group == 0 ? "entry rendezvous" : (group == 1 ? "rendezvous 1" : "rendezvous 2");
        s.group = group;
        s.phase_id = group;
        s.stream = stream;
        return s;
    };

    // every rank must build the same number of steps: derive the group / chunk counts from global minima
    size_t min_nx = nx, min_nz = g.nzc;
    for (size_t v : g.sx.size) min_nx = std::min(min_nx, v);
    for (size_t v : g.sz.size) min_nz = std::min(min_nz, v);
    const size_t NG = std::min<size_t>(size_t(inverse && p->tuned_groups_inv > 0 ? p->tuned_groups_inv : p->ovl_groups), min_nx);
    if (inverse) return build_overlapped_pencil_inverse(p, sc, NG);
    const size_t want_chunks = size_t(p->tuned_chunks[0] > 0 ? p->tuned_chunks[0] : p->ovl_chunks);
    const size_t NS = min_nz >= 32 * want_chunks ? want_chunks : (min_nz >= 32 ? 2 : 1);
    Split groups, chunks;
    groups.make(nx_i, NG);
    // blocked hand-over of the second transposition: receiver (q, j) holds [nz_j/CH][nx][ny_q][CH] (+ a plain-layout
    // tail of nz_j % CH columns), see build_schedule
    const size_t CH = size_t(p->blocked_ch);
    const size_t rem = CH ? nz_j % CH : 0, nzm = nz_j - rem;
    bool any_rem = false;  // ranks without leftover columns run the tail steps as empty passes: same step list everywhere
    if (CH) for (size_t v : g.sz.size) any_rem = any_rem || (v % CH != 0);
    if (CH) {
        Split u;
        u.make(nzm / CH, std::min<size_t>(NS, nzm / CH));
        if (u.size.size() != NS) return fail(DFFT_ERR_STATE, "internal: blocked z chunks");
        chunks.size.clear(); chunks.start.clear();
        for (size_t c = 0; c < u.size.size(); ++c) { chunks.size.push_back(u.size[c] * CH); chunks.start.push_back(u.start[c] * CH); }
    } else {
        chunks.make(nz_j, NS);
    }
    const unsigned char *tab_z = nullptr, *tab_y = nullptr;
    if (T.seg_table(g.sz, &tab_z) != cudaSuccess || T.seg_table(g.oy, &tab_y) != cudaSuccess) return fail(DFFT_ERR_CUDA, "segment table");

    sc.steps.push_back(rendezvous(0, 0));
    const int ev_entry = nev++;
    sc.steps.back().record = ev_entry;
    int rc;
    std::vector<int> ev_z(NG), ev_y(NS);
    for (size_t gi = 0; gi < NG; ++gi) {
        Step s;
        rc = new_pass(c2c ? PASS_C2C_CONTIG : PASS_R2C, c2c ? g.nz : g.nz / 2, c2c ? "z pass" : "z pass (R2C)", s);
        if (rc) return rc;
        const size_t pl0 = groups.start[gi], npl = groups.size[gi];
        const long long pitch = c2c ? (long long)g.nz : (long long)(g.nz / 2);
        s.prm.A0 = int(npl); s.prm.A1 = int(ny_j);
        s.prm.in = single_view((void*)(size_t)(pl0 * ny_j * pitch * es), pitch * (long long)ny_j, pitch, 1);
        s.in_user = 1;
        seg_view(s.prm.out, tab_z, G1, [&](int q, int r) {
            const size_t nzq = g.sz.size[q];
            return mkseg(eptr(slotp(D1, r), (pl0 * ny + y0_j) * nzq, es), (long long)(ny * nzq), (long long)nzq, 1, g.sz.start[q]);
        });
        s.stream = 0;
        s.record = ev_z[gi] = nev++;
        sc.steps.push_back(s);
    }
    for (size_t c = 0; c < NS; ++c) {
        const size_t z0 = chunks.start[c], zc = chunks.size[c];
        for (size_t gi = 0; gi < NG; ++gi) {
            const size_t pl0 = groups.start[gi], npl = groups.size[gi];
            if (c == 0) {
                Step r = rendezvous(1, 1);  // all row peers have delivered plane group gi
                r.waits.push_back(ev_z[gi]);
                if (gi == 0) r.waits.push_back(ev_entry);
                sc.steps.push_back(r);
            }
            Step s;
            rc = new_pass(PASS_C2C_TILED, ny, "y pass", s);
            if (rc) return rc;
            if (CH) {
                s.prm.A0 = int(npl); s.prm.A1 = int(zc / CH); s.prm.B = int(CH);
                s.prm.in = single_view(eptr(slotp(D1, me), pl0 * ny * nz_j + z0, es), (long long)(ny * nz_j), (long long)CH, (long long)nz_j);
                s.prm.bulk_out = G2.size() > 1 ? p->bulk_store : 0;
                seg_view(s.prm.out, tab_y, G2, [&](int q, int r) {
                    const size_t nyq = g.oy.size[q];
                    return mkseg(eptr(slotp(D2, r), ((z0 / CH) * nx + x0_i + pl0) * nyq * CH, es), (long long)(nyq * CH), (long long)(nx * nyq * CH),
                                 (long long)CH, g.oy.start[q]);
                });
            } else {
                s.prm.A0 = int(npl); s.prm.A1 = 1; s.prm.B = int(zc);
                s.prm.in = single_view(eptr(slotp(D1, me), pl0 * ny * nz_j + z0, es), (long long)(ny * nz_j), 0, (long long)nz_j);
                seg_view(s.prm.out, tab_y, G2, [&](int q, int r) {
                    const size_t nyq = g.oy.size[q];
                    return mkseg(eptr(slotp(D2, r), (x0_i + pl0) * nyq * nz_j + z0, es), (long long)(nyq * nz_j), 0, (long long)nz_j, g.oy.start[q]);
                });
            }
            s.prm.max_ctas = exchange_ctas(p, 0);
            s.prm.tile_pref = G2.size() > 1 ? p->xchg_tile_pref : 0;
            s.stream = 1;
            const bool tail_here = CH && any_rem && c + 1 == NS;
            if (gi + 1 == NG && !tail_here) s.record = ev_y[c] = nev++;
            sc.steps.push_back(s);
            if (tail_here) {
                Step t = s;
                t.label = "y pass (tail)";
                t.prm.bulk_out = 0;
                t.waits.clear();
                t.record = -1;
                t.prm.A0 = int(npl); t.prm.A1 = 1; t.prm.B = int(rem);
                t.prm.in = single_view(eptr(slotp(D1, me), pl0 * ny * nz_j + nzm, es), (long long)(ny * nz_j), 0, (long long)nz_j);
                seg_view(t.prm.out, tab_y, G2, [&](int q, int r) {
                    const size_t nyq = g.oy.size[q];
                    return mkseg(eptr(slotp(D2, r), nx * nyq * nzm + (x0_i + pl0) * nyq * rem, es), (long long)(nyq * rem), 0, (long long)rem, g.oy.start[q]);
                });
                if (gi + 1 == NG) t.record = ev_y[c] = nev++;
                sc.steps.push_back(t);
            }
        }
    }
    for (size_t c = 0; c < NS; ++c) {
        const size_t z0 = chunks.start[c], zc = chunks.size[c];
        Step r = rendezvous(2, 2);
        r.waits.push_back(ev_y[c]);
        sc.steps.push_back(r);
        Step s;
        rc = new_pass(PASS_C2C_TILED, nx, "x pass", s);
        if (rc) return rc;
        if (CH) {
            s.prm.A0 = int(oy_i); s.prm.A1 = int(zc / CH); s.prm.B = int(CH);
            s.prm.in = single_view(eptr(slotp(D2, me), (z0 / CH) * nx * oy_i * CH, es), (long long)CH, (long long)(nx * oy_i * CH), (long long)(oy_i * CH));
            s.prm.out = single_view((void*)(size_t)(z0 * es), (long long)nz_j, (long long)CH, (long long)(oy_i * nz_j));
            s.prm.tile_swz = p->x_swz;
        } else {
            s.prm.A0 = 1; s.prm.A1 = int(oy_i); s.prm.B = int(zc);
            s.prm.in = single_view(eptr(slotp(D2, me), z0, es), 0, (long long)nz_j, (long long)(oy_i * nz_j));
            s.prm.out = single_view((void*)(size_t)(z0 * es), 0, (long long)nz_j, (long long)(oy_i * nz_j));
        }
        s.out_user = 2;
        s.stream = 2;
        sc.steps.push_back(s);
        if (CH && any_rem && c + 1 == NS) {
            Step t = s;
            t.label = "x pass (tail)";
            t.prm.A0 = int(oy_i); t.prm.A1 = 1; t.prm.B = int(rem);
            t.prm.in = single_view(eptr(slotp(D2, me), nx * oy_i * nzm, es), (long long)rem, 0, (long long)(oy_i * rem));
            t.prm.out = single_view((void*)(size_t)(nzm * es), (long long)nz_j, 0, (long long)(oy_i * nz_j));
            sc.steps.push_back(t);
        }
    }
    sc.nevents = nev;
    sc.built = true;
    return DFFT_SUCCESS;
}

}  // namespace dfft


// =====================================================================================================
// memory / peer mapping
// =====================================================================================================
static int nccl_allgather_bytes(dfft_plan_s* p, const void* mine, size_t bytes, std::vector<char>& all) {
    const int P = p->P;
    all.assign(bytes * P, 0);
    if (P == 1) {
        memcpy(all.data(), mine, bytes);
        return DFFT_SUCCESS;
    }
    char* d = nullptr;
    CK_CUDA(cudaMalloc(&d, bytes * (P + 1)));
    CK_CUDA(cudaMemcpy(d + bytes * P, mine, bytes, cudaMemcpyHostToDevice));
    CK_NCCL(ncclAllGather(d + bytes * P, d, bytes, ncclChar, p->comm->nccl, p->own_stream));
    CK_CUDA(cudaStreamSynchronize(p->own_stream));
    CK_CUDA(cudaMemcpy(all.data(), d, bytes * P, cudaMemcpyDeviceToHost));
    CK_CUDA(cudaFree(d));
    return DFFT_SUCCESS;
}

// Base address and size of the allocation that contains `ptr` (driver API cuMemGetAddressRange, resolved at run time
// through the runtime so that libdfft.so does not link libcuda).
static bool alloc_range(const void* ptr, unsigned long long* base, size_t* span) {
    typedef int (*fn_t)(unsigned long long*, size_t*, unsigned long long);
    static fn_t fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
            cudaGetLastError();
            return false;
        }
        fn = (fn_t)f;
    }
    return fn(base, span, (unsigned long long)ptr) == 0;
}

static void plan_release_memory(dfft_plan_s* p) {
    for (void* q : p->opened) cudaIpcCloseMemHandle(q);
    p->opened.clear();
    if (p->work && p->work_owned) cudaFree(p->work);
    p->work = nullptr;
    p->work_owned = false;
    p->slot_ptr.clear();
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 3; ++b) p->sched[a][b] = Schedule();
    p->tabs.release_seg_tables();
}

static int plan_setup_memory(dfft_plan_s* p, void* user_device) {
    plan_release_memory(p);
    const int P = p->P, me = p->rank;
    if (user_device) {
        p->work = user_device;
        p->work_owned = false;
    } else {
        CK_CUDA(cudaMalloc(&p->work, p->work_bytes));
        p->work_owned = true;
    }
    p->slot_ptr.assign(p->nslots, std::vector<void*>(P, nullptr));
    for (int s = 0; s < p->nslots; ++s) p->slot_ptr[s][me] = (char*)p->work + size_t(s) * p->slot_bytes;
    if (p->any_direct) {
        // A caller-supplied work area may sit INSIDE a larger allocation (a tensor of a caching allocator): the IPC
        // handle then names the whole allocation and cudaIpcOpenMemHandle returns its base on the peer, so the byte
        // offset of the work area inside its allocation travels with the handle.
        cudaIpcMemHandle_t h;
        cudaError_t e = cudaIpcGetMemHandle(&h, p->work);
        int ok = (e == cudaSuccess) ? 1 : 0;
        if (!ok) cudaGetLastError();
        unsigned long long offset = 0;
        if (ok) {
            unsigned long long base = 0;
            size_t span = 0;
            if (!alloc_range(p->work, &base, &span)) ok = 0;
            else {
                offset = (unsigned long long)p->work - base;
                if (offset + p->work_bytes > span) ok = 0;  // the area does not fit its allocation
            }
        }
        struct Rec { cudaIpcMemHandle_t h; int ok; int pad; unsigned long long offset; } rec{};
        rec.h = h; rec.ok = ok; rec.offset = offset;
        std::vector<char> all;
        int rc = nccl_allgather_bytes(p, &rec, sizeof(rec), all);
        if (rc) return rc;
        for (int r = 0; r < P; ++r) {
            const Rec* rr = reinterpret_cast<const Rec*>(all.data()) + r;
            if (!rr->ok) return fail(DFFT_ERR_PEER, "rank " + std::to_string(r) + ": work area is not exportable with cudaIpcGetMemHandle "
                                                    "(it must lie inside one cudaMalloc allocation that holds getWorkSizeDevice() bytes from "
                                                    "the given pointer); use comm_method All2All");
        }
        for (int r = 0; r < P; ++r) {
            if (r == me) continue;
            const Rec* rr = reinterpret_cast<const Rec*>(all.data()) + r;
            void* mapped = nullptr;
            e = cudaIpcOpenMemHandle(&mapped, rr->h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(DFFT_ERR_PEER, std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(e));
            }
            p->opened.push_back(mapped);
            for (int s = 0; s < p->nslots; ++s) p->slot_ptr[s][r] = (char*)mapped + rr->offset + size_t(s) * p->slot_bytes;
        }
    }
    return DFFT_SUCCESS;
}

static int plan_setup_flags(dfft_plan_s* p) {
    const int P = p->P, me = p->rank;
    CK_CUDA(cudaMalloc(&p->flags, sizeof(unsigned long long) * NPHASE * P));
    CK_CUDA(cudaMemset(p->flags, 0, sizeof(unsigned long long) * NPHASE * P));
    CK_CUDA(cudaMalloc(&p->err_d, sizeof(int)));
    CK_CUDA(cudaMemset(p->err_d, 0, sizeof(int)));
    std::vector<unsigned long long*> pf(P, nullptr);
    pf[me] = p->flags;
    if (p->any_direct) {
        cudaIpcMemHandle_t h;
        CK_CUDA(cudaIpcGetMemHandle(&h, p->flags));
        std::vector<char> all;
        int rc = nccl_allgather_bytes(p, &h, sizeof(h), all);
        if (rc) return rc;
        for (int r = 0; r < P; ++r) {
            if (r == me) continue;
            void* mapped = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&mapped, reinterpret_cast<const cudaIpcMemHandle_t*>(all.data())[r], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(DFFT_ERR_PEER, std::string("cudaIpcOpenMemHandle(flags) failed: ") + cudaGetErrorString(e));
            }
            p->opened_flags.push_back(mapped);
            pf[r] = (unsigned long long*)mapped;
        }
    }
    CK_CUDA(cudaMalloc(&p->peer_flags_d, sizeof(void*) * P));
    CK_CUDA(cudaMemcpy(p->peer_flags_d, pf.data(), sizeof(void*) * P, cudaMemcpyHostToDevice));
    std::vector<int> gl(3 * P, 0);
    for (int k = 0; k < 3; ++k)
        for (size_t q = 0; q < p->grp[k].size(); ++q) gl[k * P + q] = p->grp[k][q];
    CK_CUDA(cudaMalloc(&p->groups_d, sizeof(int) * 3 * P));
    CK_CUDA(cudaMemcpy(p->groups_d, gl.data(), sizeof(int) * 3 * P, cudaMemcpyHostToDevice));
    return DFFT_SUCCESS;
}

// =====================================================================================================
// execution
// =====================================================================================================
static int run_schedule(dfft_plan_s* p, Schedule& sc, void* out, const void* in, cudaStream_t st) {
    const int P = p->P, me = p->rank;
    const size_t es = p->esize;
    p->epoch++;
    p->last_stream = st;
    int launches = 0;
    int ev = 0;
    const bool timing = p->timing;
    cudaStream_t const ax0 = p->aux[0], ax1 = p->aux[1];
    cudaStream_t streams[3] = {st, ax0, ax1};
    auto mark = [&](const char* name, int is_fft, const char* label = "") -> cudaError_t {
        if (!timing || (sc.overlapped && is_fft != -1)) return cudaSuccess;  // overlapped: only start / end are meaningful
        if (ev >= int(p->events.size())) {
            cudaEvent_t e;
            cudaError_t r = cudaEventCreate(&e);
            if (r != cudaSuccess) return r;
            p->events.push_back(e);
            p->ev_names.push_back(name);
            p->ev_is_fft.push_back(is_fft);
            p->ev_labels.push_back(label);
        }
        p->ev_names[ev] = name;
        p->ev_is_fft[ev] = is_fft;
        p->ev_labels[ev] = label;
        return cudaEventRecord(p->events[ev++], st);
    };
    if (sc.overlapped) {
        while (int(p->sync_events.size()) < sc.nevents) {
            cudaEvent_t e;
            CK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            p->sync_events.push_back(e);
        }
        CK_CUDA(cudaEventRecord(p->fork_ev, st));
        CK_CUDA(cudaStreamWaitEvent(ax0, p->fork_ev, 0));
        CK_CUDA(cudaStreamWaitEvent(ax1, p->fork_ev, 0));
    }
    CK_CUDA(mark("start", -1));
    int tl = 0;
    auto tl_mark = [&](cudaStream_t ss_, const char* label, int stream_id) -> cudaError_t {
        if (!timing) return cudaSuccess;
        if (tl >= int(p->tl_events.size())) {
            cudaEvent_t e;
            cudaError_t r = cudaEventCreate(&e);
            if (r != cudaSuccess) return r;
            p->tl_events.push_back(e);
            p->tl_labels.push_back(label);
            p->tl_streams.push_back(stream_id);
        }
        p->tl_labels[tl] = label;
        p->tl_streams[tl] = stream_id;
        return cudaEventRecord(p->tl_events[tl++], ss_);
    };
    for (Step& s : sc.steps) {
        cudaStream_t ss = streams[s.stream];
        for (int w : s.waits) CK_CUDA(cudaStreamWaitEvent(ss, p->sync_events[w], 0));
        CK_CUDA(tl_mark(ss, s.label, s.stream));
        if (s.type == STEP_PASS) {
            FftParams prm = s.prm;
            // views on the caller's buffers store a byte offset in seg[0].base
            if (s.in_user == 1) prm.in.seg[0].base = (char*)const_cast<void*>(in) + (size_t)s.prm.in.seg[0].base;
            if (s.out_user == 2) prm.out.seg[0].base = (char*)out + (size_t)s.prm.out.seg[0].base;
            cudaError_t e = p->prec == DFFT_F64 ? launch_pass_f64(s.log2n, s.kind, prm, ss) : launch_pass_f32(s.log2n, s.kind, prm, ss);
            if (e != cudaSuccess) return fail(DFFT_ERR_CUDA, std::string("FFT pass launch failed: ") + cudaGetErrorString(e));
            ++launches;
            CK_CUDA(mark(s.phase, 1, s.label));
        } else if (s.type == STEP_RENDEZVOUS) {
            const std::vector<int>& G = p->grp[s.group];
            if (G.size() > 1) {
                const unsigned long long ticket = ++p->ticket[s.phase_id];
                rendezvous_kernel<<<1, 32 * int((G.size() + 31) / 32), 0, ss>>>(p->peer_flags_d, p->flags, p->groups_d + s.group * P, int(G.size()), me, P,
                                                                              s.phase_id, ticket, p->err_d, p->rendezvous_timeout_cycles);
                CK_CUDA(cudaGetLastError());
                ++launches;
            }
            CK_CUDA(mark(s.phase, 0, s.label));
        } else {
            const std::vector<int>& G = p->grp[s.group];
            char* sb = (char*)p->slot_ptr[s.send_slot][me];
            char* rb = (char*)p->slot_ptr[s.recv_slot][me];
            CK_NCCL(ncclGroupStart());
            for (size_t q = 0; q < G.size(); ++q) {
                if (G[q] == me) continue;
                if (s.scount[q]) CK_NCCL(ncclSend(sb + s.soff[q] * es, s.scount[q] * es, ncclChar, G[q], p->comm->nccl, ss));
                if (s.rcount[q]) CK_NCCL(ncclRecv(rb + s.roff[q] * es, s.rcount[q] * es, ncclChar, G[q], p->comm->nccl, ss));
            }
            CK_NCCL(ncclGroupEnd());
            for (size_t q = 0; q < G.size(); ++q)
                if (G[q] == me && s.scount[q])
                    CK_CUDA(cudaMemcpyAsync(rb + s.roff[q] * es, sb + s.soff[q] * es, s.scount[q] * es, cudaMemcpyDeviceToDevice, ss));
            CK_CUDA(mark(s.phase, 0, s.label));
        }
        CK_CUDA(tl_mark(ss, s.label, s.stream));
        if (s.record >= 0) CK_CUDA(cudaEventRecord(p->sync_events[s.record], ss));
    }
    if (timing) p->tl_used = tl;
    if (sc.overlapped) {
        for (int a = 0; a < 2; ++a) {
            CK_CUDA(cudaEventRecord(p->join_ev[a], a == 0 ? ax0 : ax1));
            CK_CUDA(cudaStreamWaitEvent(st, p->join_ev[a], 0));
        }
    }
    CK_CUDA(mark("Run complete", -1));
    if (timing) p->n_events_used = ev;
    p->last_launches = launches;
    p->execs++;
    return DFFT_SUCCESS;
}

// The one place that decides which schedule a (direction, d) pair runs — used by exec and by dfft_plan_describe, so
// the CPU schedule emulation always validates exactly what executes.
static int get_schedule(dfft_plan_s* p, int inverse, int d, Schedule** out) {
    Schedule& sc = p->sched[inverse ? 1 : 0][d - 1];
    if (!sc.built) {
        g_view_error = false;
        const bool streams = p->cfg.send_method == DFFT_SEND_STREAMS || (p->g.decomp == DFFT_PENCIL && p->cfg.send_method2 == DFFT_SEND_STREAMS);
        const bool seq_won = p->tuned_seq[inverse ? 1 : 0] != 0;
        const bool want_overlap = streams && d == 3 && p->P > 1 && p->g.decomp == DFFT_SLAB_ZY_THEN_X && p->direct2 && p->xchg_ctas >= 0 && !seq_won;
        // pencil: forward on a grid with two real transpositions by default (measured: profiles/r02/8gpu_b); every other
        // combination (inverse, p1 or p2 == 1) only once dfft_plan_tune has measured it faster than the sequential schedule
        const bool both_multi = p->grp[1].size() > 1 && p->grp[2].size() > 1;
        const bool tuned = p->tuned_ctas[inverse ? 1 : 0] != -2;
        const bool want_pencil_overlap = streams && d == 3 && p->g.decomp == DFFT_PENCIL && p->direct1 && p->direct2 && p->P > 1 && p->xchg_ctas >= 0 &&
                                         pencil_overlap_enabled() && !seq_won && ((both_multi && !inverse) || tuned || pencil_overlap_mode() == 2);
        int rc;
        if (want_overlap) rc = build_overlapped_slab(p, inverse ? 1 : 0, sc);
        else if (want_pencil_overlap) rc = build_overlapped_pencil(p, inverse ? 1 : 0, sc);
        else rc = build_schedule(p, inverse ? 1 : 0, d, sc);
        if (rc) return rc;
        if (g_view_error) return fail(DFFT_ERR_STATE, "internal: segments of one view disagree on the axis stride");
    }
    *out = &sc;
    return DFFT_SUCCESS;
}

static int timer_gather(dfft_plan_s* p);
static int exec_common(dfft_plan_t p, void* out, const void* in, int inverse, int d, int need_transform, void* stream, bool sync) {
    // sync == true: the plain calls run on the plan's own stream; _async calls use exactly the stream given
    // (NULL = the CUDA default stream)
    if (!p) return fail(DFFT_ERR_INVALID, "null plan");
    if (p->comm->dry) return fail(DFFT_ERR_STATE, "geometry-only plan (dfft_comm_create_dry) cannot execute");
    if (!p->work) return fail(DFFT_ERR_STATE, "plan has no work area (call dfft_set_work_area)");
    if (p->g.transform != need_transform) return fail(DFFT_ERR_INVALID, need_transform == DFFT_C2C ? "plan was created for R2C/C2R" : "plan was created for C2C");
    if (d < 1 || d > 3) return fail(DFFT_ERR_INVALID, "d must be 1, 2 or 3");
    if (!out || !in) return fail(DFFT_ERR_INVALID, "null buffer");
    CK_CUDA(cudaSetDevice(p->comm->device));
    Schedule* scp = nullptr;
    {
        int rc = get_schedule(p, inverse, d, &scp);
        if (rc) return rc;
    }
    Schedule& sc = *scp;
    cudaStream_t st = sync ? p->own_stream : (cudaStream_t)stream;
    if (sync) {
        // explicit edge from the legacy default stream (belt and braces on top of the blocking-stream semantics; it
        // also covers callers whose own streams are blocking streams, which the legacy stream itself waits for)
        CK_CUDA(cudaEventRecord(p->entry_ev, cudaStreamLegacy));
        CK_CUDA(cudaStreamWaitEvent(st, p->entry_ev, 0));
    }
    int rc = run_schedule(p, sc, out, in, st);
    if (rc) return rc;
    if (!sync) return DFFT_SUCCESS;
    rc = dfft_plan_wait(p);
    if (rc) return rc;
    CK_CUDA(cudaDeviceSynchronize());  // the reference's execs end like this (mpicufft_slab.cpp:807, 870)
    // the reference gathers and appends the section times after every non-warm-up exec
    // (mpicufft_slab.cpp:817-821)
    if (!p->csv_path.empty() && p->timing && d == 3) {
        if (p->csv_warmup_left > 0) p->csv_warmup_left--;
        else return timer_gather(p);
    }
    return DFFT_SUCCESS;
}

// =====================================================================================================
// phase-timer CSV in the reference's on-disk schema (/root/reference/src/timer.cpp:58-101; file names
// mpicufft_slab.cpp:99-103, mpicufft_slab_z_then_yx.cpp:76-80, mpicufft_pencil.cpp:67-72)
// =====================================================================================================
static const char* const SECTIONS_SLAB[] = {"init", "2D FFT (Sync)", "2D FFT Y-Z-Direction", "Transpose (First Send)", "Transpose (Packing)",
    "Transpose (Start Local Transpose)", "Transpose (Start Receive)", "Transpose (First Receive)", "Transpose (Finished Receive)",
    "Transpose (Start All2All)", "Transpose (Finished All2All)", "Transpose (Unpacking)", "1D FFT X-Direction", "Run complete"};
static const char* const SECTIONS_ZYX[] = {"init", "1D FFT Z-Direction", "Transpose (First Send)", "Transpose (Packing)",
    "Transpose (Start Local Transpose)", "Transpose (Start Receive)", "Transpose (First Receive)", "Transpose (Finished Receive)",
    "Transpose (Start All2All)", "Transpose (Finished All2All)", "Transpose (Unpacking)", "2D FFT Y-X-Direction", "Run complete"};
static const char* const SECTIONS_PENCIL[] = {"init", "1D FFT Z-Direction", "First Transpose (First Send)", "First Transpose (Packing)",
    "First Transpose (Start Local Transpose)", "First Transpose (Start Receive)", "First Transpose (First Receive)",
    "First Transpose (Finished Receive)", "First Transpose (Start All2All)", "First Transpose (Finished All2All)",
    "First Transpose (Unpacking)", "First Transpose (Send Complete)", "1D FFT Y-Direction", "Second Transpose (First Send)",
    "Second Transpose (Packing)", "Second Transpose (Start Local Transpose)", "Second Transpose (Start Receive)",
    "Second Transpose (First Receive)", "Second Transpose (Finished Receive)", "Second Transpose (Start All2All)",
    "Second Transpose (Finished All2All)", "Second Transpose (Unpacking)", "1D FFT X-Direction", "Run complete"};

static void plan_sections(const dfft_plan_s* p, const char* const** list, int* n) {
    if (p->g.decomp == DFFT_PENCIL) { *list = SECTIONS_PENCIL; *n = int(sizeof(SECTIONS_PENCIL) / sizeof(char*)); }
    else if (p->g.decomp == DFFT_SLAB_Z_THEN_YX) { *list = SECTIONS_ZYX; *n = int(sizeof(SECTIONS_ZYX) / sizeof(char*)); }
    else { *list = SECTIONS_SLAB; *n = int(sizeof(SECTIONS_SLAB) / sizeof(char*)); }
}

static int timer_gather(dfft_plan_s* p) {
    if (p->n_events_used < 2) return fail(DFFT_ERR_STATE, "no timed exec to gather");
    const char* const* names = nullptr;
    int ns = 0;
    plan_sections(p, &names, &ns);
    std::vector<double> dur(ns, 0.0);
    dur[0] = p->init_ms;
    CK_CUDA(cudaEventSynchronize(p->events[p->n_events_used - 1]));
    double last = 0;
    for (int k = 1; k < p->n_events_used; ++k) {
        float f = 0;
        CK_CUDA(cudaEventElapsedTime(&f, p->events[0], p->events[k]));
        last = f;
        if (!p->ev_names[k]) continue;
        for (int i = 0; i < ns; ++i)
            if (!strcmp(names[i], p->ev_names[k])) dur[i] = f;
    }
    dur[ns - 1] = last;  // "Run complete"
    if (p->n_events_used == 2 && p->tl_used >= 2) {
        // overlapped schedule: the steps ran on several streams and only start / end were marked on the caller's.  The
        // reference's sections are cumulative times, so a section gets the time at which the LAST step of its kind ended
        // (step timeline, dfft_get_timeline).
        const bool pencil = p->g.decomp == DFFT_PENCIL;
        auto section_of = [&](const char* label) -> const char* {
            if (!strncmp(label, "z pass", 6)) return p->g.decomp == DFFT_SLAB_ZY_THEN_X ? nullptr : "1D FFT Z-Direction";
            if (!strncmp(label, "y pass", 6)) return pencil ? "1D FFT Y-Direction" : "2D FFT Y-Z-Direction";
            if (!strncmp(label, "x pass", 6)) return "1D FFT X-Direction";
            if (!strcmp(label, "rendezvous 1")) return "First Transpose (Finished Receive)";
            if (!strcmp(label, "rendezvous 2")) return pencil ? "Second Transpose (Finished Receive)" : "Transpose (Finished Receive)";
            return nullptr;
        };
        for (int i = 0; i + 1 < p->tl_used; i += 2) {
            const char* sec = section_of(p->tl_labels[i]);
            if (!sec) continue;
            float f = 0;
            CK_CUDA(cudaEventSynchronize(p->tl_events[i + 1]));
            CK_CUDA(cudaEventElapsedTime(&f, p->events[0], p->tl_events[i + 1]));
            for (int k = 0; k < ns; ++k)
                if (!strcmp(names[k], sec) && f > dur[k]) dur[k] = f;
        }
    }
    std::vector<char> all;
    int rc = nccl_allgather_bytes(p, dur.data(), sizeof(double) * ns, all);
    if (rc) return rc;
    if (p->rank != 0 || p->csv_path.empty()) return DFFT_SUCCESS;
    const double* o = reinterpret_cast<const double*>(all.data());
    struct stat sb;
    std::ofstream f;
    if (stat(p->csv_path.c_str(), &sb) != 0) {
        f.open(p->csv_path);
        f << ",";
        for (int i = 0; i < p->P; ++i) f << i << ",";
    } else {
        f.open(p->csv_path, std::ios_base::app);
    }
    if (!f) return fail(DFFT_ERR_INVALID, "cannot open " + p->csv_path);
    f << "\n";
    for (int i = 0; i < ns; ++i) {
        f << names[i] << ",";
        for (int r = 0; r < p->P; ++r) f << o[size_t(r) * ns + i] << ",";
        f << "\n";
    }
    return DFFT_SUCCESS;
}

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char* dfft_last_error_string(void) { return g_err.c_str(); }
int dfft_version(void) { return DFFT_VERSION; }

int dfft_get_unique_id(void* id) {
    if (!id) return fail(DFFT_ERR_INVALID, "null id");
    static_assert(sizeof(ncclUniqueId) <= DFFT_UNIQUE_ID_BYTES, "unique id size");
    ncclUniqueId u;
    CK_NCCL(ncclGetUniqueId(&u));
    memset(id, 0, DFFT_UNIQUE_ID_BYTES);
    memcpy(id, &u, sizeof(u));
    return DFFT_SUCCESS;
}

int dfft_comm_create(int rank, int nranks, const void* id, int device, dfft_comm_t* comm) {
    if (!comm || nranks < 1 || rank < 0 || rank >= nranks) return fail(DFFT_ERR_INVALID, "bad rank / nranks");
    if (nranks > MAXSEG) return fail(DFFT_ERR_UNSUPPORTED, "at most " + std::to_string(MAXSEG) + " ranks");
    CK_CUDA(cudaSetDevice(device));
    dfft_comm_s* c = new dfft_comm_s();
    c->rank = rank; c->nranks = nranks; c->device = device;
    if (nranks > 1) {
        if (!id) { delete c; return fail(DFFT_ERR_INVALID, "unique id required for nranks > 1"); }
        ncclUniqueId u;
        memcpy(&u, id, sizeof(u));
        ncclResult_t r = ncclCommInitRank(&c->nccl, nranks, u, rank);
        if (r != ncclSuccess) { delete c; return fail(DFFT_ERR_NCCL, std::string("ncclCommInitRank: ") + ncclGetErrorString(r)); }
    }
    *comm = c;
    return DFFT_SUCCESS;
}
int dfft_comm_create_dry(int rank, int nranks, dfft_comm_t* comm) {
    if (!comm || nranks < 1 || rank < 0 || rank >= nranks) return fail(DFFT_ERR_INVALID, "bad rank / nranks");
    if (nranks > MAXSEG) return fail(DFFT_ERR_UNSUPPORTED, "at most " + std::to_string(MAXSEG) + " ranks");
    dfft_comm_s* c = new dfft_comm_s();
    c->rank = rank; c->nranks = nranks; c->device = -1; c->dry = true;
    *comm = c;
    return DFFT_SUCCESS;
}
int dfft_comm_destroy(dfft_comm_t c) {
    if (!c) return DFFT_SUCCESS;
    if (c->nccl) ncclCommDestroy(c->nccl);
    delete c;
    return DFFT_SUCCESS;
}
int dfft_comm_rank(dfft_comm_t c) { return c ? c->rank : -1; }
int dfft_comm_size(dfft_comm_t c) { return c ? c->nranks : -1; }

int dfft_partition(size_t n, size_t parts, size_t* sizes, size_t* starts) {
    if (!parts || !sizes || !starts) return fail(DFFT_ERR_INVALID, "bad partition request");
    Split s;
    s.make(n, parts);
    for (size_t p = 0; p < parts; ++p) { sizes[p] = s.size[p]; starts[p] = s.start[p]; }
    return DFFT_SUCCESS;
}

int dfft_layout(int decomp, int transform, size_t nx, size_t ny, size_t nz, size_t p1, size_t p2, int rank, int which,
                size_t size[3], size_t start[3]) {
    Geometry g;
    int nranks = decomp == DFFT_PENCIL ? int(p1 * p2) : int(p1);
    if (!g.init(decomp, transform, nx, ny, nz, p1, p2, nranks)) return fail(DFFT_ERR_INVALID, "invalid partition");
    if (rank < 0 || rank >= nranks || which < 0 || which > 3) return fail(DFFT_ERR_INVALID, "bad rank / which");
    g.layout(rank, which, size, start);
    return DFFT_SUCCESS;
}

int dfft_plan_create(dfft_comm_t comm, const dfft_config* config, int decomp, int precision, int transform, size_t nx, size_t ny,
                     size_t nz, size_t p1, size_t p2, int allocate, dfft_plan_t* plan) {
    if (!comm || !plan) return fail(DFFT_ERR_INVALID, "null comm / plan");
    if (precision != DFFT_F32 && precision != DFFT_F64) return fail(DFFT_ERR_INVALID, "bad precision");
    if (transform != DFFT_R2C && transform != DFFT_C2C) return fail(DFFT_ERR_INVALID, "bad transform");
    const bool dry = comm->dry;
    if (!dry) CK_CUDA(cudaSetDevice(comm->device));
    const auto t_init0 = std::chrono::steady_clock::now();
    dfft_plan_s* p = new dfft_plan_s();
    p->comm = comm;
    if (config) {
        p->cfg = *config;
        if (config->benchmark_dir) p->bench_dir = config->benchmark_dir;
        p->cfg.benchmark_dir = nullptr;
    } else {
        p->cfg.comm_method = DFFT_PEER2PEER; p->cfg.comm_method2 = DFFT_PEER2PEER;
    }
    p->rank = comm->rank; p->P = comm->nranks;
    p->prec = precision;
    p->esize = precision == DFFT_F64 ? 16 : 8;
    p->tabs.prec = precision;
    p->tabs.dry = dry;
    if (!p->g.init(decomp, transform, nx, ny, nz, p1, p2, p->P)) {
        delete p;
        return fail(DFFT_ERR_INVALID, "invalid partition: P1*P2 must equal the communicator size");
    }
    const Geometry& g = p->g;
    if (ilog2_exact(nx) < 1 || ilog2_exact(ny) < 1 || ilog2_exact(nz) < 1 || ilog2_exact(nx) > MAX_LOG2N || ilog2_exact(ny) > MAX_LOG2N ||
        ilog2_exact(nz) > MAX_LOG2N + (transform == DFFT_R2C ? 1 : 0) || (transform == DFFT_R2C && nz < 4)) {
        delete p;
        return fail(DFFT_ERR_UNSUPPORTED, "Nx, Ny, Nz must be powers of two in [2, 8192] (Nz >= 4 for R2C)");
    }
    // every rank needs a non-empty share along each split axis
    if (g.sx.size.back() == 0 || g.sy.size.back() == 0 || g.sz.size.back() == 0 || g.oy.size.back() == 0) {
        delete p;
        return fail(DFFT_ERR_INVALID, "partition leaves a rank without data");
    }
    // groups
    const int me = p->rank, P = p->P;
    p->grp[0].clear();
    for (int r = 0; r < P; ++r) p->grp[0].push_back(r);
    if (decomp == DFFT_PENCIL) {
        const int i = g.pi(me), j = g.pj(me);
        for (int q = 0; q < g.P2; ++q) p->grp[1].push_back(i * g.P2 + q);
        for (int q = 0; q < g.P1; ++q) p->grp[2].push_back(q * g.P2 + j);
    } else if (decomp == DFFT_SLAB_ZY_THEN_X) {
        p->grp[1].push_back(me);
        p->grp[2] = p->grp[0];
    } else {
        p->grp[1] = p->grp[0];
        p->grp[2].push_back(me);
    }
    p->direct1 = p->cfg.comm_method != DFFT_ALL2ALL;
    p->direct2 = (decomp == DFFT_PENCIL ? p->cfg.comm_method2 : p->cfg.comm_method) != DFFT_ALL2ALL;
    if (decomp == DFFT_SLAB_ZY_THEN_X) p->direct1 = true;   // trivial group
    if (decomp == DFFT_SLAB_Z_THEN_YX) p->direct2 = true;
    const bool d1 = p->direct1 || p->grp[1].size() == 1, d2 = p->direct2 || p->grp[2].size() == 1;
    p->any_direct = (p->direct1 && p->grp[1].size() > 1) || (p->direct2 && p->grp[2].size() > 1);
    p->nslots = (d1 ? 1 : 0) + (d2 ? 1 : 0) + ((!d1 || !d2) ? 2 : 0);
    if (p->nslots < 2) p->nslots = 2;
    size_t dom = 0;
    for (int r = 0; r < P; ++r) dom = std::max(dom, g.domain_elems(r));
    p->domain_bytes = g.domain_elems(me) * p->esize;
    p->slot_bytes = ((dom * p->esize + 255) / 256) * 256;
    p->work_bytes = p->slot_bytes * p->nslots;
    {
        const char* e = getenv("DFFT_XCHG_CTAS");
        p->xchg_ctas = e ? atoi(e) : 96;  // best of {48, 72, 96} at 8 GPUs, close to best at 2 (profiles/r01_8gpu_b)
        // Blocked intermediate layout for the slab's y -> x hand-over: [Nzc/CH][Nx][Ny_q][CH].  The rows a y-pass tile
        // sends to one destination are adjacent (512-byte warp stores instead of 64-byte rows: what NVLink needs),
        // and the x pass reads rows 16 KB apart instead of one row per 2 MB page (tools/layout_probe.py).
        // The Nzc % CH leftover columns (one for R2C plans: Nzc = Nz/2+1) travel in a small plain-layout tail region
        // behind the blocked part, handled by two tiny extra launches.  DFFT_BLOCKED=0 disables.
        // NVLink store efficiency grows with the contiguous run per row: 64-byte rows reach 434 GB/s per direction,
        // 128-byte rows 700 GB/s, 2 KB runs 704 GB/s (profiles/r01_8gpu, r01_bench_n2_*): exchanging passes prefer the
        // wide tile even though it is slower as a purely local pass.  DFFT_XCHG_WIDE=0 keeps the narrow tile.
        if (const char* eg = getenv("DFFT_OVL_GROUPS")) p->ovl_groups = std::max(1, std::min(16, atoi(eg)));
        if (const char* ec = getenv("DFFT_OVL_CHUNKS")) p->ovl_chunks = std::max(1, std::min(16, atoi(ec)));
        const char* et = getenv("DFFT_RENDEZVOUS_TIMEOUT_S");
        const double tsec = et ? atof(et) : 300.0;
        p->rendezvous_timeout_cycles = tsec > 0 ? (long long)(tsec * 1.9e9) : 0;  // clock64 ticks at <= 1.965 GHz
        const char* ebs = getenv("DFFT_BULK_STORE");
        p->bulk_store = (ebs && atoi(ebs) != 0) ? 1 : 0;
        const char* ew = getenv("DFFT_XCHG_WIDE");
        p->xchg_tile_pref = (ew && atoi(ew) == 0) ? 1 : 2;
        if (const char* ebi = getenv("DFFT_BLOCKED_INV")) p->blocked_inv = atoi(ebi) != 0;
        // one rank: x innermost wins or ties up to 512-point y lines (2.01 ms either way at 512^3); for 1024-point y lines its far
        // output rows cost the y pass 50 % (4.9 vs 3.2 ms at 1024^3 R2C, profiles/r02): use the general blocked layout there
        p->single_rank_layout = ny < 1024 ? 1 : 0;
        if (const char* en1 = getenv("DFFT_N1_LAYOUT")) p->single_rank_layout = atoi(en1) != 0;
        if (const char* esw = getenv("DFFT_X_SWZ")) p->x_swz = std::max(0, std::min(4, atoi(esw)));
        const char* eb = getenv("DFFT_BLOCKED");
        // block width = the widest tile the y and x passes use for these lengths (fft_kernels.cuh: Shape::TBT —
        // 4096 points per tile, rows of at least 64 bytes, at most 32 columns), never below 8 elements
        auto tile_cols = [&](size_t n) {
            size_t w = 4096 / n, lo = precision == DFFT_F64 ? 4 : 8;
            return w < lo ? lo : (w > 32 ? size_t(32) : w);
        };
        int ch = int(std::max<size_t>(8, std::max(tile_cols(ny), tile_cols(nx))));
        if (eb) ch = atoi(eb);
        size_t min_nz = g.nzc;  // smallest z extent any rank holds between the y and x passes
        if (decomp == DFFT_PENCIL)
            for (size_t v : g.sz.size) min_nz = std::min(min_nz, v);
        p->blocked_ch = ((decomp == DFFT_SLAB_ZY_THEN_X || decomp == DFFT_PENCIL) && ch > 0 && min_nz >= size_t(4 * ch)) ? ch : 0;
    }
    if (dry) {
        // fake, rank-distinct slot addresses: ((rank + 1) << 44) + slot * slot_bytes; user buffers are offsets
        p->slot_ptr.assign(p->nslots, std::vector<void*>(P, nullptr));
        for (int s_ = 0; s_ < p->nslots; ++s_)
            for (int r = 0; r < P; ++r) p->slot_ptr[s_][r] = (void*)(((unsigned long long)(r + 1) << 44) + (unsigned long long)s_ * p->slot_bytes);
        p->work = p->slot_ptr[0][me];
        *plan = p;
        return DFFT_SUCCESS;
    }
    // The plan's own stream is a BLOCKING stream: like the reference's cuFFT execs on the legacy default stream
    // (mpicufft_slab.cpp:788-807) the synchronous dfft_exec_* calls are ordered behind everything the caller queued on
    // the default stream before the call — a pageable cudaMemcpy whose DMA is still in flight, a kernel that is still
    // filling `in` (the reference's own testcase 4 does exactly that: random_dist_default.cu:719-724).
    cudaError_t ce = cudaStreamCreate(&p->own_stream);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&p->entry_ev, cudaEventDisableTiming);
    if (ce != cudaSuccess) { delete p; return fail(DFFT_ERR_CUDA, "cudaStreamCreate failed"); }
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
        if (cudaStreamCreateWithPriority(&p->aux[0], cudaStreamNonBlocking, hi) != cudaSuccess ||
            cudaStreamCreateWithPriority(&p->aux[1], cudaStreamNonBlocking, lo) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->fork_ev, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->join_ev[0], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->join_ev[1], cudaEventDisableTiming) != cudaSuccess) {
            delete p;
            return fail(DFFT_ERR_CUDA, "stream / event creation failed");
        }
    }
    int rc = plan_setup_flags(p);
    if (rc == DFFT_SUCCESS && allocate) rc = plan_setup_memory(p, nullptr);
    if (rc != DFFT_SUCCESS) {
        std::string keep = g_err;
        dfft_plan_destroy(p);
        g_err = keep;
        return rc;
    }
    p->init_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_init0).count();
    if (!p->bench_dir.empty()) {
        // <dir>/<variant>/test_0_<comm>_<snd>[_<comm2>_<snd2>]_<Nx>_<Ny>_<Nz>_<cuda_aware>_<P | P1_P2>.csv
        const char* variant = decomp == DFFT_PENCIL ? "pencil" : (decomp == DFFT_SLAB_Z_THEN_YX ? "slab_z_then_yx" : "slab_default");
        mkdir(p->bench_dir.c_str(), 0777);
        mkdir((p->bench_dir + "/" + variant).c_str(), 0777);
        std::string f = p->bench_dir + "/" + variant + "/test_0_" + std::to_string(p->cfg.comm_method) + "_" + std::to_string(p->cfg.send_method);
        if (decomp == DFFT_PENCIL) f += "_" + std::to_string(p->cfg.comm_method2) + "_" + std::to_string(p->cfg.send_method2);
        f += "_" + std::to_string(nx) + "_" + std::to_string(ny) + "_" + std::to_string(nz) + "_" + std::to_string(p->cfg.cuda_aware ? 1 : 0);
        if (decomp == DFFT_PENCIL) f += "_" + std::to_string(g.P1) + "_" + std::to_string(g.P2);
        else f += "_" + std::to_string(P);
        p->csv_path = f + ".csv";
        p->csv_warmup_left = p->cfg.warmup_rounds;
        p->timing = true;
    }
    *plan = p;
    return DFFT_SUCCESS;
}

int dfft_plan_destroy(dfft_plan_t p) {
    if (!p) return DFFT_SUCCESS;
    if (p->comm->dry) {
        p->work = nullptr;
        p->tabs.release();
        delete p;
        return DFFT_SUCCESS;
    }
    cudaSetDevice(p->comm->device);
    cudaDeviceSynchronize();
    plan_release_memory(p);
    for (void* q : p->opened_flags) cudaIpcCloseMemHandle(q);
    if (p->flags) cudaFree(p->flags);
    if (p->peer_flags_d) cudaFree(p->peer_flags_d);
    if (p->groups_d) cudaFree(p->groups_d);
    if (p->err_d) cudaFree(p->err_d);
    p->tabs.release();
    for (cudaEvent_t e : p->events) cudaEventDestroy(e);
    for (cudaEvent_t e : p->tl_events) cudaEventDestroy(e);
    for (cudaEvent_t e : p->sync_events) cudaEventDestroy(e);
    if (p->fork_ev) cudaEventDestroy(p->fork_ev);
    for (int a = 0; a < 2; ++a) {
        if (p->join_ev[a]) cudaEventDestroy(p->join_ev[a]);
        if (p->aux[a]) cudaStreamDestroy(p->aux[a]);
    }
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    if (p->entry_ev) cudaEventDestroy(p->entry_ev);
    delete p;
    return DFFT_SUCCESS;
}

int dfft_set_work_area(dfft_plan_t p, void* device, void* host) {
    (void)host;
    if (!p) return fail(DFFT_ERR_INVALID, "null plan");
    CK_CUDA(cudaSetDevice(p->comm->device));
    CK_CUDA(cudaDeviceSynchronize());
    return plan_setup_memory(p, device);
}

int dfft_exec_r2c(dfft_plan_t p, void* out, const void* in) { return exec_common(p, out, in, 0, 3, DFFT_R2C, nullptr, true); }
int dfft_exec_c2r(dfft_plan_t p, void* out, const void* in) { return exec_common(p, out, in, 1, 3, DFFT_R2C, nullptr, true); }
int dfft_exec_c2c(dfft_plan_t p, void* out, const void* in, int direction) {
    return exec_common(p, out, in, direction > 0 ? 1 : 0, 3, DFFT_C2C, nullptr, true);
}
int dfft_exec_r2c_partial(dfft_plan_t p, void* out, const void* in, int d) { return exec_common(p, out, in, 0, d, DFFT_R2C, nullptr, true); }
int dfft_exec_c2r_partial(dfft_plan_t p, void* out, const void* in, int d) { return exec_common(p, out, in, 1, d, DFFT_R2C, nullptr, true); }
int dfft_exec_c2c_partial(dfft_plan_t p, void* out, const void* in, int direction, int d) {
    return exec_common(p, out, in, direction > 0 ? 1 : 0, d, DFFT_C2C, nullptr, true);
}
int dfft_exec_r2c_async(dfft_plan_t p, void* out, const void* in, void* stream) { return exec_common(p, out, in, 0, 3, DFFT_R2C, stream, false); }
int dfft_exec_c2r_async(dfft_plan_t p, void* out, const void* in, void* stream) { return exec_common(p, out, in, 1, 3, DFFT_R2C, stream, false); }
int dfft_exec_c2c_async(dfft_plan_t p, void* out, const void* in, int direction, void* stream) {
    return exec_common(p, out, in, direction > 0 ? 1 : 0, 3, DFFT_C2C, stream, false);
}

int dfft_plan_wait(dfft_plan_t p) {
    if (!p) return fail(DFFT_ERR_INVALID, "null plan");
    CK_CUDA(cudaStreamSynchronize(p->last_stream));
    int err = 0;
    CK_CUDA(cudaMemcpy(&err, p->err_d, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) {
        cudaMemset(p->err_d, 0, sizeof(int));
        if (err >= 100)
            return fail(DFFT_ERR_TIMEOUT, "a peer rank gave up waiting at a device rendezvous (phase " + std::to_string(err - 100) +
                                              "); results are invalid and the plan must be destroyed");
        return fail(DFFT_ERR_TIMEOUT, "device rendezvous timed out in phase " + std::to_string(err - 1) +
                                          " (a peer rank did not arrive; DFFT_RENDEZVOUS_TIMEOUT_S); results are invalid and the plan must be destroyed");
    }
    return DFFT_SUCCESS;
}

static int layout_of(dfft_plan_t p, int which, size_t size[3], size_t start[3]) {
    if (!p) return fail(DFFT_ERR_INVALID, "null plan");
    size_t s[3], o[3];
    p->g.layout(p->rank, which, s, o);
    for (int k = 0; k < 3; ++k) {
        if (size) size[k] = s[k];
        if (start) start[k] = o[k];
    }
    return DFFT_SUCCESS;
}
int dfft_get_in_size(dfft_plan_t p, size_t size[3]) { return layout_of(p, 0, size, nullptr); }
int dfft_get_in_start(dfft_plan_t p, size_t start[3]) { return layout_of(p, 0, nullptr, start); }
int dfft_get_out_size(dfft_plan_t p, size_t size[3]) { return layout_of(p, 3, size, nullptr); }
int dfft_get_out_start(dfft_plan_t p, size_t start[3]) { return layout_of(p, 3, nullptr, start); }
int dfft_get_partial_size(dfft_plan_t p, int d, size_t size[3]) {
    if (d < 1 || d > 3) return fail(DFFT_ERR_INVALID, "d must be 1..3");
    return layout_of(p, d, size, nullptr);
}
int dfft_get_partial_start(dfft_plan_t p, int d, size_t start[3]) {
    if (d < 1 || d > 3) return fail(DFFT_ERR_INVALID, "d must be 1..3");
    return layout_of(p, d, nullptr, start);
}
size_t dfft_get_domain_size(dfft_plan_t p) { return p ? p->domain_bytes : 0; }
size_t dfft_get_work_size_device(dfft_plan_t p) { return p ? p->work_bytes : 0; }
size_t dfft_get_work_size_host(dfft_plan_t) { return 0; }
void* dfft_get_work_area_device(dfft_plan_t p) { return p ? p->work : nullptr; }
int dfft_get_rank(dfft_plan_t p) { return p ? p->rank : -1; }
int dfft_get_world_size(dfft_plan_t p) { return p ? p->P : -1; }

int dfft_timer_enable(dfft_plan_t p, int enable) {
    if (!p) return fail(DFFT_ERR_INVALID, "null plan");
    p->timing = enable != 0;
    return DFFT_SUCCESS;
}
int dfft_get_phase_count(dfft_plan_t p) {
    if (!p) return 0;
    int n = 0;
    for (int k = 1; k < p->n_events_used; ++k)
        if (p->ev_names[k]) ++n;
    return n;
}
const char* dfft_get_phase_name(dfft_plan_t p, int i) {
    if (!p) return nullptr;
    int n = 0;
    for (int k = 1; k < p->n_events_used; ++k)
        if (p->ev_names[k]) {
            if (n == i) return p->ev_names[k];
            ++n;
        }
    return nullptr;
}
int dfft_get_phase_times(dfft_plan_t p, double* ms, int capacity) {
    if (!p || !ms) return fail(DFFT_ERR_INVALID, "null argument");
    if (p->n_events_used < 2) return fail(DFFT_ERR_STATE, "no timed exec yet (dfft_timer_enable)");
    CK_CUDA(cudaEventSynchronize(p->events[p->n_events_used - 1]));
    int n = 0;
    for (int k = 1; k < p->n_events_used; ++k)
        if (p->ev_names[k]) {
            if (n < capacity) {
                float f = 0;
                CK_CUDA(cudaEventElapsedTime(&f, p->events[0], p->events[k]));
                ms[n] = f;
            }
            ++n;
        }
    return n;
}
int dfft_get_last_breakdown(dfft_plan_t p, double* fft_ms, double* exchange_ms, double* total_ms) {
    if (!p) return fail(DFFT_ERR_INVALID, "null plan");
    if (p->n_events_used < 2) return fail(DFFT_ERR_STATE, "no timed exec yet (dfft_timer_enable)");
    CK_CUDA(cudaEventSynchronize(p->events[p->n_events_used - 1]));
    double f = 0, x = 0;
    for (int k = 1; k < p->n_events_used; ++k) {
        float d = 0;
        CK_CUDA(cudaEventElapsedTime(&d, p->events[k - 1], p->events[k]));
        if (p->ev_is_fft[k] == 1) f += d;
        else if (p->ev_is_fft[k] == 0) x += d;
    }
    float tot = 0;
    CK_CUDA(cudaEventElapsedTime(&tot, p->events[0], p->events[p->n_events_used - 1]));
    if (fft_ms) *fft_ms = f;
    if (exchange_ms) *exchange_ms = x;
    if (total_ms) *total_ms = tot;
    return DFFT_SUCCESS;
}
int dfft_get_last_launch_count(dfft_plan_t p) { return p ? p->last_launches : 0; }
int dfft_get_step_count(dfft_plan_t p) { return p && p->n_events_used > 1 ? p->n_events_used - 2 : 0; }
const char* dfft_get_step_label(dfft_plan_t p, int i) {
    if (!p || i < 0 || i + 1 >= p->n_events_used - 1) return nullptr;
    return p->ev_labels[i + 1];
}
int dfft_get_step_times(dfft_plan_t p, double* ms, int capacity) {
    if (!p || !ms) return fail(DFFT_ERR_INVALID, "null argument");
    if (p->n_events_used < 2) return fail(DFFT_ERR_STATE, "no timed exec yet (dfft_timer_enable)");
    CK_CUDA(cudaEventSynchronize(p->events[p->n_events_used - 1]));
    int n = 0;
    for (int k = 1; k < p->n_events_used - 1; ++k, ++n) {
        if (n >= capacity) continue;
        float f = 0;
        CK_CUDA(cudaEventElapsedTime(&f, p->events[k - 1], p->events[k]));
        ms[n] = f;
    }
    return n;
}

// Plan-time measurement of the execution schedule (the role FFTW_MEASURE plays for FFTW plans; the reference leaves
// the choice between its Sync and Streams variants to the user's benchmarks).  Runs the plan's transform on the
// caller's buffers with each candidate — the sequential schedule and overlapped schedules with different numbers of
// CTAs for the exchanging pass — and keeps the fastest one, judged by the slowest rank.  Collective; `out` is
// overwritten; the input is left intact.  Only plans created with send_method Streams have alternatives.
static std::string cand_name(int seq, int ctas, int groups, int chunks) {
    if (seq) return "sequential";
    return "overlapped/" + std::to_string(ctas) + " CTAs/" + std::to_string(groups) + " groups/" + std::to_string(chunks) + " chunks";
}
int dfft_plan_tune(dfft_plan_t p, void* out, const void* in, int inverse, int reps) {
    if (!p) return fail(DFFT_ERR_INVALID, "null plan");
    if (p->comm->dry) return fail(DFFT_ERR_STATE, "geometry-only plan");
    if (!p->work || !out || !in) return fail(DFFT_ERR_INVALID, "plan needs its work area and both buffers");
    CK_CUDA(cudaSetDevice(p->comm->device));
    const int dir = inverse ? 1 : 0;
    if (reps < 1) reps = 3;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->comm->device);
    // candidates: the sequential schedule, and overlapped schedules over a small grid of (CTAs of the exchanging pass,
    // plane groups of the z pass, z chunks of the y / x passes).  One candidate costs reps + 1 transforms (milliseconds).
    struct Cand { int seq; int ctas; int groups; int chunks; };
    std::vector<Cand> cands;
    cands.push_back({1, 0, 0, 0});
    const bool streams = p->cfg.send_method == DFFT_SEND_STREAMS || (p->g.decomp == DFFT_PENCIL && p->cfg.send_method2 == DFFT_SEND_STREAMS);
    const bool has_overlap = streams && p->P > 1 && p->xchg_ctas >= 0 &&
                             ((p->g.decomp == DFFT_SLAB_ZY_THEN_X && p->direct2) ||
                              (p->g.decomp == DFFT_PENCIL && p->direct1 && p->direct2 && pencil_overlap_enabled()));
    const bool pencil_inv = inverse && p->g.decomp == DFFT_PENCIL;
    const int groups_default = p->ovl_groups, chunks_default = p->ovl_chunks;
    if (has_overlap) {
        const int cta_list[] = {sms / 3, sms / 2, (2 * sms) / 3, (5 * sms) / 6, sms, 2 * sms};
        for (int c : cta_list) {
            if (pencil_inv) {  // plane groups of the y / z passes; the x pass is not chunked
                for (int g_ : {2, 4, 8}) cands.push_back({0, c, g_, 0});
                continue;
            }
            for (int ch : {4, 8}) {
                if (inverse) cands.push_back({0, c, 1, ch});  // the slab inverse has no plane groups (the exchanging x pass comes first)
                else
                    for (int g_ : {1, 2, 4}) cands.push_back({0, c, g_, ch});
            }
        }
    }
    p->tune_report.clear();
    if (cands.size() == 1) {
        p->tune_report = "sequential schedule (no alternatives for this plan)";
        return 0;
    }
    const bool was_timing = p->timing;
    auto apply = [&](const Cand& c) {
        p->tuned_seq[dir] = c.seq;
        p->tuned_ctas[dir] = c.seq ? -2 : c.ctas;
        p->tuned_chunks[dir] = c.seq ? 0 : c.chunks;
        if (!inverse) p->ovl_groups = (!c.seq && c.groups > 0) ? c.groups : groups_default;
        else if (pencil_inv) p->tuned_groups_inv = c.seq ? 0 : c.groups;
        p->sched[dir][2] = Schedule();
    };
    // whatever way this function is left: events destroyed, the phase timer back on, and — on an error in the middle of
    // the measurements — the plan back on the sequential schedule
    struct Cleanup {
        std::function<void()> f;
        ~Cleanup() { f(); }
    };
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool finished = false;
    Cleanup cleanup{[&] {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        p->timing = was_timing;
        if (!finished) apply(cands[0]);
    }};
    CK_CUDA(cudaEventCreate(&e0));
    CK_CUDA(cudaEventCreate(&e1));
    p->timing = false;
    int best = 0;
    double best_ms = 1e30;
    std::string rep;
    for (size_t k = 0; k < cands.size(); ++k) {
        apply(cands[k]);
        Schedule* sc = nullptr;
        int rc = get_schedule(p, inverse, 3, &sc);
        if (rc) continue;  // a candidate this geometry cannot build: skip it (every rank takes the same branch)
        for (int it = 0; it < reps + 1; ++it) {
            if (it == 1) CK_CUDA(cudaEventRecord(e0, p->own_stream));
            rc = run_schedule(p, *sc, out, in, p->own_stream);
            if (rc) return rc;
        }
        CK_CUDA(cudaEventRecord(e1, p->own_stream));
        rc = dfft_plan_wait(p);
        if (rc) return rc;
        float ms = 0;
        CK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double mine = ms / reps;
        std::vector<char> all;
        rc = nccl_allgather_bytes(p, &mine, sizeof(double), all);
        if (rc) return rc;
        double worst = 0;
        for (int r = 0; r < p->P; ++r) worst = std::max(worst, reinterpret_cast<const double*>(all.data())[r]);
        rep += (k ? ", " : "") + cand_name(cands[k].seq, cands[k].ctas, cands[k].groups, cands[k].chunks) + " " + std::to_string(worst).substr(0, 6) + " ms";
        if (worst < best_ms) { best_ms = worst; best = int(k); }
    }
    apply(cands[best]);
    finished = true;
    (void)chunks_default;
    p->tune_report = std::string(inverse ? "inverse: " : "forward: ") + rep + " -> " +
                     cand_name(cands[best].seq, cands[best].ctas, cands[best].groups, cands[best].chunks);
    return best;
}
const char* dfft_plan_tune_report(dfft_plan_t p) { return p ? p->tune_report.c_str() : nullptr; }

// Timeline of the last timed exec: step i ran on plan stream `stream[i]` (0 caller's, 1 exchange, 2 follow-up) from
// begin_ms[i] to end_ms[i] after the start of the exec.  The way to look at an overlapped (Streams) schedule.
int dfft_get_timeline(dfft_plan_t p, double* begin_ms, double* end_ms, int* stream, int capacity) {
    if (!p) return fail(DFFT_ERR_INVALID, "null plan");
    if (p->n_events_used < 2 || p->tl_used < 2) return fail(DFFT_ERR_STATE, "no timed exec yet (dfft_timer_enable)");
    CK_CUDA(cudaEventSynchronize(p->events[p->n_events_used - 1]));
    const int n = p->tl_used / 2;
    for (int i = 0; i < n && i < capacity; ++i) {
        float a = 0, b = 0;
        CK_CUDA(cudaEventSynchronize(p->tl_events[2 * i + 1]));
        CK_CUDA(cudaEventElapsedTime(&a, p->events[0], p->tl_events[2 * i]));
        CK_CUDA(cudaEventElapsedTime(&b, p->events[0], p->tl_events[2 * i + 1]));
        if (begin_ms) begin_ms[i] = a;
        if (end_ms) end_ms[i] = b;
        if (stream) stream[i] = p->tl_streams[2 * i];
    }
    return n;
}
const char* dfft_get_timeline_label(dfft_plan_t p, int i) {
    if (!p || i < 0 || 2 * i >= p->tl_used) return nullptr;
    return p->tl_labels[2 * i];
}
// JSON description of a schedule (test hook: tests/test_schedule_emulation.py replays it with numpy)
static void json_view(std::string& o, const View& v, const Tables& T) {
    o += "{\"nseg\":" + std::to_string(v.nseg) + ",\"sN\":" + std::to_string(v.sN) + ",\"segs\":[";
    for (int i = 0; i < v.nseg; ++i) {
        const Seg& g = v.seg[i];
        if (i) o += ",";
        o += "{\"base\":" + std::to_string((unsigned long long)g.base) + ",\"sA0\":" + std::to_string(g.sA0) + ",\"sA1\":" + std::to_string(g.sA1) +
             ",\"n0\":" + std::to_string(g.n0) + "}";
    }
    o += "],\"seg_of_n\":[";
    if (v.nseg > 1) {
        auto it = T.host_tabs.find(v.seg_of_n);
        if (it != T.host_tabs.end())
            for (size_t i = 0; i < it->second.size(); ++i) { if (i) o += ","; o += std::to_string(int(it->second[i])); }
    }
    o += "]}";
}
int dfft_plan_describe(dfft_plan_t p, int inverse, int d, char* buf, size_t capacity, size_t* needed) {
    if (!p || d < 1 || d > 3) return fail(DFFT_ERR_INVALID, "bad arguments");
    Schedule* scp = nullptr;
    {
        int rc = get_schedule(p, inverse, d, &scp);
        if (rc) return rc;
    }
    Schedule& sc = *scp;
    std::string o = "{\"rank\":" + std::to_string(p->rank) + ",\"esize\":" + std::to_string(p->esize) + ",\"slot_bytes\":" + std::to_string(p->slot_bytes) +
                    ",\"nslots\":" + std::to_string(p->nslots) + ",\"overlapped\":" + (sc.overlapped ? "true" : "false") + ",\"slots\":[";
    for (int s_ = 0; s_ < p->nslots; ++s_) {
        if (s_) o += ",";
        o += "[";
        for (int r = 0; r < p->P; ++r) { if (r) o += ","; o += std::to_string((unsigned long long)p->slot_ptr[s_][r]); }
        o += "]";
    }
    o += "],\"steps\":[";
    bool first = true;
    for (const Step& s_ : sc.steps) {
        if (!first) o += ",";
        first = false;
        o += "{\"type\":" + std::to_string(int(s_.type)) + ",\"label\":\"" + s_.label + "\",\"stream\":" + std::to_string(s_.stream) +
             ",\"record\":" + std::to_string(s_.record) + ",\"waits\":[";
        for (size_t w = 0; w < s_.waits.size(); ++w) o += (w ? "," : "") + std::to_string(s_.waits[w]);
        o += "]";
        if (s_.type == STEP_PASS) {
            o += ",\"kind\":" + std::to_string(int(s_.kind)) + ",\"log2n\":" + std::to_string(s_.log2n) + ",\"A0\":" + std::to_string(s_.prm.A0) +
                 ",\"A1\":" + std::to_string(s_.prm.A1) + ",\"B\":" + std::to_string(s_.prm.B) + ",\"inverse\":" + std::to_string(s_.prm.inverse) +
                 ",\"in_user\":" + std::to_string(s_.in_user) + ",\"out_user\":" + std::to_string(s_.out_user) + ",\"in\":";
            json_view(o, s_.prm.in, p->tabs);
            o += ",\"out\":";
            json_view(o, s_.prm.out, p->tabs);
        } else if (s_.type == STEP_RENDEZVOUS) {
            o += ",\"group\":" + std::to_string(s_.group) + ",\"phase_id\":" + std::to_string(s_.phase_id) + ",\"members\":[";
            const std::vector<int>& G = p->grp[s_.group];
            for (size_t q = 0; q < G.size(); ++q) o += (q ? "," : "") + std::to_string(G[q]);
            o += "]";
        } else {
            o += ",\"group\":" + std::to_string(s_.group) + ",\"send_slot\":" + std::to_string(s_.send_slot) + ",\"recv_slot\":" + std::to_string(s_.recv_slot) + ",\"peers\":[";
            const std::vector<int>& G = p->grp[s_.group];
            for (size_t q = 0; q < G.size(); ++q) {
                if (q) o += ",";
                o += "{\"rank\":" + std::to_string(G[q]) + ",\"scount\":" + std::to_string(s_.scount[q]) + ",\"soff\":" + std::to_string(s_.soff[q]) +
                     ",\"rcount\":" + std::to_string(s_.rcount[q]) + ",\"roff\":" + std::to_string(s_.roff[q]) + "}";
            }
            o += "]";
        }
        o += "}";
    }
    o += "]}";
    if (needed) *needed = o.size() + 1;
    if (buf && capacity > o.size()) memcpy(buf, o.c_str(), o.size() + 1);
    else if (buf && capacity) buf[0] = 0;
    return DFFT_SUCCESS;
}
int dfft_timer_gather(dfft_plan_t p) {
    if (!p) return fail(DFFT_ERR_INVALID, "null plan");
    return timer_gather(p);
}
const char* dfft_timer_csv_path(dfft_plan_t p) { return p ? p->csv_path.c_str() : nullptr; }

// ---- single-axis building blocks -----------------------------------------------------------------------
static std::map<std::pair<int, int>, void*> g_tw_cache, g_tw2_cache;  // (prec, log2n)

static int cached_table(bool second, int prec, int log2n, void** out) {
    int dev = 0;
    cudaGetDevice(&dev);
    auto& cache = second ? g_tw2_cache : g_tw_cache;
    auto key = std::make_pair(prec * 64 + dev, log2n);
    auto it = cache.find(key);
    if (it == cache.end()) {
        void* d = nullptr;
        size_t n = size_t(1) << log2n;
        cudaError_t e;
        if (!second) e = prec == DFFT_F64 ? make_table<double>(n, n, &d) : make_table<float>(n, n, &d);
        else e = prec == DFFT_F64 ? make_table<double>(n / 2 + 1, 2 * n, &d) : make_table<float>(n / 2 + 1, 2 * n, &d);
        if (e != cudaSuccess) return fail(DFFT_ERR_CUDA, "twiddle table allocation failed");
        it = cache.emplace(key, d).first;
    }
    *out = it->second;
    return DFFT_SUCCESS;
}

int dfft_fft1d_contig(int precision, int kind, int direction, size_t n, size_t lines, void* out, size_t out_pitch, const void* in,
                      size_t in_pitch, void* stream) {
    if (precision != DFFT_F32 && precision != DFFT_F64) return fail(DFFT_ERR_INVALID, "bad precision");
    if (kind < 0 || kind > 2) return fail(DFFT_ERR_INVALID, "bad kind");
    const size_t len = kind == 0 ? n : n / 2;
    const int l2 = ilog2_exact(len);
    if (ilog2_exact(n) < 1 || l2 < 1 || l2 > MAX_LOG2N) return fail(DFFT_ERR_UNSUPPORTED, "length must be a power of two");
    FftParams prm{};
    prm.A0 = 1; prm.A1 = int(lines); prm.B = 1;
    prm.inverse = (kind == 0 && direction > 0) ? 1 : 0;
    void *tw = nullptr, *tw2 = nullptr;
    int rc = cached_table(false, precision, l2, &tw);
    if (rc) return rc;
    if (kind != 0) { rc = cached_table(true, precision, l2, &tw2); if (rc) return rc; }
    prm.tw = tw; prm.tw2 = tw2;
    // pitches are in elements of the respective array (reals for real arrays); views are in complex units
    long long ip = (long long)in_pitch, op = (long long)out_pitch;
    if (kind == 1) { if (in_pitch & 1) return fail(DFFT_ERR_INVALID, "real pitch must be even"); ip /= 2; }
    if (kind == 2) { if (out_pitch & 1) return fail(DFFT_ERR_INVALID, "real pitch must be even"); op /= 2; }
    prm.in = single_view(const_cast<void*>(in), 0, ip, 1);
    prm.out = single_view(out, 0, op, 1);
    PassKind pk = kind == 0 ? PASS_C2C_CONTIG : (kind == 1 ? PASS_R2C : PASS_C2R);
    cudaError_t e = precision == DFFT_F64 ? launch_pass_f64(l2, pk, prm, (cudaStream_t)stream) : launch_pass_f32(l2, pk, prm, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(DFFT_ERR_CUDA, std::string("launch failed: ") + cudaGetErrorString(e));
    return DFFT_SUCCESS;
}

int dfft_fft1d_strided(int precision, int direction, size_t a, size_t n, size_t b, void* out, const void* in, void* stream) {
    if (precision != DFFT_F32 && precision != DFFT_F64) return fail(DFFT_ERR_INVALID, "bad precision");
    const int l2 = ilog2_exact(n);
    if (l2 < 1 || l2 > MAX_LOG2N) return fail(DFFT_ERR_UNSUPPORTED, "length must be a power of two");
    FftParams prm{};
    prm.A0 = int(a); prm.A1 = 1; prm.B = int(b);
    prm.inverse = direction > 0 ? 1 : 0;
    void* tw = nullptr;
    int rc = cached_table(false, precision, l2, &tw);
    if (rc) return rc;
    prm.tw = tw;
    prm.in = single_view(const_cast<void*>(in), (long long)(n * b), 0, (long long)b);
    prm.out = single_view(out, (long long)(n * b), 0, (long long)b);
    cudaError_t e = precision == DFFT_F64 ? launch_pass_f64(l2, PASS_C2C_TILED, prm, (cudaStream_t)stream)
                                          : launch_pass_f32(l2, PASS_C2C_TILED, prm, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(DFFT_ERR_CUDA, std::string("launch failed: ") + cudaGetErrorString(e));
    return DFFT_SUCCESS;
}

int dfft_fft1d_general(int precision, int direction, size_t n, size_t a0, size_t a1, size_t b, void* out, const long long out_strides[3],
                       const void* in, const long long in_strides[3], void* stream) {
    if (precision != DFFT_F32 && precision != DFFT_F64) return fail(DFFT_ERR_INVALID, "bad precision");
    const int l2 = ilog2_exact(n);
    if (l2 < 1 || l2 > MAX_LOG2N) return fail(DFFT_ERR_UNSUPPORTED, "length must be a power of two");
    FftParams prm{};
    prm.A0 = int(a0); prm.A1 = int(a1); prm.B = int(b);
    prm.inverse = direction > 0 ? 1 : 0;
    void* tw = nullptr;
    int rc = cached_table(false, precision, l2, &tw);
    if (rc) return rc;
    prm.tw = tw;
    prm.in = single_view(const_cast<void*>(in), in_strides[0], in_strides[1], in_strides[2]);
    prm.out = single_view(out, out_strides[0], out_strides[1], out_strides[2]);
    cudaError_t e = precision == DFFT_F64 ? launch_pass_f64(l2, PASS_C2C_TILED, prm, (cudaStream_t)stream)
                                          : launch_pass_f32(l2, PASS_C2C_TILED, prm, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(DFFT_ERR_CUDA, std::string("launch failed: ") + cudaGetErrorString(e));
    return DFFT_SUCCESS;
}

}  // extern "C"
