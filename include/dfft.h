/* dfft.h — C ABI of the B200-native distributed 3D FFT (libdfft.so).
 *
 * Drop-in boundary for the plan/execute surface of eggersn/DistributedFFT. The reference exposes a C++
 * class hierarchy, one object per MPI rank, one rank per GPU:
 *     MPIcuFFT<T>            /root/reference/include/mpicufft.hpp:55-105
 *     MPIcuFFT_Slab<T>       /root/reference/include/mpicufft_slab.hpp:88-125
 *     MPIcuFFT_Slab_Z_Then_YX<T>  /root/reference/include/mpicufft_slab_z_then_yx.hpp
 *     MPIcuFFT_Pencil<T>     /root/reference/include/mpicufft_pencil.hpp:71-122
 * This header is what a binding of those classes would link against: plain pointers and sizes, no C++
 * or torch types.  include/dfft.hpp re-creates the reference's class and method names on top of it.
 *
 * Process model: identical to the reference — one process (rank) per GPU.  MPI_Comm is replaced by
 * dfft_comm_t, which wraps an NCCL communicator; the 128-byte unique id is produced on rank 0 by
 * dfft_get_unique_id and broadcast by whatever launcher the caller already has (MPI_Bcast,
 * torch.distributed.broadcast_object_list, a file).
 *
 * All device pointers are CUDA device pointers on the communicator's device.  Every function returns
 * DFFT_SUCCESS (0) or a negative dfft_status; dfft_last_error_string() describes the last failure of
 * the calling thread.  (The reference prints and calls exit(EXIT_FAILURE) —
 * /root/reference/src/slab/default/mpicufft_slab.cpp:23-29 — the C++ shim keeps that behaviour.)
 */
#ifndef DFFT_H_
#define DFFT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFFT_VERSION 100
#define DFFT_UNIQUE_ID_BYTES 128

typedef enum {
    DFFT_SUCCESS = 0,
    DFFT_ERR_INVALID = -1,     /* bad argument / unsupported size */
    DFFT_ERR_CUDA = -2,        /* CUDA runtime failure */
    DFFT_ERR_NCCL = -3,        /* NCCL failure */
    DFFT_ERR_STATE = -4,       /* plan not initialised / no work area */
    DFFT_ERR_UNSUPPORTED = -5, /* valid request this build cannot serve */
    DFFT_ERR_PEER = -6,        /* peer memory mapping (CUDA IPC) unavailable */
    DFFT_ERR_TIMEOUT = -7      /* a device-side rendezvous timed out */
} dfft_status;

/* enum CommunicationMethod {Peer2Peer, All2All}  — /root/reference/include/params.hpp:83
 *   Peer2Peer: the FFT kernels store straight into the peers' receive buffers over NVLink
 *              (CUDA-IPC mapped peer memory), no pack, no send buffer, no unpack.
 *   All2All:   the FFT kernels store into per-destination send slots, NCCL grouped send/recv
 *              (all-to-all-v) moves them, the next FFT pass gathers from the receive slots. */
typedef enum { DFFT_PEER2PEER = 0, DFFT_ALL2ALL = 1 } dfft_comm_method;
/* enum SendMethod {Sync, Streams, MPI_Type} — /root/reference/include/params.hpp:84.
 * Accepted for source compatibility; Streams selects the chunk-pipelined schedule where available,
 * MPI_Type is treated like Sync (there is never a separate pack step to elide). */
typedef enum { DFFT_SEND_SYNC = 0, DFFT_SEND_STREAMS = 1, DFFT_SEND_MPI_TYPE = 2 } dfft_send_method;

/* struct Configurations — /root/reference/include/params.hpp:85-93 (same fields, same order). */
typedef struct {
    int cuda_aware;            /* kept for the CSV file name; device buffers are always used */
    int warmup_rounds;         /* execs whose phase times are not recorded */
    int comm_method;           /* dfft_comm_method, first (or only) transposition */
    int send_method;           /* dfft_send_method */
    const char* benchmark_dir; /* where the phase-timer CSV goes; NULL = no CSV */
    int comm_method2;          /* pencil: second transposition */
    int send_method2;
} dfft_config;

/* Which reference class the plan stands for. */
typedef enum {
    DFFT_SLAB_ZY_THEN_X = 0, /* MPIcuFFT_Slab: 2D (y,z) FFT, transpose x-split -> y-split, 1D x FFT */
    DFFT_SLAB_Z_THEN_YX = 1, /* MPIcuFFT_Slab_Z_Then_YX: 1D z FFT, transpose x-split -> z-split, 2D (y,x) */
    DFFT_PENCIL = 2          /* MPIcuFFT_Pencil: z FFT, transpose, y FFT, transpose, x FFT on a P1 x P2 grid */
} dfft_decomp;

typedef enum { DFFT_F32 = 0, DFFT_F64 = 1 } dfft_prec; /* cuFFT<float> / cuFFT<double>, include/cufft.hpp:22-60 */

/* The reference only offers real<->complex (execR2C/execC2R).  DFFT_C2C plans add the complex
 * transform the benchmark configurations name; the z extent of every complex array is then Nz instead
 * of Nz/2+1. */
typedef enum { DFFT_R2C = 0, DFFT_C2C = 1 } dfft_transform;

typedef struct dfft_comm_s* dfft_comm_t;
typedef struct dfft_plan_s* dfft_plan_t;

/* ---- communicator (replaces MPI_Comm / MPI_Comm_size / MPI_Comm_rank, src/mpicufft.cpp:42-50) ---- */
int dfft_get_unique_id(void* id /* DFFT_UNIQUE_ID_BYTES */);
/* nranks == 1: id may be NULL and NCCL is not touched. device = CUDA ordinal this rank drives. */
int dfft_comm_create(int rank, int nranks, const void* id, int device, dfft_comm_t* comm);
int dfft_comm_destroy(dfft_comm_t comm);
int dfft_comm_rank(dfft_comm_t comm);
int dfft_comm_size(dfft_comm_t comm);

/* ---- plan = constructor + initFFT (mpicufft_slab.cpp:97-233, mpicufft_pencil.cpp:63-342) ---------- */
/* p1,p2: Partition{P1,P2} (params.hpp:39-56); slab decompositions ignore them and use the communicator
 * size.  allocate != 0: the plan allocates (and owns) its device work area, like initFFT(.., true). */
int dfft_plan_create(dfft_comm_t comm, const dfft_config* config, int decomp /*dfft_decomp*/,
                     int precision /*dfft_prec*/, int transform /*dfft_transform*/, size_t nx, size_t ny,
                     size_t nz, size_t p1, size_t p2, int allocate, dfft_plan_t* plan);
int dfft_plan_destroy(dfft_plan_t plan);
/* setWorkArea(device, host) — mpicufft_slab.cpp:236-281. device must hold dfft_get_work_size_device()
 * bytes; with comm_method Peer2Peer and more than one rank it must be the base of a cudaMalloc
 * allocation (it is exported to the peers with CUDA IPC).  host is accepted and unused. */
int dfft_set_work_area(dfft_plan_t plan, void* device, void* host);

/* ---- execute (mpicufft_slab.cpp:771-871, mpicufft_pencil.cpp:1643-1839) ----------------------------
 * Collective over the communicator, unnormalised in both directions (inverse(forward(x)) = Nx*Ny*Nz*x).
 * Layouts, z contiguous:
 *   real/complex input  [in_size[0]][in_size[1]][in_size[2]]     (dfft_get_in_size)
 *   complex output      [out_size[0]][out_size[1]][out_size[2]]  (dfft_get_out_size)
 * `out` of exec_r2c / `in` of exec_c2r must hold dfft_get_domain_size() bytes.
 * The plain calls return after the device has finished (the reference's cudaDeviceSynchronize);
 * the _async calls enqueue on `stream` (a cudaStream_t) and return. */
int dfft_exec_r2c(dfft_plan_t plan, void* out, const void* in);
int dfft_exec_c2r(dfft_plan_t plan, void* out, const void* in);
/* direction: -1 forward (DFFT_FORWARD), +1 inverse; plan must be DFFT_C2C. */
#define DFFT_FORWARD (-1)
#define DFFT_INVERSE (1)
int dfft_exec_c2c(dfft_plan_t plan, void* out, const void* in, int direction);
/* pencil execR2C(out,in,d)/execC2R(out,in,d): transform only the first d dimensions (z; z,y; z,y,x). */
int dfft_exec_r2c_partial(dfft_plan_t plan, void* out, const void* in, int d);
int dfft_exec_c2r_partial(dfft_plan_t plan, void* out, const void* in, int d);
int dfft_exec_c2c_partial(dfft_plan_t plan, void* out, const void* in, int direction, int d);

int dfft_exec_r2c_async(dfft_plan_t plan, void* out, const void* in, void* stream);
int dfft_exec_c2r_async(dfft_plan_t plan, void* out, const void* in, void* stream);
int dfft_exec_c2c_async(dfft_plan_t plan, void* out, const void* in, int direction, void* stream);
/* Waits for the stream of the last _async call and reports a device-side rendezvous failure, if any. */
int dfft_plan_wait(dfft_plan_t plan);

/* ---- getters (mpicufft.hpp:65-78, mpicufft_slab.hpp:122-125, mpicufft_pencil.hpp:112-122) ---------- */
int dfft_get_in_size(dfft_plan_t plan, size_t size[3]);
int dfft_get_in_start(dfft_plan_t plan, size_t start[3]);
int dfft_get_out_size(dfft_plan_t plan, size_t size[3]);
int dfft_get_out_start(dfft_plan_t plan, size_t start[3]);
/* layout after d transformed dimensions (pencil partial transforms); d = 3 equals get_out_*. */
int dfft_get_partial_size(dfft_plan_t plan, int d, size_t size[3]);
int dfft_get_partial_start(dfft_plan_t plan, int d, size_t start[3]);
size_t dfft_get_domain_size(dfft_plan_t plan);      /* bytes, getDomainSize() */
size_t dfft_get_work_size_device(dfft_plan_t plan); /* bytes, getWorkSizeDevice() */
size_t dfft_get_work_size_host(dfft_plan_t plan);   /* always 0: nothing is staged through the host */
void* dfft_get_work_area_device(dfft_plan_t plan);
int dfft_get_rank(dfft_plan_t plan);
int dfft_get_world_size(dfft_plan_t plan);

/* ---- partition arithmetic without a GPU (initFFT's split rule, mpicufft_slab.cpp:112-128,
 *      mpicufft_pencil.cpp:89-110): sizes[p] = n/parts + (p < n%parts), starts = prefix sums. -------- */
int dfft_partition(size_t n, size_t parts, size_t* sizes, size_t* starts);
/* Geometry of any rank of a would-be plan, no device needed (used by the CPU tests and by callers that
 * size their buffers before creating the plan).  which: 0 in, 1 after d=1, 2 after d=2, 3 out. */
int dfft_layout(int decomp, int transform, size_t nx, size_t ny, size_t nz, size_t p1, size_t p2,
                int rank, int which, size_t size[3], size_t start[3]);

/* ---- plan-time measurement ---------------------------------------------------------------------------------
 * The reference offers Sync and Streams variants of every transposition and leaves the choice to the user's own
 * benchmark sweeps (jobs/ ** /benchmarks_base.json).  dfft_plan_tune makes it once, at plan time, like FFTW_MEASURE:
 * it runs the transform on the caller's buffers with the sequential schedule and with overlapped schedules that give the
 * exchanging pass different numbers of CTAs, and keeps the fastest (slowest rank decides).  Collective; `out` is
 * overwritten, `in` is not.  Only plans with send_method Streams have alternatives.  Returns the candidate index. */
int dfft_plan_tune(dfft_plan_t plan, void* out, const void* in, int inverse, int reps);
const char* dfft_plan_tune_report(dfft_plan_t plan);

/* ---- phase timer (include/timer.hpp, src/timer.cpp; section names mpicufft_slab.hpp:209-223,
 *      mpicufft_pencil.hpp:263-287) -------------------------------------------------------------------
 * Cumulative milliseconds since the start of the last exec, measured with CUDA events. */
int dfft_timer_enable(dfft_plan_t plan, int enable);
int dfft_get_phase_count(dfft_plan_t plan);
const char* dfft_get_phase_name(dfft_plan_t plan, int i);
int dfft_get_phase_times(dfft_plan_t plan, double* ms, int capacity);
/* Per-step GPU time of the last timed exec, in launch order: one entry per FFT pass / rendezvous / all-to-all
 * (labels such as "z pass", "y pass", "x pass", "rendezvous 2", "nccl all-to-all"). */
int dfft_get_step_count(dfft_plan_t plan);
const char* dfft_get_step_label(dfft_plan_t plan, int i);
int dfft_get_step_times(dfft_plan_t plan, double* ms, int capacity);
/* Timeline of the last timed exec (also for overlapped schedules, whose steps run on three streams): step i ran on
 * plan stream stream[i] (0 the caller's, 1 exchange, 2 follow-up) from begin_ms[i] to end_ms[i] after the exec's start.
 * Returns the number of steps. */
int dfft_get_timeline(dfft_plan_t plan, double* begin_ms, double* end_ms, int* stream, int capacity);
const char* dfft_get_timeline_label(dfft_plan_t plan, int i);
/* Collective: gathers the section times of the last timed exec to rank 0, which appends one block to the CSV
 * in the reference's schema (src/timer.cpp:58-101).  Called automatically after every non-warm-up
 * synchronous exec when Configurations::benchmark_dir is set. */
int dfft_timer_gather(dfft_plan_t plan);
const char* dfft_timer_csv_path(dfft_plan_t plan);
/* GPU time of the local FFT passes and of the exchange steps of the last timed exec (ms). */
int dfft_get_last_breakdown(dfft_plan_t plan, double* fft_ms, double* exchange_ms, double* total_ms);

/* number of kernels the last exec launched (FFT passes + rendezvous kernels; NCCL's own not counted) */
int dfft_get_last_launch_count(dfft_plan_t plan);

const char* dfft_last_error_string(void);
int dfft_version(void);

/* ---- single-axis building block (what one cufftExec* of the reference's plans computes) ------------
 * Batched 1D transform of `lines` contiguous lines of length n (power of two) on the current device.
 * kind: 0 C2C, 1 R2C (n real -> n/2+1 complex), 2 C2R.  Exposed for tests and micro-benchmarks. */
int dfft_fft1d_contig(int precision, int kind, int direction, size_t n, size_t lines, void* out,
                      size_t out_pitch, const void* in, size_t in_pitch, void* stream);
/* Batched 1D C2C along the middle axis of a [a][n][b] array (b contiguous). */
int dfft_fft1d_strided(int precision, int direction, size_t a, size_t n, size_t b, void* out, const void* in,
                       void* stream);

/* ---- test hooks -----------------------------------------------------------------------------------------
 * Geometry-only communicator: plans created on it never touch CUDA or NCCL (slot addresses are synthetic,
 * ((rank+1) << 44) + slot * slot_bytes) and cannot execute; dfft_plan_describe returns the step list of a
 * schedule (passes with their views, rendezvous, all-to-all counts) as JSON so that the CPU tests can replay
 * the data movement of any rank count with numpy (tests/test_schedule_emulation.py). */
int dfft_comm_create_dry(int rank, int nranks, dfft_comm_t* comm);
int dfft_plan_describe(dfft_plan_t plan, int inverse, int d, char* buf, size_t capacity, size_t* needed);

/* Batched 1D C2C along n of element (i0, i1, n, ib) at  i0*strides[0] + i1*strides[1] + n*strides[2] + ib
 * (elements; ib contiguous, extents a0 x a1 x n x b) — the general strided view the plan's passes use. */
int dfft_fft1d_general(int precision, int direction, size_t n, size_t a0, size_t a1, size_t b, void* out,
                       const long long out_strides[3], const void* in, const long long in_strides[3], void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFFT_H_ */
