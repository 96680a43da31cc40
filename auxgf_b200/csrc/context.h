// The opaque context of libagf2_b200 (agf2_b200_ctx): one per (process, device) or more, each with its own CUDA
// streams, scratch memory, instrumentation and -- optionally -- an NCCL communicator (one rank per context).
// Replaces the per-process OpenMP team + mpi4py communicator of the reference (auxgf/lib/libagf2/optragf2.c:95-115,
// auxgf/util/mpi.py:25-69).  Shared by agf2_b200.cu (moment build, general GEMM, communicator), ao2mo.cu (DF
// transform, J/K) and solver.cu (compression, eigensolver); the kernels themselves live in agf2_b200.cu only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/agf2_b200.h"

namespace agf2 {

struct K1Tile;
struct K2Tile;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

// ---- errors: text of the last failure of the calling thread (agf2_b200_last_error) ------------------------
std::string& last_error();
int fail(int code, const std::string& msg);
#define CU_TRY(expr)                                                                          \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return agf2::fail(_e == cudaErrorMemoryAllocation ? AGF2_B200_ENOMEM : AGF2_B200_ECUDA, \
                              std::string(#expr) + ": " + cudaGetErrorString(_e));            \
    } while (0)

// ---- tuning / debug options (process defaults are shared by the per-device default contexts) --------------
struct Options {
    int64_t ws_bytes = 384ll << 20;    // (xi|ja) workspace per ring buffer (upper bound: small calls allocate what they need)
    bool simple = false;               // naive debug kernels
    bool lpt_order = true;             // k1_form tiles of a chunk launched longest first
    bool tune_chunk = true;            // chunk size chosen for whole waves of k1_form tiles
    bool overlap = true;               // odd chunks run on a second stream: one chain's tail is filled by the other
    bool thin_tiles = true;            // ragged last x-block uses the thin-tile mapping of k1_form
    bool coulomb = true;               // Coulomb factorisation allowed ...
    bool coulomb_auto = true;          // ... and chosen per call by a flop estimate
    bool deterministic = true;         // stream-K partial tiles summed in a fixed order (k2_fixup)
    bool profile = true;               // CUDA events around every launch (agf2_b200_last_stage_ms)
    int64_t vis_bytes = 2ll << 30;     // pole-sharded builds: size of one visiting sub-block of (ix|Q) (two are kept)
    bool w_evict_last = false;         // (xi|ja) workspace stores carry the L2 evict_last hint
    int64_t l2_persist_bytes = 0;      // > 0: cudaLimitPersistingL2CacheSize set at context creation
    int64_t shard_limit = 0;           // > 0: only the first `shard_limit` sub-blocks of every ring step (benchmark sample)
};
Options& default_options();

struct EventPair { cudaEvent_t a, b; int kind; };

struct DevBuf {                        // grow-only device scratch
    double* p = nullptr;
    size_t cap = 0;                    // in doubles
};

}  // namespace agf2

struct agf2_b200_ctx {
    std::mutex mu;                     // one call at a time per context
    int device = -1;
    int num_sms = 148;
    bool is_default = false;
    agf2::Options own_opts;
    agf2::Options* opt = &own_opts;
    agf2::EncodeTiledFn encode = nullptr;
    cudaStream_t stream = nullptr, stream2 = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_join = nullptr, ev_copy = nullptr;
    // (xi|ja) workspace ring, energy weights, accumulators (one set of 4 per stream), stream-K slots (per stream)
    double* W[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    size_t w_elems = 0;
    agf2::DevBuf E[2], acc, partial[2], packed;
    // tile lists
    agf2::K1Tile* d_t1[2] = {nullptr, nullptr};
    agf2::K1Tile* h_t1[2] = {nullptr, nullptr};
    size_t t1_cap = 0;
    cudaEvent_t t1_free[2] = {nullptr, nullptr};
    void* d_small = nullptr; size_t d_small_cap = 0;    // K2 tile lists + cost prefixes of one call (bytes)
    void* h_small = nullptr; size_t h_small_cap = 0;    // pinned staging of the same
    cudaEvent_t small_free = nullptr;
    void* d_gemm = nullptr; size_t d_gemm_cap = 0;      // tile list + prefix of general GEMM launches (device-generated)
    // Coulomb factorisation scratch
    agf2::DevBuf Mbuf, Tbuf, wq[2], we;
    // pole-sharded builds: double-buffered visiting sub-blocks of (ix|Q) and their events
    agf2::DevBuf vis[2];
    cudaEvent_t vis_ready[2] = {nullptr, nullptr};          // sub-block has landed (copy stream)
    cudaEvent_t vis_done[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [buffer][compute stream]: last reader finished
    // host-pointer API staging
    agf2::DevBuf stage[8];
    double* h_result = nullptr; size_t h_result_cap = 0;   // pinned D2H staging of [vv | vev]
    double* h_energy = nullptr; size_t h_energy_cap = 0;   // pinned D2H staging of the pole energies
    std::vector<cudaEvent_t> pole_ev;
    // communicator (NCCL, one rank per context)
    void* comm = nullptr;
    int rank = 0, nranks = 1;
    // instrumentation
    int64_t launches = 0;
    int64_t chunk_counter = 0;
    std::vector<agf2::EventPair> events;
    size_t events_used = 0;
    int64_t last_plan[6] = {0, 0, 0, 0, 0, 0};   // factorised, PAIR, DIAG, OS, ASYM items, k1_form launches
    double last_ms[4] = {0, 0, 0, 0};            // k1_form, k2_moment (pairs), M build, Coulomb step (GEMM + k2)
    double flops[3] = {0, 0, 0};                 // executed FP64 tensor flops since the last reset: k1_form, k2_moment, GEMM
    int64_t last_comm_bytes[3] = {0, 0, 0};      // bytes this rank sent over NVLink in the last call: all-gather, all-reduce,
                                                 // point-to-point (visiting pole blocks of a pole-sharded build)
};

namespace agf2 {
typedef agf2_b200_ctx Ctx;

// RAII: make the context's device current for the duration of a call
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) changed = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
};

int grow(DevBuf& b, size_t need, bool zero);
int ctx_for_current_device(Ctx** out);      // the default context of the calling thread's current device
int make_map(Ctx& c, CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
             const cuuint64_t* strides_bytes, const cuuint32_t* box);

// C = A B^T on the DMMA / TMA / stream-K pipeline of k2_moment (general-GEMM instantiation); all operands f64 with
// the contraction index contiguous, leading dimensions even (16-byte TMA strides), bases 16-byte aligned.
// Used for T = ixq [M0|M1] of the Coulomb factorisation, the DF transform and the exchange matrix.
struct Gemm {
    // operand A: a_rows x k, row stride lda; third dimension (stride a_third) is a batch index or the outer K index
    const double* a = nullptr;  int64_t a_rows = 0, lda = 0, a_third = 0;
    const double* a1 = nullptr;           // second A operand of the same shape -> second output c1 (null: plain GEMM)
    const double* b = nullptr;  int64_t b_rows = 0, ldb = 0, b_third = 0;
    int64_t k = 0;                        // contraction length per outer index
    int64_t n_outer = 1;                  // contraction runs over (outer, k); a_third / b_third are its strides
    int64_t nbatch = 1;                   // > 1: a_third / b_third are batch strides (0: operand shared), n_outer == 1
    double* c = nullptr;  double* c1 = nullptr;
    int64_t ldc = 0, c_batch = 0;
    bool transpose = false;               // c[col * ldc + row]
    bool accumulate = false;              // false: c = A B^T (no memset needed), true: c += A B^T
};
int run_gemm(Ctx& c, const Gemm& g, cudaStream_t st);

}  // namespace agf2
