// libagf2_b200.so -- host side of the B200 moment build: contexts, TMA descriptors, launches, the general GEMM
// on the in-house DMMA pipeline, the NCCL communicator and the C-ABI of include/agf2_b200.h.
// Reference paths are relative to obackhouse/auxgf.  The work planner lives in planner.h (pure host code); this
// file executes its chunks.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "context.h"
#include "moment_kernels.cuh"
#include "planner.h"

using namespace agf2;

// =================================================================================================
// errors, options, contexts
// =================================================================================================
namespace agf2 {

std::string& last_error() {
    thread_local std::string e;
    return e;
}
int fail(int code, const std::string& msg) {
    last_error() = msg;
    return code;
}

// Environment overrides of the planner / scheduling knobs (read once; no CUDA needed).
Options& default_options() {
    static Options o;
    static bool applied = false;
    if (!applied) {
        applied = true;
        if (const char* e = getenv("AGF2_B200_OVERLAP")) o.overlap = atoi(e) != 0;
        if (const char* e = getenv("AGF2_B200_WORKSPACE_MB")) o.ws_bytes = (int64_t)atoll(e) << 20;
        if (const char* e = getenv("AGF2_B200_SIMPLE")) o.simple = atoi(e) != 0;
        if (const char* e = getenv("AGF2_B200_LPT")) o.lpt_order = atoi(e) != 0;
        if (const char* e = getenv("AGF2_B200_TUNE_CHUNK")) o.tune_chunk = atoi(e) != 0;
        if (const char* e = getenv("AGF2_B200_COULOMB")) { o.coulomb = atoi(e) != 0; o.coulomb_auto = atoi(e) == 2; }
        if (const char* e = getenv("AGF2_B200_THIN")) o.thin_tiles = atoi(e) != 0;
        if (const char* e = getenv("AGF2_B200_DETERMINISTIC")) o.deterministic = atoi(e) != 0;
        if (const char* e = getenv("AGF2_B200_VISITING_MB")) o.vis_bytes = (int64_t)atoll(e) << 20;
        if (const char* e = getenv("AGF2_B200_W_EVICT_LAST")) o.w_evict_last = atoi(e) != 0;
        if (const char* e = getenv("AGF2_B200_L2_PERSIST_MB")) o.l2_persist_bytes = (int64_t)atoll(e) << 20;
    }
    return o;
}

int grow(DevBuf& b, size_t need, bool zero) {
    if (need <= b.cap) return 0;
    if (b.p) {
        // the old buffer may still be read by work in flight on one of this device's streams
        CU_TRY(cudaDeviceSynchronize());
        cudaFree(b.p);
    }
    b.p = nullptr;
    b.cap = 0;
    CU_TRY(cudaMalloc(&b.p, need * sizeof(double)));
    if (zero) CU_TRY(cudaMemset(b.p, 0, need * sizeof(double)));
    b.cap = need;
    return 0;
}

int make_map(Ctx& c, CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
             const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = c.encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, const_cast<void*>(base), dims,
                          strides_bytes, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(AGF2_B200_ECUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 0;
}

}  // namespace agf2

namespace {

std::mutex g_ctx_mu;
std::map<int, Ctx*> g_default_ctx;      // device -> default context (legacy entry points without a ctx argument)
std::atomic<int64_t> g_launches{0};     // kernels launched by this library, all contexts

int ctx_init(Ctx& c, int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(AGF2_B200_ENODEV, "no CUDA device visible (libagf2_b200 has no CPU fallback)");
    if (device < 0) CU_TRY(cudaGetDevice(&device));
    if (device >= ndev) return fail(AGF2_B200_EINVAL, "no such device");
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(AGF2_B200_ENODEV, std::string("device is sm_") + std::to_string(prop.major) +
                                          std::to_string(prop.minor) + ", this library is built for sm_100a only");
    c.device = device;
    c.num_sms = prop.multiProcessorCount;
    if (default_options().l2_persist_bytes > 0)
        CU_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize,
                                  (size_t)std::min<int64_t>(default_options().l2_persist_bytes, prop.persistingL2CacheMaxSize)));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (fn == nullptr || q != cudaDriverEntryPointSuccess)
        return fail(AGF2_B200_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
    c.encode = reinterpret_cast<EncodeTiledFn>(fn);
    CU_TRY(cudaFuncSetAttribute(k1_form, cudaFuncAttributeMaxDynamicSharedMemorySize, kK1SmemBytes));
    CU_TRY(cudaFuncSetAttribute(k2_moment<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK2SmemBytes));
    CU_TRY(cudaFuncSetAttribute(k2_moment<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK2SmemBytes));
    CU_TRY(cudaFuncSetAttribute(k2_moment<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK2SmemBytesDual));
    CU_TRY(cudaFuncSetAttribute(k2_moment<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK2SmemBytesDual));
    int lo_prio = 0, hi_prio = 0;
    CU_TRY(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    CU_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking));
    // uploads and the all-gathers that follow them must not queue behind the compute kernels' CTAs
    CU_TRY(cudaStreamCreateWithPriority(&c.copy_stream, cudaStreamNonBlocking, hi_prio));
    for (int b = 0; b < 2; ++b) CU_TRY(cudaEventCreateWithFlags(&c.t1_free[b], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&c.ev_start, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&c.ev_copy, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&c.small_free, cudaEventDisableTiming));
    for (int b = 0; b < 2; ++b) {
        CU_TRY(cudaEventCreateWithFlags(&c.vis_ready[b], cudaEventDisableTiming));
        for (int s = 0; s < 2; ++s) CU_TRY(cudaEventCreateWithFlags(&c.vis_done[b][s], cudaEventDisableTiming));
    }
    return 0;
}

void free_buf(DevBuf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }

int comm_destroy(Ctx& c);

void ctx_free(Ctx* c) {
    if (!c) return;
    if (c->device >= 0) {
        DeviceGuard guard(c->device);
        cudaDeviceSynchronize();
        comm_destroy(*c);
        for (int r = 0; r < 2; ++r)
            for (int b = 0; b < 2; ++b) if (c->W[r][b]) cudaFree(c->W[r][b]);
        for (int b = 0; b < 2; ++b) {
            free_buf(c->E[b]); free_buf(c->partial[b]); free_buf(c->wq[b]); free_buf(c->vis[b]);
            if (c->vis_ready[b]) cudaEventDestroy(c->vis_ready[b]);
            for (int s = 0; s < 2; ++s) if (c->vis_done[b][s]) cudaEventDestroy(c->vis_done[b][s]);
            if (c->d_t1[b]) cudaFree(c->d_t1[b]);
            if (c->h_t1[b]) cudaFreeHost(c->h_t1[b]);
            if (c->t1_free[b]) cudaEventDestroy(c->t1_free[b]);
        }
        free_buf(c->acc); free_buf(c->packed); free_buf(c->Mbuf); free_buf(c->Tbuf); free_buf(c->we);
        for (int s = 0; s < 8; ++s) free_buf(c->stage[s]);
        if (c->d_small) cudaFree(c->d_small);
        if (c->h_small) cudaFreeHost(c->h_small);
        if (c->d_gemm) cudaFree(c->d_gemm);
        if (c->h_result) cudaFreeHost(c->h_result);
        if (c->h_energy) cudaFreeHost(c->h_energy);
        for (cudaEvent_t e : c->pole_ev) cudaEventDestroy(e);
        for (EventPair& p : c->events) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
        for (cudaEvent_t e : {c->ev_start, c->ev_join, c->ev_copy, c->small_free}) if (e) cudaEventDestroy(e);
        for (cudaStream_t s : {c->stream, c->stream2, c->copy_stream}) if (s) cudaStreamDestroy(s);
    }
    delete c;
}

}  // namespace

namespace agf2 {
int ctx_for_current_device(Ctx** out) {
    int ndev = 0, device = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(AGF2_B200_ENODEV, "no CUDA device visible (libagf2_b200 has no CPU fallback)");
    CU_TRY(cudaGetDevice(&device));
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    auto it = g_default_ctx.find(device);
    if (it == g_default_ctx.end()) {
        Ctx* c = new Ctx();
        c->is_default = true;
        c->opt = &default_options();
        if (int r = ctx_init(*c, device)) { delete c; return r; }
        it = g_default_ctx.emplace(device, c).first;
    }
    *out = it->second;
    return 0;
}
}  // namespace agf2

namespace {

// pinned host staging, grow-only
template <typename T>
int grow_pinned(T*& p, size_t& cap, size_t need) {
    if (need <= cap) return 0;
    if (p) { CU_TRY(cudaDeviceSynchronize()); cudaFreeHost(p); }
    p = nullptr; cap = 0;
    CU_TRY(cudaMallocHost(&p, need * sizeof(T)));
    cap = need;
    return 0;
}

void record(Ctx& c, cudaStream_t st, int kind, bool begin) {
    if (!c.opt->profile) return;
    if (begin) {
        if (c.events_used == c.events.size()) {
            EventPair p;
            cudaEventCreate(&p.a);
            cudaEventCreate(&p.b);
            c.events.push_back(p);
        }
        c.events[c.events_used].kind = kind;
        cudaEventRecord(c.events[c.events_used].a, st);
    } else {
        cudaEventRecord(c.events[c.events_used].b, st);
        c.events_used++;
    }
}

void count_launch(Ctx& c, int n = 1) { c.launches += n; g_launches += n; }

// =================================================================================================
// k2_moment launches (moments, M build, Coulomb step) and the general GEMM
// =================================================================================================
struct K2Launch {
    CUtensorMap a0, a1, b;
    K2Params p;
    int mode;          // 0: moments (W W^T), 1: M build (two weight vectors), 2: Coulomb step (two A operands), 3: GEMM
    long long total_cost;    // cost_prefix[ntiles] (time model, sixteenths of a full tile)
    long long total_work = -1;   // executed DMMA work in the same unit (-1: equal to total_cost)
};

K2Params k2_params() {
    K2Params p;
    memset(&p, 0, sizeof(p));
    p.a_row_step = kTileM;
    p.mask = 1;
    return p;
}

// `slot`: which of the context's stream-K scratch buffers the launch may use (one per stream of the pair loop)
int launch_k2(Ctx& c, K2Launch& L, int slot, cudaStream_t st) {
    // stream-K grid: one CTA per SM, unless the launch is too small to give each >= 8 iterations
    const long long total_it = L.total_cost * L.p.nk_inner * L.p.n_outer / 16;
    const int grid = (int)std::max<long long>(1, std::min<long long>(c.num_sms, total_it / 8));
    const bool det = c.opt->deterministic || (L.mode == 3 && L.p.store);
    if (det) {
        if (int r = grow(c.partial[slot], (size_t)c.num_sms * 2 * kTileM * kTileN, false)) return r;
        L.p.partial = c.partial[slot].p;
    } else {
        L.p.partial = nullptr;
    }
    if (L.mode == 0) k2_moment<false, false><<<grid, kThreads, kK2SmemBytes, st>>>(L.a0, L.a1, L.b, L.p);
    else if (L.mode == 1) k2_moment<false, true><<<grid, kThreads, kK2SmemBytes, st>>>(L.a0, L.a1, L.b, L.p);
    else if (L.mode == 2) k2_moment<true, false><<<grid, kThreads, kK2SmemBytesDual, st>>>(L.a0, L.a1, L.b, L.p);
    else k2_moment<true, false, true><<<grid, kThreads, kK2SmemBytesDual, st>>>(L.a0, L.a1, L.b, L.p);
    CU_TRY(cudaGetLastError());
    count_launch(c);
    // executed DMMA flops: two 128 x 64 accumulators per full tile, K padded to whole 16-element stages
    c.flops[L.mode == 3 ? 2 : 1] += (double)(L.total_work >= 0 ? L.total_work : L.total_cost) / 16.0 * 2.0 * 2.0 * kTileM * kTileN * (double)kKB *
                                    L.p.nk_inner * L.p.n_outer;
    if (det && grid > 1) {
        if (L.mode == 3) k2_fixup<true><<<dim3(grid, kFixupParts), 256, 0, st>>>(L.p, grid);
        else k2_fixup<false><<<dim3(grid, kFixupParts), 256, 0, st>>>(L.p, grid);
        CU_TRY(cudaGetLastError());
        count_launch(c);
    }
    return 0;
}

}  // namespace

namespace agf2 {

int run_gemm(Ctx& c, const Gemm& g, cudaStream_t st) {
    if (!g.a || !g.b || !g.c || g.a_rows <= 0 || g.b_rows <= 0 || g.k <= 0 || g.nbatch < 1 || g.n_outer < 1)
        return fail(AGF2_B200_EINVAL, "run_gemm: bad argument");
    if ((g.lda & 1) || (g.ldb & 1) || (g.a_third & 1) || (g.b_third & 1) ||
        (reinterpret_cast<uintptr_t>(g.a) & 15) || (reinterpret_cast<uintptr_t>(g.b) & 15) ||
        (g.a1 && (reinterpret_cast<uintptr_t>(g.a1) & 15)))
        return fail(AGF2_B200_EINVAL, "run_gemm: operands need even leading dimensions and 16-byte aligned bases");
    if (g.nbatch > 1 && g.n_outer > 1) return fail(AGF2_B200_EINVAL, "run_gemm: batch and outer-K are exclusive");
    const bool plain = g.a1 == nullptr;
    const int64_t row_step = plain ? 2 * kTileM : kTileM;
    const int64_t nxb = (g.a_rows + row_step - 1) / row_step, nyb = (g.b_rows + kTileN - 1) / kTileN;
    const int64_t ntiles = nxb * nyb * g.nbatch;
    if (ntiles > (1ll << 26) || g.a_rows > (1ll << 30) || g.b_rows > (1ll << 30))
        return fail(AGF2_B200_EINVAL, "run_gemm: too many tiles");
    const size_t need = (size_t)ntiles * sizeof(K2Tile) + (size_t)(ntiles + 1) * sizeof(int) + 64;
    if (need > c.d_gemm_cap) {
        if (c.d_gemm) { CU_TRY(cudaDeviceSynchronize()); cudaFree(c.d_gemm); }
        c.d_gemm = nullptr; c.d_gemm_cap = 0;
        CU_TRY(cudaMalloc(&c.d_gemm, need * 2));
        c.d_gemm_cap = need * 2;
    }
    K2Tile* d_tiles = reinterpret_cast<K2Tile*>(c.d_gemm);
    int* d_prefix = reinterpret_cast<int*>(d_tiles + ntiles);
    gen_gemm_tiles<<<(unsigned)((ntiles + 1 + 255) / 256), 256, 0, st>>>(d_tiles, d_prefix, (int)nxb, (int)nyb, ntiles);
    CU_TRY(cudaGetLastError());
    count_launch(c);

    const int64_t third = std::max<int64_t>(g.nbatch, g.n_outer);
    K2Launch L;
    // 3-D operand map (k, rows, third); when the third (batch / outer-K) stride is the smaller one the two outer
    // dimensions are listed in ascending stride order, (k, third, rows), and the kernel swaps the coordinates
    auto map3 = [&](CUtensorMap* m, const double* base, int64_t rows, int64_t ld, int64_t third_stride, bool swap) -> int {
        const bool has3 = third > 1 && third_stride > 0;
        if (swap) {
            cuuint64_t d[3] = {(cuuint64_t)g.k, (cuuint64_t)third, (cuuint64_t)rows};
            cuuint64_t s[2] = {(cuuint64_t)third_stride * 8, (cuuint64_t)ld * 8};
            cuuint32_t b[3] = {kKB, 1, 64};
            return make_map(c, m, base, 3, d, s, b);
        }
        cuuint64_t d[3] = {(cuuint64_t)g.k, (cuuint64_t)rows, (cuuint64_t)(has3 ? third : 1)};
        cuuint64_t s[2] = {(cuuint64_t)ld * 8, (cuuint64_t)(has3 ? third_stride : rows * ld) * 8};
        cuuint32_t b[3] = {kKB, 64, 1};
        return make_map(c, m, base, 3, d, s, b);
    };
    const bool swap_a = third > 1 && g.a_third > 0 && g.a_third < g.lda;
    const bool swap_b = third > 1 && g.b_third > 0 && g.b_third < g.ldb;
    if (int r = map3(&L.a0, g.a, g.a_rows, g.lda, g.a_third, swap_a)) return r;
    if (plain) {
        // second accumulator = rows +128 of the same operand (256 x 64 output tiles)
        if (g.a_rows > kTileM) { if (int r = map3(&L.a1, g.a + kTileM * g.lda, g.a_rows - kTileM, g.lda, g.a_third, swap_a)) return r; }
        else L.a1 = L.a0;
    } else {
        if (int r = map3(&L.a1, g.a1, g.a_rows, g.lda, g.a_third, swap_a)) return r;
    }
    if (int r = map3(&L.b, g.b, g.b_rows, g.ldb, g.b_third, swap_b)) return r;
    L.mode = 3;
    L.total_cost = 16 * ntiles;
    K2Params p = k2_params();
    p.tiles = d_tiles; p.cost_prefix = d_prefix;
    p.acc_vv = g.c;
    p.acc_vev = plain ? (g.transpose ? g.c + kTileM : g.c + kTileM * g.ldc) : (g.c1 ? g.c1 : g.c);
    p.ldacc = g.ldc;
    p.ntiles = (int)ntiles;
    p.nk_inner = (int)((g.k + kKB - 1) / kKB);
    p.n_outer = (int)g.n_outer;
    p.rows_valid = (int)g.a_rows; p.cols_valid = (int)g.b_rows;
    p.vev_rows_valid = plain ? (int)std::max<int64_t>(0, g.a_rows - kTileM) : (g.c1 ? (int)g.a_rows : 0);
    p.sym = 0; p.mask = 0;
    p.swap_a = swap_a ? 1 : 0; p.swap_b = swap_b ? 1 : 0;
    p.a_row_step = (int)row_step;
    p.transpose_out = g.transpose ? 1 : 0;
    p.store = g.accumulate ? 0 : 1;
    p.a_batch = (g.nbatch > 1 && g.a_third > 0) ? 1 : 0;
    p.b_batch = (g.nbatch > 1 && g.b_third > 0) ? 1 : 0;
    p.out_batch_stride = g.c_batch;
    L.p = p;
    return launch_k2(c, L, 0, st);
}

}  // namespace agf2

namespace {

PlanConfig plan_config(const Options& o, int num_sms) {
    PlanConfig p;
    p.ws_bytes = o.ws_bytes; p.num_sms = num_sms; p.simple = o.simple; p.lpt_order = o.lpt_order;
    p.tune_chunk = o.tune_chunk; p.thin_tiles = o.thin_tiles; p.coulomb = o.coulomb; p.coulomb_auto = o.coulomb_auto;
    return p;
}

struct Spin {            // one (Q|ja) tensor and its energies
    const double* qja = nullptr;
    int64_t ld = 0, nocc = 0, nvir = 0;
    const double* ea_dev = nullptr;
    const double* ei_dev = nullptr;
};

// Host-pointer API only: ixq is uploaded pole by pole on a copy stream by a helper thread while the GPU
// already works on the pairs whose poles have landed.  rank[i] = position of pole i in the upload order,
// ev[k] is recorded once the k-th pole is in HBM, ready = number of poles whose event has been recorded.
struct PoleGate {
    const std::vector<int>* rank;
    const std::vector<cudaEvent_t>* ev;
    std::atomic<int>* ready;
    std::atomic<int>* failed;
    bool with_qja = false;     // the (Q|ja) columns of a pole travel with it: the Coulomb stage then needs every pole
};

int gate_wait(const PoleGate* gate, int need_rank, cudaStream_t sc) {
    while (gate->ready->load(std::memory_order_acquire) <= need_rank) {
        if (gate->failed->load()) return fail(AGF2_B200_ECUDA, "ixq upload failed");
        std::this_thread::yield();
    }
    CU_TRY(cudaStreamWaitEvent(sc, (*gate->ev)[need_rank], 0));
    return 0;
}

// Small per-call tables (K2 tile lists, cost prefixes): built in pinned memory, copied with one async memcpy.
struct SmallTables {
    Ctx& c;
    size_t used = 0;
    explicit SmallTables(Ctx& ctx) : c(ctx) {}
    int reserve(size_t bytes) {
        if (bytes > c.d_small_cap) {
            CU_TRY(cudaDeviceSynchronize());
            if (c.d_small) cudaFree(c.d_small);
            if (c.h_small) cudaFreeHost(c.h_small);
            c.d_small = nullptr; c.h_small = nullptr; c.d_small_cap = c.h_small_cap = 0;
            CU_TRY(cudaMalloc(&c.d_small, bytes * 2));
            CU_TRY(cudaMallocHost(&c.h_small, bytes * 2));
            c.d_small_cap = c.h_small_cap = bytes * 2;
        } else {
            CU_TRY(cudaEventSynchronize(c.small_free));   // an earlier (asynchronous) call may still be copying from it
        }
        return 0;
    }
    template <typename T>
    const T* put(const std::vector<T>& v) {
        used = (used + 15) & ~(size_t)15;
        memcpy(static_cast<char*>(c.h_small) + used, v.data(), v.size() * sizeof(T));
        const T* d = reinterpret_cast<const T*>(static_cast<char*>(c.d_small) + used);
        used += v.size() * sizeof(T);
        return d;
    }
    int upload(cudaStream_t st) {
        CU_TRY(cudaMemcpyAsync(c.d_small, c.h_small, used, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaEventRecord(c.small_free, st));
        return 0;
    }
};

struct SymTiles { const K2Tile* tiles; const int* cost; int ntiles; int total_cost; long long total_work; };

int comm_allgather(Ctx& c, double* buf, int64_t count_per_rank, cudaStream_t st);

// Coulomb part of the moments without forming a single Coulomb integral:
//     sum_{i in slice} sum_{j,a} w_j X_ij X_ij^T (e_i + e_j - e_a)^n  =  sum_i ixq_i (e_i^n M0 + n M1) ixq_i^T ,
//     M0 = sum_ja w_j q_ja q_ja^T ,  M1 = sum_ja w_j (e_j - e_a) q_ja q_ja^T      (naux x naux)
// (k2_moment<false,true> on the (Q|ja) tensor), T = ixq_batch [M0 | M1] (general GEMM on the same pipeline) and
// k2_moment<true,false>: vv += T0 ixq^T, vev += (e_i T0 + T1) ixq^T, upper triangle only.
// M is kept as  Mt = [M0[Q', :] ; M1[Q', :]]  (rows = the B operand of the T GEMM, contraction index Q contiguous).
// Two ways to share the work between GPUs:
//   replicated ixq (gather = false): part p of nparts owns the block of rows Q' of M and the matching K range of the
//       last step, for ALL poles of the slice;
//   pole-sharded ixq (gather = true): rank p builds its block of rows of M, an all-gather gives every rank the whole
//       M, and each rank treats its own poles (ixq = the resident block of `iend - istart` poles, `pole_off` = global
//       index of its first pole) with the full K range, one k2 launch per block of rows.
int run_coulomb(Ctx& c, const double* ixq, int64_t ld_ixq, int64_t ixq_poles, const Spin& same, const Spin* other,
                int64_t nphys, int64_t naux, int64_t istart, int64_t iend, double w_same_in, double w_same_out,
                double w_other, int64_t part, int64_t nparts, bool gather, int64_t pole_off, double* S_vv, double* S_vev,
                int64_t p_pad, const SymTiles& sym, const K2Tile* d_tm, const int* d_cm, int nt_m, int cost_m,
                long long work_m, cudaStream_t st)
{
    struct Pass { const Spin* sp; double w_in, w_out; };
    std::vector<Pass> passes;
    const bool partial = !gather && (istart > 0 || iend < same.nocc);
    if (w_same_in != 0.0 || (partial && w_same_out != 0.0)) passes.push_back({&same, w_same_in, w_same_out});
    if (other && w_other != 0.0) passes.push_back({other, w_other, w_other});
    if (passes.empty() || (iend <= istart && !gather)) return 0;
    const int64_t ntq = (naux + kTileM - 1) / kTileM;
    auto blk_lo = [&](int64_t p) { return ntq * p / nparts * kTileM; };
    auto blk_hi = [&](int64_t p) { return std::min<int64_t>(naux, ntq * (p + 1) / nparts * kTileM); };
    const int64_t q_lo = blk_lo(part), q_hi = blk_hi(part);
    const int64_t nq = std::max<int64_t>(0, q_hi - q_lo);
    if (!gather && nq == 0) return 0;
    // every block lives in a slot of nq_slot rows (gather: one slot per rank, equal sizes for the all-gather)
    const int64_t nq_slot = gather ? (ntq + nparts - 1) / nparts * kTileM : rup(nq, kTileM);
    const int64_t nslots = gather ? nparts : 1, my_slot = gather ? part : 0;
    const int64_t ldm = rup(naux, kTileN);
    const int msym = (nparts == 1) ? 1 : 0;   // the whole (symmetric) M: upper-triangle tiles, mirrored afterwards
    const size_t slot_elems = (size_t)2 * nq_slot * ldm;

    // ---- M build ------------------------------------------------------------------------------------
    if (int r = grow(c.Mbuf, slot_elems * nslots, false)) return r;
    double* const Mt0 = c.Mbuf.p + slot_elems * my_slot;
    double* const Mt1 = Mt0 + nq_slot * ldm;
    CU_TRY(cudaMemsetAsync(Mt0, 0, slot_elems * 8, st));
    record(c, st, 2, true);
    for (const Pass& ps : passes) {
        if (nq == 0 || nt_m == 0) break;
        const Spin& sp = *ps.sp;
        const int nki = (int)((sp.nvir + kKB - 1) / kKB);
        const int64_t kpad = (int64_t)nki * kKB;
        for (int b = 0; b < 2; ++b) if (int r = grow(c.wq[b], (size_t)(sp.nocc * kpad), false)) return r;
        dim3 gw((unsigned)((kpad + 127) / 128), (unsigned)sp.nocc);
        const bool is_same = ps.sp == &same;
        const bool sliced = is_same && !gather;
        fill_m_weights<<<gw, 128, 0, st>>>(c.wq[0].p, c.wq[1].p, sp.ei_dev, sp.ea_dev, (int)sp.nvir, (int)kpad, ps.w_in,
                                          ps.w_out, sliced ? (int)istart : 0, sliced ? (int)iend : (int)sp.nocc);
        CU_TRY(cudaGetLastError());
        count_launch(c);
        K2Launch L;
        cuuint64_t d[3] = {(cuuint64_t)sp.nvir, (cuuint64_t)sp.nocc, (cuuint64_t)naux};
        cuuint64_t sb[2] = {(cuuint64_t)sp.ld * 8, (cuuint64_t)sp.ld * 8 * (cuuint64_t)sp.nocc};
        cuuint32_t bx[3] = {kKB, 1, 64};
        if (int r = make_map(c, &L.a0, sp.qja, 3, d, sb, bx)) return r;
        L.a1 = L.a0; L.b = L.a0;
        L.mode = 1;
        L.total_cost = cost_m;
        L.total_work = work_m;
        K2Params p = k2_params();
        p.w0 = c.wq[0].p; p.w1 = c.wq[1].p; p.tiles = d_tm; p.cost_prefix = d_cm;
        p.acc_vv = Mt0; p.acc_vev = Mt1; p.ldacc = ldm;
        p.ntiles = nt_m; p.nk_inner = nki; p.n_outer = (int)sp.nocc;
        p.swap_a = 1; p.swap_b = 1; p.a_row_off = (int)q_lo;
        p.rows_valid = (int)nq; p.cols_valid = (int)naux; p.sym = msym;
        L.p = p;
        if (int r = launch_k2(c, L, 0, st)) return r;
    }
    if (msym) {
        dim3 gm((unsigned)naux, (unsigned)((naux + 127) / 128));
        mirror_m<<<gm, 128, 0, st>>>(Mt0, Mt1, (int)naux, ldm);
        CU_TRY(cudaGetLastError());
        count_launch(c);
    }
    if (gather && nparts > 1) if (int r = comm_allgather(c, c.Mbuf.p, (int64_t)slot_elems, st)) return r;
    record(c, st, 2, false);
    if (iend <= istart) return 0;

    // ---- T = ixq_batch [M0 | M1]^T-rows, then the two moments ---------------------------------------------
    const int64_t ldt = 2 * nq_slot * nslots;
    const int64_t per_pole = nphys * ldt;
    const int64_t nbi_max = std::max<int64_t>(1, std::min<int64_t>(64, (512ll << 20) / (per_pole * 8)));
    const int64_t kpad = rup(gather ? nq_slot : nq, kKB);
    if (int r = grow(c.Tbuf, (size_t)(std::min<int64_t>(nbi_max, iend - istart) * per_pole), false)) return r;
    if (int r = grow(c.we, (size_t)(nbi_max * kpad), false)) return r;
    CUtensorMap mapIx;
    {
        cuuint64_t d[3] = {(cuuint64_t)naux, (cuuint64_t)nphys, (cuuint64_t)ixq_poles};
        cuuint64_t sb[2] = {(cuuint64_t)ld_ixq * 8, (cuuint64_t)ld_ixq * 8 * (cuuint64_t)nphys};
        cuuint32_t bx[3] = {kKB, 64, 1};
        if (int r = make_map(c, &mapIx, ixq, 3, d, sb, bx)) return r;
    }
    record(c, st, 3, true);
    for (int64_t i0 = istart; i0 < iend; i0 += nbi_max) {
        const int64_t nbi = std::min<int64_t>(nbi_max, iend - i0);
        Gemm g;
        g.a = ixq + i0 * nphys * ld_ixq; g.a_rows = nbi * nphys; g.lda = ld_ixq;
        g.b = c.Mbuf.p; g.b_rows = 2 * nq_slot * nslots; g.ldb = ldm;
        g.k = naux;
        g.c = c.Tbuf.p; g.ldc = ldt;
        if (int r = run_gemm(c, g, st)) return r;
        for (int64_t sl = 0; sl < nslots; ++sl) {
            const int64_t s_lo = gather ? blk_lo(sl) : q_lo;
            const int64_t s_nq = gather ? std::max<int64_t>(0, blk_hi(sl) - s_lo) : nq;
            if (s_nq == 0) continue;
            // weights e_i per (pole, K) column, laid out with this launch's K length per pole
            const int64_t kp = rup(s_nq, kKB);
            dim3 ge((unsigned)((kp + 127) / 128), (unsigned)nbi);
            fill_e_weights<<<ge, 128, 0, st>>>(c.we.p, same.ei_dev, (int)(pole_off + i0), (int)kp);
            CU_TRY(cudaGetLastError());
            count_launch(c);
            K2Launch L;
            cuuint64_t d[3] = {(cuuint64_t)s_nq, (cuuint64_t)nphys, (cuuint64_t)nbi};
            cuuint64_t sb[2] = {(cuuint64_t)ldt * 8, (cuuint64_t)ldt * 8 * (cuuint64_t)nphys};
            cuuint32_t bx[3] = {kKB, 64, 1};
            double* const t0 = c.Tbuf.p + 2 * nq_slot * sl;
            if (int r = make_map(c, &L.a0, t0, 3, d, sb, bx)) return r;
            if (int r = make_map(c, &L.a1, t0 + nq_slot, 3, d, sb, bx)) return r;
            L.b = mapIx;
            L.mode = 2;
            L.total_cost = sym.total_cost;
            L.total_work = sym.total_work;
            K2Params p = k2_params();
            p.w1 = c.we.p; p.tiles = sym.tiles; p.cost_prefix = sym.cost;
            p.acc_vv = S_vv; p.acc_vev = S_vev; p.ldacc = p_pad;
            p.ntiles = sym.ntiles; p.nk_inner = (int)((s_nq + kKB - 1) / kKB); p.n_outer = (int)nbi;
            p.b_k_off = (int)s_lo; p.b_outer_off = (int)i0;
            p.rows_valid = (int)nphys; p.cols_valid = (int)nphys; p.sym = 1;
            L.p = p;
            if (int r = launch_k2(c, L, 0, st)) return r;
        }
    }
    record(c, st, 3, false);
    return 0;
}

// What a moments call does after the per-GPU work: nothing, or one NCCL all-reduce of the packed [vv | vev].
struct Reduce { bool on = false; };

int comm_allreduce_sum(Ctx& c, double* buf, int64_t count, cudaStream_t st);
int comm_allreduce(Ctx& c, double* buf, int64_t count, int op, cudaStream_t st);
constexpr int kNcclMaxOp = 2;      // ncclMax
int comm_sendrecv(Ctx& c, const double* sendbuf, int64_t send_count, int dst, double* recvbuf, int64_t recv_count, int src,
                  cudaStream_t st);

// Pole-sharded build: `ixq` holds only the poles [shard_lo(o, rank, R), shard_lo(o, rank + 1, R)) of this rank of the
// context's communicator; the other blocks visit over NVLink (see planner.h, shard_steps).
struct Shard { int64_t rank, nranks; };
// poles per visiting sub-block: what fits the "visiting_bytes" budget, at most the largest block
int64_t shard_sub_block(const Options& o, int64_t nb_max, int64_t pole_elems) {
    return std::max<int64_t>(1, std::min<int64_t>(nb_max, (o.vis_bytes / 8) / pole_elems));
}

// The whole moment build on device-resident inputs; vv/vev are accumulated into (+=).
int run_moments(Ctx& c, const double* ixq, int64_t ld_ixq, const Spin& same, const Spin* other,
                const double* ei_same_host, const double* ei_other_host, int64_t nphys, int64_t naux,
                int64_t istart, int64_t iend, double c_same, double s_same, double os_other, int64_t part,
                int64_t nparts, double* vv, double* vev, cudaStream_t st, const PoleGate* gate, bool allreduce,
                const Shard* sh = nullptr)
{
    const Options& opt = *c.opt;
    const int64_t o = same.nocc, v = same.nvir;
    if (sh) {
        if (other || gate) return fail(AGF2_B200_EINVAL, "pole-sharded builds: RHF, device-resident inputs only");
        if (istart != 0 || iend != o) return fail(AGF2_B200_EINVAL, "pole-sharded builds cover the whole pole range");
        if (o < sh->nranks) return fail(AGF2_B200_EINVAL, "pole-sharded builds need at least one pole per rank");
        part = 0; nparts = 1;
    }
    // poles of the resident (ix|Q) operand, and the share of the Coulomb stage's M rows this GPU builds
    const int64_t n_res = sh ? shard_lo(o, sh->rank + 1, sh->nranks) - shard_lo(o, sh->rank, sh->nranks) : o;
    const int64_t m_part = sh ? sh->rank : part, m_nparts = sh ? sh->nranks : nparts;
    MomentPlan plan;
    if (int r = plan.init(plan_config(opt, c.num_sms), nphys, naux, o, v, other != nullptr, other ? other->nocc : 0,
                          other ? other->nvir : 0, istart, iend, c_same, s_same, os_other, part, nparts, sh != nullptr))
        return fail(r, plan.error);
    if ((ld_ixq & 1) || ld_ixq < naux || (same.ld & 1) || same.ld < v || (other && ((other->ld & 1) || other->ld < other->nvir)))
        return fail(AGF2_B200_EINVAL, "leading dimensions must be even (16-byte TMA strides) and >= the row length");
    if ((reinterpret_cast<uintptr_t>(ixq) & 15) || (reinterpret_cast<uintptr_t>(same.qja) & 15) ||
        (other && (reinterpret_cast<uintptr_t>(other->qja) & 15)))
        return fail(AGF2_B200_EINVAL, "device arrays must be 16-byte aligned");
    const int64_t p_pad = plan.p_pad, cap_cols = plan.cap_cols;
    const int nxb = plan.nxb, nyb = plan.nyb;
    const bool fact = plan.fact;
    const std::vector<Item>& sym_items = plan.sym_items;
    const std::vector<Item>& asym_items = plan.asym_items;
    const bool overlap = opt.overlap && !opt.simple;
    const int nsets = overlap ? 2 : 1;

    // ---- scratch ------------------------------------------------------------------------------
    const size_t w_need = (size_t)p_pad * cap_cols;
    if (w_need > c.w_elems) {
        CU_TRY(cudaDeviceSynchronize());
        for (int r = 0; r < 2; ++r)
            for (int b = 0; b < 2; ++b) { if (c.W[r][b]) cudaFree(c.W[r][b]); c.W[r][b] = nullptr; }
        c.w_elems = 0;
        for (int r = 0; r < 2; ++r)
            for (int b = 0; b < 2; ++b) {
                CU_TRY(cudaMalloc(&c.W[r][b], w_need * 8));
                CU_TRY(cudaMemset(c.W[r][b], 0, w_need * 8));
            }
        c.w_elems = w_need;
    }
    const int64_t ldw = cap_cols;
    for (int r = 0; r < 2; ++r) if (int rc = grow(c.E[r], (size_t)cap_cols + 64, true)) return rc;
    const size_t acc_n = (size_t)p_pad * p_pad;
    // one accumulator set [S_vv | S_vev | A_vv | A_vev] per stream: every element has ONE adder per launch and the
    // launches of a stream are ordered, so the summation order is fixed (bit-reproducible results)
    if (int r = grow(c.acc, (size_t)4 * nsets * acc_n, false)) return r;
    CU_TRY(cudaMemsetAsync(c.acc.p, 0, (size_t)4 * nsets * acc_n * 8, st));
    // Two independent chains: chunk c runs k1_form -> k2_moment on stream (c & 1) with ring buffer (c & 1),
    // so buffer reuse is ordered by the stream itself and the CTAs of one chain fill the SMs that the tail
    // of the other chain's kernel leaves idle.
    if (overlap) {
        CU_TRY(cudaEventRecord(c.ev_start, st));
        CU_TRY(cudaStreamWaitEvent(c.stream2, c.ev_start, 0));
    }
    bool used2 = false;

    // ---- TMA descriptors of the inputs ----------------------------------------------------------
    CUtensorMap mapA, mapB0, mapB1;
    {
        cuuint64_t d[3] = {(cuuint64_t)naux, (cuuint64_t)nphys, (cuuint64_t)n_res};
        cuuint64_t s[2] = {(cuuint64_t)ld_ixq * 8, (cuuint64_t)ld_ixq * 8 * (cuuint64_t)nphys};
        cuuint32_t b[3] = {kKB, kTileM, 1};
        if (int r = make_map(c, &mapA, ixq, 3, d, s, b)) return r;
    }
    {
        cuuint64_t d[3] = {(cuuint64_t)v, (cuuint64_t)o, (cuuint64_t)naux};
        cuuint64_t s[2] = {(cuuint64_t)same.ld * 8, (cuuint64_t)same.ld * 8 * (cuuint64_t)o};
        cuuint32_t b[3] = {16, 1, kKB};
        if (int r = make_map(c, &mapB0, same.qja, 3, d, s, b)) return r;
    }
    if (other) {
        cuuint64_t d[3] = {(cuuint64_t)other->nvir, (cuuint64_t)other->nocc, (cuuint64_t)naux};
        cuuint64_t s[2] = {(cuuint64_t)other->ld * 8, (cuuint64_t)other->ld * 8 * (cuuint64_t)other->nocc};
        cuuint32_t b[3] = {16, 1, kKB};
        if (int r = make_map(c, &mapB1, other->qja, 3, d, s, b)) return r;
    } else {
        mapB1 = mapB0;
    }
    const int nk1 = (int)((naux + kKB - 1) / kKB);

    c.last_plan[0] = fact ? 1 : 0;
    c.last_plan[1] = c.last_plan[2] = c.last_plan[3] = 0;
    for (const Item& it : sym_items) c.last_plan[it.kind == 1 ? 1 : (it.kind == 0 ? 2 : 3)]++;
    c.last_plan[4] = (int64_t)asym_items.size();
    c.last_plan[5] = 0;

    // K2 tile lists (upper-triangle for symmetric chunks, all tiles for non-symmetric ones) and their stream-K
    // cost prefixes: edge tiles (diagonal, padded last row / column) skip the warps that have nothing to do
    std::vector<K2Tile> t2_sym, t2_all, t2_m;
    for (int xb = 0; xb < nxb; ++xb)
        for (int yb = 0; yb < nyb; ++yb) {
            t2_all.push_back({xb, yb, 0, 0});
            if (kTileN * yb + kTileN - 1 >= kTileM * xb) t2_sym.push_back({xb, yb, 0, 0});
        }
    std::vector<int> cost_sym(1, 0), cost_all(1, 0), cost_m(1, 0);
    long long work_sym = 0, work_all = 0, work_m = 0;
    for (const K2Tile& t : t2_sym) {
        cost_sym.push_back(cost_sym.back() + k2_tile_cost(t.xb, t.yb, (int)nphys, (int)nphys, 1, 1));
        work_sym += k2_tile_work(t.xb, t.yb, (int)nphys, (int)nphys, 1, 1);
    }
    for (const K2Tile& t : t2_all) {
        cost_all.push_back(cost_all.back() + k2_tile_cost(t.xb, t.yb, (int)nphys, (int)nphys, 0, 1));
        work_all += k2_tile_work(t.xb, t.yb, (int)nphys, (int)nphys, 0, 1);
    }
    if (fact) {     // tiles of this part's block of M rows
        const int64_t ntq = (naux + kTileM - 1) / kTileM;
        const int64_t q_lo = ntq * m_part / m_nparts * kTileM;
        const int64_t q_hi = std::min<int64_t>(naux, ntq * (m_part + 1) / m_nparts * kTileM);
        const int64_t nq = std::max<int64_t>(0, q_hi - q_lo);
        const int msym = (m_nparts == 1) ? 1 : 0;
        const int nxm = (int)((nq + kTileM - 1) / kTileM), nym = (int)((naux + kTileN - 1) / kTileN);
        for (int xb = 0; xb < nxm; ++xb)
            for (int yb = 0; yb < nym; ++yb) {
                if (msym && kTileN * yb + kTileN - 1 < kTileM * xb) continue;
                t2_m.push_back({xb, yb, 0, 0});
                cost_m.push_back(cost_m.back() + k2_tile_cost(xb, yb, (int)nq, (int)naux, msym, 1));
                work_m += k2_tile_work(xb, yb, (int)nq, (int)naux, msym, 1);
            }
    }
    SmallTables small(c);
    if (int r = small.reserve((t2_sym.size() + t2_all.size() + t2_m.size()) * sizeof(K2Tile) +
                              (cost_sym.size() + cost_all.size() + cost_m.size()) * sizeof(int) + 256))
        return r;
    const K2Tile* d_t2_sym = small.put(t2_sym);
    const K2Tile* d_t2_all = small.put(t2_all);
    const K2Tile* d_t2_m = small.put(t2_m);
    const int* d_cost_sym = small.put(cost_sym);
    const int* d_cost_all = small.put(cost_all);
    const int* d_cost_m = small.put(cost_m);
    if (int r = small.upload(st)) return r;
    if (overlap) {   // stream2 reads the tables too
        CU_TRY(cudaEventRecord(c.ev_start, st));
        CU_TRY(cudaStreamWaitEvent(c.stream2, c.ev_start, 0));
    }

    const CUtensorMap* mapA2 = nullptr;     // visiting pole block (pole-sharded builds only)
    const double* ixq2 = nullptr;
    c.events_used = 0;
    // chunk c of THIS call uses stream / ring buffer / accumulator set (c & 1): the assignment (and with it the
    // summation order) does not depend on what ran before
    int64_t chunk_idx = 0;
    std::vector<K1Tile> tiles;

    auto run_chunks = [&](const std::vector<Item>& items, bool sym) -> int {
        size_t pos = 0;
        while (pos < items.size()) {
            // ---- pack items into one chunk (planner.h) ----
            int need_rank = -1;
            const int64_t col = plan.next_chunk(items, pos, ei_same_host, ei_other_host, gate ? gate->rank : nullptr,
                                                tiles, need_rank);
            if (col == 0) return fail(AGF2_B200_EINVAL, "workspace too small for one work item");

            // ---- upload the tile list (double-buffered pinned staging) ----
            const int buf = (int)(chunk_idx & 1);
            const int set = overlap ? buf : 0;
            cudaStream_t sc = (overlap && buf) ? c.stream2 : st;
            if (sc != st) used2 = true;
            if (tiles.size() > c.t1_cap) {
                CU_TRY(cudaDeviceSynchronize());
                for (int b = 0; b < 2; ++b) {
                    if (c.d_t1[b]) cudaFree(c.d_t1[b]);
                    if (c.h_t1[b]) cudaFreeHost(c.h_t1[b]);
                    c.d_t1[b] = nullptr; c.h_t1[b] = nullptr;
                }
                c.t1_cap = 0;
                const size_t cap = tiles.size() * 2 + 1024;
                for (int b = 0; b < 2; ++b) {
                    CU_TRY(cudaMalloc(&c.d_t1[b], cap * sizeof(K1Tile)));
                    CU_TRY(cudaMallocHost(&c.h_t1[b], cap * sizeof(K1Tile)));
                }
                c.t1_cap = cap;
            } else {
                // the previous user of this staging buffer (possibly an earlier, still running call)
                CU_TRY(cudaEventSynchronize(c.t1_free[buf]));
            }
            memcpy(c.h_t1[buf], tiles.data(), tiles.size() * sizeof(K1Tile));
            CU_TRY(cudaMemcpyAsync(c.d_t1[buf], c.h_t1[buf], tiles.size() * sizeof(K1Tile), cudaMemcpyHostToDevice, sc));

            // ---- host-pointer API: wait until the poles of this chunk have been uploaded ----
            if (gate && need_rank >= 0) if (int r = gate_wait(gate, need_rank, sc)) return r;

            // ---- K1: form the integral blocks of this chunk ----
            double* const Wa = c.W[buf][0];
            double* const Wb = c.W[buf][1];
            double* const Ebuf = c.E[buf].p;
            record(c, sc, 0, true);
            if (!opt.simple) {
                k1_form<<<(unsigned)tiles.size(), kThreads, kK1SmemBytes, sc>>>(
                    mapA, mapA2 ? *mapA2 : mapA, mapB0, mapB1, c.d_t1[buf], same.ea_dev, other ? other->ea_dev : same.ea_dev,
                    Wa, Wb, Ebuf, ldw, nk1, opt.w_evict_last ? kL2EvictLast : 0ull);
            } else {
                k1_form_simple<<<(unsigned)tiles.size(), 256, 0, sc>>>(
                    ixq, ixq2 ? ixq2 : ixq, ld_ixq, (int)nphys, (int)naux, same.qja, same.ld, (int)o, (int)v,
                    other ? other->qja : same.qja, other ? other->ld : same.ld, other ? (int)other->nocc : (int)o,
                    other ? (int)other->nvir : (int)v, c.d_t1[buf], same.ea_dev, other ? other->ea_dev : same.ea_dev,
                    Wa, Wb, Ebuf, ldw);
            }
            record(c, sc, 0, false);
            CU_TRY(cudaGetLastError());
            CU_TRY(cudaEventRecord(c.t1_free[buf], sc));
            count_launch(c);
            c.last_plan[5]++;
            {   // executed DMMA flops of the launch (padding rows / columns included)
                double units = 0.0;
                for (const K1Tile& t : tiles)
                    units += (t.mbv > 10 ? 16.0 * ((t.mbv + 1) / 2) : (t.mbv > 0 ? 8.0 * t.mbv : (double)kTileM)) * 8.0 * (t.nb1 + t.nb2);
                c.flops[0] += 2.0 * units * (double)kKB * nk1;
            }

            // ---- K2: accumulate the moments of this chunk ----
            double* const acc_set = c.acc.p + (size_t)4 * set * acc_n;
            double* const a_vv = acc_set + (sym ? 0 : 2) * acc_n;
            double* const a_vev = a_vv + acc_n;
            const int nk2 = (int)((col + kKB - 1) / kKB);
            record(c, sc, 1, true);
            if (!opt.simple) {
                K2Launch L;
                cuuint64_t d[3] = {(cuuint64_t)col, (cuuint64_t)p_pad, 1};
                cuuint64_t s[2] = {(cuuint64_t)ldw * 8, (cuuint64_t)ldw * 8 * (cuuint64_t)p_pad};
                cuuint32_t b[3] = {kKB, 64, 1};
                if (int r = make_map(c, &L.a0, Wa, 3, d, s, b)) return r;
                if (int r = make_map(c, &L.b, sym ? Wa : Wb, 3, d, s, b)) return r;
                L.a1 = L.a0;
                L.mode = 0;
                L.total_cost = sym ? cost_sym.back() : cost_all.back();
                L.total_work = sym ? work_sym : work_all;
                K2Params p = k2_params();
                p.w1 = Ebuf; p.tiles = sym ? d_t2_sym : d_t2_all; p.cost_prefix = sym ? d_cost_sym : d_cost_all;
                p.acc_vv = a_vv; p.acc_vev = a_vev; p.ldacc = p_pad;
                p.ntiles = (int)(sym ? t2_sym.size() : t2_all.size()); p.nk_inner = nk2; p.n_outer = 1;
                p.rows_valid = (int)nphys; p.cols_valid = (int)nphys; p.sym = sym ? 1 : 0;
                L.p = p;
                if (int r = launch_k2(c, L, set, sc)) return r;
            } else {
                const size_t nt = sym ? t2_sym.size() : t2_all.size();
                k2_moment_simple<<<(unsigned)nt, 256, 0, sc>>>(Wa, sym ? Wa : Wb, ldw, Ebuf, (int)col,
                                                               sym ? d_t2_sym : d_t2_all, a_vv, a_vev, p_pad);
                CU_TRY(cudaGetLastError());
                count_launch(c);
            }
            record(c, sc, 1, false);
            ++chunk_idx;
        }
        return 0;
    };

    int rc = 0;
    if (!sh) {
        rc = run_chunks(sym_items, true);
        if (!rc) rc = run_chunks(asym_items, false);
    } else {
        // ---- pole-sharded pair loop: own block first, then the visiting sub-blocks of the other ranks' blocks,
        // each received on the copy stream (NCCL send/recv over NVLink) while the previous one is being worked on
        const int64_t R = sh->nranks, rk = sh->rank;
        const int64_t pole_elems = nphys * ld_ixq;
        int64_t nb_max = 0;
        for (int64_t r = 0; r < R; ++r) nb_max = std::max(nb_max, shard_lo(o, r + 1, R) - shard_lo(o, r, R));
        const int64_t sb = shard_sub_block(*c.opt, nb_max, pole_elems);
        if (R > 1) {
            // every rank must walk the same schedule: a different sub-block size or sample limit on one rank would
            // leave unmatched sends / receives behind (a hang, not an error) -- compare them first
            double chk[6] = {(double)sb, -(double)sb, (double)c.opt->shard_limit, -(double)c.opt->shard_limit, (double)o, -(double)o};
            if (int r = grow(c.packed, 8, false)) return r;
            CU_TRY(cudaMemcpyAsync(c.packed.p, chk, sizeof(chk), cudaMemcpyHostToDevice, st));
            CU_TRY(cudaStreamSynchronize(st));          // (chk is pageable stack memory)
            if (int r = comm_allreduce(c, c.packed.p, 6, kNcclMaxOp, st)) return r;
            CU_TRY(cudaMemcpyAsync(chk, c.packed.p, sizeof(chk), cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
            if (chk[0] != -chk[1] || chk[2] != -chk[3] || chk[4] != -chk[5])
                return fail(AGF2_B200_EINVAL, "pole-sharded build: the ranks disagree on nocc, \"visiting_bytes\" or "
                                              "\"shard_step_limit\"");
            c.last_comm_bytes[1] = 0;                   // (not part of the build's traffic)
        }
        const std::vector<ShardStep> steps = shard_steps(o, rk, R, sb, c.opt->shard_limit);
        CUtensorMap mapVis[2];
        if (R > 1) {
            for (int b = 0; b < 2; ++b) {
                if (int r = grow(c.vis[b], (size_t)(sb * pole_elems), false)) return r;
                cuuint64_t d[3] = {(cuuint64_t)naux, (cuuint64_t)nphys, (cuuint64_t)sb};
                cuuint64_t s[2] = {(cuuint64_t)ld_ixq * 8, (cuuint64_t)ld_ixq * 8 * (cuuint64_t)nphys};
                cuuint32_t bx[3] = {kKB, kTileM, 1};
                if (int r = make_map(c, &mapVis[b], c.vis[b].p, 3, d, s, bx)) return r;
            }
        }
        // the resident block must be complete before a peer reads it: order the copy stream behind the caller's
        CU_TRY(cudaEventRecord(c.ev_copy, st));
        CU_TRY(cudaStreamWaitEvent(c.copy_stream, c.ev_copy, 0));
        std::vector<size_t> xfer;          // indices of the steps that move data, in order
        for (size_t q = 0; q < steps.size(); ++q) if (steps[q].recv || steps[q].send) xfer.push_back(q);
        bool used_buf[2] = {false, false};
        auto enqueue_transfer = [&](size_t x) -> int {
            const ShardStep& s = steps[xfer[x]];
            const int b = (int)(x & 1);
            if (used_buf[b])       // the previous tenant of this buffer has been consumed (both compute streams)
                for (int q = 0; q < nsets; ++q) CU_TRY(cudaStreamWaitEvent(c.copy_stream, c.vis_done[b][q], 0));
            if (int r = comm_sendrecv(c, s.send ? ixq + (s.send_lo - shard_lo(o, rk, R)) * pole_elems : nullptr,
                                      s.send ? (s.send_hi - s.send_lo) * pole_elems : 0, s.dst,
                                      s.recv ? c.vis[b].p : nullptr, s.recv ? (s.vis_hi - s.vis_lo) * pole_elems : 0, s.src,
                                      c.copy_stream))
                return r;
            CU_TRY(cudaEventRecord(c.vis_ready[b], c.copy_stream));
            return 0;
        };
        if (!xfer.empty()) rc = enqueue_transfer(0);
        size_t x = 0;                      // next transfer step to be consumed
        for (size_t q = 0; q < steps.size() && !rc; ++q) {
            const ShardStep& s = steps[q];
            const bool moves = s.recv || s.send;
            int b = -1;
            if (moves) {
                b = (int)(x & 1);
                if (x + 1 < xfer.size()) rc = enqueue_transfer(x + 1);
                ++x;
                if (rc) break;
            }
            if (s.vis_hi > s.vis_lo) {
                plan.set_shard_step(rk, R, s);
                if (s.recv) {
                    CU_TRY(cudaStreamWaitEvent(st, c.vis_ready[b], 0));
                    if (overlap) CU_TRY(cudaStreamWaitEvent(c.stream2, c.vis_ready[b], 0));
                    mapA2 = &mapVis[b];
                    ixq2 = c.vis[b].p;
                }
                c.last_plan[1] += (int64_t)plan.sym_items.size();
                rc = run_chunks(plan.sym_items, true);
            }
            if (moves && s.recv) {
                CU_TRY(cudaEventRecord(c.vis_done[b][0], st));
                if (overlap) CU_TRY(cudaEventRecord(c.vis_done[b][1], c.stream2));
                used_buf[b] = true;
            }
        }
    }
    if (!rc && fact) {
        // host-pointer API: every pole of the slice must have landed
        if (gate) rc = gate_wait(gate, gate->with_qja ? (int)o - 1 : (int)(iend - istart) - 1, st);
        // weights of the Coulomb-type terms: same-spin tensor (c - s) [+ s for j outside the slice], other spin os
        if (!rc) {
            const SymTiles symt{d_t2_sym, d_cost_sym, (int)t2_sym.size(), cost_sym.back(), work_sym};
            if (sh)
                rc = run_coulomb(c, ixq, ld_ixq, n_res, same, nullptr, nphys, naux, 0, n_res, c_same - s_same, c_same, 0.0,
                                 sh->rank, sh->nranks, true, shard_lo(o, sh->rank, sh->nranks), c.acc.p, c.acc.p + acc_n,
                                 p_pad, symt, d_t2_m, d_cost_m, (int)t2_m.size(), cost_m.back(), work_m, st);
            else
                rc = run_coulomb(c, ixq, ld_ixq, o, same, other, nphys, naux, istart, iend, c_same - s_same, c_same,
                                 os_other, part, nparts, false, 0, c.acc.p, c.acc.p + acc_n, p_pad, symt, d_t2_m, d_cost_m,
                                 (int)t2_m.size(), cost_m.back(), work_m, st);
        }
    }
    // join the second chain (also on error paths: nothing may be left in flight behind the caller's back)
    if (used2) {
        cudaEventRecord(c.ev_join, c.stream2);
        cudaStreamWaitEvent(st, c.ev_join, 0);
    }
    if (rc) return rc;

    dim3 g3((unsigned)((nphys + 127) / 128), (unsigned)nphys);
    if (allreduce && c.nranks > 1) {
        // one NCCL all-reduce of the packed [vv | vev] of this call on the compute stream, then += into the caller's
        const size_t n2 = (size_t)nphys * nphys;
        if (int r = grow(c.packed, 2 * n2, false)) return r;
        CU_TRY(cudaMemsetAsync(c.packed.p, 0, 2 * n2 * 8, st));
        k3_finalize<<<g3, 128, 0, st>>>(c.acc.p, (long long)acc_n, nsets, asym_items.empty() ? 0 : 1, p_pad, (int)nphys,
                                        c.packed.p, c.packed.p + n2);
        CU_TRY(cudaGetLastError());
        if (int r = comm_allreduce_sum(c, c.packed.p, (int64_t)(2 * n2), st)) return r;
        add_packed<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(c.packed.p, (long long)n2, vv, vev);
        CU_TRY(cudaGetLastError());
        count_launch(c, 2);
    } else {
        k3_finalize<<<g3, 128, 0, st>>>(c.acc.p, (long long)acc_n, nsets, asym_items.empty() ? 0 : 1, p_pad, (int)nphys,
                                        vv, vev);
        CU_TRY(cudaGetLastError());
        count_launch(c);
    }
    return 0;
}

int finish_profile(Ctx& c, cudaStream_t st) {
    CU_TRY(cudaStreamSynchronize(st));
    for (int k = 0; k < 4; ++k) c.last_ms[k] = 0;
    for (size_t k = 0; k < c.events_used; ++k) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c.events[k].a, c.events[k].b);
        c.last_ms[c.events[k].kind] += ms;
    }
    c.events_used = 0;
    return 0;
}

}  // namespace

#include "comm.inl"
#include "host_api.inl"
