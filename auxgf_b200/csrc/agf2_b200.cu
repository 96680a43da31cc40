// libagf2_b200.so -- host side of the B200 moment build: work planner, TMA descriptors, launches,
// and the C-ABI of include/agf2_b200.h.  Reference paths are relative to obackhouse/auxgf.
//
// The work planner lives in planner.h (pure host code); this file executes its chunks.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/agf2_b200.h"
#include "moment_kernels.cuh"
#include "planner.h"

using namespace agf2;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CU_TRY(expr)                                                                          \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(_e == cudaErrorMemoryAllocation ? AGF2_B200_ENOMEM : AGF2_B200_ECUDA, \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                  \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

struct EventPair { cudaEvent_t a, b; int kind; };

struct Ctx {
    bool init = false;
    bool env_applied = false;
    int device = -1;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    EncodeTiledFn encode = nullptr;
    int64_t ws_bytes = 40ll << 20;
    bool simple = false;
    bool profile = true;
    // scratch
    double* W[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [ring buffer][W0 | W1]
    size_t w_elems = 0;
    double* E[2] = {nullptr, nullptr};
    size_t e_elems[2] = {0, 0};
    bool lpt_order = true;             // k1_form tiles of a chunk launched longest first
    bool balanced_blocks = false;      // column blocks of an item of nearly equal width (measured: no better than 64 + rest with LPT)
    bool tune_chunk = true;            // chunk size chosen for whole waves of k1_form tiles
    bool overlap = true;               // odd chunks run on a second stream: one chain's tail is filled by the other
    cudaStream_t stream2 = nullptr;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> pole_ev;
    cudaEvent_t ev_k1[2] = {nullptr, nullptr}, ev_k2[2] = {nullptr, nullptr}, ev_start = nullptr;
    double* acc = nullptr;   // 4 accumulators: S_vv, S_vev, A_vv, A_vev
    size_t acc_elems = 0;
    K1Tile* d_t1[2] = {nullptr, nullptr};
    K1Tile* h_t1[2] = {nullptr, nullptr};
    size_t t1_cap = 0;
    cudaEvent_t t1_free[2] = {nullptr, nullptr};
    K2Tile* d_t2 = nullptr;
    size_t t2_cap = 0;
    int* d_cost = nullptr;             // stream-K cost prefixes of the K2 tile lists
    size_t cost_cap = 0;
    // Coulomb factorisation scratch: M = [M0 | M1], T = ixq_batch . M, per-column weights
    bool thin_tiles = true;            // last x-block with <= 80 valid rows uses the thin-tile mapping of k1_form
    bool coulomb = true;               // Coulomb factorisation allowed ...
    bool coulomb_auto = true;          // ... and chosen per call by a flop estimate (false: always when allowed)
    int k2_mask = 1;                   // edge tiles of k2_moment skip the column blocks they do not need
    cublasHandle_t blas = nullptr;
    double* Mbuf = nullptr; size_t m_elems = 0;
    double* Tbuf[2] = {nullptr, nullptr}; size_t t_elems[2] = {0, 0};
    double* wq[2] = {nullptr, nullptr}; size_t wq_elems[2] = {0, 0};     // w0 | w1 of the M build
    double* we[2] = {nullptr, nullptr}; size_t we_elems[2] = {0, 0};     // e_i per (i, Q) column of a T batch
    K2Tile* d_tm = nullptr; size_t tm_cap = 0;                           // tile list of the M build
    int* d_cm = nullptr; size_t cm_cap = 0;

    // staging for the host-pointer API
    double* stage[8] = {nullptr};
    size_t stage_elems[8] = {0};
    // instrumentation
    int64_t launches = 0;
    int64_t chunk_counter = 0;
    std::vector<EventPair> events;
    size_t events_used = 0;
    int64_t last_plan[6] = {0, 0, 0, 0, 0, 0};   // factorised, PAIR, DIAG, OS, ASYM items, k1_form launches
    double last_ms[4] = {0, 0, 0, 0};   // k1_form, k2_moment (pairs), M build, Coulomb step (DGEMM + k2)
};
Ctx g_ctx;

// Environment overrides of the planner / scheduling knobs (read once; no CUDA needed).
void apply_env() {
    Ctx& c = g_ctx;
    if (c.env_applied) return;
    c.env_applied = true;
    if (const char* e = getenv("AGF2_B200_OVERLAP")) c.overlap = atoi(e) != 0;
    if (const char* e = getenv("AGF2_B200_WORKSPACE_MB")) c.ws_bytes = (int64_t)atoll(e) << 20;
    if (const char* e = getenv("AGF2_B200_SIMPLE")) c.simple = atoi(e) != 0;
    if (const char* e = getenv("AGF2_B200_LPT")) c.lpt_order = atoi(e) != 0;
    if (const char* e = getenv("AGF2_B200_BALANCED")) c.balanced_blocks = atoi(e) != 0;
    if (const char* e = getenv("AGF2_B200_TUNE_CHUNK")) c.tune_chunk = atoi(e) != 0;
    if (const char* e = getenv("AGF2_B200_COULOMB")) { c.coulomb = atoi(e) != 0; c.coulomb_auto = atoi(e) == 2; }
    if (const char* e = getenv("AGF2_B200_K2_MASK")) c.k2_mask = atoi(e) != 0;
    if (const char* e = getenv("AGF2_B200_THIN")) c.thin_tiles = atoi(e) != 0;
}

int ensure_ctx() {
    Ctx& c = g_ctx;
    if (c.init) return 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(AGF2_B200_ENODEV, "no CUDA device visible (libagf2_b200 has no CPU fallback)");
    if (c.device < 0) CU_TRY(cudaGetDevice(&c.device));
    CU_TRY(cudaSetDevice(c.device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, c.device));
    if (prop.major != 10)
        return fail(AGF2_B200_ENODEV, std::string("device is sm_") + std::to_string(prop.major) +
                                          std::to_string(prop.minor) + ", this library is built for sm_100a only");
    c.num_sms = prop.multiProcessorCount;
    CU_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
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
    for (int b = 0; b < 2; ++b) {
        CU_TRY(cudaEventCreateWithFlags(&c.t1_free[b], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c.ev_k1[b], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c.ev_k2[b], cudaEventDisableTiming));
    }
    CU_TRY(cudaEventCreateWithFlags(&c.ev_start, cudaEventDisableTiming));
    CU_TRY(cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    apply_env();
    if (cublasCreate(&c.blas) != CUBLAS_STATUS_SUCCESS) return fail(AGF2_B200_ESOLVER, "cublasCreate failed");
    c.init = true;
    return 0;
}

template <typename T>
int grow(T*& p, size_t& cap, size_t need, bool zero) {
    if (need <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    CU_TRY(cudaMalloc(&p, need * sizeof(T)));
    if (zero) CU_TRY(cudaMemset(p, 0, need * sizeof(T)));
    cap = need;
    return 0;
}

int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
             const cuuint32_t* box) {
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = g_ctx.encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, const_cast<void*>(base), dims,
                              strides_bytes, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(AGF2_B200_ECUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 0;
}

struct Spin {            // one (Q|ja) tensor and its energies
    const double* qja = nullptr;
    int64_t ld = 0, nocc = 0, nvir = 0;
    const double* ei = nullptr;   // host copies of the energies (planner needs e_i + e_j)
    const double* ea_dev = nullptr;
    const double* ei_dev = nullptr;
};

// Host-pointer API only: ixq is uploaded pole by pole on a copy stream by a helper thread while the GPU
// already works on the pairs whose poles have landed.  rank[i] = position of pole i in the upload order,
// ev[k] is recorded after the k-th upload, ready = number of uploads whose event has been recorded.
struct PoleGate {
    const std::vector<int>* rank;
    const std::vector<cudaEvent_t>* ev;
    std::atomic<int>* ready;
    std::atomic<int>* failed;
};

void record(cudaStream_t st, int kind, bool begin) {
    Ctx& c = g_ctx;
    if (!c.profile) return;
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

struct K2Launch {
    CUtensorMap a0, a1, b;
    K2Params p;
    int mode;          // 0: moments (W W^T), 1: M build (two weight vectors), 2: Coulomb step (two A operands)
    int total_cost;    // cost_prefix[ntiles]
};

int launch_k2(const K2Launch& L, cudaStream_t st) {
    Ctx& c = g_ctx;
    // stream-K grid: one CTA per SM, unless the launch is too small to give each >= 8 iterations
    const long long total_it = (long long)L.total_cost * L.p.nk_inner * L.p.n_outer / 16;
    const int grid = (int)std::max<long long>(1, std::min<long long>(c.num_sms, total_it / 8));
    if (L.mode == 0) k2_moment<false, false><<<grid, kThreads, kK2SmemBytes, st>>>(L.a0, L.a1, L.b, L.p);
    else if (L.mode == 1) k2_moment<false, true><<<grid, kThreads, kK2SmemBytes, st>>>(L.a0, L.a1, L.b, L.p);
    else k2_moment<true, false><<<grid, kThreads, kK2SmemBytesDual, st>>>(L.a0, L.a1, L.b, L.p);
    CU_TRY(cudaGetLastError());
    c.launches++;
    return 0;
}

PlanConfig plan_config() {
    const Ctx& c = g_ctx;
    PlanConfig p;
    p.ws_bytes = c.ws_bytes; p.num_sms = c.num_sms; p.simple = c.simple; p.lpt_order = c.lpt_order;
    p.balanced_blocks = c.balanced_blocks; p.tune_chunk = c.tune_chunk; p.thin_tiles = c.thin_tiles;
    p.coulomb = c.coulomb; p.coulomb_auto = c.coulomb_auto;
    return p;
}

// Coulomb part of the moments without forming a single Coulomb integral:
//     sum_{i in slice} sum_{j,a} w_j X_ij X_ij^T (e_i + e_j - e_a)^n  =  sum_i ixq_i (e_i^n M0 + n M1) ixq_i^T ,
//     M0 = sum_ja w_j q_ja q_ja^T ,  M1 = sum_ja w_j (e_j - e_a) q_ja q_ja^T      (naux x naux)
// (k2_moment<false,true> on the (Q|ja) tensor), T = ixq_batch [M0 | M1] (cuBLAS DGEMM) and
// k2_moment<true,false>: vv += T0 ixq^T, vev += (e_i T0 + T1) ixq^T, upper triangle only.  With nparts GPUs each
// one owns a block of columns Q' of M (and the matching K range of the last step).
int run_coulomb(const double* ixq, int64_t ld_ixq, const Spin& same, const Spin* other, int64_t nphys, int64_t naux,
                int64_t istart, int64_t iend, double w_same_in, double w_same_out, double w_other, int64_t part,
                int64_t nparts, double* S_vv, double* S_vev, int64_t p_pad, const K2Tile* d_t2_sym,
                const int* d_cost_sym, int nt_sym, int cost_sym, cudaStream_t st)
{
    Ctx& c = g_ctx;
    struct Pass { const Spin* sp; double w_in, w_out; };
    std::vector<Pass> passes;
    const bool partial = istart > 0 || iend < same.nocc;
    if (w_same_in != 0.0 || (partial && w_same_out != 0.0)) passes.push_back({&same, w_same_in, w_same_out});
    if (other && w_other != 0.0) passes.push_back({other, w_other, w_other});
    if (passes.empty() || iend <= istart) return 0;
    const int64_t ntq = (naux + kTileN - 1) / kTileN;
    const int64_t q_lo = ntq * part / nparts * kTileN;
    const int64_t q_hi = std::min<int64_t>(naux, ntq * (part + 1) / nparts * kTileN);
    if (q_hi <= q_lo) return 0;
    const int64_t nq = q_hi - q_lo, nq_pad = rup(nq, kTileN);
    const int64_t n_pad = rup(naux, kTileM);
    const int64_t ldm = 2 * nq_pad;

    // ---- M build ------------------------------------------------------------------------------------
    { int r = grow(c.Mbuf, c.m_elems, (size_t)n_pad * ldm, false); if (r) return r; }
    CU_TRY(cudaMemsetAsync(c.Mbuf, 0, (size_t)n_pad * ldm * 8, st));
    const int nxm = (int)(n_pad / kTileM), nym = (int)(nq_pad / kTileN);
    const int msym = (nparts == 1) ? 1 : 0;   // the whole (symmetric) M: upper-triangle tiles, mirrored afterwards
    std::vector<K2Tile> tm;
    std::vector<int> cm(1, 0);
    for (int xb = 0; xb < nxm; ++xb)
        for (int yb = 0; yb < nym; ++yb) {
            if (msym && kTileN * yb + kTileN - 1 < kTileM * xb) continue;
            tm.push_back({xb, yb});
            cm.push_back(cm.back() + k2_tile_cost(xb, yb, (int)naux, (int)nq, msym, c.k2_mask));
        }
    { int r = grow(c.d_tm, c.tm_cap, tm.size(), false); if (r) return r; }
    { int r = grow(c.d_cm, c.cm_cap, cm.size(), false); if (r) return r; }
    CU_TRY(cudaMemcpyAsync(c.d_tm, tm.data(), tm.size() * sizeof(K2Tile), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(c.d_cm, cm.data(), cm.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaStreamSynchronize(st));   // tm / cm are pageable host memory
    record(st, 2, true);
    for (const Pass& ps : passes) {
        const Spin& sp = *ps.sp;
        const int nki = (int)((sp.nvir + kKB - 1) / kKB);
        const int64_t kpad = (int64_t)nki * kKB;
        for (int b = 0; b < 2; ++b) { int r = grow(c.wq[b], c.wq_elems[b], (size_t)(sp.nocc * kpad), false); if (r) return r; }
        dim3 gw((unsigned)((kpad + 127) / 128), (unsigned)sp.nocc);
        const bool is_same = ps.sp == &same;
        fill_m_weights<<<gw, 128, 0, st>>>(c.wq[0], c.wq[1], sp.ei_dev, sp.ea_dev, (int)sp.nvir, (int)kpad, ps.w_in,
                                          ps.w_out, is_same ? (int)istart : 0, is_same ? (int)iend : (int)sp.nocc);
        CU_TRY(cudaGetLastError());
        c.launches++;
        K2Launch L;
        cuuint64_t d[3] = {(cuuint64_t)sp.nvir, (cuuint64_t)sp.nocc, (cuuint64_t)naux};
        cuuint64_t sb[2] = {(cuuint64_t)sp.ld * 8, (cuuint64_t)sp.ld * 8 * (cuuint64_t)sp.nocc};
        cuuint32_t bx[3] = {kKB, 1, 64};
        if (int r = make_map(&L.a0, sp.qja, 3, d, sb, bx)) return r;
        L.a1 = L.a0; L.b = L.a0;
        L.mode = 1;
        L.total_cost = cm.back();
        L.p = K2Params{c.wq[0], c.wq[1], c.d_tm, c.d_cm, c.Mbuf, c.Mbuf + nq_pad, ldm, (int)tm.size(), nki,
                       (int)sp.nocc, 1, 1, 0, 0, (int)q_lo, (int)naux, (int)nq, msym, c.k2_mask};
        if (int r = launch_k2(L, st)) return r;
    }
    if (msym) {
        dim3 gm((unsigned)naux, (unsigned)((naux + 127) / 128));
        mirror_m<<<gm, 128, 0, st>>>(c.Mbuf, (int)naux, nq_pad, ldm);
        CU_TRY(cudaGetLastError());
        c.launches++;
    }
    record(st, 2, false);

    // ---- T = ixq_batch [M0 | M1], then the two moments ------------------------------------------------------
    const int64_t per_pole = nphys * ldm;
    const int64_t nbi_max = std::max<int64_t>(1, std::min<int64_t>(64, (512ll << 20) / (per_pole * 8)));
    const int nkq = (int)((nq + kKB - 1) / kKB);
    const int64_t kpad = (int64_t)nkq * kKB;
    { int r = grow(c.Tbuf[0], c.t_elems[0], (size_t)(std::min<int64_t>(nbi_max, iend - istart) * per_pole), false); if (r) return r; }
    { int r = grow(c.we[0], c.we_elems[0], (size_t)(nbi_max * kpad), false); if (r) return r; }
    if (cublasSetStream(c.blas, st) != CUBLAS_STATUS_SUCCESS) return fail(AGF2_B200_ESOLVER, "cublasSetStream failed");
    CUtensorMap mapIx;
    {
        cuuint64_t d[3] = {(cuuint64_t)naux, (cuuint64_t)nphys, (cuuint64_t)same.nocc};
        cuuint64_t sb[2] = {(cuuint64_t)ld_ixq * 8, (cuuint64_t)ld_ixq * 8 * (cuuint64_t)nphys};
        cuuint32_t bx[3] = {kKB, 64, 1};
        if (int r = make_map(&mapIx, ixq, 3, d, sb, bx)) return r;
    }
    record(st, 3, true);
    for (int64_t i0 = istart; i0 < iend; i0 += nbi_max) {
        const int64_t nbi = std::min<int64_t>(nbi_max, iend - i0);
        const double one = 1.0, zero = 0.0;
        // row-major T (nbi*P x ldm) = IXQ (nbi*P x naux) . M (naux x ldm)   ==   column-major T^T = M^T IXQ^T
        cublasStatus_t bs = cublasDgemm(c.blas, CUBLAS_OP_N, CUBLAS_OP_N, (int)ldm, (int)(nbi * nphys), (int)naux, &one,
                                        c.Mbuf, (int)ldm, ixq + i0 * nphys * ld_ixq, (int)ld_ixq, &zero, c.Tbuf[0],
                                        (int)ldm);
        if (bs != CUBLAS_STATUS_SUCCESS) return fail(AGF2_B200_ESOLVER, "cublasDgemm (Coulomb T) failed with status " + std::to_string((int)bs));
        c.launches++;
        dim3 ge((unsigned)((kpad + 127) / 128), (unsigned)nbi);
        fill_e_weights<<<ge, 128, 0, st>>>(c.we[0], same.ei_dev, (int)i0, (int)kpad);
        CU_TRY(cudaGetLastError());
        c.launches++;
        K2Launch L;
        cuuint64_t d[3] = {(cuuint64_t)nq, (cuuint64_t)nphys, (cuuint64_t)nbi};
        cuuint64_t sb[2] = {(cuuint64_t)ldm * 8, (cuuint64_t)ldm * 8 * (cuuint64_t)nphys};
        cuuint32_t bx[3] = {kKB, 64, 1};
        if (int r = make_map(&L.a0, c.Tbuf[0], 3, d, sb, bx)) return r;
        if (int r = make_map(&L.a1, c.Tbuf[0] + nq_pad, 3, d, sb, bx)) return r;
        L.b = mapIx;
        L.mode = 2;
        L.total_cost = cost_sym;
        L.p = K2Params{nullptr, c.we[0], d_t2_sym, d_cost_sym, S_vv, S_vev, p_pad, nt_sym, nkq, (int)nbi, 0, 0,
                       (int)q_lo, (int)i0, 0, (int)nphys, (int)nphys, 1, c.k2_mask};
        if (int r = launch_k2(L, st)) return r;
    }
    record(st, 3, false);
    return 0;
}

// The whole moment build on device-resident inputs.
int run_moments(const double* ixq, int64_t ld_ixq, const Spin& same, const Spin* other,
                const double* ei_same_host, const double* ei_other_host, int64_t nphys, int64_t naux,
                int64_t istart, int64_t iend, double c_same, double s_same, double os_other, int64_t part,
                int64_t nparts, double* vv, double* vev, cudaStream_t st, const PoleGate* gate = nullptr)
{
    Ctx& c = g_ctx;
    const int64_t o = same.nocc, v = same.nvir;
    MomentPlan plan;
    if (int r = plan.init(plan_config(), nphys, naux, o, v, other != nullptr, other ? other->nocc : 0,
                          other ? other->nvir : 0, istart, iend, c_same, s_same, os_other, part, nparts))
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
    for (int r = 0; r < 2; ++r) { int rc = grow(c.E[r], c.e_elems[r], (size_t)cap_cols + 64, true); if (rc) return rc; }
    { int r = grow(c.acc, c.acc_elems, (size_t)4 * p_pad * p_pad, false); if (r) return r; }
    const size_t acc_n = (size_t)p_pad * p_pad;
    double* S_vv = c.acc; double* S_vev = c.acc + acc_n; double* A_vv = c.acc + 2 * acc_n; double* A_vev = c.acc + 3 * acc_n;
    CU_TRY(cudaMemsetAsync(c.acc, 0, 4 * acc_n * 8, st));
    const bool overlap = c.overlap && !c.simple;
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
        cuuint64_t d[3] = {(cuuint64_t)naux, (cuuint64_t)nphys, (cuuint64_t)o};
        cuuint64_t s[2] = {(cuuint64_t)ld_ixq * 8, (cuuint64_t)ld_ixq * 8 * (cuuint64_t)nphys};
        cuuint32_t b[3] = {kKB, kTileM, 1};
        if (int r = make_map(&mapA, ixq, 3, d, s, b)) return r;
    }
    {
        cuuint64_t d[3] = {(cuuint64_t)v, (cuuint64_t)o, (cuuint64_t)naux};
        cuuint64_t s[2] = {(cuuint64_t)same.ld * 8, (cuuint64_t)same.ld * 8 * (cuuint64_t)o};
        cuuint32_t b[3] = {16, 1, kKB};
        if (int r = make_map(&mapB0, same.qja, 3, d, s, b)) return r;
    }
    if (other) {
        cuuint64_t d[3] = {(cuuint64_t)other->nvir, (cuuint64_t)other->nocc, (cuuint64_t)naux};
        cuuint64_t s[2] = {(cuuint64_t)other->ld * 8, (cuuint64_t)other->ld * 8 * (cuuint64_t)other->nocc};
        cuuint32_t b[3] = {16, 1, kKB};
        if (int r = make_map(&mapB1, other->qja, 3, d, s, b)) return r;
    } else {
        mapB1 = mapB0;
    }
    const int nk1 = (int)((naux + kKB - 1) / kKB);

    c.last_plan[0] = fact ? 1 : 0;
    c.last_plan[1] = c.last_plan[2] = c.last_plan[3] = 0;
    for (const Item& it : sym_items) c.last_plan[it.kind == 1 ? 1 : (it.kind == 0 ? 2 : 3)]++;
    c.last_plan[4] = (int64_t)asym_items.size();
    c.last_plan[5] = 0;

    // K2 tile lists (upper-triangle for symmetric chunks, all tiles for non-symmetric ones)
    std::vector<K2Tile> t2_sym, t2_all;
    for (int xb = 0; xb < nxb; ++xb)
        for (int yb = 0; yb < nyb; ++yb) {
            t2_all.push_back({xb, yb});
            if (kTileN * yb + kTileN - 1 >= kTileM * xb) t2_sym.push_back({xb, yb});
        }
    // stream-K cost prefixes: edge tiles (diagonal, padded last row / column) skip the DMMAs they do not need
    std::vector<int> cost_sym(1, 0), cost_all(1, 0);
    for (const K2Tile& t : t2_sym) cost_sym.push_back(cost_sym.back() + k2_tile_cost(t.xb, t.yb, (int)nphys, (int)nphys, 1, c.k2_mask));
    for (const K2Tile& t : t2_all) cost_all.push_back(cost_all.back() + k2_tile_cost(t.xb, t.yb, (int)nphys, (int)nphys, 0, c.k2_mask));
    { int r = grow(c.d_t2, c.t2_cap, t2_sym.size() + t2_all.size(), false); if (r) return r; }
    { int r = grow(c.d_cost, c.cost_cap, cost_sym.size() + cost_all.size(), false); if (r) return r; }
    CU_TRY(cudaMemcpyAsync(c.d_t2, t2_sym.data(), t2_sym.size() * sizeof(K2Tile), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(c.d_t2 + t2_sym.size(), t2_all.data(), t2_all.size() * sizeof(K2Tile), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(c.d_cost, cost_sym.data(), cost_sym.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(c.d_cost + cost_sym.size(), cost_all.data(), cost_all.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaStreamSynchronize(st));   // these vectors are pageable host memory

    c.events_used = 0;
    int64_t& chunk_idx = c.chunk_counter;   // global: tile staging buffers alternate across calls too
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
            cudaStream_t sc = (overlap && buf) ? c.stream2 : st;
            if (sc != st) used2 = true;
            if (tiles.size() > c.t1_cap) {
                CU_TRY(cudaStreamSynchronize(st));
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
            if (gate && need_rank >= 0) {
                while (gate->ready->load(std::memory_order_acquire) <= need_rank) {
                    if (gate->failed->load()) return fail(AGF2_B200_ECUDA, "ixq upload failed");
                    std::this_thread::yield();
                }
                CU_TRY(cudaStreamWaitEvent(sc, (*gate->ev)[need_rank], 0));
            }
            // ---- K1: form the integral blocks of this chunk ----
            double* const Wa = c.W[buf][0];
            double* const Wb = c.W[buf][1];
            double* const Ebuf = c.E[buf];
            record(sc, 0, true);
            if (!c.simple) {
                k1_form<<<(unsigned)tiles.size(), kThreads, kK1SmemBytes, sc>>>(
                    mapA, mapB0, mapB1, c.d_t1[buf], same.ea_dev, other ? other->ea_dev : same.ea_dev, Wa, Wb,
                    Ebuf, ldw, nk1);
            } else {
                k1_form_simple<<<(unsigned)tiles.size(), 256, 0, sc>>>(
                    ixq, ld_ixq, (int)nphys, (int)naux, same.qja, same.ld, (int)o, (int)v,
                    other ? other->qja : same.qja, other ? other->ld : same.ld, other ? (int)other->nocc : (int)o,
                    other ? (int)other->nvir : (int)v, c.d_t1[buf], same.ea_dev, other ? other->ea_dev : same.ea_dev,
                    Wa, Wb, Ebuf, ldw);
            }
            record(sc, 0, false);
            CU_TRY(cudaGetLastError());
            CU_TRY(cudaEventRecord(c.t1_free[buf], sc));
            c.launches++;
            c.last_plan[5]++;

            // ---- K2: accumulate the moments of this chunk ----
            const std::vector<K2Tile>& t2 = sym ? t2_sym : t2_all;
            const K2Tile* d_t2 = sym ? c.d_t2 : c.d_t2 + t2_sym.size();
            const int nk2 = (int)((col + kKB - 1) / kKB);
            record(sc, 1, true);
            if (!c.simple) {
                K2Launch L;
                cuuint64_t d[3] = {(cuuint64_t)col, (cuuint64_t)p_pad, 1};
                cuuint64_t s[2] = {(cuuint64_t)ldw * 8, (cuuint64_t)ldw * 8 * (cuuint64_t)p_pad};
                cuuint32_t b[3] = {kKB, 64, 1};
                if (int r = make_map(&L.a0, Wa, 3, d, s, b)) return r;
                if (int r = make_map(&L.b, sym ? Wa : Wb, 3, d, s, b)) return r;
                L.a1 = L.a0;
                L.mode = 0;
                L.total_cost = sym ? cost_sym.back() : cost_all.back();
                L.p = K2Params{nullptr, Ebuf, d_t2, sym ? c.d_cost : c.d_cost + cost_sym.size(), sym ? S_vv : A_vv,
                               sym ? S_vev : A_vev, p_pad, (int)t2.size(), nk2, 1, 0, 0, 0, 0, 0, (int)nphys, (int)nphys,
                               sym ? 1 : 0, c.k2_mask};
                if (int r = launch_k2(L, sc)) return r;
            } else {
                k2_moment_simple<<<(unsigned)t2.size(), 256, 0, sc>>>(Wa, sym ? Wa : Wb, ldw, Ebuf, (int)col,
                                                                      d_t2, sym ? S_vv : A_vv, sym ? S_vev : A_vev, p_pad);
                CU_TRY(cudaGetLastError());
                c.launches++;
            }
            record(sc, 1, false);
            ++chunk_idx;
        }
        return 0;
    };

    if (int r = run_chunks(sym_items, true)) return r;
    if (int r = run_chunks(asym_items, false)) return r;
    if (fact) {
        if (gate) {   // host-pointer API: every pole of the slice must have landed
            const int need = (int)(iend - istart) - 1;
            while (gate->ready->load(std::memory_order_acquire) <= need) {
                if (gate->failed->load()) return fail(AGF2_B200_ECUDA, "ixq upload failed");
                std::this_thread::yield();
            }
            CU_TRY(cudaStreamWaitEvent(st, (*gate->ev)[need], 0));
        }
        // weights of the Coulomb-type terms: same-spin tensor (c - s) [+ s for j outside the slice], other spin os
        if (int r = run_coulomb(ixq, ld_ixq, same, other, nphys, naux, istart, iend, c_same - s_same, c_same, os_other,
                                part, nparts, S_vv, S_vev, p_pad, c.d_t2, c.d_cost, (int)t2_sym.size(), cost_sym.back(), st))
            return r;
    }

    if (used2) {
        CU_TRY(cudaEventRecord(c.ev_k2[0], c.stream2));
        CU_TRY(cudaStreamWaitEvent(st, c.ev_k2[0], 0));
    }
    dim3 g3((unsigned)((nphys + 127) / 128), (unsigned)nphys);
    k3_finalize<<<g3, 128, 0, st>>>(S_vv, S_vev, asym_items.empty() ? nullptr : A_vv, asym_items.empty() ? nullptr : A_vev,
                                    p_pad, (int)nphys, vv, vev);
    CU_TRY(cudaGetLastError());
    c.launches++;
    return 0;
}

int finish_profile(cudaStream_t st) {
    Ctx& c = g_ctx;
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

// copy a host matrix (rows x cols, contiguous) into a device staging slot with an even pitch
int stage_in(int slot, const double* host, int64_t rows, int64_t cols, int64_t* ld_out, cudaStream_t st) {
    Ctx& c = g_ctx;
    const int64_t ld = rup(cols, 2);
    if (int r = grow(c.stage[slot], c.stage_elems[slot], (size_t)std::max<int64_t>(rows * ld, 2), false)) return r;
    CU_TRY(cudaMemcpy2DAsync(c.stage[slot], ld * 8, host, cols * 8, cols * 8, rows, cudaMemcpyHostToDevice, st));
    *ld_out = ld;
    return 0;
}

// Pipelined upload of ixq (nocc poles of nphys x naux) for the host-pointer API.
struct Uploader {
    std::thread th;
    std::atomic<int> ready{0}, failed{0};
    std::vector<int> rank;
    PoleGate g;

    int start(const double* host, int slot, int64_t nocc, int64_t nphys, int64_t naux, int64_t istart, int64_t iend,
              int64_t* ld_out) {
        Ctx& c = g_ctx;
        const int64_t ld = rup(naux, 2);
        if (int r = grow(c.stage[slot], c.stage_elems[slot], (size_t)std::max<int64_t>(nocc * nphys * ld, 2), false)) return r;
        *ld_out = ld;
        while ((int64_t)c.pole_ev.size() < nocc) {
            cudaEvent_t e;
            CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c.pole_ev.push_back(e);
        }
        std::vector<int> order;
        for (int64_t i = istart; i < iend; ++i) order.push_back((int)i);      // the poles the pair loop needs first
        for (int64_t i = 0; i < istart; ++i) order.push_back((int)i);
        for (int64_t i = iend; i < nocc; ++i) order.push_back((int)i);
        rank.assign((size_t)nocc, 0);
        for (size_t k = 0; k < order.size(); ++k) rank[order[k]] = (int)k;
        g.rank = &rank; g.ev = &c.pole_ev; g.ready = &ready; g.failed = &failed;
        double* dev = c.stage[slot];
        const int device = c.device;
        cudaStream_t cs = c.copy_stream;
        th = std::thread([=]() {
            if (cudaSetDevice(device) != cudaSuccess) { failed.store(1); return; }
            for (size_t k = 0; k < order.size(); ++k) {
                const int64_t i = order[k];
                cudaError_t e = cudaMemcpy2DAsync(dev + i * nphys * ld, ld * 8, host + i * nphys * naux, naux * 8,
                                                  naux * 8, nphys, cudaMemcpyHostToDevice, cs);
                if (e == cudaSuccess) e = cudaEventRecord(g_ctx.pole_ev[k], cs);
                if (e != cudaSuccess) { failed.store(1); return; }
                ready.store((int)k + 1, std::memory_order_release);
            }
        });
        return 0;
    }
    void join() { if (th.joinable()) th.join(); }
    ~Uploader() { join(); }
};

}  // namespace

// =================================================================================================
extern "C" {

const char* agf2_b200_last_error(void) { return g_err.c_str(); }
const char* agf2_b200_version(void) { return "agf2_b200 0.1 (sm_100a; DMMA.8x8x4 + TMA)"; }

int agf2_b200_set_device(int device) {
    if (g_ctx.init && g_ctx.device != device) return fail(AGF2_B200_EINVAL, "context already bound to another device");
    g_ctx.device = device;
    return 0;
}

int agf2_b200_set_workspace_bytes(int64_t bytes) {
    if (bytes < (1 << 20)) return fail(AGF2_B200_EINVAL, "workspace must be at least 1 MiB");
    g_ctx.ws_bytes = bytes;
    return 0;
}

int agf2_b200_set_simple_kernels(int on) {
    g_ctx.simple = on != 0;
    return 0;
}

int agf2_b200_set_overlap(int on) {
    g_ctx.overlap = on != 0;
    return 0;
}

int64_t agf2_b200_launch_count(int reset) {
    const int64_t n = g_ctx.launches;
    if (reset) g_ctx.launches = 0;
    return n;
}

int agf2_b200_last_stage_ms(double* ms4) {
    if (!ms4) return fail(AGF2_B200_EINVAL, "null pointer");
    for (int k = 0; k < 4; ++k) ms4[k] = g_ctx.last_ms[k];
    return 0;
}

int agf2_b200_plan_stats(int64_t nphys, int64_t nocc, int64_t nvir, int64_t naux, int64_t nocc_b, int64_t nvir_b,
                         int64_t istart, int64_t iend, double os_factor, double ss_factor, int uhf, int64_t part,
                         int64_t nparts, int64_t* cover, int64_t* stats8)
{
    if (!stats8) return fail(AGF2_B200_EINVAL, "null pointer");
    apply_env();
    MomentPlan plan;
    const bool has_b = uhf && nocc_b > 0 && nvir_b > 0;
    // same coefficient convention as the moments entry points (RHF: c = os + ss, s = ss; UHF: c = s = ss, other os)
    const double c_same = uhf ? ss_factor : os_factor + ss_factor, os_other = uhf ? os_factor : 0.0;
    if (int r = plan.init(plan_config(), nphys, naux, nocc, nvir, has_b, nocc_b, nvir_b, istart, iend, c_same, ss_factor,
                          os_other, part, nparts))
        return fail(r, plan.error);
    std::vector<double> e_same((size_t)nocc, 0.0), e_other((size_t)std::max<int64_t>(nocc_b, 1), 0.0);
    std::vector<K1Tile> tiles;
    int64_t n_chunks = 0, n_tiles = 0, n_thin = 0, max_cols = 0, max_items = 0;
    const int64_t ncov = nocc + (has_b ? nocc_b : 0);
    for (int pass = 0; pass < 2; ++pass) {
        const std::vector<Item>& items = pass ? plan.asym_items : plan.sym_items;
        size_t pos = 0;
        while (pos < items.size()) {
            const size_t pos0 = pos;
            int need_rank;
            const int64_t col = plan.next_chunk(items, pos, e_same.data(), e_other.data(), nullptr, tiles, need_rank);
            if (col == 0) return fail(AGF2_B200_EINVAL, "workspace too small for one work item");
            ++n_chunks;
            n_tiles += (int64_t)tiles.size();
            max_cols = std::max(max_cols, col);
            max_items = std::max<int64_t>(max_items, (int64_t)(pos - pos0));
            for (const K1Tile& t : tiles) {
                if (t.mbv > 0) ++n_thin;
                const int nbo2 = (t.p21 != 0.0) ? t.nb1 : t.nb2;   // width of out2, as in the k1_form epilogue
                if ((t.ws1 >= 0 && (t.col1 < 0 || t.col1 + 8 * t.nb1 > plan.cap_cols)) ||
                    (t.ws2 >= 0 && (t.col2 < 0 || t.col2 + 8 * nbo2 > plan.cap_cols)))
                    return fail(AGF2_B200_EINVAL, "planner bug: tile outside the workspace");
                if (cover) {   // 8-column blocks of (x-block, i, j) formed: same spin in [0, nocc), other spin after
                    if (t.nb1) cover[(int64_t)t.i1 * ncov + (t.bsel1 ? nocc + t.j1 : t.j1)] += t.nb1;
                    if (t.nb2) cover[(int64_t)t.i2 * ncov + (t.bsel2 ? nocc + t.j2 : t.j2)] += t.nb2;
                }
            }
        }
    }
    stats8[0] = plan.fact ? 1 : 0;
    stats8[1] = n_chunks;
    stats8[2] = n_tiles;
    stats8[3] = n_thin;
    stats8[4] = max_cols;
    stats8[5] = plan.cap_cols;
    stats8[6] = max_items;
    stats8[7] = (int64_t)plan.nxb * (plan.vpad / 8);   // 8-column blocks of one complete (xi|ja) block
    return 0;
}

int agf2_b200_last_plan(int64_t* info6) {
    if (!info6) return fail(AGF2_B200_EINVAL, "null pointer");
    for (int k = 0; k < 6; ++k) info6[k] = g_ctx.last_plan[k];
    return 0;
}

int agf2_b200_set_coulomb_factorization(int mode) {
    if (mode < 0 || mode > 2) return fail(AGF2_B200_EINVAL, "mode must be 0 (off), 1 (always) or 2 (automatic)");
    g_ctx.coulomb = mode != 0;
    g_ctx.coulomb_auto = mode == 2;
    return 0;
}

int agf2_b200_last_kernel_ms(double* k1, double* k2) {
    if (k1) *k1 = g_ctx.last_ms[0];
    if (k2) *k2 = g_ctx.last_ms[1];
    return 0;
}

static int moments_rhf_impl(const double* ixq, int64_t ld_ixq, const double* qja, int64_t ld_qja, const double* ei,
                            const double* ea, int64_t nphys, int64_t nocc, int64_t nvir, int64_t naux,
                            int64_t istart, int64_t iend, double os_factor, double ss_factor, int64_t part,
                            int64_t nparts, double* vv, double* vev, void* stream, int sync, const PoleGate* gate)
{
    if (int r = ensure_ctx()) return r;
    if (!ixq || !qja || !ei || !ea || !vv || !vev) return fail(AGF2_B200_EINVAL, "null pointer");
    if (nocc <= 0) return fail(AGF2_B200_EINVAL, "empty dimension");
    cudaStream_t st = (cudaStream_t)stream;   // used as given: 0 is the CUDA default stream
    std::vector<double> ei_h((size_t)nocc);
    CU_TRY(cudaMemcpyAsync(ei_h.data(), ei, nocc * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    Spin same;
    same.qja = qja; same.ld = ld_qja; same.nocc = nocc; same.nvir = nvir; same.ea_dev = ea; same.ei_dev = ei;
    int r = run_moments(ixq, ld_ixq, same, nullptr, ei_h.data(), nullptr, nphys, naux, istart, iend,
                        os_factor + ss_factor, ss_factor, 0.0, part, nparts, vv, vev, st, gate);
    if (r) return r;
    if (sync) return finish_profile(st);
    return 0;
}

int agf2_b200_moments_rhf_dev(const double* ixq, int64_t ld_ixq, const double* qja, int64_t ld_qja, const double* ei,
                              const double* ea, int64_t nphys, int64_t nocc, int64_t nvir, int64_t naux,
                              int64_t istart, int64_t iend, double os_factor, double ss_factor, int64_t part,
                              int64_t nparts, double* vv, double* vev, void* stream, int sync)
{
    return moments_rhf_impl(ixq, ld_ixq, qja, ld_qja, ei, ea, nphys, nocc, nvir, naux, istart, iend, os_factor,
                            ss_factor, part, nparts, vv, vev, stream, sync, nullptr);
}

static int moments_uhf_impl(const double* ixq_a, int64_t ld_ixq, const double* qja_a, int64_t ld_qja_a,
                            const double* qja_b, int64_t ld_qja_b, const double* ei_a, const double* ei_b,
                            const double* ea_a, const double* ea_b, int64_t nphys, int64_t nocc_a, int64_t nocc_b,
                            int64_t nvir_a, int64_t nvir_b, int64_t naux, int64_t istart, int64_t iend,
                            double os_factor, double ss_factor, int64_t part, int64_t nparts, double* vv,
                            double* vev, void* stream, int sync, const PoleGate* gate)
{
    if (int r = ensure_ctx()) return r;
    if (!ixq_a || !qja_a || !ei_a || !ea_a || !vv || !vev) return fail(AGF2_B200_EINVAL, "null pointer");
    if (nocc_a <= 0) return fail(AGF2_B200_EINVAL, "empty dimension");
    const bool has_b = nocc_b > 0 && nvir_b > 0;
    if (has_b && (!qja_b || !ei_b || !ea_b)) return fail(AGF2_B200_EINVAL, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;   // used as given: 0 is the CUDA default stream
    std::vector<double> eia((size_t)nocc_a), eib((size_t)std::max<int64_t>(nocc_b, 1));
    CU_TRY(cudaMemcpyAsync(eia.data(), ei_a, nocc_a * 8, cudaMemcpyDeviceToHost, st));
    if (has_b) CU_TRY(cudaMemcpyAsync(eib.data(), ei_b, nocc_b * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    Spin same, other;
    same.qja = qja_a; same.ld = ld_qja_a; same.nocc = nocc_a; same.nvir = nvir_a; same.ea_dev = ea_a; same.ei_dev = ei_a;
    other.qja = qja_b; other.ld = ld_qja_b; other.nocc = nocc_b; other.nvir = nvir_b; other.ea_dev = ea_b; other.ei_dev = ei_b;
    // same-spin: Coulomb ss, exchange ss (optuagf2.c:211-256); opposite spin: Coulomb os
    int r = run_moments(ixq_a, ld_ixq, same, has_b ? &other : nullptr, eia.data(), eib.data(), nphys, naux, istart,
                        iend, ss_factor, ss_factor, os_factor, part, nparts, vv, vev, st, gate);
    if (r) return r;
    if (sync) return finish_profile(st);
    return 0;
}

int agf2_b200_moments_uhf_dev(const double* ixq_a, int64_t ld_ixq, const double* qja_a, int64_t ld_qja_a,
                              const double* qja_b, int64_t ld_qja_b, const double* ei_a, const double* ei_b,
                              const double* ea_a, const double* ea_b, int64_t nphys, int64_t nocc_a, int64_t nocc_b,
                              int64_t nvir_a, int64_t nvir_b, int64_t naux, int64_t istart, int64_t iend,
                              double os_factor, double ss_factor, int64_t part, int64_t nparts, double* vv,
                              double* vev, void* stream, int sync)
{
    return moments_uhf_impl(ixq_a, ld_ixq, qja_a, ld_qja_a, qja_b, ld_qja_b, ei_a, ei_b, ea_a, ea_b, nphys, nocc_a,
                            nocc_b, nvir_a, nvir_b, naux, istart, iend, os_factor, ss_factor, part, nparts, vv, vev,
                            stream, sync, nullptr);
}

int agf2_b200_build_part_loop_rhf(const double* ixq, const double* qja, const double* ei, const double* ea,
                                  int64_t nphys, int64_t nocc, int64_t nvir, int64_t naux, int64_t istart,
                                  int64_t iend, double os_factor, double ss_factor, double* vv, double* vev)
{
    return agf2_b200_build_part_loop_rhf_sharded(ixq, qja, ei, ea, nphys, nocc, nvir, naux, istart, iend, os_factor,
                                                 ss_factor, 0, 1, vv, vev);
}

int agf2_b200_build_part_loop_rhf_sharded(const double* ixq, const double* qja, const double* ei, const double* ea,
                                          int64_t nphys, int64_t nocc, int64_t nvir, int64_t naux, int64_t istart,
                                          int64_t iend, double os_factor, double ss_factor, int64_t part,
                                          int64_t nparts, double* vv, double* vev)
{
    if (int r = ensure_ctx()) return r;
    if (!ixq || !qja || !ei || !ea || !vv || !vev) return fail(AGF2_B200_EINVAL, "null pointer");
    if (nphys <= 0 || nocc <= 0 || nvir <= 0 || naux <= 0) return fail(AGF2_B200_EINVAL, "empty dimension");
    if (istart >= iend) return 0;
    Ctx& c = g_ctx;
    cudaStream_t st = c.stream;
    int64_t ld_ixq, ld_qja, ld1;
    if (int r = stage_in(1, qja, naux * nocc, nvir, &ld_qja, st)) return r;   // (naux*nocc) rows of nvir
    if (int r = stage_in(2, ei, 1, nocc, &ld1, st)) return r;
    if (int r = stage_in(3, ea, 1, nvir, &ld1, st)) return r;
    if (int r = grow(c.stage[4], c.stage_elems[4], (size_t)2 * nphys * nphys, false)) return r;
    CU_TRY(cudaMemsetAsync(c.stage[4], 0, (size_t)2 * nphys * nphys * 8, st));
    double* dvv = c.stage[4];
    double* dvev = c.stage[4] + nphys * nphys;
    Uploader up;      // ixq streams in pole by pole while the first pairs are already being processed
    if (int r = up.start(ixq, 0, nocc, nphys, naux, istart, iend, &ld_ixq)) return r;
    int r = moments_rhf_impl(c.stage[0], ld_ixq, c.stage[1], ld_qja, c.stage[2], c.stage[3], nphys, nocc, nvir,
                             naux, istart, iend, os_factor, ss_factor, part, nparts, dvv, dvev, st, 1, &up.g);
    up.join();
    if (r) return r;
    std::vector<double> h((size_t)2 * nphys * nphys);
    CU_TRY(cudaMemcpy(h.data(), c.stage[4], h.size() * 8, cudaMemcpyDeviceToHost));
    const size_t n = (size_t)nphys * nphys;
    for (size_t k = 0; k < n; ++k) { vv[k] += h[k]; vev[k] += h[n + k]; }   // `+=` as optragf2.c:305-309
    return 0;
}

int agf2_b200_build_part_loop_uhf(const double* ixq_a, const double* qja_a, const double* qja_b, const double* ei_a,
                                  const double* ei_b, const double* ea_a, const double* ea_b, int64_t nphys,
                                  int64_t nocc_a, int64_t nocc_b, int64_t nvir_a, int64_t nvir_b, int64_t naux,
                                  int64_t istart, int64_t iend, double os_factor, double ss_factor, double* vv,
                                  double* vev)
{
    return agf2_b200_build_part_loop_uhf_sharded(ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b, nphys, nocc_a, nocc_b,
                                                 nvir_a, nvir_b, naux, istart, iend, os_factor, ss_factor, 0, 1, vv, vev);
}

int agf2_b200_build_part_loop_uhf_sharded(const double* ixq_a, const double* qja_a, const double* qja_b,
                                          const double* ei_a, const double* ei_b, const double* ea_a,
                                          const double* ea_b, int64_t nphys, int64_t nocc_a, int64_t nocc_b,
                                          int64_t nvir_a, int64_t nvir_b, int64_t naux, int64_t istart, int64_t iend,
                                          double os_factor, double ss_factor, int64_t part, int64_t nparts,
                                          double* vv, double* vev)
{
    if (int r = ensure_ctx()) return r;
    if (!ixq_a || !qja_a || !ei_a || !ea_a || !vv || !vev) return fail(AGF2_B200_EINVAL, "null pointer");
    if (nphys <= 0 || nocc_a <= 0 || nvir_a <= 0 || naux <= 0) return fail(AGF2_B200_EINVAL, "empty dimension");
    if (istart >= iend) return 0;
    Ctx& c = g_ctx;
    cudaStream_t st = c.stream;
    const bool has_b = nocc_b > 0 && nvir_b > 0;
    int64_t ld_ixq, ld_qa, ld_qb = 2, ld1;
    if (int r = stage_in(1, qja_a, naux * nocc_a, nvir_a, &ld_qa, st)) return r;
    if (int r = stage_in(2, ei_a, 1, nocc_a, &ld1, st)) return r;
    if (int r = stage_in(3, ea_a, 1, nvir_a, &ld1, st)) return r;
    if (has_b) {
        if (int r = stage_in(5, qja_b, naux * nocc_b, nvir_b, &ld_qb, st)) return r;
        if (int r = stage_in(6, ei_b, 1, nocc_b, &ld1, st)) return r;
        if (int r = stage_in(7, ea_b, 1, nvir_b, &ld1, st)) return r;
    }
    if (int r = grow(c.stage[4], c.stage_elems[4], (size_t)2 * nphys * nphys, false)) return r;
    CU_TRY(cudaMemsetAsync(c.stage[4], 0, (size_t)2 * nphys * nphys * 8, st));
    double* dvv = c.stage[4];
    double* dvev = c.stage[4] + nphys * nphys;
    Uploader up;
    if (int r = up.start(ixq_a, 0, nocc_a, nphys, naux, istart, iend, &ld_ixq)) return r;
    int r = moments_uhf_impl(c.stage[0], ld_ixq, c.stage[1], ld_qa, has_b ? c.stage[5] : nullptr, ld_qb,
                             c.stage[2], has_b ? c.stage[6] : nullptr, c.stage[3], has_b ? c.stage[7] : nullptr,
                             nphys, nocc_a, has_b ? nocc_b : 0, nvir_a, has_b ? nvir_b : 0, naux, istart, iend,
                             os_factor, ss_factor, part, nparts, dvv, dvev, st, 1, &up.g);
    up.join();
    if (r) return r;
    std::vector<double> h((size_t)2 * nphys * nphys);
    CU_TRY(cudaMemcpy(h.data(), c.stage[4], h.size() * 8, cudaMemcpyDeviceToHost));
    const size_t n = (size_t)nphys * nphys;
    for (size_t k = 0; k < n; ++k) { vv[k] += h[k]; vev[k] += h[n + k]; }
    return 0;
}

// ---- drop-in symbols (reference signatures) ---------------------------------------------------------
static void die(const char* fn) {
    fprintf(stderr, "libagf2_b200: %s failed: %s (no CPU fallback)\n", fn, g_err.c_str());
    abort();
}

void build_part_loop_rhf(double* ixq, double* qja, double* ei, double* ea, uint32_t nphys, uint32_t nocc,
                         uint32_t nvir, uint32_t naux, uint32_t istart, uint32_t iend, double* vv, double* vev)
{
    if (agf2_b200_build_part_loop_rhf(ixq, qja, ei, ea, nphys, nocc, nvir, naux, istart, iend, 1.0, 1.0, vv, vev))
        die("build_part_loop_rhf");
}

void build_part_loop_uhf(double* ixq_a, double* qja_a, double* qja_b, double* ei_a, double* ei_b, double* ea_a,
                         double* ea_b, uint32_t nphys, uint32_t nocc_a, uint32_t nocc_b, uint32_t nvir_a,
                         uint32_t nvir_b, uint32_t naux, uint32_t istart, uint32_t iend, double* vv, double* vev)
{
    if (agf2_b200_build_part_loop_uhf(ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b, nphys, nocc_a, nocc_b, nvir_a,
                                      nvir_b, naux, istart, iend, 1.0, 1.0, vv, vev))
        die("build_part_loop_uhf");
}

}  // extern "C"
