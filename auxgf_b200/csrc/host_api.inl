// The C-ABI of the moment build (include/agf2_b200.h): contexts, options, the device-resident and host-pointer
// entry points and the reference's own two symbols.  Part of agf2_b200.cu (included there).
namespace {

// copy a host matrix (rows x cols, contiguous) into a device staging slot with an even pitch.  With a communicator
// (shard): every rank copies rows [rows r/N, rows (r+1)/N) and an all-gather over NVLink replicates them, so that the
// PCIe traffic of the job is one copy of the matrix instead of N.
int stage_in(Ctx& c, int slot, const double* host, int64_t rows, int64_t cols, int64_t* ld_out, bool shard, cudaStream_t st) {
    const int64_t ld = rup(cols, 2);
    if (int r = grow(c.stage[slot], (size_t)std::max<int64_t>(rows * ld, 2), false)) return r;
    *ld_out = ld;
    double* dev = c.stage[slot].p;
    const int64_t per = shard ? rows / c.nranks : 0;
    if (per > 0 && per * ld * 8 >= (1 << 20)) {
        const int64_t r0 = per * c.rank;
        CU_TRY(cudaMemcpy2DAsync(dev + r0 * ld, ld * 8, host + r0 * cols, cols * 8, cols * 8, per, cudaMemcpyHostToDevice, st));
        if (int r = comm_allgather(c, dev, per * ld, st)) return r;
        const int64_t done = per * c.nranks;       // ragged tail: every rank copies it
        if (rows > done)
            CU_TRY(cudaMemcpy2DAsync(dev + done * ld, ld * 8, host + done * cols, cols * 8, cols * 8, rows - done,
                                     cudaMemcpyHostToDevice, st));
    } else {
        CU_TRY(cudaMemcpy2DAsync(dev, ld * 8, host, cols * 8, cols * 8, rows, cudaMemcpyHostToDevice, st));
    }
    return 0;
}

// Pipelined upload of ixq (nocc poles of nphys x naux) for the host-pointer API: a helper thread copies pole after
// pole on the copy stream while the pair loop already runs on the poles that have landed.  With a communicator the
// poles are handled in groups of N * sb: rank r copies its sb poles of the group from host memory and an in-place
// all-gather brings in the others over NVLink.
struct Uploader {
    std::thread th;
    std::atomic<int> ready{0}, failed{0};
    std::vector<int> rank;
    PoleGate g;
    std::string error;

    // host_q (optional; only without a communicator): (Q|ja) is not staged up front but travels with the poles -- pole i
    // brings its columns qja[:, i, :] (naux rows of nvir doubles) along, so that the pair loop starts after the first
    // two poles instead of after the whole tensor
    int start(Ctx& c, const double* host, int slot, int64_t nocc, int64_t nphys, int64_t naux, int64_t istart,
              int64_t iend, bool shard, int64_t* ld_out, const double* host_q = nullptr, double* dev_q = nullptr,
              int64_t nvir = 0, int64_t ld_q = 0) {
        const int64_t ld = rup(naux, 2);
        if (int r = grow(c.stage[slot], (size_t)std::max<int64_t>(nocc * nphys * ld, 2), false)) return r;
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
        g.rank = &rank; g.ev = &c.pole_ev; g.ready = &ready; g.failed = &failed; g.with_qja = host_q != nullptr;
        double* dev = c.stage[slot].p;
        const int device = c.device;
        cudaStream_t cs = c.copy_stream;
        const int64_t pole = nphys * ld;
        // sharded upload only for the natural pole order (the all-gather needs contiguous groups)
        const bool gather = shard && c.nranks > 1 && c.comm && istart == 0 && iend == nocc && nocc >= c.nranks;
        const int64_t sb = gather ? std::max<int64_t>(1, std::min<int64_t>((32ll << 20) / (pole * 8), nocc / c.nranks)) : 0;
        const int64_t group = sb * c.nranks;
        Ctx* cp = &c;
        th = std::thread([=]() {
            if (cudaSetDevice(device) != cudaSuccess) { failed.store(1); return; }
            auto copy_pole = [&](int64_t i) {
                cudaError_t e = cudaMemcpy2DAsync(dev + i * pole, ld * 8, host + i * nphys * naux, naux * 8, naux * 8, nphys,
                                                  cudaMemcpyHostToDevice, cs);
                if (e == cudaSuccess && host_q)
                    e = cudaMemcpy2DAsync(dev_q + i * ld_q, nocc * ld_q * 8, host_q + i * nvir, nocc * nvir * 8, nvir * 8,
                                          naux, cudaMemcpyHostToDevice, cs);
                return e;
            };
            size_t k = 0;
            while (k < order.size()) {
                size_t k_end = k + 1;
                cudaError_t e = cudaSuccess;
                if (gather && k + (size_t)group <= order.size()) {
                    k_end = k + (size_t)group;
                    for (int64_t s = 0; s < sb && e == cudaSuccess; ++s) e = copy_pole((int64_t)k + cp->rank * sb + s);
                    if (e == cudaSuccess && comm_allgather(*cp, dev + (int64_t)k * pole, sb * pole, cs)) {
                        error = last_error();
                        failed.store(1);
                        return;
                    }
                } else {
                    e = copy_pole(order[k]);
                }
                for (size_t q = k; q < k_end && e == cudaSuccess; ++q) e = cudaEventRecord(cp->pole_ev[q], cs);
                if (e != cudaSuccess) { error = cudaGetErrorString(e); failed.store(1); return; }
                ready.store((int)k_end, std::memory_order_release);
                k = k_end;
            }
        });
        return 0;
    }
    void join() { if (th.joinable()) th.join(); }
    ~Uploader() { join(); }
};

// pole energies for the planner: the caller's host copy, or a D2H copy through pinned staging
int host_energies(Ctx& c, const double* dev, const double* host, int64_t n, size_t off, const double** out, cudaStream_t st,
                  bool* need_sync) {
    if (host) { *out = host; return 0; }
    CU_TRY(cudaMemcpyAsync(c.h_energy + off, dev, n * 8, cudaMemcpyDeviceToHost, st));
    *out = c.h_energy + off;
    *need_sync = true;
    return 0;
}

int moments_rhf_impl(Ctx& c, const double* ixq, int64_t ld_ixq, const double* qja, int64_t ld_qja, const double* ei,
                     const double* ea, const double* ei_host, int64_t nphys, int64_t nocc, int64_t nvir, int64_t naux,
                     int64_t istart, int64_t iend, double os_factor, double ss_factor, int64_t part, int64_t nparts,
                     double* vv, double* vev, cudaStream_t st, int sync, const PoleGate* gate, bool allreduce)
{
    if (!ixq || !qja || !ei || !ea || !vv || !vev) return fail(AGF2_B200_EINVAL, "null pointer");
    if (nocc <= 0) return fail(AGF2_B200_EINVAL, "empty dimension");
    if (int r = grow_pinned(c.h_energy, c.h_energy_cap, (size_t)nocc)) return r;
    const double* ei_h = nullptr;
    bool need_sync = false;
    if (int r = host_energies(c, ei, ei_host, nocc, 0, &ei_h, st, &need_sync)) return r;
    if (need_sync) CU_TRY(cudaStreamSynchronize(st));
    Spin same;
    same.qja = qja; same.ld = ld_qja; same.nocc = nocc; same.nvir = nvir; same.ea_dev = ea; same.ei_dev = ei;
    int r = run_moments(c, ixq, ld_ixq, same, nullptr, ei_h, nullptr, nphys, naux, istart, iend,
                        os_factor + ss_factor, ss_factor, 0.0, part, nparts, vv, vev, st, gate, allreduce);
    if (r) return r;
    if (sync) return finish_profile(c, st);
    return 0;
}

int moments_uhf_impl(Ctx& c, const double* ixq_a, int64_t ld_ixq, const double* qja_a, int64_t ld_qja_a,
                     const double* qja_b, int64_t ld_qja_b, const double* ei_a, const double* ei_b,
                     const double* ea_a, const double* ea_b, const double* ei_a_host, const double* ei_b_host,
                     int64_t nphys, int64_t nocc_a, int64_t nocc_b, int64_t nvir_a, int64_t nvir_b, int64_t naux,
                     int64_t istart, int64_t iend, double os_factor, double ss_factor, int64_t part, int64_t nparts,
                     double* vv, double* vev, cudaStream_t st, int sync, const PoleGate* gate, bool allreduce)
{
    if (!ixq_a || !qja_a || !ei_a || !ea_a || !vv || !vev) return fail(AGF2_B200_EINVAL, "null pointer");
    if (nocc_a <= 0) return fail(AGF2_B200_EINVAL, "empty dimension");
    const bool has_b = nocc_b > 0 && nvir_b > 0;
    if (has_b && (!qja_b || !ei_b || !ea_b)) return fail(AGF2_B200_EINVAL, "null pointer");
    if (int r = grow_pinned(c.h_energy, c.h_energy_cap, (size_t)(nocc_a + std::max<int64_t>(nocc_b, 1)))) return r;
    const double* eia = nullptr;
    const double* eib = nullptr;
    bool need_sync = false;
    if (int r = host_energies(c, ei_a, ei_a_host, nocc_a, 0, &eia, st, &need_sync)) return r;
    if (has_b) if (int r = host_energies(c, ei_b, ei_b_host, nocc_b, (size_t)nocc_a, &eib, st, &need_sync)) return r;
    if (need_sync) CU_TRY(cudaStreamSynchronize(st));
    Spin same, other;
    same.qja = qja_a; same.ld = ld_qja_a; same.nocc = nocc_a; same.nvir = nvir_a; same.ea_dev = ea_a; same.ei_dev = ei_a;
    other.qja = qja_b; other.ld = ld_qja_b; other.nocc = nocc_b; other.nvir = nvir_b; other.ea_dev = ea_b; other.ei_dev = ei_b;
    // same-spin: Coulomb ss, exchange ss (optuagf2.c:211-256); opposite spin: Coulomb os
    int r = run_moments(c, ixq_a, ld_ixq, same, has_b ? &other : nullptr, eia, eib, nphys, naux, istart, iend,
                        ss_factor, ss_factor, os_factor, part, nparts, vv, vev, st, gate, allreduce);
    if (r) return r;
    if (sync) return finish_profile(c, st);
    return 0;
}

// result of a host-pointer call: [vv | vev] from the staging slot through pinned memory, added into the caller's
int add_result_to_host(Ctx& c, const double* dev, int64_t nphys, double* vv, double* vev, cudaStream_t st) {
    const size_t n = (size_t)nphys * nphys;
    if (int r = grow_pinned(c.h_result, c.h_result_cap, 2 * n)) return r;
    CU_TRY(cudaMemcpyAsync(c.h_result, dev, 2 * n * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    const double* h = c.h_result;
    for (size_t k = 0; k < n; ++k) { vv[k] += h[k]; vev[k] += h[n + k]; }   // `+=` as optragf2.c:305-309
    return 0;
}

// `comm`: use the context's communicator (part = rank, sharded upload + all-gather, one all-reduce) instead of the
// explicit part / nparts
int host_rhf(Ctx& c, const double* ixq, const double* qja, const double* ei, const double* ea, int64_t nphys,
             int64_t nocc, int64_t nvir, int64_t naux, int64_t istart, int64_t iend, double os_factor, double ss_factor,
             int64_t part, int64_t nparts, bool comm, double* vv, double* vev)
{
    if (!ixq || !qja || !ei || !ea || !vv || !vev) return fail(AGF2_B200_EINVAL, "null pointer");
    if (nphys <= 0 || nocc <= 0 || nvir <= 0 || naux <= 0) return fail(AGF2_B200_EINVAL, "empty dimension");
    if (istart >= iend) return 0;
    std::lock_guard<std::mutex> lk(c.mu);
    DeviceGuard guard(c.device);
    cudaStream_t st = c.stream;
    const bool multi = comm && c.nranks > 1;
    if (multi) { part = c.rank; nparts = c.nranks; }
    c.last_comm_bytes[0] = c.last_comm_bytes[1] = c.last_comm_bytes[2] = 0;
    int64_t ld_ixq, ld_qja, ld1;
    // qja travels on the copy stream too: with a communicator its all-gather must precede the uploader's on every rank;
    // without one its columns ride along with the poles of ixq (Uploader)
    const bool q_with_poles = !multi;
    if (q_with_poles) {
        ld_qja = rup(nvir, 2);
        if (int r = grow(c.stage[1], (size_t)std::max<int64_t>(naux * nocc * ld_qja, 2), false)) return r;
    } else if (int r = stage_in(c, 1, qja, naux * nocc, nvir, &ld_qja, multi, c.copy_stream)) return r;   // (naux*nocc) rows of nvir
    if (int r = stage_in(c, 2, ei, 1, nocc, &ld1, false, c.copy_stream)) return r;
    if (int r = stage_in(c, 3, ea, 1, nvir, &ld1, false, c.copy_stream)) return r;
    CU_TRY(cudaEventRecord(c.ev_copy, c.copy_stream));
    CU_TRY(cudaStreamWaitEvent(st, c.ev_copy, 0));
    if (int r = grow(c.stage[4], (size_t)2 * nphys * nphys, false)) return r;
    CU_TRY(cudaMemsetAsync(c.stage[4].p, 0, (size_t)2 * nphys * nphys * 8, st));
    double* dvv = c.stage[4].p;
    double* dvev = c.stage[4].p + nphys * nphys;
    Uploader up;      // ixq streams in pole by pole while the first pairs are already being processed
    if (int r = up.start(c, ixq, 0, nocc, nphys, naux, istart, iend, multi, &ld_ixq, q_with_poles ? qja : nullptr,
                         c.stage[1].p, nvir, ld_qja))
        return r;
    int r = moments_rhf_impl(c, c.stage[0].p, ld_ixq, c.stage[1].p, ld_qja, c.stage[2].p, c.stage[3].p, ei, nphys, nocc,
                             nvir, naux, istart, iend, os_factor, ss_factor, part, nparts, dvv, dvev, st, 0, &up.g, false);
    up.join();
    if (up.failed.load() && !r) r = fail(AGF2_B200_ECUDA, "ixq upload failed: " + up.error);
    if (r) { cudaDeviceSynchronize(); return r; }
    if (multi) if (int rr = comm_allreduce_sum(c, c.stage[4].p, 2 * nphys * nphys, st)) return rr;
    if (int rr = add_result_to_host(c, c.stage[4].p, nphys, vv, vev, st)) return rr;
    return finish_profile(c, st);
}

int host_uhf(Ctx& c, const double* ixq_a, const double* qja_a, const double* qja_b, const double* ei_a,
             const double* ei_b, const double* ea_a, const double* ea_b, int64_t nphys, int64_t nocc_a, int64_t nocc_b,
             int64_t nvir_a, int64_t nvir_b, int64_t naux, int64_t istart, int64_t iend, double os_factor,
             double ss_factor, int64_t part, int64_t nparts, bool comm, double* vv, double* vev)
{
    if (!ixq_a || !qja_a || !ei_a || !ea_a || !vv || !vev) return fail(AGF2_B200_EINVAL, "null pointer");
    if (nphys <= 0 || nocc_a <= 0 || nvir_a <= 0 || naux <= 0) return fail(AGF2_B200_EINVAL, "empty dimension");
    if (istart >= iend) return 0;
    std::lock_guard<std::mutex> lk(c.mu);
    DeviceGuard guard(c.device);
    cudaStream_t st = c.stream;
    const bool multi = comm && c.nranks > 1;
    if (multi) { part = c.rank; nparts = c.nranks; }
    c.last_comm_bytes[0] = c.last_comm_bytes[1] = c.last_comm_bytes[2] = 0;
    const bool has_b = nocc_b > 0 && nvir_b > 0;
    int64_t ld_ixq, ld_qa, ld_qb = 2, ld1;
    if (int r = stage_in(c, 1, qja_a, naux * nocc_a, nvir_a, &ld_qa, multi, c.copy_stream)) return r;
    if (int r = stage_in(c, 2, ei_a, 1, nocc_a, &ld1, false, c.copy_stream)) return r;
    if (int r = stage_in(c, 3, ea_a, 1, nvir_a, &ld1, false, c.copy_stream)) return r;
    if (has_b) {
        if (int r = stage_in(c, 5, qja_b, naux * nocc_b, nvir_b, &ld_qb, multi, c.copy_stream)) return r;
        if (int r = stage_in(c, 6, ei_b, 1, nocc_b, &ld1, false, c.copy_stream)) return r;
        if (int r = stage_in(c, 7, ea_b, 1, nvir_b, &ld1, false, c.copy_stream)) return r;
    }
    CU_TRY(cudaEventRecord(c.ev_copy, c.copy_stream));
    CU_TRY(cudaStreamWaitEvent(st, c.ev_copy, 0));
    if (int r = grow(c.stage[4], (size_t)2 * nphys * nphys, false)) return r;
    CU_TRY(cudaMemsetAsync(c.stage[4].p, 0, (size_t)2 * nphys * nphys * 8, st));
    double* dvv = c.stage[4].p;
    double* dvev = c.stage[4].p + nphys * nphys;
    Uploader up;
    if (int r = up.start(c, ixq_a, 0, nocc_a, nphys, naux, istart, iend, multi, &ld_ixq)) return r;
    int r = moments_uhf_impl(c, c.stage[0].p, ld_ixq, c.stage[1].p, ld_qa, has_b ? c.stage[5].p : nullptr, ld_qb,
                             c.stage[2].p, has_b ? c.stage[6].p : nullptr, c.stage[3].p, has_b ? c.stage[7].p : nullptr,
                             ei_a, has_b ? ei_b : nullptr, nphys, nocc_a, has_b ? nocc_b : 0, nvir_a, has_b ? nvir_b : 0,
                             naux, istart, iend, os_factor, ss_factor, part, nparts, dvv, dvev, st, 0, &up.g, false);
    up.join();
    if (up.failed.load() && !r) r = fail(AGF2_B200_ECUDA, "ixq upload failed: " + up.error);
    if (r) { cudaDeviceSynchronize(); return r; }
    if (multi) if (int rr = comm_allreduce_sum(c, c.stage[4].p, 2 * nphys * nphys, st)) return rr;
    if (int rr = add_result_to_host(c, c.stage[4].p, nphys, vv, vev, st)) return rr;
    return finish_profile(c, st);
}

// first failure of a void drop-in call (they cannot return a status)
std::mutex g_sticky_mu;
int g_sticky_status = 0;
std::string g_sticky_msg;

void drop_in_failed(const char* fn, int status) {
    {
        std::lock_guard<std::mutex> lk(g_sticky_mu);
        if (g_sticky_status == 0) { g_sticky_status = status; g_sticky_msg = std::string(fn) + ": " + last_error(); }
    }
    fprintf(stderr, "libagf2_b200: %s failed with status %d: %s (no CPU fallback; vv/vev left untouched)\n", fn, status,
            last_error().c_str());
    const char* e = getenv("AGF2_B200_ABORT_ON_ERROR");
    if (e && atoi(e) != 0) abort();
}

}  // namespace

// =================================================================================================
extern "C" {

const char* agf2_b200_last_error(void) { return last_error().c_str(); }
const char* agf2_b200_version(void) { return "agf2_b200 0.2 (sm_100a; DMMA.8x8x4 + TMA; NCCL communicator)"; }

int agf2_b200_sticky_status(int clear) {
    std::lock_guard<std::mutex> lk(g_sticky_mu);
    const int s = g_sticky_status;
    if (s) last_error() = g_sticky_msg;
    if (clear) { g_sticky_status = 0; g_sticky_msg.clear(); }
    return s;
}

// ---- contexts -----------------------------------------------------------------------------------
int agf2_b200_ctx_create(int device, agf2_b200_ctx** out) {
    if (!out) return fail(AGF2_B200_EINVAL, "null pointer");
    Ctx* c = new Ctx();
    c->own_opts = default_options();
    if (int r = ctx_init(*c, device)) { delete c; return r; }
    *out = c;
    return 0;
}

int agf2_b200_ctx_destroy(agf2_b200_ctx* ctx) {
    if (!ctx) return 0;
    if (ctx->is_default) return fail(AGF2_B200_EINVAL, "the default context of a device cannot be destroyed");
    ctx_free(ctx);
    return 0;
}

int agf2_b200_ctx_get_device(agf2_b200_ctx* ctx, int* device) {
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    if (device) *device = c->device;
    return 0;
}

int agf2_b200_ctx_synchronize(agf2_b200_ctx* ctx) {
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    DeviceGuard guard(c->device);
    CU_TRY(cudaDeviceSynchronize());
    return 0;
}

static int set_option(Options& o, const char* name, int64_t value) {
    const std::string n = name ? name : "";
    if (n == "workspace_bytes") {
        if (value < (1 << 20)) return fail(AGF2_B200_EINVAL, "workspace must be at least 1 MiB");
        o.ws_bytes = value;
    } else if (n == "simple_kernels") o.simple = value != 0;
    else if (n == "overlap") o.overlap = value != 0;
    else if (n == "lpt_order") o.lpt_order = value != 0;
    else if (n == "tune_chunk") o.tune_chunk = value != 0;
    else if (n == "thin_tiles") o.thin_tiles = value != 0;
    else if (n == "deterministic") o.deterministic = value != 0;
    else if (n == "profile") o.profile = value != 0;
    else if (n == "visiting_bytes") {
        if (value < 1) return fail(AGF2_B200_EINVAL, "visiting_bytes must be positive");
        o.vis_bytes = value;
    } else if (n == "shard_step_limit") o.shard_limit = value < 0 ? 0 : value;
    else if (n == "coulomb_factorization") {
        if (value < 0 || value > 2) return fail(AGF2_B200_EINVAL, "mode must be 0 (off), 1 (always) or 2 (automatic)");
        o.coulomb = value != 0;
        o.coulomb_auto = value == 2;
    } else return fail(AGF2_B200_EINVAL, "unknown option '" + n + "'");
    return 0;
}

/* ctx == NULL: the process defaults, shared by the default contexts of all devices */
int agf2_b200_ctx_set_option(agf2_b200_ctx* ctx, const char* name, int64_t value) {
    return set_option(ctx ? *ctx->opt : default_options(), name, value);
}

int agf2_b200_set_device(int device) {
    CU_TRY(cudaSetDevice(device));
    return 0;
}
int agf2_b200_set_workspace_bytes(int64_t bytes) { return set_option(default_options(), "workspace_bytes", bytes); }
int agf2_b200_set_simple_kernels(int on) { return set_option(default_options(), "simple_kernels", on); }
int agf2_b200_set_overlap(int on) { return set_option(default_options(), "overlap", on); }
int agf2_b200_set_coulomb_factorization(int mode) { return set_option(default_options(), "coulomb_factorization", mode); }
int agf2_b200_set_deterministic(int on) { return set_option(default_options(), "deterministic", on); }

int64_t agf2_b200_launch_count(int reset) {
    const int64_t n = g_launches.load();
    if (reset) g_launches = 0;
    return n;
}

int agf2_b200_ctx_last_stage_ms(agf2_b200_ctx* ctx, double* ms4) {
    if (!ms4) return fail(AGF2_B200_EINVAL, "null pointer");
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    for (int k = 0; k < 4; ++k) ms4[k] = c->last_ms[k];
    return 0;
}
int agf2_b200_last_stage_ms(double* ms4) { return agf2_b200_ctx_last_stage_ms(nullptr, ms4); }

int agf2_b200_ctx_last_plan(agf2_b200_ctx* ctx, int64_t* info6) {
    if (!info6) return fail(AGF2_B200_EINVAL, "null pointer");
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    for (int k = 0; k < 6; ++k) info6[k] = c->last_plan[k];
    return 0;
}
int agf2_b200_last_plan(int64_t* info6) { return agf2_b200_ctx_last_plan(nullptr, info6); }

int agf2_b200_ctx_last_comm_bytes(agf2_b200_ctx* ctx, int64_t* bytes2) {
    if (!bytes2) return fail(AGF2_B200_EINVAL, "null pointer");
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    bytes2[0] = c->last_comm_bytes[0];
    bytes2[1] = c->last_comm_bytes[1];
    return 0;
}

// sub-block size (poles) a pole-sharded build of this shape uses on this context
int64_t agf2_b200_ctx_shard_sub_block(agf2_b200_ctx* ctx, int64_t nocc, int64_t nphys, int64_t ld_ixq, int64_t nranks) {
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c || nranks < 1 || nocc < nranks) return -1;
    int64_t nb_max = 0;
    for (int64_t r = 0; r < nranks; ++r) nb_max = std::max(nb_max, shard_lo(nocc, r + 1, nranks) - shard_lo(nocc, r, nranks));
    return shard_sub_block(*c->opt, nb_max, nphys * ld_ixq);
}

int64_t agf2_b200_ctx_last_p2p_bytes(agf2_b200_ctx* ctx) {
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    return c ? c->last_comm_bytes[2] : -1;
}

int agf2_b200_pole_block(int64_t nocc, int64_t rank, int64_t nranks, int64_t* lo, int64_t* hi) {
    if (nocc < 0 || nranks < 1 || rank < 0 || rank >= nranks || !lo || !hi) return fail(AGF2_B200_EINVAL, "bad argument");
    *lo = shard_lo(nocc, rank, nranks);
    *hi = shard_lo(nocc, rank + 1, nranks);
    return 0;
}

// Host-only replay of the pole-sharded schedule of all ranks (tests): cover[i * nocc + j] += 1 for every pair item
// (i, j) any rank forms, stats[2 r] = pairs of rank r, stats[2 r + 1] = poles rank r sends.
int agf2_b200_shard_cover(int64_t nocc, int64_t nranks, int64_t sub_block, int64_t limit, int64_t* cover, int64_t* stats) {
    if (nocc < nranks || nranks < 1 || sub_block < 1 || !cover || !stats) return fail(AGF2_B200_EINVAL, "bad argument");
    for (int64_t r = 0; r < nranks; ++r) {
        MomentPlan plan;
        if (int rc = plan.init(plan_config(default_options(), 148), 8, 8, nocc, 8, false, 0, 0, 0, nocc, 2.0, 1.0, 0.0, 0, 1, true))
            return fail(rc, plan.error);
        for (const ShardStep& s : shard_steps(nocc, r, nranks, sub_block, limit)) {
            if (s.send) stats[2 * r + 1] += s.send_hi - s.send_lo;
            if (s.vis_hi <= s.vis_lo) continue;
            plan.set_shard_step(r, nranks, s);
            for (const Item& it : plan.sym_items) {
                const int64_t a = std::max(it.i, it.j), b = std::min(it.i, it.j);
                cover[a * nocc + b] += 1;
                stats[2 * r] += 1;
            }
        }
    }
    return 0;
}

int agf2_b200_ctx_flop_counters(agf2_b200_ctx* ctx, double* flops3, int reset) {
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    for (int k = 0; k < 3; ++k) {
        if (flops3) flops3[k] = c->flops[k];
        if (reset) c->flops[k] = 0.0;
    }
    return 0;
}

int agf2_b200_last_kernel_ms(double* k1, double* k2) {
    int s;
    Ctx* c = resolve_ctx(nullptr, &s);
    if (!c) return s;
    if (k1) *k1 = c->last_ms[0];
    if (k2) *k2 = c->last_ms[1];
    return 0;
}

int agf2_b200_plan_stats(int64_t nphys, int64_t nocc, int64_t nvir, int64_t naux, int64_t nocc_b, int64_t nvir_b,
                         int64_t istart, int64_t iend, double os_factor, double ss_factor, int uhf, int64_t part,
                         int64_t nparts, int64_t* cover, int64_t* stats8)
{
    if (!stats8) return fail(AGF2_B200_EINVAL, "null pointer");
    MomentPlan plan;
    const bool has_b = uhf && nocc_b > 0 && nvir_b > 0;
    // same coefficient convention as the moments entry points (RHF: c = os + ss, s = ss; UHF: c = s = ss, other os)
    const double c_same = uhf ? ss_factor : os_factor + ss_factor, os_other = uhf ? os_factor : 0.0;
    if (int r = plan.init(plan_config(default_options(), 148), nphys, naux, nocc, nvir, has_b, nocc_b, nvir_b, istart, iend,
                          c_same, ss_factor, os_other, part, nparts))
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

// ---- device-resident moments --------------------------------------------------------------------
int agf2_b200_moments_rhf_dev(const double* ixq, int64_t ld_ixq, const double* qja, int64_t ld_qja, const double* ei,
                              const double* ea, int64_t nphys, int64_t nocc, int64_t nvir, int64_t naux,
                              int64_t istart, int64_t iend, double os_factor, double ss_factor, int64_t part,
                              int64_t nparts, double* vv, double* vev, void* stream, int sync)
{
    int s;
    Ctx* c = resolve_ctx(nullptr, &s);
    if (!c) return s;
    std::lock_guard<std::mutex> lk(c->mu);
    return moments_rhf_impl(*c, ixq, ld_ixq, qja, ld_qja, ei, ea, nullptr, nphys, nocc, nvir, naux, istart, iend, os_factor,
                            ss_factor, part, nparts, vv, vev, (cudaStream_t)stream, sync, nullptr, false);
}

int agf2_b200_moments_uhf_dev(const double* ixq_a, int64_t ld_ixq, const double* qja_a, int64_t ld_qja_a,
                              const double* qja_b, int64_t ld_qja_b, const double* ei_a, const double* ei_b,
                              const double* ea_a, const double* ea_b, int64_t nphys, int64_t nocc_a, int64_t nocc_b,
                              int64_t nvir_a, int64_t nvir_b, int64_t naux, int64_t istart, int64_t iend,
                              double os_factor, double ss_factor, int64_t part, int64_t nparts, double* vv,
                              double* vev, void* stream, int sync)
{
    int s;
    Ctx* c = resolve_ctx(nullptr, &s);
    if (!c) return s;
    std::lock_guard<std::mutex> lk(c->mu);
    return moments_uhf_impl(*c, ixq_a, ld_ixq, qja_a, ld_qja_a, qja_b, ld_qja_b, ei_a, ei_b, ea_a, ea_b, nullptr, nullptr,
                            nphys, nocc_a, nocc_b, nvir_a, nvir_b, naux, istart, iend, os_factor, ss_factor, part, nparts,
                            vv, vev, (cudaStream_t)stream, sync, nullptr, false);
}

int agf2_b200_ctx_moments_rhf_dev(agf2_b200_ctx* ctx, const double* ixq, int64_t ld_ixq, const double* qja,
                                  int64_t ld_qja, const double* ei, const double* ea, const double* ei_host,
                                  int64_t nphys, int64_t nocc, int64_t nvir, int64_t naux, int64_t istart, int64_t iend,
                                  double os_factor, double ss_factor, double* vv, double* vev, void* stream, int sync)
{
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard guard(c->device);
    c->last_comm_bytes[0] = c->last_comm_bytes[1] = c->last_comm_bytes[2] = 0;
    return moments_rhf_impl(*c, ixq, ld_ixq, qja, ld_qja, ei, ea, ei_host, nphys, nocc, nvir, naux, istart, iend, os_factor,
                            ss_factor, c->rank, c->nranks, vv, vev, (cudaStream_t)stream, sync, nullptr, true);
}

int agf2_b200_ctx_moments_rhf_dev_sharded(agf2_b200_ctx* ctx, const double* ixq_block, int64_t ld_ixq, const double* qja,
                                          int64_t ld_qja, const double* ei, const double* ea, const double* ei_host,
                                          int64_t nphys, int64_t nocc, int64_t nvir, int64_t naux, double os_factor,
                                          double ss_factor, double* vv, double* vev, void* stream, int sync)
{
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    if (!ixq_block || !qja || !ei || !ea || !vv || !vev) return fail(AGF2_B200_EINVAL, "null pointer");
    if (nocc <= 0) return fail(AGF2_B200_EINVAL, "empty dimension");
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard guard(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    c->last_comm_bytes[0] = c->last_comm_bytes[1] = c->last_comm_bytes[2] = 0;
    if (int r = grow_pinned(c->h_energy, c->h_energy_cap, (size_t)nocc)) return r;
    const double* ei_h = nullptr;
    bool need_sync = false;
    if (int r = host_energies(*c, ei, ei_host, nocc, 0, &ei_h, st, &need_sync)) return r;
    if (need_sync) CU_TRY(cudaStreamSynchronize(st));
    Spin same;
    same.qja = qja; same.ld = ld_qja; same.nocc = nocc; same.nvir = nvir; same.ea_dev = ea; same.ei_dev = ei;
    const Shard sh{c->rank, c->nranks};
    if (int r = run_moments(*c, ixq_block, ld_ixq, same, nullptr, ei_h, nullptr, nphys, naux, 0, nocc,
                            os_factor + ss_factor, ss_factor, 0.0, 0, 1, vv, vev, st, nullptr, true, &sh))
        return r;
    if (sync) return finish_profile(*c, st);
    return 0;
}

int agf2_b200_ctx_moments_uhf_dev(agf2_b200_ctx* ctx, const double* ixq_a, int64_t ld_ixq, const double* qja_a,
                                  int64_t ld_qja_a, const double* qja_b, int64_t ld_qja_b, const double* ei_a,
                                  const double* ei_b, const double* ea_a, const double* ea_b, const double* ei_a_host,
                                  const double* ei_b_host, int64_t nphys, int64_t nocc_a, int64_t nocc_b,
                                  int64_t nvir_a, int64_t nvir_b, int64_t naux, int64_t istart, int64_t iend,
                                  double os_factor, double ss_factor, double* vv, double* vev, void* stream, int sync)
{
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard guard(c->device);
    c->last_comm_bytes[0] = c->last_comm_bytes[1] = c->last_comm_bytes[2] = 0;
    return moments_uhf_impl(*c, ixq_a, ld_ixq, qja_a, ld_qja_a, qja_b, ld_qja_b, ei_a, ei_b, ea_a, ea_b, ei_a_host,
                            ei_b_host, nphys, nocc_a, nocc_b, nvir_a, nvir_b, naux, istart, iend, os_factor, ss_factor,
                            c->rank, c->nranks, vv, vev, (cudaStream_t)stream, sync, nullptr, true);
}

// ---- host-pointer moments -----------------------------------------------------------------------
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
    int s;
    Ctx* c = resolve_ctx(nullptr, &s);
    if (!c) return s;
    return host_rhf(*c, ixq, qja, ei, ea, nphys, nocc, nvir, naux, istart, iend, os_factor, ss_factor, part, nparts,
                    false, vv, vev);
}

int agf2_b200_ctx_build_part_loop_rhf(agf2_b200_ctx* ctx, const double* ixq, const double* qja, const double* ei,
                                      const double* ea, int64_t nphys, int64_t nocc, int64_t nvir, int64_t naux,
                                      int64_t istart, int64_t iend, double os_factor, double ss_factor, double* vv,
                                      double* vev)
{
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    return host_rhf(*c, ixq, qja, ei, ea, nphys, nocc, nvir, naux, istart, iend, os_factor, ss_factor, 0, 1, true, vv, vev);
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
    int s;
    Ctx* c = resolve_ctx(nullptr, &s);
    if (!c) return s;
    return host_uhf(*c, ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b, nphys, nocc_a, nocc_b, nvir_a, nvir_b, naux, istart,
                    iend, os_factor, ss_factor, part, nparts, false, vv, vev);
}

int agf2_b200_ctx_build_part_loop_uhf(agf2_b200_ctx* ctx, const double* ixq_a, const double* qja_a, const double* qja_b,
                                      const double* ei_a, const double* ei_b, const double* ea_a, const double* ea_b,
                                      int64_t nphys, int64_t nocc_a, int64_t nocc_b, int64_t nvir_a, int64_t nvir_b,
                                      int64_t naux, int64_t istart, int64_t iend, double os_factor, double ss_factor,
                                      double* vv, double* vev)
{
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    return host_uhf(*c, ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b, nphys, nocc_a, nocc_b, nvir_a, nvir_b, naux, istart,
                    iend, os_factor, ss_factor, 0, 1, true, vv, vev);
}

// Page-lock / unlock a caller-owned host buffer so that the uploads of the host-pointer entry points are true
// asynchronous DMA transfers (optional; pageable memory works, staged by the driver).
int agf2_b200_host_register(const void* ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return fail(AGF2_B200_EINVAL, "bad argument");
    CU_TRY(cudaHostRegister(const_cast<void*>(ptr), (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterReadOnly));
    return 0;
}
int agf2_b200_host_unregister(const void* ptr) {
    if (!ptr) return fail(AGF2_B200_EINVAL, "bad argument");
    CU_TRY(cudaHostUnregister(const_cast<void*>(ptr)));
    return 0;
}

// C (a_rows x b_rows, row-major, or its transpose) = / += A (a_rows x k) . B (b_rows x k)^T on the library's own
// DMMA / TMA / stream-K pipeline; all pointers device memory, k contiguous in A and B.
int agf2_b200_ctx_gemm(agf2_b200_ctx* ctx, const double* a, int64_t a_rows, int64_t lda, const double* b, int64_t b_rows,
                       int64_t ldb, int64_t k, double* c_out, int64_t ldc, int transpose, int accumulate, void* stream)
{
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard guard(c->device);
    Gemm g;
    g.a = a; g.a_rows = a_rows; g.lda = lda;
    g.b = b; g.b_rows = b_rows; g.ldb = ldb;
    g.k = k; g.c = c_out; g.ldc = ldc;
    g.transpose = transpose != 0;
    g.accumulate = accumulate != 0;
    return run_gemm(*c, g, (cudaStream_t)stream);
}

// ---- drop-in symbols (reference signatures) ---------------------------------------------------------
void build_part_loop_rhf(double* ixq, double* qja, double* ei, double* ea, uint32_t nphys, uint32_t nocc,
                         uint32_t nvir, uint32_t naux, uint32_t istart, uint32_t iend, double* vv, double* vev)
{
    if (int s = agf2_b200_build_part_loop_rhf(ixq, qja, ei, ea, nphys, nocc, nvir, naux, istart, iend, 1.0, 1.0, vv, vev))
        drop_in_failed("build_part_loop_rhf", s);
}

void build_part_loop_uhf(double* ixq_a, double* qja_a, double* qja_b, double* ei_a, double* ei_b, double* ea_a,
                         double* ea_b, uint32_t nphys, uint32_t nocc_a, uint32_t nocc_b, uint32_t nvir_a,
                         uint32_t nvir_b, uint32_t naux, uint32_t istart, uint32_t iend, double* vv, double* vev)
{
    if (int s = agf2_b200_build_part_loop_uhf(ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b, nphys, nocc_a, nocc_b, nvir_a,
                                              nvir_b, naux, istart, iend, 1.0, 1.0, vv, vev))
        drop_in_failed("build_part_loop_uhf", s);
}

}  // extern "C"
