// The communicator of a context: NCCL over NVLink / NVSwitch, one rank per context (= per GPU), loaded at run time
// with dlopen so that the library has no link-time NCCL dependency and shares the copy a host framework may already
// have loaded.  Replaces mpi4py's Reduce + Bcast of auxgf/util/mpi.py:25-69 (used by auxgf/mpi/agf2/optragf2.py:
// 394-395) with ONE all-reduce of the packed [vv | vev] per sector, plus the all-gather that replicates the DF
// blocks every rank uploaded 1/N of.  Part of agf2_b200.cu (included there).
namespace {

struct NcclId { char internal[128]; };

struct NcclApi {
    void* handle = nullptr;
    int (*GetVersion)(int*) = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string where;
};
constexpr int kNcclDouble = 8;                 // ncclFloat64
constexpr int kNcclSum = 0, kNcclMax = 2, kNcclMin = 3;

NcclApi g_nccl;
std::mutex g_nccl_mu;

int nccl_load() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.handle) return 0;
    std::vector<std::string> names;
    if (const char* e = getenv("AGF2_B200_NCCL_LIB")) names.push_back(e);
    names.push_back("libnccl.so.2");       // an already loaded copy (same soname) is reused
    names.push_back("libnccl.so");
    std::string tried;
    for (const std::string& n : names) {
        void* h = dlopen(n.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!h) { tried += n + " (" + (dlerror() ? "not loadable" : "?") + ") "; continue; }
        NcclApi a;
        a.handle = h;
        a.where = n;
#define AGF2_NCCL_SYM(field, sym) a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, sym))
        AGF2_NCCL_SYM(GetVersion, "ncclGetVersion");
        AGF2_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
        AGF2_NCCL_SYM(CommInitRank, "ncclCommInitRank");
        AGF2_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        AGF2_NCCL_SYM(AllReduce, "ncclAllReduce");
        AGF2_NCCL_SYM(AllGather, "ncclAllGather");
        AGF2_NCCL_SYM(Broadcast, "ncclBroadcast");
        AGF2_NCCL_SYM(Send, "ncclSend");
        AGF2_NCCL_SYM(Recv, "ncclRecv");
        AGF2_NCCL_SYM(GroupStart, "ncclGroupStart");
        AGF2_NCCL_SYM(GroupEnd, "ncclGroupEnd");
        AGF2_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef AGF2_NCCL_SYM
        if (a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.AllGather && a.Broadcast && a.Send &&
            a.Recv && a.GroupStart && a.GroupEnd && a.GetErrorString) {
            g_nccl = a;
            return 0;
        }
        dlclose(h);
        tried += n + " (symbols missing) ";
    }
    return fail(AGF2_B200_ECOMM, "NCCL library not found: " + tried + "-- set AGF2_B200_NCCL_LIB");
}

#define NCCL_TRY(expr)                                                                                   \
    do {                                                                                                 \
        int _r = (expr);                                                                                 \
        if (_r != 0)                                                                                     \
            return fail(AGF2_B200_ECOMM, std::string(#expr) + ": " + g_nccl.GetErrorString(_r));         \
    } while (0)

int comm_destroy(Ctx& c) {
    if (c.comm && g_nccl.handle) g_nccl.CommDestroy(c.comm);
    c.comm = nullptr;
    c.rank = 0;
    c.nranks = 1;
    return 0;
}

int comm_allreduce(Ctx& c, double* buf, int64_t count, int op, cudaStream_t st) {
    if (c.nranks <= 1) return 0;
    if (!c.comm) return fail(AGF2_B200_ECOMM, "context has no communicator");
    NCCL_TRY(g_nccl.AllReduce(buf, buf, (size_t)count, kNcclDouble, op, c.comm, st));
    // ring all-reduce: every rank sends 2 (N-1)/N of the buffer
    c.last_comm_bytes[1] += (int64_t)(2.0 * (c.nranks - 1) / c.nranks * count * 8);
    return 0;
}
int comm_allreduce_sum(Ctx& c, double* buf, int64_t count, cudaStream_t st) {
    return comm_allreduce(c, buf, count, kNcclSum, st);
}

// in place: rank r's piece lives at buf + r * count_per_rank
int comm_allgather(Ctx& c, double* buf, int64_t count_per_rank, cudaStream_t st) {
    if (c.nranks <= 1) return 0;
    if (!c.comm) return fail(AGF2_B200_ECOMM, "context has no communicator");
    NCCL_TRY(g_nccl.AllGather(buf + (size_t)c.rank * count_per_rank, buf, (size_t)count_per_rank, kNcclDouble, c.comm, st));
    c.last_comm_bytes[0] += (int64_t)(c.nranks - 1) * count_per_rank * 8;
    return 0;
}

// One step of the pole-block exchange: send `send_count` doubles to rank dst and receive `recv_count` from rank src
// (either may be empty), as one NCCL group on `st`.
int comm_sendrecv(Ctx& c, const double* sendbuf, int64_t send_count, int dst, double* recvbuf, int64_t recv_count, int src,
                  cudaStream_t st) {
    if (send_count <= 0 && recv_count <= 0) return 0;
    if (!c.comm) return fail(AGF2_B200_ECOMM, "context has no communicator");
    NCCL_TRY(g_nccl.GroupStart());
    int r1 = 0, r2 = 0;
    if (send_count > 0) r1 = g_nccl.Send(sendbuf, (size_t)send_count, kNcclDouble, dst, c.comm, st);
    if (recv_count > 0) r2 = g_nccl.Recv(recvbuf, (size_t)recv_count, kNcclDouble, src, c.comm, st);
    NCCL_TRY(g_nccl.GroupEnd());
    NCCL_TRY(r1);
    NCCL_TRY(r2);
    if (send_count > 0) c.last_comm_bytes[2] += send_count * 8;
    return 0;
}

Ctx* resolve_ctx(agf2_b200_ctx* ctx, int* status) {
    *status = 0;
    if (ctx) return ctx;
    Ctx* c = nullptr;
    *status = ctx_for_current_device(&c);
    return c;
}

}  // namespace

extern "C" {

int agf2_b200_comm_get_unique_id(void* id128) {
    if (!id128) return fail(AGF2_B200_EINVAL, "null pointer");
    if (int r = nccl_load()) return r;
    NcclId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int agf2_b200_comm_init(agf2_b200_ctx* ctx, const void* id128, int rank, int nranks) {
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(AGF2_B200_EINVAL, "bad rank / nranks / id");
    if (int r = nccl_load()) return r;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard guard(c->device);
    comm_destroy(*c);
    NcclId id;
    memcpy(&id, id128, sizeof(id));
    void* comm = nullptr;
    NCCL_TRY(g_nccl.CommInitRank(&comm, nranks, id, rank));
    c->comm = comm;
    c->rank = rank;
    c->nranks = nranks;
    return 0;
}

int agf2_b200_comm_info(agf2_b200_ctx* ctx, int* rank, int* nranks) {
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    return 0;
}

int agf2_b200_comm_allreduce(agf2_b200_ctx* ctx, double* buf, int64_t count, int op, void* stream) {
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    if (!buf || count < 0 || op < 0 || op > 2) return fail(AGF2_B200_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard guard(c->device);
    return comm_allreduce(*c, buf, count, op == 0 ? kNcclSum : (op == 1 ? kNcclMax : kNcclMin), (cudaStream_t)stream);
}

int agf2_b200_comm_allgather(agf2_b200_ctx* ctx, double* buf, int64_t count_per_rank, void* stream) {
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    if (!buf || count_per_rank < 0) return fail(AGF2_B200_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard guard(c->device);
    return comm_allgather(*c, buf, count_per_rank, (cudaStream_t)stream);
}

int agf2_b200_comm_broadcast(agf2_b200_ctx* ctx, double* buf, int64_t count, int root, void* stream) {
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    if (c->nranks <= 1) return 0;
    if (!buf || count < 0 || root < 0 || root >= c->nranks) return fail(AGF2_B200_EINVAL, "bad argument");
    if (!c->comm) return fail(AGF2_B200_ECOMM, "context has no communicator");
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard guard(c->device);
    NCCL_TRY(g_nccl.Broadcast(buf, buf, (size_t)count, kNcclDouble, root, c->comm, (cudaStream_t)stream));
    return 0;
}

// All ranks have reached this point and the work queued on the context's streams has finished.
int agf2_b200_comm_barrier(agf2_b200_ctx* ctx) {
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard guard(c->device);
    if (c->nranks > 1) {
        if (int r = grow(c->packed, 2, false)) return r;
        CU_TRY(cudaMemsetAsync(c->packed.p, 0, 16, c->stream));
        if (int r = comm_allreduce(*c, c->packed.p, 1, kNcclSum, c->stream)) return r;
    }
    CU_TRY(cudaDeviceSynchronize());
    return 0;
}

int agf2_b200_comm_destroy(agf2_b200_ctx* ctx) {
    int s;
    Ctx* c = resolve_ctx(ctx, &s);
    if (!c) return s;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard guard(c->device);
    cudaDeviceSynchronize();
    return comm_destroy(*c);
}

}  // extern "C"
