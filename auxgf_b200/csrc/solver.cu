// Dense FP64 linear algebra of the pole compression and the extended-Fock eigensolve on the device
// (cuSOLVER / cuBLAS): replaces np.linalg.cholesky / inv / eigh of auxgf/agf2/optragf2.py:270-278 and
// np.linalg.eigh of auxgf/aux/eig.py:23-33.
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "context.h"

namespace {
thread_local std::string s_err;
// cuSOLVER / cuBLAS handles, stream and scratch of agf2_b200_compress / agf2_b200_eigh: one set per device (created on
// first use for the calling thread's current device); the two entry points are serialised by s_mu
struct SolverState {
    cusolverDnHandle_t solver = nullptr;
    cublasHandle_t blas = nullptr;
    cudaStream_t stream = nullptr;
    double* buf = nullptr;
    size_t buf_elems = 0;
    int* info = nullptr;
    double* work = nullptr;
    int work_elems = 0;
};
std::mutex s_mu;
std::map<int, SolverState> s_states;
SolverState* S = nullptr;      // state of the call in progress (valid while s_mu is held)

int sfail(int code, const std::string& m) { s_err = m; agf2::last_error() = m; return code; }
#define S_CU(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return sfail(AGF2_B200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)
#define S_SOL(expr) do { cusolverStatus_t _e = (expr); if (_e != CUSOLVER_STATUS_SUCCESS) return sfail(AGF2_B200_ESOLVER, std::string(#expr) + " -> cusolver status " + std::to_string((int)_e)); } while (0)
#define S_BLAS(expr) do { cublasStatus_t _e = (expr); if (_e != CUBLAS_STATUS_SUCCESS) return sfail(AGF2_B200_ESOLVER, std::string(#expr) + " -> cublas status " + std::to_string((int)_e)); } while (0)

int ensure() {
    int ndev = 0, device = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return sfail(AGF2_B200_ENODEV, "no CUDA device visible");
    S_CU(cudaGetDevice(&device));
    S = &s_states[device];
    if (S->solver) return 0;
    S_CU(cudaStreamCreateWithFlags(&S->stream, cudaStreamNonBlocking));
    S_SOL(cusolverDnCreate(&S->solver));
    S_SOL(cusolverDnSetStream(S->solver, S->stream));
    S_BLAS(cublasCreate(&S->blas));
    S_BLAS(cublasSetStream(S->blas, S->stream));
    S_CU(cudaMalloc(&S->info, sizeof(int)));
    return 0;
}
int need_buf(size_t elems) {
    if (elems <= S->buf_elems) return 0;
    if (S->buf) cudaFree(S->buf);
    S->buf = nullptr; S->buf_elems = 0;
    S_CU(cudaMalloc(&S->buf, elems * 8));
    S->buf_elems = elems;
    return 0;
}
int need_work(int elems) {
    if (elems <= S->work_elems) return 0;
    if (S->work) cudaFree(S->work);
    S->work = nullptr; S->work_elems = 0;
    S_CU(cudaMalloc(&S->work, (size_t)elems * 8));
    S->work_elems = elems;
    return 0;
}
}  // namespace

extern "C" const char* agf2_b200_solver_last_error(void) { return s_err.c_str(); }

extern "C" int agf2_b200_compress(const double* vv, const double* vev, int64_t nphys, double* e, double* c,
                                  int on_device)
{
    std::lock_guard<std::mutex> lk(s_mu);
    if (int r = ensure()) return r;
    if (!vv || !vev || !e || !c || nphys <= 0 || nphys > 46000) return sfail(AGF2_B200_EINVAL, "bad argument");
    const int n = (int)nphys;
    const size_t nn = (size_t)n * n;
    if (int r = need_buf(3 * nn + n)) return r;
    double* L = S->buf;            // vv -> Cholesky factor (column-major lower)
    double* M = S->buf + nn;       // vev -> L^-1 vev L^-T -> eigenvectors -> L U
    double* T = S->buf + 2 * nn;   // transposed output
    double* w = S->buf + 3 * nn;
    const cudaMemcpyKind in = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const cudaMemcpyKind out = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    S_CU(cudaMemcpyAsync(L, vv, nn * 8, in, S->stream));
    S_CU(cudaMemcpyAsync(M, vev, nn * 8, in, S->stream));
    int lwork1 = 0, lwork2 = 0;
    S_SOL(cusolverDnDpotrf_bufferSize(S->solver, CUBLAS_FILL_MODE_LOWER, n, L, n, &lwork1));
    S_SOL(cusolverDnDsyevd_bufferSize(S->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, M, n, w, &lwork2));
    if (int r = need_work(lwork1 > lwork2 ? lwork1 : lwork2)) return r;
    // b = chol(vv)^T                                                  (optragf2.py:270)
    S_SOL(cusolverDnDpotrf(S->solver, CUBLAS_FILL_MODE_LOWER, n, L, n, S->work, lwork1, S->info));
    int info = 0;
    S_CU(cudaMemcpyAsync(&info, S->info, sizeof(int), cudaMemcpyDeviceToHost, S->stream));
    S_CU(cudaStreamSynchronize(S->stream));
    if (info != 0) return sfail(AGF2_B200_ENOTSPD, "vv is not positive definite (potrf info = " + std::to_string(info) + ")");
    // m = b^-T vev b^-1 = L^-1 vev L^-T  by two triangular solves (no explicit inverse; :271-273)
    const double one = 1.0, zero = 0.0;
    S_BLAS(cublasDtrsm(S->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n, n, &one, L, n, M, n));
    S_BLAS(cublasDtrsm(S->blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, n, n, &one, L, n, M, n));
    // e, u = eigh(m)                                                   (:275)
    S_SOL(cusolverDnDsyevd(S->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, M, n, w, S->work, lwork2, S->info));
    S_CU(cudaMemcpyAsync(&info, S->info, sizeof(int), cudaMemcpyDeviceToHost, S->stream));
    S_CU(cudaStreamSynchronize(S->stream));
    if (info != 0) return sfail(AGF2_B200_ESOLVER, "syevd info = " + std::to_string(info));
    // c = b^T u = L u                                                  (:276)
    S_BLAS(cublasDtrmm(S->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n, n, &one, L, n, M, n, M, n));
    // column-major (x, k) -> row-major c[x][k]
    S_BLAS(cublasDgeam(S->blas, CUBLAS_OP_T, CUBLAS_OP_N, n, n, &one, M, n, &zero, M, n, T, n));
    S_CU(cudaMemcpyAsync(c, T, nn * 8, out, S->stream));
    S_CU(cudaMemcpyAsync(e, w, (size_t)n * 8, out, S->stream));
    S_CU(cudaStreamSynchronize(S->stream));
    return 0;
}

extern "C" int agf2_b200_eigh(const double* a, int64_t n64, double* w_out, double* v_out, int on_device)
{
    std::lock_guard<std::mutex> lk(s_mu);
    if (int r = ensure()) return r;
    if (!a || !w_out || !v_out || n64 <= 0 || n64 > 46000) return sfail(AGF2_B200_EINVAL, "bad argument");
    const int n = (int)n64;
    const size_t nn = (size_t)n * n;
    if (int r = need_buf(2 * nn + n)) return r;
    double* A = S->buf; double* T = S->buf + nn; double* w = S->buf + 2 * nn;
    const cudaMemcpyKind in = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const cudaMemcpyKind out = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    S_CU(cudaMemcpyAsync(A, a, nn * 8, in, S->stream));
    int lwork = 0;
    S_SOL(cusolverDnDsyevd_bufferSize(S->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, w, &lwork));
    if (int r = need_work(lwork)) return r;
    S_SOL(cusolverDnDsyevd(S->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, w, S->work, lwork, S->info));
    int info = 0;
    S_CU(cudaMemcpyAsync(&info, S->info, sizeof(int), cudaMemcpyDeviceToHost, S->stream));
    S_CU(cudaStreamSynchronize(S->stream));
    if (info != 0) return sfail(AGF2_B200_ESOLVER, "syevd info = " + std::to_string(info));
    const double one = 1.0, zero = 0.0;
    S_BLAS(cublasDgeam(S->blas, CUBLAS_OP_T, CUBLAS_OP_N, n, n, &one, A, n, &zero, A, n, T, n));
    S_CU(cudaMemcpyAsync(v_out, T, nn * 8, out, S->stream));
    S_CU(cudaMemcpyAsync(w_out, w, (size_t)n * 8, out, S->stream));
    S_CU(cudaStreamSynchronize(S->stream));
    return 0;
}

// =================================================================================================
// Extended-Fock eigensolver with resident poles (the inner loop of auxgf/agf2/fock.py:119-143 and
// auxgf/agf2/chempot.py:11-49, which diagonalise  [[F, v], [v^T, diag(e - x)]]  hundreds of times per AGF2
// iteration with the same couplings v): v stays in HBM, the extended matrix is assembled on the device, cuSOLVER
// dsyevd runs in a reused workspace and only the eigenvalues and the PHYSICAL rows of the eigenvectors -- all the
// Fock loop, the density and the chemical-potential gradient need -- travel back through pinned memory.
// =================================================================================================
namespace {

__global__ void build_ext_kernel(double* __restrict__ h, const double* __restrict__ fock, const double* __restrict__ v,
                                 const double* __restrict__ e, double shift, int nphys, int naux)
{
    const int n = nphys + naux;
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= n) return;
    double x = 0.0;
    if (r < nphys && c < nphys) x = fock[(long long)r * nphys + c];
    else if (r < nphys) x = v[(long long)r * naux + (c - nphys)];
    else if (c < nphys) x = v[(long long)c * naux + (r - nphys)];
    else if (r == c) x = e[r - nphys] - shift;
    h[(long long)r * n + c] = x;
}

// out[x, j] = V[x + j n]  (physical rows of the column-major eigenvector matrix, row-major (nphys, n))
__global__ void phys_rows_kernel(const double* __restrict__ vec, double* __restrict__ out, int nphys, int n)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, x = blockIdx.y;
    if (j < n) out[(long long)x * n + j] = vec[(long long)j * n + x];
}

// e2b = sign * sum_{l,k} ovl[l,k]^2 / (ge[l] - se[k]): one partial per block, summed in block order on the host side
__global__ void e2b_reduce_kernel(const double* __restrict__ ovl, long long ld, const double* __restrict__ ge,
                                  const double* __restrict__ se, int nl, int nk, double* __restrict__ partial)
{
    const int l = blockIdx.x;
    double s = 0.0;
    for (int k = threadIdx.x; k < nk; k += blockDim.x) {
        const double o = ovl[(long long)l * ld + k];
        s += o * o / (ge[l] - se[k]);
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[l] = sh[0];
}

}  // namespace

struct agf2_b200_fock {
    int device = 0;
    cudaStream_t stream = nullptr;
    cusolverDnHandle_t solver = nullptr;
    int64_t nphys = 0, naux = 0;
    double* d_v = nullptr;        // couplings (nphys, naux)
    double* d_in = nullptr;       // [fock (nphys^2) | e (naux)]
    double* d_h = nullptr;        // extended matrix / eigenvectors (n^2)
    double* d_w = nullptr;        // [w (n) | v_phys (nphys n)]
    double* d_work = nullptr;
    int* d_info = nullptr;
    int lwork = 0;
    double* h_in = nullptr;       // pinned mirrors
    double* h_out = nullptr;
    size_t cap_n = 0, cap_v = 0;
    double* user_w = nullptr;     // destination of the solve in flight
    double* user_v = nullptr;
};

static void fock_free_buffers(agf2_b200_fock* h) {
    for (double** p : {&h->d_v, &h->d_in, &h->d_h, &h->d_w, &h->d_work}) { if (*p) cudaFree(*p); *p = nullptr; }
    for (double** p : {&h->h_in, &h->h_out}) { if (*p) cudaFreeHost(*p); *p = nullptr; }
    h->cap_n = h->cap_v = 0;
    h->lwork = 0;
}

extern "C" int agf2_b200_fock_create(agf2_b200_fock** out)
{
    if (!out) return sfail(AGF2_B200_EINVAL, "null pointer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return sfail(AGF2_B200_ENODEV, "no CUDA device visible");
    agf2_b200_fock* h = new agf2_b200_fock();
    S_CU(cudaGetDevice(&h->device));
    S_CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    S_SOL(cusolverDnCreate(&h->solver));
    S_SOL(cusolverDnSetStream(h->solver, h->stream));
    S_CU(cudaMalloc(&h->d_info, sizeof(int)));
    *out = h;
    return 0;
}

extern "C" int agf2_b200_fock_destroy(agf2_b200_fock* h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    fock_free_buffers(h);
    if (h->d_info) cudaFree(h->d_info);
    if (h->solver) cusolverDnDestroy(h->solver);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

/* v: host (nphys, naux) row-major.  Uploaded once; every solve that follows uses it. */
extern "C" int agf2_b200_fock_set_couplings(agf2_b200_fock* h, const double* v, int64_t nphys, int64_t naux)
{
    if (!h || nphys <= 0 || naux < 0 || (naux > 0 && !v) || nphys + naux > 46000) return sfail(AGF2_B200_EINVAL, "bad argument");
    S_CU(cudaSetDevice(h->device));
    S_CU(cudaStreamSynchronize(h->stream));
    const size_t n = (size_t)(nphys + naux), nv = (size_t)std::max<int64_t>(nphys * naux, 1);
    if (n > h->cap_n || nv > h->cap_v) {
        fock_free_buffers(h);
        S_CU(cudaMalloc(&h->d_v, nv * 8));
        S_CU(cudaMalloc(&h->d_in, ((size_t)nphys * nphys + n) * 8));
        S_CU(cudaMalloc(&h->d_h, n * n * 8));
        S_CU(cudaMalloc(&h->d_w, (n + (size_t)nphys * n) * 8));
        S_CU(cudaMallocHost(&h->h_in, ((size_t)nphys * nphys + n) * 8));
        S_CU(cudaMallocHost(&h->h_out, (n + (size_t)nphys * n) * 8));
        int lwork = 0;
        S_SOL(cusolverDnDsyevd_bufferSize(h->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, h->d_h, (int)n,
                                          h->d_w, &lwork));
        S_CU(cudaMalloc(&h->d_work, (size_t)lwork * 8));
        h->lwork = lwork;
        h->cap_n = n; h->cap_v = nv;
    } else if (n != (size_t)(h->nphys + h->naux)) {
        int lwork = 0;
        S_SOL(cusolverDnDsyevd_bufferSize(h->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, h->d_h, (int)n,
                                          h->d_w, &lwork));
        if (lwork > h->lwork) {
            cudaFree(h->d_work); h->d_work = nullptr;
            S_CU(cudaMalloc(&h->d_work, (size_t)lwork * 8));
            h->lwork = lwork;
        }
    }
    h->nphys = nphys; h->naux = naux;
    if (naux > 0) S_CU(cudaMemcpyAsync(h->d_v, v, (size_t)nphys * naux * 8, cudaMemcpyHostToDevice, h->stream));
    S_CU(cudaStreamSynchronize(h->stream));
    return 0;
}

/* Queue one eigensolve of [[fock, v], [v^T, diag(e - shift)]]: fock (nphys, nphys) and e (naux) are host arrays (copied
 * before the call returns); results land in w (nphys + naux) and v_phys (nphys, nphys + naux; row x = component x of
 * every eigenvector) when agf2_b200_fock_wait returns.  Solves on different handles overlap on the device. */
extern "C" int agf2_b200_fock_solve_async(agf2_b200_fock* h, const double* fock, const double* e, double shift, double* w,
                                          double* v_phys)
{
    if (!h || !fock || !w || !v_phys || h->nphys <= 0 || (h->naux > 0 && !e)) return sfail(AGF2_B200_EINVAL, "bad argument");
    if (h->user_w) return sfail(AGF2_B200_EINVAL, "a solve is already in flight on this handle");
    S_CU(cudaSetDevice(h->device));
    const int P = (int)h->nphys, A = (int)h->naux, n = P + A;
    const size_t nin = (size_t)P * P + A;
    memcpy(h->h_in, fock, (size_t)P * P * 8);
    if (A) memcpy(h->h_in + (size_t)P * P, e, (size_t)A * 8);
    S_CU(cudaMemcpyAsync(h->d_in, h->h_in, nin * 8, cudaMemcpyHostToDevice, h->stream));
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)n);
    build_ext_kernel<<<grid, 128, 0, h->stream>>>(h->d_h, h->d_in, h->d_v, h->d_in + (size_t)P * P, shift, P, A);
    S_CU(cudaGetLastError());
    S_SOL(cusolverDnDsyevd(h->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, h->d_h, n, h->d_w, h->d_work, h->lwork,
                           h->d_info));
    dim3 g2((unsigned)((n + 127) / 128), (unsigned)P);
    phys_rows_kernel<<<g2, 128, 0, h->stream>>>(h->d_h, h->d_w + n, P, n);
    S_CU(cudaGetLastError());
    S_CU(cudaMemcpyAsync(h->h_out, h->d_w, ((size_t)n + (size_t)P * n) * 8, cudaMemcpyDeviceToHost, h->stream));
    h->user_w = w; h->user_v = v_phys;
    return 0;
}

extern "C" int agf2_b200_fock_wait(agf2_b200_fock* h)
{
    if (!h || !h->user_w) return sfail(AGF2_B200_EINVAL, "no solve in flight");
    S_CU(cudaSetDevice(h->device));
    int info = 0;
    S_CU(cudaMemcpyAsync(&info, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    double* w = h->user_w; double* v = h->user_v;
    h->user_w = h->user_v = nullptr;
    S_CU(cudaStreamSynchronize(h->stream));
    if (info != 0) return sfail(AGF2_B200_ESOLVER, "syevd info = " + std::to_string(info));
    const size_t n = (size_t)(h->nphys + h->naux);
    memcpy(w, h->h_out, n * 8);
    memcpy(v, h->h_out + n, (size_t)h->nphys * n * 8);
    return 0;
}

extern "C" int agf2_b200_fock_solve(agf2_b200_fock* h, const double* fock, const double* e, double shift, double* w, double* v_phys)
{
    if (int r = agf2_b200_fock_solve_async(h, fock, e, shift, w, v_phys)) return r;
    return agf2_b200_fock_wait(h);
}


extern "C" int agf2_b200_energy_2body(const double* gv, const double* ge, int64_t nl, const double* sv, const double* se,
                                      int64_t nk, int64_t nphys, double* out)
{
    if (!out || nl < 0 || nk < 0 || nphys <= 0) return sfail(AGF2_B200_EINVAL, "bad argument");
    *out = 0.0;
    if (nl == 0 || nk == 0) return 0;
    if (!gv || !ge || !sv || !se) return sfail(AGF2_B200_EINVAL, "null pointer");
    agf2::Ctx* cp = nullptr;
    if (int r = agf2::ctx_for_current_device(&cp)) { s_err = agf2::last_error(); return r; }
    agf2::Ctx& c = *cp;
    std::lock_guard<std::mutex> lk(c.mu);
    cudaStream_t st = c.stream;
    const int64_t ldx = (nphys + 1) / 2 * 2, ldo = (nk + 1) / 2 * 2;
    const size_t need = (size_t)(nl + nk) * ldx + (size_t)nl * ldo + (size_t)(2 * nl + nk) + 8;
    if (int r = agf2::grow(c.stage[7], need, false)) { s_err = agf2::last_error(); return r; }
    double* d_gv = c.stage[7].p; double* d_sv = d_gv + nl * ldx; double* d_ovl = d_sv + nk * ldx;
    double* d_ge = d_ovl + nl * ldo; double* d_se = d_ge + nl; double* d_part = d_se + nk;
    S_CU(cudaMemcpy2DAsync(d_gv, ldx * 8, gv, nphys * 8, nphys * 8, nl, cudaMemcpyHostToDevice, st));
    S_CU(cudaMemcpy2DAsync(d_sv, ldx * 8, sv, nphys * 8, nphys * 8, nk, cudaMemcpyHostToDevice, st));
    S_CU(cudaMemcpyAsync(d_ge, ge, nl * 8, cudaMemcpyHostToDevice, st));
    S_CU(cudaMemcpyAsync(d_se, se, nk * 8, cudaMemcpyHostToDevice, st));
    agf2::Gemm g;                                    // ovl[l, k] = sum_x gv[l, x] sv[k, x]
    g.a = d_gv; g.a_rows = nl; g.lda = ldx;
    g.b = d_sv; g.b_rows = nk; g.ldb = ldx;
    g.k = nphys; g.c = d_ovl; g.ldc = ldo;
    if (int r = agf2::run_gemm(c, g, st)) { s_err = agf2::last_error(); return r; }
    e2b_reduce_kernel<<<(unsigned)nl, 256, 0, st>>>(d_ovl, ldo, d_ge, d_se, (int)nl, (int)nk, d_part);
    S_CU(cudaGetLastError());
    std::vector<double> part((size_t)nl);
    S_CU(cudaMemcpyAsync(part.data(), d_part, nl * 8, cudaMemcpyDeviceToHost, st));
    S_CU(cudaStreamSynchronize(st));
    double s = 0.0;
    for (double x : part) s += x;                    // fixed order
    *out = s;
    return 0;
}
