// Dense FP64 linear algebra of the pole compression and the extended-Fock eigensolve on the device
// (cuSOLVER / cuBLAS): replaces np.linalg.cholesky / inv / eigh of auxgf/agf2/optragf2.py:270-278 and
// np.linalg.eigh of auxgf/aux/eig.py:23-33.
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <string>

#include "../../include/agf2_b200.h"

namespace {
thread_local std::string s_err;
cusolverDnHandle_t s_solver = nullptr;
cublasHandle_t s_blas = nullptr;
cudaStream_t s_stream = nullptr;
double* s_buf = nullptr;
size_t s_buf_elems = 0;
int* s_info = nullptr;
double* s_work = nullptr;
int s_work_elems = 0;

int sfail(int code, const std::string& m) { s_err = m; return code; }
#define S_CU(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return sfail(AGF2_B200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)
#define S_SOL(expr) do { cusolverStatus_t _e = (expr); if (_e != CUSOLVER_STATUS_SUCCESS) return sfail(AGF2_B200_ESOLVER, std::string(#expr) + " -> cusolver status " + std::to_string((int)_e)); } while (0)
#define S_BLAS(expr) do { cublasStatus_t _e = (expr); if (_e != CUBLAS_STATUS_SUCCESS) return sfail(AGF2_B200_ESOLVER, std::string(#expr) + " -> cublas status " + std::to_string((int)_e)); } while (0)

int ensure() {
    if (s_solver) return 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return sfail(AGF2_B200_ENODEV, "no CUDA device visible");
    S_CU(cudaStreamCreateWithFlags(&s_stream, cudaStreamNonBlocking));
    S_SOL(cusolverDnCreate(&s_solver));
    S_SOL(cusolverDnSetStream(s_solver, s_stream));
    S_BLAS(cublasCreate(&s_blas));
    S_BLAS(cublasSetStream(s_blas, s_stream));
    S_CU(cudaMalloc(&s_info, sizeof(int)));
    return 0;
}
int need_buf(size_t elems) {
    if (elems <= s_buf_elems) return 0;
    if (s_buf) cudaFree(s_buf);
    s_buf = nullptr; s_buf_elems = 0;
    S_CU(cudaMalloc(&s_buf, elems * 8));
    s_buf_elems = elems;
    return 0;
}
int need_work(int elems) {
    if (elems <= s_work_elems) return 0;
    if (s_work) cudaFree(s_work);
    s_work = nullptr; s_work_elems = 0;
    S_CU(cudaMalloc(&s_work, (size_t)elems * 8));
    s_work_elems = elems;
    return 0;
}
}  // namespace

extern "C" const char* agf2_b200_solver_last_error(void) { return s_err.c_str(); }

extern "C" int agf2_b200_compress(const double* vv, const double* vev, int64_t nphys, double* e, double* c,
                                  int on_device)
{
    if (int r = ensure()) return r;
    if (!vv || !vev || !e || !c || nphys <= 0 || nphys > 46000) return sfail(AGF2_B200_EINVAL, "bad argument");
    const int n = (int)nphys;
    const size_t nn = (size_t)n * n;
    if (int r = need_buf(3 * nn + n)) return r;
    double* L = s_buf;            // vv -> Cholesky factor (column-major lower)
    double* M = s_buf + nn;       // vev -> L^-1 vev L^-T -> eigenvectors -> L U
    double* T = s_buf + 2 * nn;   // transposed output
    double* w = s_buf + 3 * nn;
    const cudaMemcpyKind in = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const cudaMemcpyKind out = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    S_CU(cudaMemcpyAsync(L, vv, nn * 8, in, s_stream));
    S_CU(cudaMemcpyAsync(M, vev, nn * 8, in, s_stream));
    int lwork1 = 0, lwork2 = 0;
    S_SOL(cusolverDnDpotrf_bufferSize(s_solver, CUBLAS_FILL_MODE_LOWER, n, L, n, &lwork1));
    S_SOL(cusolverDnDsyevd_bufferSize(s_solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, M, n, w, &lwork2));
    if (int r = need_work(lwork1 > lwork2 ? lwork1 : lwork2)) return r;
    // b = chol(vv)^T                                                  (optragf2.py:270)
    S_SOL(cusolverDnDpotrf(s_solver, CUBLAS_FILL_MODE_LOWER, n, L, n, s_work, lwork1, s_info));
    int info = 0;
    S_CU(cudaMemcpyAsync(&info, s_info, sizeof(int), cudaMemcpyDeviceToHost, s_stream));
    S_CU(cudaStreamSynchronize(s_stream));
    if (info != 0) return sfail(AGF2_B200_ENOTSPD, "vv is not positive definite (potrf info = " + std::to_string(info) + ")");
    // m = b^-T vev b^-1 = L^-1 vev L^-T  by two triangular solves (no explicit inverse; :271-273)
    const double one = 1.0, zero = 0.0;
    S_BLAS(cublasDtrsm(s_blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n, n, &one, L, n, M, n));
    S_BLAS(cublasDtrsm(s_blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, n, n, &one, L, n, M, n));
    // e, u = eigh(m)                                                   (:275)
    S_SOL(cusolverDnDsyevd(s_solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, M, n, w, s_work, lwork2, s_info));
    S_CU(cudaMemcpyAsync(&info, s_info, sizeof(int), cudaMemcpyDeviceToHost, s_stream));
    S_CU(cudaStreamSynchronize(s_stream));
    if (info != 0) return sfail(AGF2_B200_ESOLVER, "syevd info = " + std::to_string(info));
    // c = b^T u = L u                                                  (:276)
    S_BLAS(cublasDtrmm(s_blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n, n, &one, L, n, M, n, M, n));
    // column-major (x, k) -> row-major c[x][k]
    S_BLAS(cublasDgeam(s_blas, CUBLAS_OP_T, CUBLAS_OP_N, n, n, &one, M, n, &zero, M, n, T, n));
    S_CU(cudaMemcpyAsync(c, T, nn * 8, out, s_stream));
    S_CU(cudaMemcpyAsync(e, w, (size_t)n * 8, out, s_stream));
    S_CU(cudaStreamSynchronize(s_stream));
    return 0;
}

extern "C" int agf2_b200_eigh(const double* a, int64_t n64, double* w_out, double* v_out, int on_device)
{
    if (int r = ensure()) return r;
    if (!a || !w_out || !v_out || n64 <= 0 || n64 > 46000) return sfail(AGF2_B200_EINVAL, "bad argument");
    const int n = (int)n64;
    const size_t nn = (size_t)n * n;
    if (int r = need_buf(2 * nn + n)) return r;
    double* A = s_buf; double* T = s_buf + nn; double* w = s_buf + 2 * nn;
    const cudaMemcpyKind in = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const cudaMemcpyKind out = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    S_CU(cudaMemcpyAsync(A, a, nn * 8, in, s_stream));
    int lwork = 0;
    S_SOL(cusolverDnDsyevd_bufferSize(s_solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, w, &lwork));
    if (int r = need_work(lwork)) return r;
    S_SOL(cusolverDnDsyevd(s_solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, w, s_work, lwork, s_info));
    int info = 0;
    S_CU(cudaMemcpyAsync(&info, s_info, sizeof(int), cudaMemcpyDeviceToHost, s_stream));
    S_CU(cudaStreamSynchronize(s_stream));
    if (info != 0) return sfail(AGF2_B200_ESOLVER, "syevd info = " + std::to_string(info));
    const double one = 1.0, zero = 0.0;
    S_BLAS(cublasDgeam(s_blas, CUBLAS_OP_T, CUBLAS_OP_N, n, n, &one, A, n, &zero, A, n, T, n));
    S_CU(cudaMemcpyAsync(v_out, T, nn * 8, out, s_stream));
    S_CU(cudaMemcpyAsync(w_out, w, (size_t)n * 8, out, s_stream));
    S_CU(cudaStreamSynchronize(s_stream));
    return 0;
}
