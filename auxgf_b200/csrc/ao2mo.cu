// Device-resident DF tensor and the per-sector DF transform ("ao2mo") on the GPU.
//
// Replaces, for the moment build, the two `OptRAGF2.ao2mo` calls of auxgf/agf2/optragf2.py:243-244
// (PySCF `_ao2mo.nr_e2`, 's2' packed input -> 's1' unpacked output, optragf2.py:175-214):
//     ixq[i, x, Q] = sum_p   L[Q, x, p] C_occ[p, i]
//     qja[Q, j, a] = sum_x   ixq[j, x, Q] C_vir[x, a]          (= sum_pq L[Q,p,q] C_occ[p,j] C_vir[q,a])
// L is unpacked once to (naux, nphys, nphys) in HBM ("staged once to HBM") and both transforms are plain
// strided-batched library DGEMMs (cuBLAS) writing straight into the layouts the moment kernels consume.
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "../../include/agf2_b200.h"

namespace {
thread_local std::string a_err;
int afail(int code, const std::string& m) { a_err = m; return code; }
#define A_CU(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return afail(_e == cudaErrorMemoryAllocation ? AGF2_B200_ENOMEM : AGF2_B200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)
#define A_BLAS(expr) do { cublasStatus_t _e = (expr); if (_e != CUBLAS_STATUS_SUCCESS) return afail(AGF2_B200_ESOLVER, std::string(#expr) + " -> cublas status " + std::to_string((int)_e)); } while (0)

cublasHandle_t a_blas = nullptr;

// Lp[q][Q][s] = L[Q0 + Q][q][s] for one block of auxiliary functions (p-major copy for the exchange GEMM)
__global__ void permute_block_kernel(const double* __restrict__ full, double* __restrict__ lp, int nphys, int nq)
{
    const int Q = blockIdx.z, q = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nphys) return;
    lp[((long long)q * nq + Q) * nphys + s] = full[((long long)Q * nphys + q) * nphys + s];
}

// packed lower triangle (row p, col q <= p at p(p+1)/2 + q; pyscf.lib.pack_tril) -> full symmetric
__global__ void unpack_tril_kernel(const double* __restrict__ packed, double* __restrict__ full, int nphys, long long npair)
{
    const long long q = blockIdx.z;          // auxiliary index
    const int p = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nphys) return;
    const int hi = p > c ? p : c, lo = p > c ? c : p;
    full[(q * nphys + p) * nphys + c] = packed[q * npair + (long long)hi * (hi + 1) / 2 + lo];
}
}  // namespace

struct agf2_b200_eri {
    int64_t naux = 0, nphys = 0;
    double* full = nullptr;           // (naux, nphys, nphys)
    double* ixq = nullptr;            // scratch (nocc, nphys, ld_ixq)
    size_t ixq_elems = 0;
    double* qja = nullptr;            // scratch (naux, nocc, ld_qja)
    size_t qja_elems = 0;
    double* coef = nullptr;           // staged coefficient matrices
    size_t coef_elems = 0;
    double* jk = nullptr;             // get_jk scratch: D, rho, J, K, T', L' blocks
    size_t jk_elems = 0;
};

extern "C" const char* agf2_b200_ao2mo_last_error(void) { return a_err.c_str(); }

extern "C" int agf2_b200_eri_create(const double* eri, int64_t naux, int64_t nphys, int packed, int on_device,
                                    agf2_b200_eri** out)
{
    if (!eri || !out || naux <= 0 || nphys <= 0 || nphys > 65535) return afail(AGF2_B200_EINVAL, "bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return afail(AGF2_B200_ENODEV, "no CUDA device visible");
    if (!a_blas) A_BLAS(cublasCreate(&a_blas));
    agf2_b200_eri* h = new agf2_b200_eri();
    h->naux = naux; h->nphys = nphys;
    const size_t nfull = (size_t)naux * nphys * nphys;
    const long long npair = (long long)nphys * (nphys + 1) / 2;
    cudaError_t e = cudaMalloc(&h->full, nfull * 8);
    if (e != cudaSuccess) { delete h; return afail(AGF2_B200_ENOMEM, std::string("cudaMalloc(full DF tensor): ") + cudaGetErrorString(e)); }
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (!packed) {
        e = cudaMemcpy(h->full, eri, nfull * 8, kind);
        if (e != cudaSuccess) { cudaFree(h->full); delete h; return afail(AGF2_B200_ECUDA, cudaGetErrorString(e)); }
    } else {
        // stream the packed tensor through a bounded staging buffer, unpacking block by block
        const int64_t blk = std::max<int64_t>(1, std::min<int64_t>(naux, (int64_t)((256ll << 20) / (npair * 8))));
        double* stage = nullptr;
        e = cudaMalloc(&stage, (size_t)blk * npair * 8);
        if (e != cudaSuccess) { cudaFree(h->full); delete h; return afail(AGF2_B200_ENOMEM, cudaGetErrorString(e)); }
        for (int64_t q0 = 0; q0 < naux; q0 += blk) {
            const int64_t nq = std::min(blk, naux - q0);
            e = cudaMemcpy(stage, eri + q0 * npair, (size_t)nq * npair * 8, kind);
            if (e == cudaSuccess) {
                dim3 grid((unsigned)((nphys + 127) / 128), (unsigned)nphys, (unsigned)nq);
                unpack_tril_kernel<<<grid, 128>>>(stage, h->full + (size_t)q0 * nphys * nphys, (int)nphys, npair);
                e = cudaDeviceSynchronize();
            }
            if (e != cudaSuccess) { cudaFree(stage); cudaFree(h->full); delete h; return afail(AGF2_B200_ECUDA, cudaGetErrorString(e)); }
        }
        cudaFree(stage);
    }
    *out = h;
    return 0;
}

extern "C" int agf2_b200_eri_destroy(agf2_b200_eri* h)
{
    if (!h) return 0;
    cudaFree(h->full); cudaFree(h->ixq); cudaFree(h->qja); cudaFree(h->coef); cudaFree(h->jk);
    delete h;
    return 0;
}

extern "C" int agf2_b200_eri_full_ptr(agf2_b200_eri* h, double** full)
{
    if (!h || !full) return afail(AGF2_B200_EINVAL, "null handle");
    *full = h->full;
    return 0;
}

template <typename T>
static int agrow(T*& p, size_t& cap, size_t need)
{
    if (need <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    A_CU(cudaMalloc(&p, need * sizeof(T)));
    cap = need;
    return 0;
}

extern "C" int agf2_b200_ao2mo_sector(agf2_b200_eri* h, const double* v_occ, int64_t nocc, const double* v_vir,
                                      int64_t nvir, int on_device, void* stream, double** ixq_out, int64_t* ld_ixq_out,
                                      double** qja_out, int64_t* ld_qja_out)
{
    if (!h || !v_occ || !v_vir || nocc <= 0 || nvir <= 0 || !ixq_out || !qja_out || !ld_ixq_out || !ld_qja_out)
        return afail(AGF2_B200_EINVAL, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t P = h->nphys, N = h->naux;
    const int64_t ldn = (N + 1) / 2 * 2, ldv = (nvir + 1) / 2 * 2;
    if (int r = agrow(h->ixq, h->ixq_elems, (size_t)nocc * P * ldn)) return r;
    if (int r = agrow(h->qja, h->qja_elems, (size_t)N * nocc * ldv)) return r;
    if (int r = agrow(h->coef, h->coef_elems, (size_t)P * (nocc + nvir))) return r;
    double* co = h->coef;
    double* cv = h->coef + P * nocc;
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    A_CU(cudaMemcpyAsync(co, v_occ, (size_t)P * nocc * 8, kind, st));
    A_CU(cudaMemcpyAsync(cv, v_vir, (size_t)P * nvir * 8, kind, st));
    if (ldn != N) A_CU(cudaMemsetAsync(h->ixq, 0, (size_t)nocc * P * ldn * 8, st));
    if (ldv != nvir) A_CU(cudaMemsetAsync(h->qja, 0, (size_t)N * nocc * ldv * 8, st));
    A_BLAS(cublasSetStream(a_blas, st));
    const double one = 1.0, zero = 0.0;
    // (1) for every physical index x:  ixq[:, x, :]^T (naux x nocc, ld = nphys*ldn) = L[:, x, :] (naux x nphys) . C_occ
    A_BLAS(cublasDgemmStridedBatched(a_blas, CUBLAS_OP_T, CUBLAS_OP_T, (int)N, (int)nocc, (int)P, &one,
                                     h->full, (int)(P * P), P,            /* A_x(p, Q) = L[Q, x, p], lda = P^2 */
                                     co, (int)nocc, 0,                     /* C_occ row-major (P, nocc)          */
                                     &zero, h->ixq, (int)(P * ldn), ldn, (int)P));
    // (2) for every occupied pole j:  qja[:, j, :]^T (nvir x naux, ld = nocc*ldv) = C_vir^T (nvir x nphys) . ixq[j]^T
    A_BLAS(cublasDgemmStridedBatched(a_blas, CUBLAS_OP_N, CUBLAS_OP_T, (int)nvir, (int)N, (int)P, &one,
                                     cv, (int)nvir, 0,                     /* C_vir row-major (P, nvir)          */
                                     h->ixq, (int)ldn, P * ldn,            /* ixq[j](x, Q), Q fastest             */
                                     &zero, h->qja, (int)(nocc * ldv), ldv, (int)nocc));
    *ixq_out = h->ixq; *ld_ixq_out = ldn;
    *qja_out = h->qja; *ld_qja_out = ldv;
    return 0;
}


// Coulomb and exchange matrices from the resident DF tensor (replaces the blocked host loop of
// auxgf/agf2/optragf2.py:358-398 / optuagf2.py:317-366):
//     J[p,q] = sum_Q L[Q,p,q] (sum_rs L[Q,r,s] D[r,s]),     K[p,q] = sum_Q sum_rs L[Q,p,r] D[r,s] L[Q,q,s]
// rdm1, j, k: (nphys, nphys) row-major; host or device memory (on_device).  Library GEMV/GEMMs (cuBLAS).
extern "C" int agf2_b200_get_jk(agf2_b200_eri* h, const double* rdm1, double* j_out, double* k_out, int on_device,
                                void* stream)
{
    if (!h || !rdm1 || !j_out || !k_out) return afail(AGF2_B200_EINVAL, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t P = h->nphys, N = h->naux, PP = P * P;
    // block of auxiliary functions sized so that T' and L' stay within ~512 MiB each
    const int64_t qblk = std::max<int64_t>(1, std::min<int64_t>(N, (int64_t)((512ll << 20) / (PP * 8))));
    const size_t need = (size_t)(3 * PP + N + 2 * qblk * PP);
    if (int r = agrow(h->jk, h->jk_elems, need)) return r;
    double* D = h->jk; double* J = D + PP; double* K = J + PP; double* rho = K + PP;
    double* Tp = rho + N; double* Lp = Tp + qblk * PP;
    const cudaMemcpyKind in = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const cudaMemcpyKind out = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    A_CU(cudaMemcpyAsync(D, rdm1, (size_t)PP * 8, in, st));
    A_CU(cudaMemsetAsync(K, 0, (size_t)PP * 8, st));
    A_BLAS(cublasSetStream(a_blas, st));
    const double one = 1.0, zero = 0.0;
    // rho[Q] = sum_pq L[Q,pq] D[pq] ;  J[pq] = sum_Q rho[Q] L[Q,pq]       (L as column-major (P^2 x N))
    A_BLAS(cublasDgemv(a_blas, CUBLAS_OP_T, (int)PP, (int)N, &one, h->full, (int)PP, D, 1, &zero, rho, 1));
    A_BLAS(cublasDgemv(a_blas, CUBLAS_OP_N, (int)PP, (int)N, &one, h->full, (int)PP, rho, 1, &zero, J, 1));
    for (int64_t q0 = 0; q0 < N; q0 += qblk) {
        const int64_t nq = std::min(qblk, N - q0);
        const double* Lb = h->full + q0 * PP;
        // T'[p][Q][s] = sum_r L[Q,p,r] D[r,s]   (batched over p; written p-major)
        A_BLAS(cublasDgemmStridedBatched(a_blas, CUBLAS_OP_N, CUBLAS_OP_N, (int)P, (int)nq, (int)P, &one,
                                         D, (int)P, 0, Lb, (int)PP, P, &zero, Tp, (int)P, nq * P, (int)P));
        dim3 grid((unsigned)((P + 127) / 128), (unsigned)P, (unsigned)nq);
        permute_block_kernel<<<grid, 128, 0, st>>>(Lb, Lp, (int)P, (int)nq);
        A_CU(cudaGetLastError());
        // K[p,q] += sum_{(Q,s)} T'[p,(Q,s)] L'[q,(Q,s)]
        A_BLAS(cublasDgemm(a_blas, CUBLAS_OP_T, CUBLAS_OP_N, (int)P, (int)P, (int)(nq * P), &one, Lp, (int)(nq * P),
                           Tp, (int)(nq * P), &one, K, (int)P));
    }
    A_CU(cudaMemcpyAsync(j_out, J, (size_t)PP * 8, out, st));
    A_CU(cudaMemcpyAsync(k_out, K, (size_t)PP * 8, out, st));
    A_CU(cudaStreamSynchronize(st));
    return 0;
}
