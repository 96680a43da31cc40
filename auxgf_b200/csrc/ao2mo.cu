// Device-resident DF tensor, the per-sector DF transform ("ao2mo") and the Coulomb / exchange matrices on the GPU.
//
// Replaces, for the moment build, the two `OptRAGF2.ao2mo` calls of auxgf/agf2/optragf2.py:243-244
// (PySCF `_ao2mo.nr_e2`, 's2' packed input -> 's1' unpacked output, optragf2.py:175-214):
//     ixq[i, x, Q] = sum_p   L[Q, x, p] C_occ[p, i]
//     qja[Q, j, a] = sum_pq  L[Q, p, q] C_occ[p, j] C_vir[q, a]
// L is unpacked once to (naux, nphys, ldp) in HBM ("staged once to HBM").  Every contraction runs on the library's
// own FP64 DMMA / TMA / stream-K pipeline (agf2::run_gemm, the general-GEMM instantiation of k2_moment) and writes
// straight into the layouts the moment kernels consume; no library GEMM is involved.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "context.h"

using namespace agf2;

namespace {

inline int64_t rup2(int64_t x) { return (x + 1) / 2 * 2; }

// packed lower triangle (row p, col q <= p at p(p+1)/2 + q; pyscf.lib.pack_tril) -> full symmetric, pitch ldp
__global__ void unpack_tril_kernel(const double* __restrict__ packed, double* __restrict__ full, int nphys, int ldp,
                                   long long npair)
{
    const long long q = blockIdx.z;          // auxiliary index
    const int p = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ldp) return;
    double v = 0.0;
    if (c < nphys) {
        const int hi = p > c ? p : c, lo = p > c ? c : p;
        v = packed[q * npair + (long long)hi * (hi + 1) / 2 + lo];
    }
    full[(q * nphys + p) * ldp + c] = v;
}

// dst (cols x ld_dst, zero padded) = src (rows x cols, row-major)^T
__global__ void transpose_pad_kernel(const double* __restrict__ src, double* __restrict__ dst, int rows, int cols, int ld_dst)
{
    const int c = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ld_dst) return;
    dst[(long long)c * ld_dst + r] = r < rows ? src[(long long)r * cols + c] : 0.0;
}

// rho[Q] = sum_pq L[Q,p,q] D[p,q]   (one block per Q)
__global__ void jk_rho_kernel(const double* __restrict__ full, const double* __restrict__ d, double* __restrict__ rho,
                              int nphys, int ldp)
{
    const long long Q = blockIdx.x;
    const double* l = full + Q * nphys * ldp;
    double s = 0.0;
    for (long long k = threadIdx.x; k < (long long)nphys * ldp; k += blockDim.x) {
        const int p = (int)(k / ldp), q = (int)(k % ldp);
        if (q < nphys) s += l[k] * d[(long long)p * ldp + q];
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {       // fixed tree: bit-reproducible
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) rho[Q] = sh[0];
}

// J[p,q] = sum_Q rho[Q] L[Q,p,q]
__global__ void jk_j_kernel(const double* __restrict__ full, const double* __restrict__ rho, double* __restrict__ j,
                            int naux, int nphys, int ldp)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (long long)nphys * ldp) return;
    double s = 0.0;
    for (int Q = 0; Q < naux; ++Q) s += rho[Q] * full[(long long)Q * nphys * ldp + k];
    j[k] = s;
}

// host / device (rows x cols, contiguous) <-> device (rows x ld) copies
cudaError_t copy_in(double* dst, int64_t ld, const double* src, int64_t rows, int64_t cols, bool on_device, cudaStream_t st) {
    return cudaMemcpy2DAsync(dst, ld * 8, src, cols * 8, cols * 8, rows, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st);
}
cudaError_t copy_out(double* dst, const double* src, int64_t ld, int64_t rows, int64_t cols, bool on_device, cudaStream_t st) {
    return cudaMemcpy2DAsync(dst, cols * 8, src, ld * 8, cols * 8, rows, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st);
}

}  // namespace

struct agf2_b200_eri {
    int device = 0;
    int64_t naux = 0, nphys = 0, ldp = 0;
    double* full = nullptr;           // (naux, nphys, ldp)
    DevBuf ixq;                       // scratch (nocc, nphys, ld_ixq)
    DevBuf qja;                       // scratch (naux, nocc, ld_qja)
    DevBuf coef;                      // staged coefficient matrices, transposed: C_occ^T | C_vir^T
    DevBuf half;                      // half-transformed block Z[Q, a, p]
    DevBuf jk;                        // get_jk scratch: D, J, K, rho, T blocks
};

extern "C" const char* agf2_b200_ao2mo_last_error(void) { return last_error().c_str(); }

extern "C" int agf2_b200_eri_create(const double* eri, int64_t naux, int64_t nphys, int packed, int on_device,
                                    agf2_b200_eri** out)
{
    if (!eri || !out || naux <= 0 || nphys <= 0 || nphys > 65535) return fail(AGF2_B200_EINVAL, "bad argument");
    Ctx* c = nullptr;
    if (int r = ctx_for_current_device(&c)) return r;
    agf2_b200_eri* h = new agf2_b200_eri();
    h->device = c->device;
    h->naux = naux; h->nphys = nphys; h->ldp = rup2(nphys);
    const int64_t ldp = h->ldp;
    const size_t nfull = (size_t)naux * nphys * ldp;
    const long long npair = (long long)nphys * (nphys + 1) / 2;
    cudaError_t e = cudaMalloc(&h->full, nfull * 8);
    if (e != cudaSuccess) { delete h; return fail(AGF2_B200_ENOMEM, std::string("cudaMalloc(full DF tensor): ") + cudaGetErrorString(e)); }
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (!packed) {
        e = cudaMemcpy2D(h->full, ldp * 8, eri, nphys * 8, nphys * 8, naux * nphys, kind);
        if (e == cudaSuccess && ldp != nphys)
            e = cudaMemset2D(h->full + nphys, ldp * 8, 0, (ldp - nphys) * 8, naux * nphys);
        if (e != cudaSuccess) { cudaFree(h->full); delete h; return fail(AGF2_B200_ECUDA, cudaGetErrorString(e)); }
    } else {
        // stream the packed tensor through a bounded staging buffer, unpacking block by block
        const int64_t blk = std::max<int64_t>(1, std::min<int64_t>(naux, (int64_t)((256ll << 20) / (npair * 8))));
        double* stage = nullptr;
        e = cudaMalloc(&stage, (size_t)blk * npair * 8);
        if (e != cudaSuccess) { cudaFree(h->full); delete h; return fail(AGF2_B200_ENOMEM, cudaGetErrorString(e)); }
        for (int64_t q0 = 0; q0 < naux; q0 += blk) {
            const int64_t nq = std::min(blk, naux - q0);
            e = cudaMemcpy(stage, eri + q0 * npair, (size_t)nq * npair * 8, kind);
            if (e == cudaSuccess) {
                dim3 grid((unsigned)((ldp + 127) / 128), (unsigned)nphys, (unsigned)nq);
                unpack_tril_kernel<<<grid, 128>>>(stage, h->full + (size_t)q0 * nphys * ldp, (int)nphys, (int)ldp, npair);
                e = cudaDeviceSynchronize();
            }
            if (e != cudaSuccess) { cudaFree(stage); cudaFree(h->full); delete h; return fail(AGF2_B200_ECUDA, cudaGetErrorString(e)); }
        }
        cudaFree(stage);
    }
    *out = h;
    return 0;
}

extern "C" int agf2_b200_eri_destroy(agf2_b200_eri* h)
{
    if (!h) return 0;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    cudaFree(h->full);
    for (DevBuf* b : {&h->ixq, &h->qja, &h->coef, &h->half, &h->jk}) if (b->p) cudaFree(b->p);
    delete h;
    return 0;
}

/* (naux, nphys, ldp) with ldp = nphys rounded up to even; the pad column is zero */
extern "C" int agf2_b200_eri_full_ptr(agf2_b200_eri* h, double** full)
{
    if (!h || !full) return fail(AGF2_B200_EINVAL, "null handle");
    *full = h->full;
    return 0;
}

extern "C" int agf2_b200_ao2mo_sector(agf2_b200_eri* h, const double* v_occ, int64_t nocc, const double* v_vir,
                                      int64_t nvir, int on_device, void* stream, double** ixq_out, int64_t* ld_ixq_out,
                                      double** qja_out, int64_t* ld_qja_out)
{
    if (!h || !v_occ || !v_vir || nocc <= 0 || nvir <= 0 || !ixq_out || !qja_out || !ld_ixq_out || !ld_qja_out)
        return fail(AGF2_B200_EINVAL, "bad argument");
    DeviceGuard guard(h->device);
    Ctx* cp = nullptr;
    if (int r = ctx_for_current_device(&cp)) return r;
    Ctx& c = *cp;
    std::lock_guard<std::mutex> lk(c.mu);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t P = h->nphys, N = h->naux, ldp = h->ldp;
    const int64_t ldn = rup2(N), ldv = rup2(nvir);
    if (int r = grow(h->ixq, (size_t)nocc * P * ldn, false)) return r;
    if (int r = grow(h->qja, (size_t)N * nocc * ldv, false)) return r;
    if (int r = grow(h->coef, (size_t)(rup2(P * nocc) + rup2(P * nvir)) + (size_t)ldp * (nocc + nvir), false)) return r;
    // coefficients: staged row-major (P, n), then transposed to (n, ldp) so that the contraction index is contiguous
    // (every piece starts on a 16-byte boundary: TMA operand bases)
    double* co_raw = h->coef.p;
    double* cv_raw = co_raw + rup2(P * nocc);
    double* coT = cv_raw + rup2(P * nvir);
    double* cvT = coT + nocc * ldp;
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CU_TRY(cudaMemcpyAsync(co_raw, v_occ, (size_t)P * nocc * 8, kind, st));
    CU_TRY(cudaMemcpyAsync(cv_raw, v_vir, (size_t)P * nvir * 8, kind, st));
    transpose_pad_kernel<<<dim3((unsigned)((ldp + 127) / 128), (unsigned)nocc), 128, 0, st>>>(co_raw, coT, (int)P, (int)nocc, (int)ldp);
    transpose_pad_kernel<<<dim3((unsigned)((ldp + 127) / 128), (unsigned)nvir), 128, 0, st>>>(cv_raw, cvT, (int)P, (int)nvir, (int)ldp);
    CU_TRY(cudaGetLastError());
    if (ldn != N) CU_TRY(cudaMemsetAsync(h->ixq.p, 0, (size_t)nocc * P * ldn * 8, st));
    if (ldv != nvir) CU_TRY(cudaMemsetAsync(h->qja.p, 0, (size_t)N * nocc * ldv * 8, st));

    // (1) for every physical index x (batch):  ixq[:, x, :] (nocc x naux, written transposed) = L[:, x, :] (naux x P) . C_occ
    {
        Gemm g;
        g.a = h->full; g.a_rows = N; g.lda = P * ldp; g.a_third = ldp;      // A_x[Q, p] = L[Q, x, p]
        g.b = coT; g.b_rows = nocc; g.ldb = ldp; g.b_third = 0;             // B[i, p] = C_occ[p, i]
        g.k = P; g.nbatch = P;
        g.c = h->ixq.p; g.ldc = P * ldn; g.c_batch = ldn; g.transpose = true;   // C_x[Q, i] -> ixq[i, x, Q]
        if (int r = run_gemm(c, g, st)) return r;
    }
    // (2) blocks of auxiliary functions:  Z[Q, a, p] = sum_x C_vir[x, a] L[Q, x, p]  (L symmetric in x, p), then
    //     qja[Q, j, a] = sum_p C_occ[p, j] Z[Q, a, p]
    const int64_t qblk = std::max<int64_t>(1, std::min<int64_t>(N, (512ll << 20) / (nvir * ldp * 8)));
    if (int r = grow(h->half, (size_t)qblk * nvir * ldp, false)) return r;
    for (int64_t q0 = 0; q0 < N; q0 += qblk) {
        const int64_t nq = std::min(qblk, N - q0);
        Gemm g1;
        g1.a = cvT; g1.a_rows = nvir; g1.lda = ldp; g1.a_third = 0;                          // A[a, x] = C_vir[x, a]
        g1.b = h->full + q0 * P * ldp; g1.b_rows = P; g1.ldb = ldp; g1.b_third = P * ldp;    // B_Q[p, x] = L[Q, p, x]
        g1.k = P; g1.nbatch = nq;
        g1.c = h->half.p; g1.ldc = ldp; g1.c_batch = nvir * ldp;                             // Z[Q, a, p]
        if (int r = run_gemm(c, g1, st)) return r;
        Gemm g2;
        g2.a = coT; g2.a_rows = nocc; g2.lda = ldp; g2.a_third = 0;                          // A[j, p] = C_occ[p, j]
        g2.b = h->half.p; g2.b_rows = nvir; g2.ldb = ldp; g2.b_third = nvir * ldp;           // B_Q[a, p] = Z[Q, a, p]
        g2.k = P; g2.nbatch = nq;
        g2.c = h->qja.p + q0 * nocc * ldv; g2.ldc = ldv; g2.c_batch = nocc * ldv;            // qja[Q, j, a]
        if (int r = run_gemm(c, g2, st)) return r;
    }
    *ixq_out = h->ixq.p; *ld_ixq_out = ldn;
    *qja_out = h->qja.p; *ld_qja_out = ldv;
    return 0;
}


// Coulomb and exchange matrices from the resident DF tensor (replaces the blocked host loop of
// auxgf/agf2/optragf2.py:358-398 / optuagf2.py:317-366):
//     J[p,q] = sum_Q L[Q,p,q] (sum_rs L[Q,r,s] D[r,s]),     K[p,q] = sum_Q sum_rs L[Q,p,r] D[r,s] L[Q,q,s]
// rdm1, j, k: (nphys, nphys) row-major; host or device memory (on_device).  J: two memory-bound reductions over the
// DF tensor; K: T_Q = L_Q D^T-rows (batched over Q) and K = sum_(Q,s) T[Q,p,s] L[Q,q,s] (outer-K GEMM), both on the
// in-house pipeline.
extern "C" int agf2_b200_get_jk(agf2_b200_eri* h, const double* rdm1, double* j_out, double* k_out, int on_device,
                                void* stream)
{
    if (!h || !rdm1 || !j_out || !k_out) return fail(AGF2_B200_EINVAL, "bad argument");
    DeviceGuard guard(h->device);
    Ctx* cp = nullptr;
    if (int r = ctx_for_current_device(&cp)) return r;
    Ctx& c = *cp;
    std::lock_guard<std::mutex> lk(c.mu);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t P = h->nphys, N = h->naux, ldp = h->ldp, PP = P * ldp;
    // block of auxiliary functions sized so that T stays within ~512 MiB
    const int64_t qblk = std::max<int64_t>(1, std::min<int64_t>(N, (int64_t)((512ll << 20) / (PP * 8))));
    const size_t need = (size_t)(4 * PP + rup2(N) + qblk * PP);
    if (int r = grow(h->jk, need, false)) return r;
    double* D = h->jk.p; double* Dt = D + PP; double* J = Dt + PP; double* K = J + PP; double* rho = K + PP;
    double* T = rho + rup2(N);
    CU_TRY(cudaMemsetAsync(D, 0, (size_t)PP * 8, st));
    CU_TRY(copy_in(D, ldp, rdm1, P, P, on_device != 0, st));
    // Dt[s, r] = D[r, s]: the contraction index r of T_Q = L_Q D must be contiguous in the B operand
    transpose_pad_kernel<<<dim3((unsigned)((ldp + 127) / 128), (unsigned)P), 128, 0, st>>>(D, Dt, (int)P, (int)ldp, (int)ldp);
    jk_rho_kernel<<<(unsigned)N, 256, 0, st>>>(h->full, D, rho, (int)P, (int)ldp);
    jk_j_kernel<<<(unsigned)((PP + 255) / 256), 256, 0, st>>>(h->full, rho, J, (int)N, (int)P, (int)ldp);
    CU_TRY(cudaGetLastError());
    for (int64_t q0 = 0; q0 < N; q0 += qblk) {
        const int64_t nq = std::min(qblk, N - q0);
        const double* Lb = h->full + q0 * PP;
        Gemm g1;                                       // T[(Q, p), s] = sum_r L[(Q, p), r] D[r, s]: the rows (Q, p) of the
        g1.a = Lb; g1.a_rows = nq * P; g1.lda = ldp;   // block are equally spaced, so this is ONE tall GEMM (a batch of nq
        g1.b = Dt; g1.b_rows = P; g1.ldb = ldp;        // small ones would pad every nphys rows to the 256-row tile)
        g1.k = P;
        g1.c = T; g1.ldc = ldp;
        if (int r = run_gemm(c, g1, st)) return r;
        Gemm g2;                                       // K[p, q] (+)= sum_{Q, s} T[Q, p, s] L[Q, q, s]
        g2.a = T; g2.a_rows = P; g2.lda = ldp; g2.a_third = PP;
        g2.b = Lb; g2.b_rows = P; g2.ldb = ldp; g2.b_third = PP;
        g2.k = P; g2.n_outer = nq;
        g2.c = K; g2.ldc = ldp;
        g2.accumulate = q0 > 0;
        if (int r = run_gemm(c, g2, st)) return r;
    }
    CU_TRY(copy_out(j_out, J, ldp, P, P, on_device != 0, st));
    CU_TRY(copy_out(k_out, K, ldp, P, P, on_device != 0, st));
    CU_TRY(cudaStreamSynchronize(st));
    return 0;
}
