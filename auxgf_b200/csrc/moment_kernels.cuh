// Device kernels of the DF-AGF2(None,0) moment build for sm_100a.
//
//   k1_form     : (xi|ja) formation.  One CTA = one 128-row (x) tile of up to two integral blocks
//                 acc1 = ixq[i1] . qja[:, j1, a1:a1+8*nb1],  acc2 = ixq[i2] . qja[:, j2, a2:...]
//                 (FP64 DMMA.8x8x4, operands staged by TMA into 128B-swizzled shared memory through a
//                 4-stage mbarrier pipeline), then the fused epilogue
//                     out1 = p11*acc1 + p12*acc2,  out2 = p21*acc1 + p22*acc2
//                 written to the L2-resident workspace W[x][col] together with the energy weights
//                 E[col] = e_i + e_j - e_a.  For a pole pair i>j, (acc1, acc2) = (X_ij, X_ji) and
//                 (out1, out2) = (sqrt(os/2+ss) (X_ij - X_ji), sqrt(os/2) (X_ij + X_ji)): the
//                 2(xi|ja)-(xj|ia) exchange combination in pair-symmetric (SYRK) form, each integral
//                 formed once.  Replaces auxgf/lib/libagf2/optragf2.c:121-199.
//   k2_moment   : vv += W W^T, vev += W diag(E) W^T for one (128 x 64) output tile and one K slice
//                 (same DMMA/TMA pipeline; E applied to the A fragments in registers), reduced into
//                 the accumulators with red.global.add.f64.  Symmetric chunks only visit upper-triangle
//                 tiles.  Replaces optragf2.c:206-290.
//   k3_finalize : mirrors the upper triangle, adds the non-symmetric part, accumulates into vv/vev.
//   *_simple    : naive one-thread-per-element versions of k1/k2 driven by the same tile lists (debug).
#pragma once
#include "ptx.cuh"

namespace agf2 {

constexpr int kTileM = 128;        // rows (x) per CTA tile
constexpr int kTileN = 64;         // max columns per integral block / moment tile
constexpr int kKB = 16;            // K elements per pipeline stage (128 B rows -> SWIZZLE_128B)
constexpr int kConsumerWarps = 8;  // two consumer warpgroups; each warp owns 16 rows x 64 cols
// + one producer warpgroup (only its first lane issues TMA).  12 warps cap the launch at 168 registers per
// thread; setmaxnreg then moves registers from the producer warpgroup (40) to the consumers (232), which
// hold 128 accumulator registers plus double-buffered fragments.
constexpr int kThreads = (kConsumerWarps + 4) * 32;
constexpr int kProducerRegs = 40, kConsumerRegs = 232;
__device__ __forceinline__ void reg_dealloc_producer() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProducerRegs)); }
__device__ __forceinline__ void reg_alloc_consumer() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kConsumerRegs)); }

// ---- K1 ---------------------------------------------------------------------------------------
struct K1Tile {
    int i1, j1, a1, nb1;   // block 1: rows ixq[i1], columns qja{bsel1}[:, j1, a1 : a1 + 8*nb1)
    int i2, j2, a2, nb2;   // block 2 (nb2 == 0: absent)
    int xb;                // 128-row block of x
    int bsel1, bsel2;      // which (Q|ja) tensor (0: same spin / RHF, 1: opposite spin)
    int ws1, ws2;          // destination workspace of out1 / out2 (0 or 1; -1: not written)
    int esel1, esel2;      // which virtual-energy vector the E weights of out1 / out2 use
    int nv1, nv2;          // number of real (non-padding) columns of out1 / out2
    long long col1, col2;  // destination column of out1 / out2
    double p11, p12, p21, p22;
    double e1, e2;         // e_i + e_j of out1 / out2
};

constexpr int kK1StageBytes = 2 * kTileM * kKB * 8 + 2 * kTileN * kKB * 8;  // 49152
constexpr int kK1Stages = 4;
constexpr int kK1SmemBytes = kK1Stages * kK1StageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

// pi: logical fragment row -> physical row inside an 8-row group, chosen so that the two rows read
// by one quarter-warp (LDS.128) fall into different halves of the 128B swizzle atom.
__device__ __forceinline__ int pi8(int g) { return (g >> 1) | ((g & 1) << 2); }


// Main loop of k1_form for compile-time block widths (NB1, NB2 in units of 8 columns).
// One pipeline stage (16 Q) = 8 "units" (ks, e, set): the fragments of unit u+1 are loaded while the
// DMMAs of unit u issue (they belong to the other block, so the registers are free), including across
// the stage boundary, so shared-memory latency never sits between dependent instructions.
template <int NB1, int NB2>
__device__ __forceinline__ void k1_mainloop(const uint8_t* __restrict__ gbase, uint32_t bar_full, uint32_t bar_empty,
                                            int nk, int warp, int lane, double (&acc1)[2][8][2],
                                            double (&acc2)[2][8][2])
{
    const int gid = lane >> 2, tig = lane & 3;
    const int pg = pi8(gid);
    // A-type (K-contiguous rows, LDS.128): row = 16*warp + 8*mb + pg, chunk = 4*ks + tig
    const uint32_t a_row = (uint32_t)((16 * warp + pg) * 128);
    const uint32_t a_off[2] = {a_row + (uint32_t)(((tig) ^ pg) << 4), a_row + (uint32_t)(((4 + tig) ^ pg) << 4)};
    // B-type (rows = Q, 16 columns per 2 KB box, LDS.64): row = 8*ks + 2*tig + e,
    // column-in-box = 8*(nb&1) + gid  ->  chunk = 4*(nb&1) + gid/2, half = gid&1
    uint32_t b_off[2][2];
#pragma unroll
    for (int par = 0; par < 2; ++par)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int r = 2 * tig + e;
            const int chunk = 4 * par + (gid >> 1);
            b_off[par][e] = (uint32_t)(r * 128 + ((chunk ^ r) << 4) + (gid & 1) * 8);
        }

    double2 fa[2][2];   // [set][mb]: A fragments of the current ks (x: e = 0, y: e = 1)
    double fb[2][8];    // [set][nb]: B fragments of the current (ks, e)
    constexpr bool kDual = NB2 > 0;
    constexpr int kUnits = kDual ? 8 : 4;

    // unit u -> (ks, e, block).  Register buffers: B fragments alternate with u (so the prefetch of unit
    // u+1 never touches the registers unit u is reading); A fragments live for the two units of one
    // (ks, block) and use buffer `block` (dual) or `ks & 1` (single block).
    auto load_unit = [&](const uint8_t* st, int u) {
        const int set = kDual ? (u & 1) : 0;
        const int e = kDual ? ((u >> 1) & 1) : (u & 1);
        const int ks = kDual ? (u >> 2) : (u >> 1);
        const int abuf = kDual ? set : (ks & 1);
        const int bbuf = u & 1;
        if (e == 0) {
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
                fa[abuf][mb] = *reinterpret_cast<const double2*>(st + set * 16384 + a_off[ks] + mb * 1024);
        }
#pragma unroll
        for (int nb = 0; nb < 8; ++nb)
            if (nb < (set ? NB2 : NB1))
                fb[bbuf][nb] = *reinterpret_cast<const double*>(st + 32768 + set * 8192 + b_off[nb & 1][e] +
                                                                (nb >> 1) * 2048 + ks * 1024);
    };
    auto compute_unit = [&](int u) {
        const int set = kDual ? (u & 1) : 0;
        const int e = kDual ? ((u >> 1) & 1) : (u & 1);
        const int ks = kDual ? (u >> 2) : (u >> 1);
        const int abuf = kDual ? set : (ks & 1);
        const int bbuf = u & 1;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb)
            if (nb < (set ? NB2 : NB1)) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    const double a = e ? fa[abuf][mb].y : fa[abuf][mb].x;
                    if (set) dmma884(acc2[mb][nb][0], acc2[mb][nb][1], a, fb[bbuf][nb]);
                    else dmma884(acc1[mb][nb][0], acc1[mb][nb][1], a, fb[bbuf][nb]);
                }
            }
    };

    mbar_wait(bar_full, 0);
    load_unit(gbase, 0);
    for (int it = 0; it < nk; ++it) {
        const int s = it % kK1Stages;
        const uint8_t* st = gbase + s * kK1StageBytes;
#pragma unroll
        for (int u = 0; u < kUnits; ++u) {
            if (u + 1 < kUnits) {
                load_unit(st, u + 1);
            } else if (it + 1 < nk) {
                const int s2 = (it + 1) % kK1Stages;
                mbar_wait(bar_full + 8 * s2, ((it + 1) / kK1Stages) & 1);
                load_unit(gbase + s2 * kK1StageBytes, 0);
            }
            compute_unit(u);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);
    }
}

__global__ void __launch_bounds__(kThreads, 1)
k1_form(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB0,
        const __grid_constant__ CUtensorMap mapB1, const K1Tile* __restrict__ tiles,
        const double* __restrict__ ea0, const double* __restrict__ ea1,
        double* __restrict__ W0, double* __restrict__ W1, double* __restrict__ E0,
        long long ldw, int nk)
{
    extern __shared__ uint8_t smem_raw[];
    // 1024 B alignment: the swizzle pattern is a function of absolute shared-memory address bits
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_full = base + kK1Stages * kK1StageBytes;
    const uint32_t bar_empty = bar_full + 8 * kK1Stages;
    uint8_t* const gbase = smem_raw + (base - smem_u32(smem_raw));

    __shared__ K1Tile t;   // descriptor stays in shared memory: the epilogue fields are not kept in registers
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        t = tiles[blockIdx.x];
        for (int s = 0; s < kK1Stages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, kConsumerWarps);
        }
        fence_barrier_init();
    }
    __syncthreads();

    // ===== producer warpgroup: one lane streams the operand tiles through the mbarrier ring =====
    if (warp >= kConsumerWarps) {
        reg_dealloc_producer();
        if (warp == kConsumerWarps && lane == 0) {
            tma_prefetch_desc(&mapA);
            tma_prefetch_desc(t.bsel1 ? &mapB1 : &mapB0);
            const int ng1 = (t.nb1 + 1) >> 1, ng2 = (t.nb2 + 1) >> 1;   // 16-column TMA boxes
            const uint32_t stage_tx = (uint32_t)((t.nb1 ? kTileM * kKB * 8 : 0) + (t.nb2 ? kTileM * kKB * 8 : 0) +
                                                 (ng1 + ng2) * 16 * kKB * 8);
            const CUtensorMap* mB1 = t.bsel1 ? &mapB1 : &mapB0;
            const CUtensorMap* mB2 = t.bsel2 ? &mapB1 : &mapB0;
            const int nb1 = t.nb1, nb2 = t.nb2, x0 = t.xb * kTileM, i1 = t.i1, i2 = t.i2, j1 = t.j1, j2 = t.j2,
                      a1 = t.a1, a2 = t.a2;
            for (int j = 0; j < nk; ++j) {
                const int s = j % kK1Stages;
                if (j >= kK1Stages) mbar_wait(bar_empty + 8 * s, ((j / kK1Stages) - 1) & 1);
                const uint32_t st = base + s * kK1StageBytes;
                const uint32_t fb = bar_full + 8 * s;
                mbar_expect_tx(fb, stage_tx);
                const int q0 = j * kKB;
                if (nb1) tma_load_3d(st, &mapA, q0, x0, i1, fb);
                if (nb2) tma_load_3d(st + 16384, &mapA, q0, x0, i2, fb);
                for (int g = 0; g < ng1; ++g) tma_load_3d(st + 32768 + g * 2048, mB1, a1 + 16 * g, j1, q0, fb);
                for (int g = 0; g < ng2; ++g) tma_load_3d(st + 40960 + g * 2048, mB2, a2 + 16 * g, j2, q0, fb);
            }
        }
        return;
    }
    reg_alloc_consumer();

    // ===== consumers: 8 warps x (16 rows x 64 cols x 2 blocks) =====
    const int gid = lane >> 2, tig = lane & 3;
    const int pg = pi8(gid);
    double acc1[2][8][2], acc2[2][8][2];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            acc1[mb][nb][0] = acc1[mb][nb][1] = 0.0;
            acc2[mb][nb][0] = acc2[mb][nb][1] = 0.0;
        }
    const int nb1 = t.nb1, nb2 = t.nb2;
    switch (nb1 * 16 + nb2) {
#define AGF2_K1_CASE(A, B) case (A) * 16 + (B): k1_mainloop<A, B>(gbase, bar_full, bar_empty, nk, warp, lane, acc1, acc2); break;
        AGF2_K1_CASE(8, 8) AGF2_K1_CASE(7, 7) AGF2_K1_CASE(6, 6) AGF2_K1_CASE(5, 5)
        AGF2_K1_CASE(4, 4) AGF2_K1_CASE(3, 3) AGF2_K1_CASE(2, 2) AGF2_K1_CASE(1, 1)
        AGF2_K1_CASE(8, 0) AGF2_K1_CASE(7, 0) AGF2_K1_CASE(6, 0) AGF2_K1_CASE(5, 0)
        AGF2_K1_CASE(4, 0) AGF2_K1_CASE(3, 0) AGF2_K1_CASE(2, 0) AGF2_K1_CASE(1, 0)
#undef AGF2_K1_CASE
        default: __trap();   // the planner only emits (n, n) and (n, 0) tiles
    }

    // ===== fused epilogue: exchange combination + scaling, write W (never the raw integrals) =====
    const long long x0 = (long long)t.xb * kTileM + 16 * warp;
    double* const Wd1 = t.ws1 == 0 ? W0 : (t.ws1 == 1 ? W1 : nullptr);
    double* const Wd2 = t.ws2 == 0 ? W0 : (t.ws2 == 1 ? W1 : nullptr);
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
        const long long row = x0 + 8 * mb + pg;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const int c = 8 * nb + 2 * tig;
            if (Wd1 != nullptr && nb < nb1) {
                double2 o;
                o.x = t.p11 * acc1[mb][nb][0] + t.p12 * acc2[mb][nb][0];
                o.y = t.p11 * acc1[mb][nb][1] + t.p12 * acc2[mb][nb][1];
                *reinterpret_cast<double2*>(Wd1 + row * ldw + t.col1 + c) = o;
            }
            const int nbo2 = (t.p21 != 0.0) ? nb1 : nb2;   // width of out2: block 1's if it mixes block 1
            if (Wd2 != nullptr && nb < nbo2) {
                double2 o;
                o.x = t.p21 * acc1[mb][nb][0] + t.p22 * acc2[mb][nb][0];
                o.y = t.p21 * acc1[mb][nb][1] + t.p22 * acc2[mb][nb][1];
                *reinterpret_cast<double2*>(Wd2 + row * ldw + t.col2 + c) = o;
            }
        }
    }
    // energy weights E[col] = e_i + e_j - e_a (only the xb == 0 tile of each block writes them;
    // only workspace 0 carries weights -- k2 scales its A operand, which always comes from W0)
    if (t.xb == 0 && warp == 0 && gid == 0) {
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = 8 * nb + 2 * tig + h;
                if (t.ws1 == 0 && nb < nb1) {
                    const double* ea = t.esel1 ? ea1 : ea0;
                    E0[t.col1 + c] = (c < t.nv1) ? (t.e1 - ea[t.a1 + c]) : 0.0;
                }
                const int nbo2 = (t.p21 != 0.0) ? nb1 : nb2;
                if (t.ws2 == 0 && nb < nbo2) {
                    const double* ea = t.esel2 ? ea1 : ea0;
                    const int a0 = (t.p21 != 0.0) ? t.a1 : t.a2;
                    E0[t.col2 + c] = (c < t.nv2) ? (t.e2 - ea[a0 + c]) : 0.0;
                }
            }
        }
    }
}

// ---- K2 ---------------------------------------------------------------------------------------
struct K2Tile {
    int xb, yb;   // rows [128*xb, +128) of operand A, rows [64*yb, +64) of operand B
};

constexpr int kK2StageBytes = kTileM * kKB * 8 + kTileN * kKB * 8 + 1024;  // A 16K + B 8K + E(128B, padded)
constexpr int kK2Stages = 7;
constexpr int kK2SmemBytes = kK2Stages * kK2StageBytes + 1024 + 256;

__global__ void __launch_bounds__(kThreads, 1)
k2_moment(const __grid_constant__ CUtensorMap mapWA, const __grid_constant__ CUtensorMap mapWB,
          const double* __restrict__ E, const K2Tile* __restrict__ tiles,
          double* __restrict__ acc_vv, double* __restrict__ acc_vev, long long ldacc,
          int nk_total, int ntiles)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_full = base + kK2Stages * kK2StageBytes;
    const uint32_t bar_empty = bar_full + 8 * kK2Stages;
    uint8_t* const gbase = smem_raw + (base - smem_u32(smem_raw));

    // stream-K: the (tile, k-iteration) space is linearised tile-major and cut into gridDim.x equal
    // ranges, so every SM gets the same number of DMMA iterations whatever the tile count; a CTA
    // flushes its accumulators (red.global.add) whenever its range crosses a tile boundary.
    const long long total = (long long)ntiles * nk_total;
    const long long L0 = total * blockIdx.x / gridDim.x;
    const long long L1 = total * (blockIdx.x + 1) / gridDim.x;
    const int nk = (int)(L1 - L0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kK2Stages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, kConsumerWarps);
        }
        fence_barrier_init();
    }
    __syncthreads();
    if (nk <= 0) return;

    // ===== producer warpgroup =====
    if (warp >= kConsumerWarps) {
        reg_dealloc_producer();
        if (warp == kConsumerWarps && lane == 0) {
            tma_prefetch_desc(&mapWA);
            tma_prefetch_desc(&mapWB);
            constexpr uint32_t bytes = kTileM * kKB * 8 + kTileN * kKB * 8 + kKB * 8;
            int tile = (int)(L0 / nk_total), kit = (int)(L0 % nk_total);
            K2Tile t = tiles[tile];
            for (int j = 0; j < nk; ++j) {
                const int s = j % kK2Stages;
                if (j >= kK2Stages) mbar_wait(bar_empty + 8 * s, ((j / kK2Stages) - 1) & 1);
                const uint32_t st = base + s * kK2StageBytes;
                const uint32_t fb = bar_full + 8 * s;
                mbar_expect_tx(fb, bytes);
                const int k0 = kit * kKB;
                tma_load_2d(st, &mapWA, k0, t.xb * kTileM, fb);
                tma_load_2d(st + 8192, &mapWA, k0, t.xb * kTileM + 64, fb);
                tma_load_2d(st + 16384, &mapWB, k0, t.yb * kTileN, fb);
                bulk_load_1d(st + 24576, E + k0, kKB * 8, fb);
                if (++kit == nk_total && j + 1 < nk) { kit = 0; t = tiles[++tile]; }
            }
        }
        return;
    }
    reg_alloc_consumer();

    const int gid = lane >> 2, tig = lane & 3;
    const int pg = pi8(gid);
    const uint32_t sw[2] = {(uint32_t)(((tig) ^ pg) << 4), (uint32_t)(((4 + tig) ^ pg) << 4)};
    const uint32_t a_row = (uint32_t)((16 * warp + pg) * 128);
    const uint32_t b_row = (uint32_t)(pg * 128);

    double cvv[2][8][2], cvev[2][8][2];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            cvv[mb][nb][0] = cvv[mb][nb][1] = 0.0;
            cvev[mb][nb][0] = cvev[mb][nb][1] = 0.0;
        }

    // One stage = 4 units (ks, half): unit u+1's fragments are loaded while unit u's 32 DMMAs issue.
    double2 fa[2][2], fe[2][2];   // [ks parity][mb]: A fragments and their energy-weighted copies
    double2 fb[2][4];             // [half][j]: B fragments of n-blocks 4*half + j
    auto load_unit = [&](const uint8_t* st, int u) {
        const int ks = u >> 1, half = u & 1;
        if (half == 0) {
            const double2 ev = *reinterpret_cast<const double2*>(st + 24576 + (4 * ks + tig) * 16);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                fa[ks][mb] = *reinterpret_cast<const double2*>(st + a_row + mb * 1024 + sw[ks]);
                fe[ks][mb].x = fa[ks][mb].x * ev.x;
                fe[ks][mb].y = fa[ks][mb].y * ev.y;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            fb[half][j] = *reinterpret_cast<const double2*>(st + 16384 + b_row + (4 * half + j) * 1024 + sw[ks]);
    };
    auto compute_unit = [&](int u) {
        const int ks = u >> 1, half = u & 1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nb = 4 * half + j;
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                dmma884(cvv[mb][nb][0], cvv[mb][nb][1], fa[ks][mb].x, fb[half][j].x);
                dmma884(cvev[mb][nb][0], cvev[mb][nb][1], fe[ks][mb].x, fb[half][j].x);
            }
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                dmma884(cvv[mb][nb][0], cvv[mb][nb][1], fa[ks][mb].y, fb[half][j].y);
                dmma884(cvev[mb][nb][0], cvev[mb][nb][1], fe[ks][mb].y, fb[half][j].y);
            }
        }
    };

    int tile = (int)(L0 / nk_total), kit = (int)(L0 % nk_total);
    mbar_wait(bar_full, 0);
    load_unit(gbase, 0);
    for (int it = 0; it < nk; ++it) {
        const int s = it % kK2Stages;
        const uint8_t* st = gbase + s * kK2StageBytes;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (u < 3) {
                load_unit(st, u + 1);
            } else if (it + 1 < nk) {
                const int s2 = (it + 1) % kK2Stages;
                mbar_wait(bar_full + 8 * s2, ((it + 1) / kK2Stages) & 1);
                load_unit(gbase + s2 * kK2StageBytes, 0);
            }
            compute_unit(u);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);
        if (++kit == nk_total || it + 1 == nk) {
            // split-K / cross-chunk reduction: warp-level fragments straight into the global accumulators
            const K2Tile t = tiles[tile];
            const long long x0 = (long long)t.xb * kTileM + 16 * warp;
            const long long y0 = (long long)t.yb * kTileN;
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                const long long row = x0 + 8 * mb + pg;
#pragma unroll
                for (int nb = 0; nb < 8; ++nb)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const long long col = y0 + 8 * nb + pi8(2 * tig + h);
                        red_add_f64(acc_vv + row * ldacc + col, cvv[mb][nb][h]);
                        red_add_f64(acc_vev + row * ldacc + col, cvev[mb][nb][h]);
                        cvv[mb][nb][h] = 0.0;
                        cvev[mb][nb][h] = 0.0;
                    }
            }
            kit = 0;
            ++tile;
        }
    }
}

// ---- K3: finalize ------------------------------------------------------------------------------
// out[x][y] += (x <= y ? S[x][y] : S[y][x]) + A[x][y]      (S: symmetric-chunk accumulator, upper
// triangle valid; A: accumulator of the non-symmetric chunks, may be null)
__global__ void k3_finalize(const double* __restrict__ s_vv, const double* __restrict__ s_vev,
                            const double* __restrict__ a_vv, const double* __restrict__ a_vev,
                            long long ldacc, int nphys, double* __restrict__ vv, double* __restrict__ vev)
{
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    const int x = blockIdx.y;
    if (y >= nphys) return;
    const long long u = (x <= y) ? (x * ldacc + y) : (y * ldacc + x);
    double r0 = s_vv[u], r1 = s_vev[u];
    if (a_vv != nullptr) {
        r0 += a_vv[x * ldacc + y];
        r1 += a_vev[x * ldacc + y];
    }
    vv[(long long)x * nphys + y] += r0;
    vev[(long long)x * nphys + y] += r1;
}

// ---- naive debug kernels (same tile lists, no tensor cores, no TMA) ------------------------------
__global__ void k1_form_simple(const double* __restrict__ ixq, long long ld_ixq, int nphys, int naux,
                               const double* __restrict__ qja0, long long ld_q0, int nocc0, int nvir0,
                               const double* __restrict__ qja1, long long ld_q1, int nocc1, int nvir1,
                               const K1Tile* __restrict__ tiles, const double* __restrict__ ea0,
                               const double* __restrict__ ea1, double* __restrict__ W0,
                               double* __restrict__ W1, double* __restrict__ E0, long long ldw)
{
    const K1Tile t = tiles[blockIdx.x];
    for (int idx = threadIdx.x; idx < kTileM * kTileN; idx += blockDim.x) {
        const int r = idx / kTileN, c = idx % kTileN;
        const long long x = (long long)t.xb * kTileM + r;
        double s1 = 0.0, s2 = 0.0;
        if (x < nphys) {
            if (c < 8 * t.nb1) {
                const double* q = t.bsel1 ? qja1 : qja0;
                const long long ldq = t.bsel1 ? ld_q1 : ld_q0;
                const int no = t.bsel1 ? nocc1 : nocc0, nv = t.bsel1 ? nvir1 : nvir0;
                if (t.a1 + c < nv)
                    for (int k = 0; k < naux; ++k)
                        s1 += ixq[((long long)t.i1 * nphys + x) * ld_ixq + k] *
                              q[((long long)k * no + t.j1) * ldq + t.a1 + c];
            }
            if (c < 8 * t.nb2) {
                const double* q = t.bsel2 ? qja1 : qja0;
                const long long ldq = t.bsel2 ? ld_q1 : ld_q0;
                const int no = t.bsel2 ? nocc1 : nocc0, nv = t.bsel2 ? nvir1 : nvir0;
                if (t.a2 + c < nv)
                    for (int k = 0; k < naux; ++k)
                        s2 += ixq[((long long)t.i2 * nphys + x) * ld_ixq + k] *
                              q[((long long)k * no + t.j2) * ldq + t.a2 + c];
            }
        }
        const int nbo2 = (t.p21 != 0.0) ? t.nb1 : t.nb2;
        if (t.ws1 >= 0 && c < 8 * t.nb1)
            (t.ws1 ? W1 : W0)[x * ldw + t.col1 + c] = t.p11 * s1 + t.p12 * s2;
        if (t.ws2 >= 0 && c < 8 * nbo2)
            (t.ws2 ? W1 : W0)[x * ldw + t.col2 + c] = t.p21 * s1 + t.p22 * s2;
        if (t.xb == 0 && r == 0) {
            if (t.ws1 == 0 && c < 8 * t.nb1)
                E0[t.col1 + c] = (c < t.nv1) ? (t.e1 - (t.esel1 ? ea1 : ea0)[t.a1 + c]) : 0.0;
            if (t.ws2 == 0 && c < 8 * nbo2) {
                const int a0 = (t.p21 != 0.0) ? t.a1 : t.a2;
                E0[t.col2 + c] = (c < t.nv2) ? (t.e2 - (t.esel2 ? ea1 : ea0)[a0 + c]) : 0.0;
            }
        }
    }
}

__global__ void k2_moment_simple(const double* __restrict__ WA, const double* __restrict__ WB,
                                 long long ldw, const double* __restrict__ E, int kc,
                                 const K2Tile* __restrict__ tiles, double* __restrict__ acc_vv,
                                 double* __restrict__ acc_vev, long long ldacc)
{
    const K2Tile t = tiles[blockIdx.x];
    for (int idx = threadIdx.x; idx < kTileM * kTileN; idx += blockDim.x) {
        const long long x = (long long)t.xb * kTileM + idx / kTileN;
        const long long y = (long long)t.yb * kTileN + idx % kTileN;
        double s0 = 0.0, s1 = 0.0;
        for (int k = 0; k < kc; ++k) {
            const double p = WA[x * ldw + k] * WB[y * ldw + k];
            s0 += p;
            s1 += p * E[k];
        }
        acc_vv[x * ldacc + y] += s0;
        acc_vev[x * ldacc + y] += s1;
    }
}

}  // namespace agf2
