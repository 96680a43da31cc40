// Device kernels of the DF-AGF2(None,0) moment build for sm_100a.
//
//   k1_form     : (xi|ja) formation.  One CTA = one 128-row (x) tile of up to two integral blocks
//                 acc1 = ixq[i1] . qja[:, j1, a1:a1+8*nb1],  acc2 = ixq[i2] . qja[:, j2, a2:...]
//                 (FP64 DMMA.8x8x4, operands staged by TMA into 128B-swizzled shared memory through a
//                 4-stage mbarrier pipeline), then the fused epilogue
//                     out1 = p11*acc1 + p12*acc2,  out2 = p21*acc1 + p22*acc2
//                 written to the chunk workspace W[x][col] together with the energy weights
//                 E[col] = e_i + e_j - e_a.  For a pole pair i>j, (acc1, acc2) = (X_ij, X_ji) and
//                 out1 = sqrt(ss) (X_ij - X_ji)  [all-pairs route: also out2 = sqrt(os/2) (X_ij + X_ji)]:
//                 the 2(xi|ja)-(xj|ia) exchange combination in pair-symmetric (SYRK) form, each integral
//                 formed once.  A ragged last x-block uses the thin-tile (<= 80 valid rows, k1_mainloop_thin) or
//                 quad-tile (81..112 rows, k1_quad_tile) warp mapping.  Replaces auxgf/lib/libagf2/optragf2.c:121-199.
//   k2_moment   : acc_vv += A' B^T, acc_vev += A'' B^T on 128 x 64 output tiles, stream-K over the SMs (same
//                 DMMA/TMA pipeline; weights applied to the A fragments in registers).  The segment that starts a
//                 tile adds into the accumulators, every other segment goes to a per-CTA slot that k2_fixup adds in
//                 CTA order (bit-reproducible sums).  Uses: pair-loop moments W W^T / W diag(E) W^T
//                 (optragf2.c:206-290, upper-triangle tiles only), the M0/M1 build and the Coulomb step of the
//                 Coulomb factorisation (see the comment above the kernel), and -- kGen -- the general GEMM behind
//                 T = ixq [M0|M1], the DF transform, DF-JK and the two-body energy.
//   k3_finalize : mirrors the upper triangle, adds the non-symmetric part, accumulates into vv/vev.
//   *_simple    : naive one-thread-per-element versions of k1/k2 driven by the same tile lists (debug).
#pragma once
#include <type_traits>
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
    int mbv;               // > 0: only the first mbv 8-row blocks of the x-block exist: thin tile (<= 10, k1_mainloop_thin)
                           // or quad tile (11..14, k1_quad_tile)
    int asel1, asel2;      // which (ix|Q) tensor i1 / i2 index: 0 the resident one, 1 the visiting pole block (pole-sharded
    int pad_;              // builds: ixq is split over the GPUs and blocks of poles travel over NVLink)
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

// energy weights E[col] = e_i + e_j - e_a (only the xb == 0 tile of each block writes them; only workspace 0
// carries weights -- k2 scales its A operand, which always comes from W0)
__device__ __forceinline__ void k1_write_weights(const K1Tile& t, int warp, int gid, int tig,
                                                 const double* __restrict__ ea0, const double* __restrict__ ea1,
                                                 double* __restrict__ E0)
{
    if (t.xb == 0 && warp == 0 && gid == 0) {
        const int nb1 = t.nb1, nb2 = t.nb2;
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

// Thin tile: the last x-block of a shape whose nphys is not a multiple of 128 holds only MBV <= 10 valid 8-row
// blocks.  Here warp w takes column block w of both integral blocks and ALL row blocks (instead of 16 rows and all
// column blocks), so the tile costs MBV / (nb1 + nb2) of a full one instead of the same.  One pipeline stage =
// 4 units (ks, set); the fragments of unit u+1 (the other set) are loaded while unit u's DMMAs issue.
template <int MBV>
__device__ __forceinline__ void k1_mainloop_thin(const uint8_t* __restrict__ gbase, uint32_t bar_full, uint32_t bar_empty,
                                                 int nk, int warp, int lane, bool act1, bool act2,
                                                 double (&acc1)[MBV][2], double (&acc2)[MBV][2])
{
    const int gid = lane >> 2, tig = lane & 3;
    const int pg = pi8(gid);
    const uint32_t a_off[2] = {(uint32_t)(pg * 128 + (((tig) ^ pg) << 4)), (uint32_t)(pg * 128 + (((4 + tig) ^ pg) << 4))};
    uint32_t b_off[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int r = 2 * tig + e;
        const int chunk = 4 * (warp & 1) + (gid >> 1);
        b_off[e] = (uint32_t)(r * 128 + ((chunk ^ r) << 4) + (gid & 1) * 8 + (warp >> 1) * 2048);
    }
    double2 fa[2][MBV];   // [set][mb]
    double fb[2][2];      // [set][e]
    auto load_unit = [&](const uint8_t* st, int u) {
        const int ks = u >> 1, set = u & 1;
        if (set ? act2 : act1) {
#pragma unroll
            for (int mb = 0; mb < MBV; ++mb)
                fa[set][mb] = *reinterpret_cast<const double2*>(st + set * 16384 + a_off[ks] + mb * 1024);
#pragma unroll
            for (int e = 0; e < 2; ++e)
                fb[set][e] = *reinterpret_cast<const double*>(st + 32768 + set * 8192 + b_off[e] + ks * 1024);
        }
    };
    auto compute_unit = [&](int u) {
        const int set = u & 1;
        if (set ? act2 : act1) {
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int mb = 0; mb < MBV; ++mb) {
                    const double a = e ? fa[set][mb].y : fa[set][mb].x;
                    if (set) dmma884(acc2[mb][0], acc2[mb][1], a, fb[set][e]);
                    else dmma884(acc1[mb][0], acc1[mb][1], a, fb[set][e]);
                }
        }
    };
    mbar_wait(bar_full, 0);
    load_unit(gbase, 0);
    for (int it = 0; it < nk; ++it) {
        const int s = it % kK1Stages;
        const uint8_t* st = gbase + s * kK1StageBytes;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (u < 3) {
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

// thin-tile main loop + fused epilogue for one compile-time row-block count
template <int MBV>
__device__ __forceinline__ void k1_thin_tile(const K1Tile& t, const uint8_t* __restrict__ gbase, uint32_t bar_full,
                                             uint32_t bar_empty, int nk, int warp, int lane, double* __restrict__ W0,
                                             double* __restrict__ W1, long long ldw, unsigned long long w_policy)
{
    const int gid = lane >> 2, tig = lane & 3;
    const int pg = pi8(gid);
    double acc1[MBV][2], acc2[MBV][2];
#pragma unroll
    for (int mb = 0; mb < MBV; ++mb) { acc1[mb][0] = acc1[mb][1] = 0.0; acc2[mb][0] = acc2[mb][1] = 0.0; }
    const bool act1 = warp < t.nb1, act2 = warp < t.nb2;
    k1_mainloop_thin<MBV>(gbase, bar_full, bar_empty, nk, warp, lane, act1, act2, acc1, acc2);
    double* const Wd1 = t.ws1 == 0 ? W0 : (t.ws1 == 1 ? W1 : nullptr);
    double* const Wd2 = t.ws2 == 0 ? W0 : (t.ws2 == 1 ? W1 : nullptr);
    const int nbo2 = (t.p21 != 0.0) ? t.nb1 : t.nb2;
    const int c = 8 * warp + 2 * tig;
#pragma unroll
    for (int mb = 0; mb < MBV; ++mb) {
        const long long row = (long long)t.xb * kTileM + 8 * mb + pg;
        if (Wd1 != nullptr && act1) {
            double2 o;
            o.x = t.p11 * acc1[mb][0] + t.p12 * acc2[mb][0];
            o.y = t.p11 * acc1[mb][1] + t.p12 * acc2[mb][1];
            st_hint_v2f64(Wd1 + row * ldw + t.col1 + c, o, w_policy);
        }
        if (Wd2 != nullptr && warp < nbo2) {
            double2 o;
            o.x = t.p21 * acc1[mb][0] + t.p22 * acc2[mb][0];
            o.y = t.p21 * acc1[mb][1] + t.p22 * acc2[mb][1];
            st_hint_v2f64(Wd2 + row * ldw + t.col2 + c, o, w_policy);
        }
    }
}

// Quad tile: the ragged last x-block with 11..14 valid 8-row blocks (e.g. nphys = 367: 111 rows).  The thin mapping
// would need all MBV A fragments of both integral blocks in registers; here the 8 warps form a 2 x 4 grid instead --
// warp (r, c) takes the row blocks [r MBW, (r+1) MBW), MBW = ceil(MBV / 2), and the column blocks 2c, 2c+1 of BOTH
// integral blocks (the fused epilogue needs both in the same thread) -- so the tile costs 2 MBW / 16 of a full one
// when nb = 8.  One pipeline stage = 4 units (ks, set) of up to 4 MBW DMMAs; unit u+1's fragments (the other set) are
// loaded while unit u's DMMAs issue.  Shared-memory traffic per DMMA is that of the full mapping.
template <int MBW>
__device__ __forceinline__ void k1_quad_tile(const K1Tile& t, const uint8_t* __restrict__ gbase, uint32_t bar_full,
                                             uint32_t bar_empty, int nk, int warp, int lane, double* __restrict__ W0,
                                             double* __restrict__ W1, long long ldw, unsigned long long w_policy)
{
    const int gid = lane >> 2, tig = lane & 3;
    const int pg = pi8(gid);
    const int r = warp >> 2, c = warp & 3;
    const int mb0 = r * MBW;
    // column blocks of this warp inside each integral block
    const int n1 = min(2, max(0, t.nb1 - 2 * c)), n2 = min(2, max(0, t.nb2 - 2 * c));
    const uint32_t a_off[2] = {(uint32_t)((8 * mb0 + pg) * 128 + (((tig) ^ pg) << 4)),
                               (uint32_t)((8 * mb0 + pg) * 128 + (((4 + tig) ^ pg) << 4))};
    uint32_t b_off[2][2];   // [j][e]: column block 2c + j (parity j), K row 2 tig + e
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int kr = 2 * tig + e;
            const int chunk = 4 * j + (gid >> 1);
            b_off[j][e] = (uint32_t)(kr * 128 + ((chunk ^ kr) << 4) + (gid & 1) * 8 + c * 2048);
        }
    double acc[2][MBW][2][2];   // [set][mb][j][half]
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int m = 0; m < MBW; ++m)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[s][m][j][0] = acc[s][m][j][1] = 0.0;
    double2 fa[2][MBW];     // [set][mb]
    double fb[2][2][2];     // [set][e][j]
    auto load_unit = [&](const uint8_t* st, int u) {
        const int ks = u >> 1, set = u & 1;
        const int n = set ? n2 : n1;
        if (n > 0) {
#pragma unroll
            for (int m = 0; m < MBW; ++m)
                fa[set][m] = *reinterpret_cast<const double2*>(st + set * 16384 + a_off[ks] + m * 1024);
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    if (j < n) fb[set][e][j] = *reinterpret_cast<const double*>(st + 32768 + set * 8192 + b_off[j][e] + ks * 1024);
        }
    };
    auto compute_unit = [&](int u) {
        const int set = u & 1;
        const int n = set ? n2 : n1;
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (j < n) {
#pragma unroll
                    for (int m = 0; m < MBW; ++m) {
                        const double a = e ? fa[set][m].y : fa[set][m].x;
                        if (set) dmma884(acc[1][m][j][0], acc[1][m][j][1], a, fb[1][e][j]);
                        else dmma884(acc[0][m][j][0], acc[0][m][j][1], a, fb[0][e][j]);
                    }
                }
    };
    mbar_wait(bar_full, 0);
    load_unit(gbase, 0);
    for (int it = 0; it < nk; ++it) {
        const int s = it % kK1Stages;
        const uint8_t* st = gbase + s * kK1StageBytes;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (u < 3) {
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
    double* const Wd1 = t.ws1 == 0 ? W0 : (t.ws1 == 1 ? W1 : nullptr);
    double* const Wd2 = t.ws2 == 0 ? W0 : (t.ws2 == 1 ? W1 : nullptr);
    const int nbo2 = (t.p21 != 0.0) ? t.nb1 : t.nb2;
#pragma unroll
    for (int m = 0; m < MBW; ++m) {
        const long long row = (long long)t.xb * kTileM + 8 * (mb0 + m) + pg;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int nb = 2 * c + j;
            const int col = 8 * nb + 2 * tig;
            if (Wd1 != nullptr && nb < t.nb1) {
                double2 o;
                o.x = t.p11 * acc[0][m][j][0] + t.p12 * acc[1][m][j][0];
                o.y = t.p11 * acc[0][m][j][1] + t.p12 * acc[1][m][j][1];
                st_hint_v2f64(Wd1 + row * ldw + t.col1 + col, o, w_policy);
            }
            if (Wd2 != nullptr && nb < nbo2) {
                double2 o;
                o.x = t.p21 * acc[0][m][j][0] + t.p22 * acc[1][m][j][0];
                o.y = t.p21 * acc[0][m][j][1] + t.p22 * acc[1][m][j][1];
                st_hint_v2f64(Wd2 + row * ldw + t.col2 + col, o, w_policy);
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads, 1)
k1_form(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapA2,
        const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
        const K1Tile* __restrict__ tiles,
        const double* __restrict__ ea0, const double* __restrict__ ea1,
        double* __restrict__ W0, double* __restrict__ W1, double* __restrict__ E0,
        long long ldw, int nk, unsigned long long w_policy)
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
            const CUtensorMap* mA1 = t.asel1 ? &mapA2 : &mapA;
            const CUtensorMap* mA2 = t.asel2 ? &mapA2 : &mapA;
            tma_prefetch_desc(mA1);
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
                if (nb1) tma_load_3d(st, mA1, q0, x0, i1, fb);
                if (nb2) tma_load_3d(st + 16384, mA2, q0, x0, i2, fb);
                for (int g = 0; g < ng1; ++g) tma_load_3d(st + 32768 + g * 2048, mB1, a1 + 16 * g, j1, q0, fb);
                for (int g = 0; g < ng2; ++g) tma_load_3d(st + 40960 + g * 2048, mB2, a2 + 16 * g, j2, q0, fb);
            }
        }
        return;
    }
    reg_alloc_consumer();

    // ===== consumers: 8 warps x (16 rows x 64 cols x 2 blocks), or the thin-tile mapping =====
    const int gid = lane >> 2, tig = lane & 3;
    const int pg = pi8(gid);
    if (t.mbv > 0) {
        switch (t.mbv) {
#define AGF2_K1_THIN(M) case M: k1_thin_tile<M>(t, gbase, bar_full, bar_empty, nk, warp, lane, W0, W1, ldw, w_policy); break;
            AGF2_K1_THIN(1) AGF2_K1_THIN(2) AGF2_K1_THIN(3) AGF2_K1_THIN(4) AGF2_K1_THIN(5)
            AGF2_K1_THIN(6) AGF2_K1_THIN(7) AGF2_K1_THIN(8) AGF2_K1_THIN(9) AGF2_K1_THIN(10)
#undef AGF2_K1_THIN
            case 11: case 12: k1_quad_tile<6>(t, gbase, bar_full, bar_empty, nk, warp, lane, W0, W1, ldw, w_policy); break;
            case 13: case 14: k1_quad_tile<7>(t, gbase, bar_full, bar_empty, nk, warp, lane, W0, W1, ldw, w_policy); break;
            default: __trap();
        }
        k1_write_weights(t, warp, gid, tig, ea0, ea1, E0);
        return;
    }
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
                st_hint_v2f64(Wd1 + row * ldw + t.col1 + c, o, w_policy);
            }
            const int nbo2 = (t.p21 != 0.0) ? nb1 : nb2;   // width of out2: block 1's if it mixes block 1
            if (Wd2 != nullptr && nb < nbo2) {
                double2 o;
                o.x = t.p21 * acc1[mb][nb][0] + t.p22 * acc2[mb][nb][0];
                o.y = t.p21 * acc1[mb][nb][1] + t.p22 * acc2[mb][nb][1];
                st_hint_v2f64(Wd2 + row * ldw + t.col2 + c, o, w_policy);
            }
        }
    }
    k1_write_weights(t, warp, gid, tig, ea0, ea1, E0);
}

// ---- K2 ---------------------------------------------------------------------------------------
// One kernel, three uses (all: acc_vv += A' B^T, acc_vev += A'' B^T on 128 x 64 output tiles, K streamed through a
// TMA/mbarrier ring, stream-K over the SMs):
//   moments   <false,false>  A = B = W (workspace chunk), A'' = A diag(w1), w1 = E          (optragf2.c:206-290)
//   M build   <false,true >  A = B = (Q|ja) with rows Q and K = (j, a): A' = A diag(w0), A'' = A diag(w1);
//                            M0 = sum_ja w0 q q^T, M1 = sum_ja w0 (e_j - e_a) q q^T           (Coulomb factorisation)
//   Coulomb   <true ,false>  A = T0 = ixq_i M0, A1 = T1 = ixq_i M1, B = ixq_i, K = (i, Q):
//                            vv += T0 B^T, vev += (e_i T0 + T1) B^T, w1 = e_i
// Operands are 3-D tensor maps (k, row, outer) -- or (k, outer, row) when swap is set -- with 16 x 64 x 1 boxes;
// the K iteration `kit` addresses outer = kit / nk_inner, k = 16 (kit % nk_inner); weights are indexed 16 kit + k.
struct K2Tile {
    int xb, yb;   // rows [a_row_step*xb, +128) of operand A, rows [64*yb, +64) of operand B
    int zb;       // batch index (general GEMM launches only: outer TMA coordinate of the batched operands)
    int pad_;
};

struct K2Params {
    const double* w0;        // per-column weights of the vv operand (kW0 only)
    const double* w1;        // per-column weights of the vev operand
    const K2Tile* tiles;
    const int* cost_prefix;  // [ntiles + 1] cumulative tile cost (sixteenths of a full tile)
    double* acc_vv;
    double* acc_vev;
    long long ldacc;
    double* partial;         // stream-K scratch, gridDim.x slots of 2 x 128 x 64 doubles (nullptr: every partial
                             // tile goes straight into the accumulators with red.global.add -- order not fixed)
    int ntiles, nk_inner, n_outer;
    int swap_a, swap_b;      // operand maps are (k, outer, row) instead of (k, row, outer)
    int a_row_off;
    int b_k_off, b_outer_off, b_row_off;
    int rows_valid, cols_valid;   // logical extent of the output (masks out padded warps / column blocks)
    int sym;                 // only the blocks touching the upper triangle are computed
    int mask;                // 0: every warp computes all 8 column blocks of every tile; 1: warps with nothing
                             // to do on an edge tile skip it
    // ---- general GEMM launches (kGen instantiation only) ----
    int a_row_step;          // rows of operand A per unit of xb: 128, or 256 when the second accumulator is the
                             // product of rows +128 of the same operand (plain C = A B^T on 256 x 64 tiles)
    int transpose_out;       // output element (row, col) lives at acc[col * ldacc + row]
    int store;               // the segment that starts a tile stores instead of accumulating (no memset needed)
    int a_batch, b_batch;    // add the tile's zb to the outer coordinate of operand A / B
    long long out_batch_stride;   // output offset per unit of zb
    int vev_rows_valid;      // valid rows of the second accumulator (a_row_step == 256: rows_valid - 128 - 256 xb)
};

constexpr int kK2StageBytes = kTileM * kKB * 8 + kTileN * kKB * 8 + 1024;  // A 16K + B 8K + weights (2 x 128 B, padded)
constexpr int kK2StageBytesDual = kK2StageBytes + kTileM * kKB * 8;        // + A1 16K
constexpr int kK2Stages = 7, kK2StagesDual = 5;
constexpr int kK2SmemBytes = kK2Stages * kK2StageBytes + 1024 + 256;
constexpr int kK2SmemBytesDual = kK2StagesDual * kK2StageBytesDual + 1024 + 256;

// Column blocks [lo, hi) of a tile that warp `warp` (rows 16 warp .. +15 of the tile) has to compute.
__host__ __device__ inline void k2_warp_range(int xb, int yb, int warp, int rows_valid, int cols_valid, int sym,
                                              int mask, int& lo, int& hi)
{
    if (!mask) { lo = 0; hi = 8; return; }
    const int row0 = xb * kTileM + 16 * warp, col0 = yb * kTileN;
    hi = (cols_valid - col0 + 7) / 8;
    hi = hi < 0 ? 0 : (hi > 8 ? 8 : hi);
    lo = 0;
    if (row0 >= rows_valid) hi = 0;
    if (sym && row0 > col0) lo = (row0 - col0) / 8;   // blocks entirely below the diagonal are skipped
    if (lo >= hi) lo = hi = 0;
    else { lo = 0; hi = 8; }      // a warp either computes all 8 column blocks or skips the tile
}
// Time of a tile in sixteenths of a full one (what stream-K balances): the slowest SM sub-partition (warps w and
// w + 4 share one; a fully masked warp branches over the unit).  A sub-partition with ONE active warp does not run at
// half the time of one with two: a single warp issues DMMA.8x8x4 at ~57 % of the pipe rate (measured: the SM whose
// stream-K range held the diagonal tiles executed 62 % of the average flops in 111 % of the average time,
// profiles/r02_k2_moment_ncu.csv), so it costs 14, not 8.
__host__ __device__ inline int k2_tile_cost(int xb, int yb, int rows_valid, int cols_valid, int sym, int mask)
{
    int worst = 0;
    for (int s = 0; s < 4; ++s) {
        int active = 0;
        for (int w = s; w < 8; w += 4) {
            int lo, hi;
            k2_warp_range(xb, yb, w, rows_valid, cols_valid, sym, mask, lo, hi);
            active += hi > lo ? 1 : 0;
        }
        const int t = active == 2 ? 16 : (active == 1 ? 14 : 0);
        worst = t > worst ? t : worst;
    }
    return worst < 4 ? 4 : worst;   // shared-memory traffic and barriers do not shrink with the mask
}
// DMMA work of a tile in sixteenths of a full one (executed-flop accounting): two per active warp
__host__ __device__ inline int k2_tile_work(int xb, int yb, int rows_valid, int cols_valid, int sym, int mask)
{
    int active = 0;
    for (int w = 0; w < 8; ++w) {
        int lo, hi;
        k2_warp_range(xb, yb, w, rows_valid, cols_valid, sym, mask, lo, hi);
        active += hi > lo ? 1 : 0;
    }
    return 2 * active;
}

// stream-K boundary b of G: position (tile, kit) at cost  total * b / G
__device__ __forceinline__ void k2_boundary(const int* __restrict__ prefix, int ntiles, int nk, int b, int G,
                                            int& tile, int& kit)
{
    const long long total = (long long)prefix[ntiles] * nk;
    const long long u = total * b / G;
    int lo = 0, hi = ntiles;   // largest t with prefix[t] * nk <= u
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((long long)prefix[mid] * nk <= u) lo = mid; else hi = mid - 1;
    }
    tile = lo;
    kit = 0;
    if (tile < ntiles) kit = (int)((u - (long long)prefix[tile] * nk) / (prefix[tile + 1] - prefix[tile]));
}

// Address of output element (r, c) of tile t (r in [0,128), c in [0,64)) of accumulator `which` (0: vv, 1: vev);
// general launches: nullptr when the element lies outside the logical output.
template <bool kGen>
__device__ __forceinline__ double* k2_out_ptr(const K2Params& p, const K2Tile& t, int r, int c, int which)
{
    double* base = which ? p.acc_vev : p.acc_vv;
    if (!kGen) return base + ((long long)t.xb * kTileM + r) * p.ldacc + (long long)t.yb * kTileN + c;
    const long long row = (long long)t.xb * p.a_row_step + r;
    const long long col = (long long)t.yb * kTileN + c;
    if (row >= (which ? p.vev_rows_valid : p.rows_valid) || col >= p.cols_valid) return nullptr;
    base += (long long)t.zb * p.out_batch_stride;
    return p.transpose_out ? base + col * p.ldacc + row : base + row * p.ldacc + col;
}

// kGen: general GEMM launches (guarded / transposed / batched output, 256-row plain-GEMM tiles) -- instantiated
// with kDual only; the moment launches keep the unguarded epilogue on padded accumulators.
template <bool kDual, bool kW0, bool kGen = false>
__global__ void __launch_bounds__(kThreads, 1)
k2_moment(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
          const __grid_constant__ CUtensorMap mapB, const __grid_constant__ K2Params p)
{
    constexpr int kStages = kDual ? kK2StagesDual : kK2Stages;
    constexpr int kStageBytes = kDual ? kK2StageBytesDual : kK2StageBytes;
    constexpr uint32_t kOffB = 16384, kOffW = 24576, kOffA1 = 25600;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_full = base + kStages * kStageBytes;
    const uint32_t bar_empty = bar_full + 8 * kStages;
    uint8_t* const gbase = smem_raw + (base - smem_u32(smem_raw));

    // stream-K: the (tile, k-iteration) space, weighted by tile cost, is cut into gridDim.x equal ranges so
    // that every SM gets the same amount of DMMA work whatever the tile count.  A CTA flushes its accumulators
    // whenever its range crosses a tile boundary: the segment that STARTS a tile (k = 0) adds into the
    // accumulators (one adder per element and launch), every other segment -- at most one per CTA, its first --
    // goes to the CTA's slot of the `partial` scratch and k2_fixup adds the slots in CTA order afterwards, so
    // that the summation order is fixed (bit-reproducible moments).
    const int nk_total = p.nk_inner * p.n_outer;
    int tile, kit, tile_end, kit_end;
    k2_boundary(p.cost_prefix, p.ntiles, nk_total, blockIdx.x, gridDim.x, tile, kit);
    k2_boundary(p.cost_prefix, p.ntiles, nk_total, blockIdx.x + 1, gridDim.x, tile_end, kit_end);
    const int nk = (tile_end - tile) * nk_total + kit_end - kit;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
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
            tma_prefetch_desc(&mapA0);
            tma_prefetch_desc(&mapB);
            if (kDual) tma_prefetch_desc(&mapA1);
            constexpr uint32_t bytes = kTileM * kKB * 8 + kTileN * kKB * 8 + (kGen ? 0 : kKB * 8) +
                                       (kW0 ? kKB * 8 : 0) + (kDual ? kTileM * kKB * 8 : 0);
            K2Tile t = p.tiles[tile];
            int outer = kit / p.nk_inner, ki = kit % p.nk_inner;
            for (int j = 0; j < nk; ++j) {
                const int s = j % kStages;
                if (j >= kStages) mbar_wait(bar_empty + 8 * s, ((j / kStages) - 1) & 1);
                const uint32_t st = base + s * kStageBytes;
                const uint32_t fb = bar_full + 8 * s;
                mbar_expect_tx(fb, bytes);
                const int k0 = ki * kKB;
                const int ra = t.xb * (kGen ? p.a_row_step : kTileM) + p.a_row_off;
                const int rb = t.yb * kTileN + p.b_row_off;
                const int oa = outer + ((kGen && p.a_batch) ? t.zb : 0);
                const int ob = outer + p.b_outer_off + ((kGen && p.b_batch) ? t.zb : 0);
                if (p.swap_a) {
                    tma_load_3d(st, &mapA0, k0, oa, ra, fb);
                    tma_load_3d(st + 8192, &mapA0, k0, oa, ra + 64, fb);
                } else {
                    tma_load_3d(st, &mapA0, k0, ra, oa, fb);
                    tma_load_3d(st + 8192, &mapA0, k0, ra + 64, oa, fb);
                }
                if (kDual) {
                    if (p.swap_a) {
                        tma_load_3d(st + kOffA1, &mapA1, k0, oa, ra, fb);
                        tma_load_3d(st + kOffA1 + 8192, &mapA1, k0, oa, ra + 64, fb);
                    } else {
                        tma_load_3d(st + kOffA1, &mapA1, k0, ra, oa, fb);
                        tma_load_3d(st + kOffA1 + 8192, &mapA1, k0, ra + 64, oa, fb);
                    }
                }
                if (p.swap_b) tma_load_3d(st + kOffB, &mapB, k0 + p.b_k_off, ob, rb, fb);
                else tma_load_3d(st + kOffB, &mapB, k0 + p.b_k_off, rb, ob, fb);
                const long long wi = ((long long)outer * p.nk_inner + ki) * kKB;
                if (!kGen) bulk_load_1d(st + kOffW, p.w1 + wi, kKB * 8, fb);   // (general GEMM: no weights)
                if (kW0) bulk_load_1d(st + kOffW + 128, p.w0 + wi, kKB * 8, fb);
                if (++ki == p.nk_inner) {
                    ki = 0;
                    if (++outer == p.n_outer) { outer = 0; if (j + 1 < nk) t = p.tiles[++tile]; }
                }
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
    double2 fa[2][2], fe[2][2];   // [ks parity][mb]: fragments of the vv operand and of the vev operand
    double2 fb[2][4];             // [half][j]: B fragments of n-blocks 4*half + j
    auto load_unit = [&](const uint8_t* st, int u) {
        const int ks = u >> 1, half = u & 1;
        if (half == 0) {
            double2 ev = make_double2(0.0, 0.0);
            if (!kGen) ev = *reinterpret_cast<const double2*>(st + kOffW + (4 * ks + tig) * 16);
            double2 ev0;
            if (kW0) ev0 = *reinterpret_cast<const double2*>(st + kOffW + 128 + (4 * ks + tig) * 16);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                double2 a = *reinterpret_cast<const double2*>(st + a_row + mb * 1024 + sw[ks]);
                if (kDual) {
                    const double2 a1 = *reinterpret_cast<const double2*>(st + kOffA1 + a_row + mb * 1024 + sw[ks]);
                    if (kGen) {
                        fe[ks][mb] = a1;
                    } else {
                        fe[ks][mb].x = fma(a.x, ev.x, a1.x);
                        fe[ks][mb].y = fma(a.y, ev.y, a1.y);
                    }
                } else {
                    fe[ks][mb].x = a.x * ev.x;
                    fe[ks][mb].y = a.y * ev.y;
                }
                if (kW0) { a.x *= ev0.x; a.y *= ev0.y; }
                fa[ks][mb] = a;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            fb[half][j] = *reinterpret_cast<const double2*>(st + kOffB + b_row + (4 * half + j) * 1024 + sw[ks]);
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

    K2Tile t = p.tiles[tile];
    int lo, hi;
    k2_warp_range(t.xb, t.yb, warp, p.rows_valid, p.cols_valid, p.sym, kGen ? 0 : p.mask, lo, hi);
    mbar_wait(bar_full, 0);
    load_unit(gbase, 0);
    // One iteration = one pipeline stage.  kMode 0: this warp computes the tile, 2: it has nothing to compute on
    // this (edge) tile and only keeps the pipeline moving.
    auto iteration = [&](int it, auto mode) {
        constexpr int kMode = decltype(mode)::value;
        const int s = it % kStages;
        const uint8_t* st = gbase + s * kStageBytes;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (kMode != 2) {
                if (u < 3) {
                    load_unit(st, u + 1);
                } else if (it + 1 < nk) {
                    const int s2 = (it + 1) % kStages;
                    mbar_wait(bar_full + 8 * s2, ((it + 1) / kStages) & 1);
                    load_unit(gbase + s2 * kStageBytes, 0);
                }
                compute_unit(u);
            }
        }
        if (kMode == 2 && it + 1 < nk) {
            const int s2 = (it + 1) % kStages;
            mbar_wait(bar_full + 8 * s2, ((it + 1) / kStages) & 1);
            load_unit(gbase + s2 * kStageBytes, 0);   // the next tile may need this warp again
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);
    };
    int it = 0;
    while (it < nk) {
        const int seg_end = min(nk, it + nk_total - kit);   // this tile's share of the range
        const bool owner = kit == 0;                        // this segment starts the tile
        if (hi > lo) { for (; it < seg_end; ++it) iteration(it, std::integral_constant<int, 0>()); }
        else { for (; it < seg_end; ++it) iteration(it, std::integral_constant<int, 2>()); }
        if (!owner && p.partial != nullptr) {
            // partial tile: fragments in register order into this CTA's slot (coalesced); masked blocks as zeros
            double* const slot = p.partial + (size_t)blockIdx.x * (2 * kTileM * kTileN) + lane;
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int nb = 0; nb < 8; ++nb) {
                    const bool on = nb >= lo && nb < hi;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int idx = (((warp * 2 + mb) * 8 + nb) * 2 + h) * 32;
                        slot[idx] = on ? cvv[mb][nb][h] : 0.0;
                        slot[kTileM * kTileN + idx] = on ? cvev[mb][nb][h] : 0.0;
                        cvv[mb][nb][h] = 0.0;
                        cvev[mb][nb][h] = 0.0;
                    }
                }
        } else {
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                const int r = 16 * warp + 8 * mb + pg;
#pragma unroll
                for (int nb = 0; nb < 8; ++nb)
                    if (nb >= lo && nb < hi) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int c = 8 * nb + pi8(2 * tig + h);
                            double* const d0 = k2_out_ptr<kGen>(p, t, r, c, 0);
                            double* const d1 = k2_out_ptr<kGen>(p, t, r, c, 1);
                            if (kGen) {
                                if (p.store && owner) {
                                    if (d0) *d0 = cvv[mb][nb][h];
                                    if (d1) *d1 = cvev[mb][nb][h];
                                } else {
                                    if (d0) red_add_f64(d0, cvv[mb][nb][h]);
                                    if (d1) red_add_f64(d1, cvev[mb][nb][h]);
                                }
                            } else {
                                red_add_f64(d0, cvv[mb][nb][h]);
                                red_add_f64(d1, cvev[mb][nb][h]);
                            }
                            cvv[mb][nb][h] = 0.0;
                            cvev[mb][nb][h] = 0.0;
                        }
                    }
            }
        }
        kit = 0;
        if (it < nk) {
            t = p.tiles[++tile];
            k2_warp_range(t.xb, t.yb, warp, p.rows_valid, p.cols_valid, p.sym, kGen ? 0 : p.mask, lo, hi);
        }
    }
}

// Adds the partial-tile slots of a k2_moment launch (grid G) into the accumulators, slot after slot in CTA order.
// CTA b of this kernel handles the tile CTA b of k2_moment started in mid-tile, unless CTA b - 1 did so too (then
// that CTA's fixup block walks all the slots of the tile).
constexpr int kFixupParts = 8;     // CTAs per tile of k2_fixup (grid.y): 1024 output elements each
template <bool kGen>
__global__ void __launch_bounds__(256)
k2_fixup(const __grid_constant__ K2Params p, int G)
{
    __shared__ int s_slots[160];
    __shared__ int s_n;
    const int nk_total = p.nk_inner * p.n_outer;
    const int b = blockIdx.x;
    int tile, kit;
    k2_boundary(p.cost_prefix, p.ntiles, nk_total, b, G, tile, kit);
    if (kit == 0 || tile >= p.ntiles) return;
    if (b > 0) {
        int tp, kp;
        k2_boundary(p.cost_prefix, p.ntiles, nk_total, b - 1, G, tp, kp);
        if (tp == tile && kp > 0) return;      // CTA b - 1 walks this tile's slots
    }
    if (threadIdx.x == 0) {
        int n = 0, tc = tile, kc = kit;
        for (int s = b; s < G; ++s) {           // slots of the CTAs that started inside this tile, in CTA order
            int tn, kn;
            k2_boundary(p.cost_prefix, p.ntiles, nk_total, s + 1, G, tn, kn);
            if ((tn != tc || kn != kc) && n < 160) s_slots[n++] = s;   // (an empty range wrote no slot)
            if (tn != tile || kn == 0) break;
            tc = tn; kc = kn;
        }
        s_n = n;
    }
    __syncthreads();
    const int n = s_n;
    const K2Tile t = p.tiles[tile];
    // 4 elements x 2 accumulators per thread, loads of one slot issued together (the kernel is latency bound)
    constexpr int kPer = kTileM * kTileN / kFixupParts / 256;
    double* d[2 * kPer];
    double v[2 * kPer];
    int off[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int idx = blockIdx.y * (kTileM * kTileN / kFixupParts) + j * 256 + threadIdx.x;
        const int lane = idx & 31, h = (idx >> 5) & 1, nb = (idx >> 6) & 7, mb = (idx >> 9) & 1, warp = idx >> 10;
        const int r = 16 * warp + 8 * mb + pi8(lane >> 2);
        const int c = 8 * nb + pi8(2 * (lane & 3) + h);
        off[j] = idx;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            d[2 * j + which] = k2_out_ptr<kGen>(p, t, r, c, which);
            v[2 * j + which] = d[2 * j + which] ? *d[2 * j + which] : 0.0;
        }
    }
    for (int k = 0; k < n; ++k) {
        const double* slot = p.partial + (size_t)s_slots[k] * (2 * kTileM * kTileN);
#pragma unroll
        for (int j = 0; j < kPer; ++j)
#pragma unroll
            for (int which = 0; which < 2; ++which) v[2 * j + which] += slot[which * (kTileM * kTileN) + off[j]];
    }
#pragma unroll
    for (int q = 0; q < 2 * kPer; ++q)
        if (d[q]) *d[q] = v[q];
}

// ---- per-column weights of the Coulomb factorisation ---------------------------------------------
// M build: w0[j, a] = w_in (j in [istart, iend)) or w_out, 0 in the K padding; w1 = w0 (e_j - e_a)
__global__ void fill_m_weights(double* __restrict__ w0, double* __restrict__ w1, const double* __restrict__ ej,
                               const double* __restrict__ ea, int v, int kpad, double w_in, double w_out,
                               int istart, int iend)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (a >= kpad) return;
    const double w = (a < v) ? ((j >= istart && j < iend) ? w_in : w_out) : 0.0;
    w0[(long long)j * kpad + a] = w;
    w1[(long long)j * kpad + a] = (a < v) ? w * (ej[j] - ea[a]) : 0.0;
}
// M0, M1 (n x n, pitch ld) were accumulated on their upper triangles: fill the lower ones
__global__ void mirror_m(double* __restrict__ m0, double* __restrict__ m1, int n, long long ld)
{
    const int c = blockIdx.y * blockDim.x + threadIdx.x, r = blockIdx.x;   // rows on grid.x: naux may exceed 65535
    if (c < r && c < n) {
        m0[(long long)r * ld + c] = m0[(long long)c * ld + r];
        m1[(long long)r * ld + c] = m1[(long long)c * ld + r];
    }
}
// Coulomb step: we[il, k] = e_{i0 + il}
__global__ void fill_e_weights(double* __restrict__ we, const double* __restrict__ ei, int i0, int kpad)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x, il = blockIdx.y;
    if (k < kpad) we[(long long)il * kpad + k] = ei[i0 + il];
}

// ---- K3: finalize ------------------------------------------------------------------------------
// out[x][y] += sum over accumulator sets s (one per stream, fixed order) of
//              (x <= y ? S_s[x][y] : S_s[y][x]) + A_s[x][y]
// (S: symmetric-chunk accumulator, upper triangle valid; A: accumulator of the non-symmetric chunks).
// Set s lives at acc + 4 s n: [S_vv | S_vev | A_vv | A_vev], n = ldacc^2.
__global__ void k3_finalize(const double* __restrict__ acc, long long n, int nsets, int has_asym,
                            long long ldacc, int nphys, double* __restrict__ vv, double* __restrict__ vev)
{
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    const int x = blockIdx.y;
    if (y >= nphys) return;
    const long long u = (x <= y) ? (x * ldacc + y) : (y * ldacc + x);
    double r0 = 0.0, r1 = 0.0;
    for (int s = 0; s < nsets; ++s) {
        const double* a = acc + 4 * s * n;
        r0 += a[u];
        r1 += a[n + u];
        if (has_asym) {
            r0 += a[2 * n + x * ldacc + y];
            r1 += a[3 * n + x * ldacc + y];
        }
    }
    vv[(long long)x * nphys + y] += r0;
    vev[(long long)x * nphys + y] += r1;
}

// vv += packed[0 : n], vev += packed[n : 2n]   (after the all-reduce of the packed moments)
__global__ void add_packed(const double* __restrict__ packed, long long n, double* __restrict__ vv, double* __restrict__ vev)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { vv[k] += packed[k]; vev[k] += packed[n + k]; }
}

// Tile list of a general GEMM launch, generated on the device: tile t = (xb, yb, zb) with xb fastest, uniform cost.
__global__ void gen_gemm_tiles(K2Tile* __restrict__ tiles, int* __restrict__ cost_prefix, int nxb, int nyb, long long ntiles)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    cost_prefix[t] = (int)(16 * t);
    if (t == ntiles) return;
    K2Tile k;
    k.xb = (int)(t % nxb);
    k.yb = (int)((t / nxb) % nyb);
    k.zb = (int)(t / ((long long)nxb * nyb));
    k.pad_ = 0;
    tiles[t] = k;
}

// ---- naive debug kernels (same tile lists, no tensor cores, no TMA) ------------------------------
__global__ void k1_form_simple(const double* __restrict__ ixq, const double* __restrict__ ixq2, long long ld_ixq, int nphys, int naux,
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
                        s1 += (t.asel1 ? ixq2 : ixq)[((long long)t.i1 * nphys + x) * ld_ixq + k] *
                              q[((long long)k * no + t.j1) * ldq + t.a1 + c];
            }
            if (c < 8 * t.nb2) {
                const double* q = t.bsel2 ? qja1 : qja0;
                const long long ldq = t.bsel2 ? ld_q1 : ld_q0;
                const int no = t.bsel2 ? nocc1 : nocc0, nv = t.bsel2 ? nvir1 : nvir0;
                if (t.a2 + c < nv)
                    for (int k = 0; k < naux; ++k)
                        s2 += (t.asel2 ? ixq2 : ixq)[((long long)t.i2 * nphys + x) * ld_ixq + k] *
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
