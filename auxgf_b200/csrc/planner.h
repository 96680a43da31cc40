// Work planner of the moment build -- pure host code, no CUDA calls, so that it can be exercised without a GPU
// (agf2_b200_plan_stats, tests/test_planner.py).  Replaces the `for i` loop of auxgf/lib/libagf2/optragf2.c:118 and
// its MPI split in auxgf/mpi/agf2/optragf2.py:370-374:
//   work items   PAIR(i>j)    X_ij, X_ji          -> sqrt(ss) (X_ij - X_ji)   [+ sqrt(os/2) (X_ij + X_ji) on the all-pairs route]
//                DIAG(i)      X_ii                -> sqrt(os) X_ii            (all-pairs route only)
//                OS(i,J)      X_iJ (other spin)   -> sqrt(os) X_iJ            (all-pairs route only, UHF)
//                ASYM(i,j)    j outside [istart,iend): U = X_ij -> W0, V = (c X_ij - s X_ji) -> W1
//   items are dealt round-robin to `nparts` (one per GPU), kind by kind; each part packs its items into chunks whose
//   column blocks fit the L2-sized workspace and whose CTA tiles fill whole waves of SMs; per chunk: k1_form
//   (integrals -> W), k2_moment (W -> vv, vev).  Coulomb-type terms go through the M0/M1 factorisation when it pays.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "moment_kernels.cuh"

namespace agf2 {

struct Item { int kind, i, j; };   // kind: 0 DIAG, 1 PAIR, 2 OS, 3 ASYM
struct SubTile { int i, j, a0, nb, bsel, esel, nv; long long col; double p, eij; };

inline int64_t rup(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// Column blocks (a0, nb) of one item of padded width w: ceil(w/64) blocks of 64 columns + the rest, e.g. 104 -> 64 + 40
// (with the longest-first launch order this measured no worse than nearly equal blocks).
inline std::vector<std::pair<int, int>> split_blocks(int64_t w) {
    std::vector<std::pair<int, int>> out;
    const int units = (int)(w / 8);
    const int nblk = (units + 7) / 8;
    int a0 = 0;
    for (int b = 0; b < nblk; ++b) {
        const int nb = std::min(8, units - 8 * b);
        out.push_back({a0, nb});
        a0 += 8 * nb;
    }
    return out;
}

// Greedy in-order dispatch of CTA tiles (durations in units of one 8-column block) over nsm SMs: the time the
// last SM finishes.  This is how the hardware block scheduler hands out the CTAs of one k1_form launch.
inline double dispatch_makespan(const std::vector<double>& durs, int nsm) {
    std::vector<double> h((size_t)nsm, 0.0);   // min-heap of SM finish times
    auto cmp = [](double a, double b) { return a > b; };
    for (double d : durs) {
        std::pop_heap(h.begin(), h.end(), cmp);
        h.back() += d;
        std::push_heap(h.begin(), h.end(), cmp);
    }
    return *std::max_element(h.begin(), h.end());
}

// How many identical work items one chunk should hold (<= n_max, what the workspace allows) so that the tiles
// of the k1_form launch fill an integer number of waves: maximises work / (makespan + launch overhead).
inline int pick_chunk_items(const std::vector<double>& item_tiles, int n_max, int nsm, bool lpt, double launch_ovh) {
    if (n_max <= 1) return 1;
    int best_n = n_max;
    double best = -1.0;
    const int lo = std::max(1, n_max / 2), step = std::max(1, (n_max - lo) / 48);
    double per_item = 0.0;
    for (double d : item_tiles) per_item += d;
    for (int n = n_max; n >= lo; n -= step) {
        std::vector<double> durs;
        durs.reserve(item_tiles.size() * (size_t)n);
        for (int k = 0; k < n; ++k) durs.insert(durs.end(), item_tiles.begin(), item_tiles.end());
        if (lpt) std::stable_sort(durs.begin(), durs.end(), [](double a, double b) { return a > b; });
        const double eff = per_item * n / nsm / (dispatch_makespan(durs, nsm) + launch_ovh);
        if (eff > best * 1.005) { best = eff; best_n = n; }   // prefer the larger chunk unless clearly worse
    }
    return best_n;
}

// ---- pole-sharded builds --------------------------------------------------------------------------
// (ix|Q) is split over the R GPUs by contiguous pole blocks, block r = [o r / R, o (r+1) / R); (Q|ja) is replicated.
// Rank r forms every pair inside its own block (step t = 0) and, for t = 1 .. R-1, the pairs (i in its block, j in
// the block of rank src = r - t) whose index sum has the parity assigned to it -- the lower rank of a block pair
// takes the even sums, the higher one the odd sums -- while the poles of block src visit in sub-blocks of `sb`
// poles (sent by their owner over NVLink, double-buffered under the pair loop).  Every unordered pair is formed
// exactly once and every rank does the same amount of work without any rank holding more than its own block plus
// two sub-blocks.  Replaces the replicated inputs of auxgf/mpi/agf2/optragf2.py:370-374 for shapes whose (ix|Q)
// does not fit one GPU.
inline int64_t shard_lo(int64_t o, int64_t r, int64_t R) { return o * r / R; }

struct ShardStep {
    int t = 0, k = 0;             // ring step, sub-block index within the step
    int src = 0, dst = 0;         // owner of the visiting block, receiver of this rank's block
    int64_t vis_lo = 0, vis_hi = 0;     // visiting poles (global indices; t = 0: a sub-range of the own block)
    int64_t send_lo = 0, send_hi = 0;   // own poles sent to dst in this step (empty: nothing to send)
    bool recv = false, send = false;
};

// limit > 0: only the first `limit` sub-blocks of every ring step (a benchmark sample of a build; the same on every rank)
inline std::vector<ShardStep> shard_steps(int64_t o, int64_t rank, int64_t R, int64_t sb, int64_t limit = 0) {
    std::vector<ShardStep> out;
    const int64_t my_lo = shard_lo(o, rank, R), my_hi = shard_lo(o, rank + 1, R);
    for (int64_t t = 0; t < R; ++t) {
        const int64_t src = (rank - t + R) % R, dst = (rank + t) % R;
        const int64_t s_lo = shard_lo(o, src, R), s_hi = shard_lo(o, src + 1, R);
        const int64_t k_recv = (s_hi - s_lo + sb - 1) / sb, k_send = t ? (my_hi - my_lo + sb - 1) / sb : 0;
        for (int64_t k = 0; k < std::max(k_recv, k_send); ++k) {
            if (limit > 0 && k >= limit) break;
            ShardStep s;
            s.t = (int)t; s.k = (int)k; s.src = (int)src; s.dst = (int)dst;
            if (k < k_recv) { s.vis_lo = s_lo + k * sb; s.vis_hi = std::min(s_hi, s.vis_lo + sb); s.recv = t > 0; }
            if (k < k_send) { s.send_lo = my_lo + k * sb; s.send_hi = std::min(my_hi, s.send_lo + sb); s.send = true; }
            out.push_back(s);
        }
    }
    return out;
}

struct PlanConfig {
    int64_t ws_bytes = 384ll << 20;  // workspace per ring buffer (upper bound)
    int num_sms = 148;
    bool simple = false;             // naive debug kernels: all-pairs route, no tuning, no thin tiles
    bool lpt_order = true;           // k1_form tiles of a chunk launched longest first
    bool tune_chunk = true;          // chunk size chosen for whole waves of k1_form tiles
    bool thin_tiles = true;          // ragged last x-block: thin-tile (<= 80 rows) or quad-tile (81..112 rows) mapping
    bool coulomb = true;             // Coulomb factorisation allowed ...
    bool coulomb_auto = true;        // ... and chosen per call by a flop estimate
};

struct MomentPlan {
    PlanConfig cfg;
    // the call
    int64_t nphys = 0, naux = 0, o = 0, v = 0, ob = 0, vb = 0, istart = 0, iend = 0, part = 0, nparts = 1;
    bool has_other = false;
    double c_same = 0, s_same = 0, os_other = 0;
    // derived
    int64_t p_pad = 0, vpad = 0, vbpad = 0, cap_cols = 0;
    int nxb = 0, nyb = 0;
    bool fact = false, want_asym = true;
    double ca = 0, cs = 0, cd = 0, co = 0, asym_c = 0;
    std::vector<Item> sym_items, asym_items;
    std::vector<std::pair<int, int>> blk_same, blk_other;
    int chunk_items[4] = {0, 0, 0, 0};   // items per chunk, per kind (0 = not yet chosen)
    const char* error = "";
    // pole-sharded builds: resident block [res_lo, res_hi) (operand map 0, indexed from res_lo), visiting poles
    // [vis_lo, vis_hi) (operand map 1, indexed from vis_lo)
    bool sharded = false;
    int64_t res_lo = 0, res_hi = 0, vis_lo = 0, vis_hi = 0;

    // Same-spin tensor (o poles x v), optional other-spin tensor (ob x vb); coefficients: vv = c C_same - s Xc_same
    // + os_other C_other.  Returns 0, or 3 (AGF2_B200_EINVAL) with `error` set.
    int init(const PlanConfig& config, int64_t nphys_, int64_t naux_, int64_t o_, int64_t v_, bool has_other_,
             int64_t ob_, int64_t vb_, int64_t istart_, int64_t iend_, double c_same_, double s_same_,
             double os_other_, int64_t part_, int64_t nparts_, bool sharded_ = false)
    {
        cfg = config;
        nphys = nphys_; naux = naux_; o = o_; v = v_; has_other = has_other_; ob = has_other_ ? ob_ : 0;
        vb = has_other_ ? vb_ : 0; istart = istart_; iend = iend_; c_same = c_same_; s_same = s_same_;
        os_other = os_other_; part = part_; nparts = nparts_;
        if (nphys <= 0 || naux <= 0 || o <= 0 || v <= 0) { error = "empty dimension"; return 3; }
        if (istart < 0 || iend > o || istart > iend) { error = "bad [istart, iend)"; return 3; }
        if (nparts < 1 || part < 0 || part >= nparts) { error = "bad part/nparts"; return 3; }
        if (c_same - s_same < 0 || c_same + s_same < 0 || os_other < 0) {
            error = "os_factor and ss_factor must be non-negative"; return 3;
        }
        if (nphys > (1 << 30) || naux > (1 << 30) || o * v > (1ll << 40) || o > 65535 || ob > 65535) {
            error = "dimension too large"; return 3;
        }
        p_pad = rup(nphys, kTileM);
        nxb = (int)(p_pad / kTileM);
        nyb = (int)((nphys + kTileN - 1) / kTileN);
        vpad = rup(v, 8);
        vbpad = has_other ? rup(vb, 8) : 0;
        // Coulomb factorisation (default): the pair loop only carries the exchange-type term  s sum_{i>j} A A^T,
        // A = X_ij - X_ji; every Coulomb-type term goes through the M0/M1 route.  Otherwise (debug / simple
        // kernels): pairs carry sqrt(c/2+s/2) A and sqrt(c/2-s/2) S, plus DIAG and OS items.
        fact = cfg.coulomb && !cfg.simple;
        if (fact && cfg.coulomb_auto) {
            // flops the factorisation removes from the pair loop (S blocks, DIAG and OS items) vs the flops it adds
            const double P = (double)nphys, N = (double)naux, ni = (double)(iend - istart), tri = 0.61;
            const double ov = (double)o * (double)v;
            const double obvb = has_other ? (double)ob * (double)vb : 0.0;
            const double k2col = 2.0 * tri * 2.0 * P * P;                       // both moments, per workspace column
            double direct = 0.5 * ni * (ni - 1.0) * v * k2col * (c_same != s_same ? 1.0 : 0.0);   // S blocks
            direct += (c_same != s_same ? ni * ((double)v * k2col + 2.0 * P * v * N) : 0.0);         // DIAG items
            direct += os_other != 0.0 ? ni * (obvb * k2col + 2.0 * P * obvb * N) : 0.0;             // OS items
            const bool partial_slice = istart > 0 || iend < o;
            double mcols = 0.0;                                                  // (j, a) columns of the M build
            if (c_same != s_same || (partial_slice && c_same != 0.0)) mcols += ov;
            if (os_other != 0.0) mcols += obvb;
            const double fwork = (nparts == 1 ? 0.5 : 1.0) * 4.0 * N * N * mcols + ni * (4.0 * P * N * N + k2col * N);
            fact = mcols > 0.0 && fwork < 0.95 * direct;
        }
        sharded = sharded_;
        if (sharded) {     // the pair loop only carries the exchange-type term; items come from set_shard_step()
            if (cfg.simple) { error = "pole-sharded builds need the tensor-core kernels (simple_kernels = 0)"; return 3; }
            fact = true;
        }
        ca = fact ? std::sqrt(s_same) : std::sqrt(0.5 * (c_same + s_same));
        cs = fact ? 0.0 : std::sqrt(0.5 * (c_same - s_same));
        cd = fact ? 0.0 : std::sqrt(c_same - s_same);
        co = fact ? 0.0 : std::sqrt(os_other);
        asym_c = fact ? 0.0 : c_same;
        want_asym = !(fact && s_same == 0.0);
        cap_cols = std::max<int64_t>(cfg.ws_bytes / (p_pad * 8), 2 * std::max(vpad, vbpad));
        cap_cols = rup(cap_cols, 16);

        // ---- items of this part ---------------------------------------------------------------------
        // (kinds are kept contiguous -- PAIR, then DIAG, then OS -- so that the chunks are homogeneous and their
        // size can be tuned to whole waves of CTA tiles; within a kind the order is i-major, which is also the
        // order in which the host-pointer API uploads the poles)
        sym_items.clear();
        asym_items.clear();
        int64_t n_diag = 0, n_pair = 0, n_os = 0, n_asym = 0;
        std::vector<Item> diag_items, os_items;
        for (int64_t i = istart; i < iend && !sharded; ++i) {
            if (cd != 0.0 && (n_diag++ % nparts) == part) diag_items.push_back({0, (int)i, (int)i});
            if (ca != 0.0 || cs != 0.0)
                for (int64_t j = istart; j < i; ++j)
                    if ((n_pair++ % nparts) == part) sym_items.push_back({1, (int)i, (int)j});
            if (has_other && co != 0.0)
                for (int64_t J = 0; J < ob; ++J)
                    if ((n_os++ % nparts) == part) os_items.push_back({2, (int)i, (int)J});
            if (want_asym)
                for (int64_t j = 0; j < o; ++j)
                    if ((j < istart || j >= iend) && (n_asym++ % nparts) == part) asym_items.push_back({3, (int)i, (int)j});
        }
        sym_items.insert(sym_items.end(), diag_items.begin(), diag_items.end());
        sym_items.insert(sym_items.end(), os_items.begin(), os_items.end());
        blk_same = split_blocks(vpad);
        blk_other = split_blocks(vbpad);
        for (int k = 0; k < 4; ++k) chunk_items[k] = 0;
        if (!sharded) {
            // a call never needs more workspace columns than all of its items together
            int64_t need = 0;
            for (const Item& it : sym_items)
                need += it.kind == 2 ? vbpad : (it.kind == 1 ? ((ca != 0.0) + (cs != 0.0)) * vpad : vpad);
            need = std::max<int64_t>(need, (int64_t)asym_items.size() * vpad);
            cap_cols = std::min(cap_cols, rup(std::max<int64_t>(need, 2 * std::max(vpad, vbpad)), 16));
        }
        return 0;
    }

    // Pole-sharded build: the PAIR items of one step of shard_steps() for `rank` of R (replaces sym_items).
    void set_shard_step(int64_t rank, int64_t R, const ShardStep& s) {
        sharded = true;
        res_lo = shard_lo(o, rank, R); res_hi = shard_lo(o, rank + 1, R);
        vis_lo = s.vis_lo; vis_hi = s.vis_hi;
        sym_items.clear();
        asym_items.clear();
        if (ca == 0.0) return;
        const int64_t want = rank < s.src ? 0 : 1;     // parity of i + j this rank takes of the block pair
        for (int64_t i = res_lo; i < res_hi; ++i)
            for (int64_t j = vis_lo; j < vis_hi; ++j) {
                if (s.t == 0 ? (j >= i) : (((i + j) & 1) != want)) continue;
                sym_items.push_back({1, (int)i, (int)j});
            }
    }
    // operand index of pole `pole` of the (ix|Q) tensor: resident map (asel 0) or visiting map (asel 1)
    void localize(int& pole, int& asel) const {
        if (!sharded) return;
        if (pole >= res_lo && pole < res_hi) { pole -= (int)res_lo; asel = 0; }
        else { pole -= (int)vis_lo; asel = 1; }
    }

    // thin-tile mapping of the last x-block (k1_mainloop_thin): number of valid 8-row blocks, 0 = full tile
    int thin_mbv(int xb, int nb1, int nb2) const {
        if (!cfg.thin_tiles || cfg.simple || xb != nxb - 1) return 0;
        const int mbv = (int)((nphys - (int64_t)xb * kTileM + 7) / 8);
        if (mbv > 14) return 0;
        const int nsets = (nb1 > 0) + (nb2 > 0);
        if (mbv > 10)   // quad tile (2 x 4 warp grid): ceil(mbv / 2) row blocks x 2 column blocks of each integral block
            return (1.03 * ((mbv + 1) / 2) * 2 * nsets < 2.0 * (nb1 + nb2)) ? mbv : 0;
        return (1.15 * mbv * nsets < 2.0 * (nb1 + nb2)) ? mbv : 0;   // per-warp DMMA count: thin vs full
    }
    // per-warp DMMAs per k-step (x 2): the tile's duration
    static int tile_work(const K1Tile& t) {
        const int nsets = (t.nb1 > 0) + (t.nb2 > 0);
        if (t.mbv > 10) return ((t.mbv + 1) / 2) * 2 * nsets;
        return t.mbv > 0 ? t.mbv * nsets : 2 * (t.nb1 + t.nb2);
    }

    int items_per_chunk(int kind) {
        if (chunk_items[kind]) return chunk_items[kind];
        const int64_t w = kind == 2 ? vbpad : vpad;
        const int nblk = kind == 1 ? ((ca != 0.0) + (cs != 0.0)) : 1;
        const int n_max = (int)std::max<int64_t>(1, cap_cols / std::max<int64_t>(1, nblk * w));
        int n = n_max;
        if (cfg.tune_chunk && !cfg.simple) {
            // tile durations of one item, in 8-column units; ~0.15 units of fill + epilogue per tile at nk = 250
            const double ovh = 0.15 * 250.0 / std::max(1, (int)((naux + kKB - 1) / kKB));
            std::vector<double> t;
            const auto& blks = kind == 2 ? blk_other : blk_same;
            if (kind == 1 || kind == 3) {
                for (int xb = 0; xb < nxb; ++xb)
                    for (auto& b : blks) {
                        const int mbv = thin_mbv(xb, b.second, b.second);
                        t.push_back((mbv > 10 ? 2.0 * ((mbv + 1) / 2) : (mbv ? (double)mbv : 2.0 * b.second)) + ovh);
                    }
            } else {
                // plain blocks pair up two by two across items: model one item as half-tiles
                for (int xb = 0; xb < nxb; ++xb)
                    for (auto& b : blks) {
                        const int mbv = thin_mbv(xb, b.second, b.second);
                        t.push_back((mbv > 10 ? (double)((mbv + 1) / 2) : (mbv ? 0.5 * mbv : (double)b.second)) + 0.5 * ovh);
                    }
            }
            n = pick_chunk_items(t, n_max, cfg.num_sms, cfg.lpt_order, 0.5);
        }
        chunk_items[kind] = n;
        if (getenv("AGF2_B200_VERBOSE"))
            fprintf(stderr, "agf2_b200: kind %d: %d items per chunk (workspace allows %d), %d column blocks/item\n", kind, n,
                    n_max, (int)(kind == 2 ? blk_other : blk_same).size());
        return n;
    }

    // Packs the next chunk of `items` (starting at pos, which is advanced) into CTA tiles.  ei_*: pole energies on the
    // host; pole_rank (optional): upload order of the poles, need_rank = the last upload this chunk waits for.
    // Returns the number of workspace columns the chunk uses (0: not even one item fits the workspace).
    int64_t next_chunk(const std::vector<Item>& items, size_t& pos, const double* ei_same, const double* ei_other,
                       const std::vector<int>* pole_rank, std::vector<K1Tile>& tiles, int& need_rank)
    {
        std::vector<SubTile> plain;
        tiles.clear();
        int64_t col = 0;
        need_rank = -1;
        const size_t pos_end = std::min(items.size(), pos + (size_t)items_per_chunk(items[pos].kind));
        while (pos < pos_end) {
            const Item& it = items[pos];
            const int64_t w = it.kind == 2 ? vbpad : vpad;
            const int nblk = it.kind == 1 ? ((ca != 0.0) + (cs != 0.0)) : 1;
            if (col + nblk * w > cap_cols) break;
            if (pole_rank) {
                need_rank = std::max(need_rank, (*pole_rank)[it.i]);
                if (it.kind != 2) need_rank = std::max(need_rank, (*pole_rank)[it.j]);
            }
            if (it.kind == 1 || it.kind == 3) {
                const double eij = ei_same[it.i] + ei_same[it.j];
                const long long c1 = col, c2 = (it.kind == 1 && ca != 0.0 && cs != 0.0) ? col + w : col;
                // tile order (pair, x-block, a-block): consecutive CTAs share the ixq rows of their
                // x-block and the (Q|ja) columns of their pair while those are still in L2
                for (int xb = 0; xb < nxb; ++xb) {
                    for (const auto& blk : blk_same) {
                        const int64_t a0 = blk.first;
                        const int nb = blk.second;
                        const int nv = (int)std::max<int64_t>(0, std::min<int64_t>(8 * nb, v - a0));
                        K1Tile t;
                        memset(&t, 0, sizeof(t));
                        t.i1 = it.i; t.j1 = it.j; t.a1 = (int)a0; t.nb1 = nb;
                        t.i2 = it.j; t.j2 = it.i; t.a2 = (int)a0; t.nb2 = nb;
                        t.xb = xb; t.nv1 = t.nv2 = nv; t.e1 = t.e2 = eij;
                        t.col1 = c1 + a0; t.col2 = c2 + a0;
                        if (it.kind == 1) {
                            t.ws1 = ca != 0.0 ? 0 : -1; t.p11 = ca; t.p12 = -ca;
                            t.ws2 = cs != 0.0 ? 0 : -1; t.p21 = cs; t.p22 = cs;
                        } else {
                            t.ws1 = 0; t.p11 = 1.0; t.p12 = 0.0;
                            t.ws2 = 1; t.p21 = asym_c; t.p22 = -s_same;   // (p21 = 0: out2 width falls back to nb2 == nb1)
                        }
                        t.mbv = thin_mbv(xb, t.nb1, t.nb2);
                        localize(t.i1, t.asel1);
                        localize(t.i2, t.asel2);
                        tiles.push_back(t);
                    }
                }
            } else {
                const bool os = it.kind == 2;
                const int64_t vv_ = os ? vb : v;
                const double eij = ei_same[it.i] + (os ? ei_other[it.j] : ei_same[it.j]);
                for (const auto& blk : (os ? blk_other : blk_same)) {
                    const int64_t a0 = blk.first;
                    const int nb = blk.second;
                    const int nv = (int)std::max<int64_t>(0, std::min<int64_t>(8 * nb, vv_ - a0));
                    plain.push_back({it.i, it.j, (int)a0, nb, os ? 1 : 0, os ? 1 : 0, nv, (long long)(col + a0),
                                     os ? co : cd, eij});
                }
            }
            col += nblk * w;
            ++pos;
        }
        if (col == 0) return 0;
        // two independent plain blocks share one CTA tile
        std::stable_sort(plain.begin(), plain.end(), [](const SubTile& a, const SubTile& b) { return a.nb > b.nb; });
        for (size_t k = 0; k < plain.size();) {
            const SubTile& s1 = plain[k];
            // only equal-width blocks share a CTA tile: k1_form is instantiated for (n, n) and (n, 0)
            const SubTile* s2 = (k + 1 < plain.size() && plain[k + 1].nb == s1.nb) ? &plain[k + 1] : nullptr;
            k += s2 ? 2 : 1;
            for (int xb = 0; xb < nxb; ++xb) {
                K1Tile t;
                memset(&t, 0, sizeof(t));
                t.xb = xb;
                t.i1 = s1.i; t.j1 = s1.j; t.a1 = s1.a0; t.nb1 = s1.nb; t.bsel1 = s1.bsel; t.esel1 = s1.esel;
                t.nv1 = s1.nv; t.col1 = s1.col; t.p11 = s1.p; t.e1 = s1.eij; t.ws1 = 0;
                t.ws2 = -1;
                if (s2) {
                    t.i2 = s2->i; t.j2 = s2->j; t.a2 = s2->a0; t.nb2 = s2->nb; t.bsel2 = s2->bsel; t.esel2 = s2->esel;
                    t.nv2 = s2->nv; t.col2 = s2->col; t.p22 = s2->p; t.e2 = s2->eij; t.ws2 = 0;
                }
                t.mbv = thin_mbv(xb, t.nb1, t.nb2);
                tiles.push_back(t);
            }
        }
        // longest tiles first: the block scheduler then packs the short ones into the tail of the launch
        if (cfg.lpt_order)
            std::stable_sort(tiles.begin(), tiles.end(),
                             [](const K1Tile& a, const K1Tile& b) { return tile_work(a) > tile_work(b); });
        return col;
    }
};

}  // namespace agf2
