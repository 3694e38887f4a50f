// eik_fast.cuh -- warp-synchronous Podvin-Lecomte solver: 32 solves per warp, the live part of
// every time field in shared memory, one node update per lane and iteration.
//
// What is kept from the generic solver (eik_core.cuh): the arithmetic of every stencil and the
// visiting order inside a solve (reference src/time_2d.c:959-1147, 1186-1373), hence the results.
// What is different:
//   * The expanding box only ever reads its own perimeter: the top row, the bottom row and the
//     right column (the source sits on the left edge, X0 == 0).  Those three lines live in shared
//     memory, lane-interleaved ([index][lane], bank == lane: conflict-free whatever index a lane
//     is at), and a side sweep overwrites the past line IN PLACE with the new one: the values the
//     stencils still need (past and current time at the previous node of the walk) are carried in
//     registers, the one node a later walk revisits (the local maximum between two minima) keeps
//     its past value in a register.
//   * After the box spans the full depth range the solve is a pure march over columns: past and
//     current column ping-pong between two buffers, both directions of a column run as two
//     independent chains of one lock-step loop (march_sweep3); receiver rows are emitted as the
//     march passes.  (eik_march.cuh holds the same march with the columns in tensor memory.)
//   * All lanes of a warp run the same instruction stream: the per-lane walk position and
//     direction are data.  Lanes are grouped by source depth and by the layering round the source
//     (eikonal.cu: eik_order_tasks), so their boxes grow in step.
//   * Head waves along a row (a faster layer on the far side, src/time_2d.c:1029-1059) are
//     detected, not handled, by the fast sweep: the lane re-does that line with the generic
//     sweep (eik_core.cuh) on the global-memory copy of the box that every fast sweep writes
//     through to, including the reverse propagation, and reloads its perimeter.  The same slow
//     path takes the few other irregular cases (exact ties on a plateau, pre-set nodes after the
//     minimal initialisation, boxes that touch the right edge).
//
// Compiles for the device and, with a 1-lane "warp", for the host (tests/emu): same order, same arithmetic
// (bit-identical to the generic core on the host; on the device the radicands' square roots are MUFU.SQRT).
#pragma once
#include "eik_core.cuh"
#ifdef EIKF_TRACE      // debug builds (tools/ab_build.py ... -DEIKF_TRACE=<solve>): one solve's box phase is printed
#include <cstdio>
#endif

namespace eikf {

using eik::kInf;
using eik::kFuzz;
using eik::kInitMin;
using eik::kRsqrt2;
using eik::kSqrt2;

#ifdef __CUDA_ARCH__
constexpr int LS = 32;   // lane stride of the interleaved arrays
#define EIKF_ANY(p) __any_sync(0xffffffffu, (p))
#define EIKF_SYNC() __syncwarp()
#define EIKF_MIN(v) __reduce_min_sync(0xffffffffu, (v))
#define EIKF_MAX(v) __reduce_max_sync(0xffffffffu, (v))
#elif defined(EIKF_HOST_WARP)
// tests/emu/eik_emu_mt.cpp: a "warp" of host threads in lock-step, collectives by barrier (defines LS and the four macros)
#include EIKF_HOST_WARP
#else
constexpr int LS = 1;
#define EIKF_ANY(p) (p)
#define EIKF_SYNC() ((void)0)
#define EIKF_MIN(v) (v)
#define EIKF_MAX(v) (v)
#endif

struct Dims {
    int nx, nz;          // coarse plane
    int wx;              // columns of the global box window: min(nx, max(nz + 3, 13))
    int col_len;         // column buffer: indices -1 .. col_len (col_len >= max(nz, 43); -1 and nz hold sentinels)
    int row_len;         // row buffer: top row from the front, bottom row from the back.  While both are needed they hold
                         // X1 + 1 nodes each, and X1 <= (nz + 2) / 2 then (a seed box that failed on both sides in the same
                         // round of seed_search is 2.5 nodes wider than half its height): nz + 4 nodes for an even nz,
                         // nz + 5 for an odd one (tests/test_emu_cpu.py: test_row_buffer_*); the refined grid needs 2 x 22
    int row_march;       // rows whose past times do not decrease away from the axis are swept in lock-step (row_march)
    int lock_cols;       // shared-memory slice with a second column buffer: the columns of the growing box are swept by the
                         // two-chain loop over the union of the lanes' ranges (as in GM mode) instead of the per-lane
                         // in-place walk.  Faster box phases, fewer slices per SM: pays for launches a few tasks per warp deep
};

EIK_HD Dims make_dims(int nx, int nz)
{
    Dims D;
    D.nx = nx; D.nz = nz;
    D.wx = (nz + 3 > 13) ? nz + 3 : 13;
    if (D.wx > nx) D.wx = nx;
    D.col_len = nz > 43 ? nz : 43;
    D.row_len = (nz + 4 + (nz & 1) > 48) ? nz + 4 + (nz & 1) : 48;
    D.row_march = 1;
    D.lock_cols = 0;
    return D;
}
// floats per lane of the three shared arrays: S[-1..nz-1], COL[-1..col_len], ROW[0..row_len-1] (+ COL2 with lock_cols)
EIK_HD int smem_floats_per_lane(const Dims& D) { return (D.nz + 1) + (D.col_len + 2) + D.row_len + (D.lock_cols ? D.col_len + 2 : 0); }
// the same for a slice in global memory (carve_global): always with the second column buffer
EIK_HD int gmem_floats_per_lane(const Dims& D) { return (D.nz + 1) + (D.col_len + 2) + D.row_len + (D.col_len + 2); }

// One lane's view of the storage.
struct Lane {
    float* S;      // shared: slowness per coarse depth cell, S[my] = INF (masked dummy row), S[-1] = INF
    float* COL;    // shared: right column of the box, in place
    float* ROW;    // shared: top row at [x], bottom row at [row_len-1-x]
    float* W;      // global: coarse box window, node (x,y) at W[(x*nz+y)*LS]
    float* WF;     // global: refined grid, node (x,y) at WF[(x*nyf+y)*LS]
    float* COL2;   // second column buffer (same shape as COL) when the slice lives in global memory (GM mode), else nullptr
};

// Medium of the grid being solved, for the fast sweeps: depth cells come from the shared array
// either directly (coarse) or through the half-spacing map (refined, src/time_2d.c:865-871).
struct Med {
    const float* S;
    int fine, j0, hy;
    EIK_HD float cell(int cy) const
    {
        return fine ? 0.5f * S[(size_t)(j0 + ((cy + hy) >> 1)) * LS] : S[(size_t)cy * LS];
    }
};

// Box state of one lane on one grid.
struct Box {
    int nx, ny, mx, my;   // this grid
    int ys;               // source (0, ys)
    int X1, Y0, Y1;
    int preset_up;        // row above the box holds pre-set nodes (minimal initialisation)
    int active;
    // GM mode: the homogeneous seed box [0, sX1] x [sY0, sY1] (slowness shs0) has only its outline and the receiver rows in
    // the window so far; the interior is filled when a slow path first needs the window (fill_seed_interior)
    int seed_pending, sX1, sY0, sY1;
    float shs0;
#ifdef EIKF_TRACE
    int trace;            // debug builds: this lane's solve is the one being traced (printf)
#endif
};

// Square root of a stencil radicand.  On the device this is the hardware approximation (one MUFU.SQRT, relative
// error <= 2^-23, i.e. at most one ulp of a term that is itself <= one cell's slowness, ~0.5 s: < 6e-8 s) instead of
// the correctly rounded rsqrt + Newton sequence (five instructions).  The sum t + sqrt() is rounded to an ulp of t,
// which is 60 times coarser at t = 30 s, so the field moves by far less than the FP32-vs-reference difference the
// parity tolerance max(1e-4 s, 2e-6 T) is stated for (tests/test_eikonal_gpu.py measures it).  A non-positive argument
// only occurs where the stencil is not selected and yields an ignored NaN.  EIKF_EXACT_SQRT restores the Newton form.
EIK_HD float sqrt_pos(float r)
{
#if defined(__CUDA_ARCH__) && !defined(EIKF_EXACT_SQRT)
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(r));
    return y;
#elif defined(__CUDA_ARCH__)
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(r));
    const float s = r * y;
    const float h = 0.5f * y;
    const float e = fmaf(-s, s, r);
    return fmaf(e, h, s);
#else
    return sqrtf(r);
#endif
}

// 0 <= x < lim for lim > 0: non-negative floats order like their bit patterns and a negative x has its sign bit set, so one
// unsigned compare does both (x is a difference of two times or sentinels here: never NaN, never -0.0, which x - x is not).
EIK_HD bool in_zero_lim(float x, float lim)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(x) < __float_as_uint(lim);
#else
    return x >= 0.f && x < lim;
#endif
}

// Write-through of a value to the lane's global window: a streaming store (evict-first), the window is only read back on
// the rare slow path and by the output copy, and must not push the lanes' live arrays out of L2 (GM mode).
EIK_HD void wt_store(float* p, float v)
{
#ifdef __CUDA_ARCH__
    __stcs(p, v);
#else
    *p = v;
#endif
}

// ---- one node of a walk ------------------------------------------------------------------------
// pk: past time at the node, pn/cn: past and current time at the neighbour towards the minimum,
// c: current value of the node so far, hs0: cell between node and neighbour, hs1: next cell away
// from the minimum (use3 == false where the reference has no such cell: k == 0 walking backwards).
// All candidates are finite or INF when they reach fminf (never NaN: an unselected stencil is replaced by INF
// first), so fminf == the reference's "if (est < t) t = est".
EIK_HD float node_update(float c, float pk, float pn, float cn, float hs0, float hs1, bool use3)
{
    const float lim = hs0 * kRsqrt2;
    const float hs0sq = hs0 * hs0;
    const float dt = pk - pn;
    float est = pk + sqrt_pos(fmaf(-dt, dt, hs0sq));              // plane wave through the past side
    c = fminf(c, (dt < lim) ? est : kInf);
    const float dt2 = cn - pn;
    est = cn + sqrt_pos(fmaf(-dt2, dt2, hs0sq));                  // plane wave through the lateral side
    c = fminf(c, in_zero_lim(dt2, lim) ? est : kInf);
    c = fminf(c, use3 ? pk + hs1 : kInf);                         // 1-D transmission towards the future
    c = fminf(c, fmaf(hs0, kSqrt2, pn));                          // corner diffraction
    return c;
}

// Would the head-wave stencil of src/time_2d.c:1029-1059 change anything at this node?
EIK_HD bool headwave_fires(float c, float cn, float hs2)
{
    float est = cn + hs2;
    if ((c - est) > kFuzz * c) return true;
    est = c + hs2;
    return (cn - est) > kFuzz * cn;
}

// ---- fast sweep of one line -----------------------------------------------------------------------------
// P: the past line, C: where the new line goes (element k at X[k*stride]).  Two storage disciplines:
//   in place (C == P, inplace = true): the line is overwritten as the walk passes; the one node a later
//     walk revisits (the local maximum between two minima) keeps its past value in a register.  A walk that
//     would have to go further back over overwritten values (an exact tie there) gives up -> slow path.
//   ping-pong (C != P): nothing is lost, every case of the reference's walk is reproduced; used on the
//     march, where no global copy exists to fall back on.
// ROW sweeps have a constant strip slowness c (and far-side slowness c2 for the head-wave test); COL
// sweeps read the strip slowness per depth cell from `med`.
// The walk of a lane is ONE loop: a lane is either looking for its next local minimum (SEG) or at a node of
// its walk, first towards kb (d = -1) then towards ke (d = +1); position and direction are data, so lanes whose
// minima sit at different depths do not wait for each other.
// Wt/wstride: write-through target of node k (Wt[k*wstride]), or nullptr.
// hint (may be nullptr): in: index of the first local minimum of P if known (>= 0); out: the same for C, found
// as C is written (valid when the next sweep covers the same [kb, ke]: the march), or -1.
// Returns true for lanes that must re-do the line on the slow path.
// kedge (rows of the coarse grid only, else -1): node of the row whose cell on the far side along the row is the masked
// dummy column (src/time_2d.c:489-496): the row reaches the right edge of the grid, X1 == mx.
// report_ties (per lane, ping-pong discipline): return true where the in-place walk would.
template <bool ROW>
EIK_HD bool fast_sweep(bool act, const float* P, float* C, int stride, bool inplace, int kb, int ke, const Med& med,
                       float c, float c2, float* Wt, long wstride, int* hint, int kedge = -1, bool report_ties = false)
{
    enum { SEG = 0, WALK = 1, DONE = 2 };
    const bool hw = ROW && (c2 < c);
    int st = act ? SEG : DONE;
    bool slow = false;
    int k = kb;            // where the search for the next local minimum starts
    int vhi = kb - 1;      // nodes kb..vhi of the new line have been written
    int kk = 0, d = -1, kmin = kb, nseg = 0;
    float pmin = 0.f, cmin = 0.f, pn = 0.f, cn = 0.f, s0 = 0.f, sp = 0.f;
    float plast = 0.f;     // past-line time of node vhi (in-place discipline)
    bool eq_tail = false;  // the last forward walk ended on an exact tie
    int ans_b = 0x7fffffff, ans_f = ke;   // first local minimum of the new line, tracked as it is written

    while (EIKF_ANY(st != DONE)) {
        if (st == SEG) {
            float pk;
            if (hint && *hint >= 0 && nseg == 0) {
                k = *hint;
                pk = P[(long)k * stride];
            } else {
                pk = P[(long)k * stride];
                while (k < ke) {
                    const float pnx = P[(long)(k + 1) * stride];
                    if (!(pnx < pk)) break;
                    pk = pnx;
                    k++;
                }
            }
            nseg++;
            kmin = k;
            pmin = pk;
            float sm;
            if (ROW) { sp = (k == kedge) ? kInf : c; sm = (k == 0) ? kInf : c; }
            else { sp = med.cell(k); sm = (k == 0) ? kInf : med.cell(k - 1); }
            const float est = pk + eik::fmin_ref(sm, sp);      // 1-D transmission in front of the minimum
            cmin = (est < kInf) ? est : kInf;
            C[(long)k * stride] = cmin;
            if (Wt) Wt[(long)k * wstride] = cmin;
            kk = kmin - 1; d = -1; pn = pmin; cn = cmin; s0 = sm;
            st = WALK;
        } else if (st == WALK) {
            bool ok = (d < 0) ? (kk >= kb) : (kk <= ke);
            bool seen = false;
            float pk2 = 0.f;
            if (ok) {
                seen = kk <= vhi;
                pk2 = (inplace && seen) ? plast : P[(long)kk * stride];
                ok = (pk2 - pn >= 0.f);
            }
            if (!ok) {
                if (d < 0 && kmin < ke) {       // turn round: walk towards ke from the minimum
                    d = 1; kk = kmin + 1; pn = pmin; cn = cmin; s0 = sp;
                    vhi = kmin; plast = pmin; eq_tail = false;
                    seen = false;
                    pk2 = P[(long)kk * stride];
                    ok = (pk2 - pn >= 0.f);
                    if (!ok) { k = kk; st = SEG; }
                } else if (d < 0) {
                    st = DONE;                  // the minimum was the last node of the line
                } else {                        // this segment is finished
                    k = kk;
                    st = (k <= ke) ? SEG : DONE;
                }
            }
            if (ok) {
                // the walk would go on over overwritten values (report_ties: a ping-pong walk that tells where the in-place
                // walk would have given up, so that its caller takes the slow path for the same lines)
                if ((inplace || report_ties) && seen && eq_tail) slow = true;
                const bool use3 = (d > 0) || (kk != 0);
                float hs0, hs1;
                if (ROW) { hs0 = c; hs1 = (d > 0 && kk == kedge) ? kInf : c; }
                else { hs0 = s0; hs1 = use3 ? med.cell((d > 0) ? kk : kk - 1) : 0.f; }
                const float cold = seen ? C[(long)kk * stride] : kInf;
                const float cv = node_update(cold, pk2, pn, cn, hs0, hs1, use3);
                if (hw && headwave_fires(cv, cn, c2)) slow = true;
                C[(long)kk * stride] = cv;
                if (Wt) Wt[(long)kk * wstride] = cv;
                if (d < 0) { if (cn >= cv) ans_b = kk; }
                else {
                    if (cv >= cn && kk - 1 < ans_f) ans_f = kk - 1;
                    vhi = kk; plast = pk2; eq_tail = (pk2 - pn == 0.f);
                }
                pn = pk2; cn = cv; s0 = hs1;
                kk += d;
                if (d < 0 && inplace && seen) kk = kb - 1;   // an in-place walk stops after the node it revisits
                if (slow) st = DONE;
            }
        }
    }
    if (hint) *hint = (nseg == 1 && !slow) ? ((ans_b != 0x7fffffff) ? ans_b : ans_f) : -1;
    return slow;
}

#ifdef EIKF_STATS
static long g_stats[12];
#endif

// ---- rows in lock-step ---------------------------------------------------------------------------------------
// In a laterally homogeneous medium the first arrival at a given depth never comes earlier further away from the
// source axis, so the past row of a row sweep (nodes 0 .. ke, the axis at 0) is non-decreasing -- up to rounding,
// which is why it is checked.  The reference's walk over such a row is one segment: the local minimum it finds is
// node 0, timed by 1-D transmission, then one pass outwards (fast_sweep<true>: SEG at k = 0, the walk towards kb ends
// at once, the walk towards ke runs through).  All lanes are at the same node at the same time: straight-line code,
// no per-lane state machine.  In place like fast_sweep's row discipline; a lane whose head-wave test fires stops
// writing and reports the line for the slow path, exactly where fast_sweep would.
// R: node k at R[k*stride]; ke per lane.  *mono: the NEW row is non-decreasing as well (known for free).
EIK_HD bool row_is_monotone(bool act, const float* R, int stride, int ke)
{
    bool ok = true;
    float prev = act ? R[0] : 0.f;
    for (int k = 1; EIKF_ANY(act && k <= ke); k++) {
        if (act && k <= ke) {
            const float cur = R[(long)k * stride];
            ok = ok && (cur - prev >= 0.f);
            prev = cur;
        }
    }
    return ok;
}

// EDGE: some lane's row reaches the right edge of the coarse grid (kedge == its last node, else -1): the cell beyond that
// node is the masked dummy column, so it has no 1-D transmission towards the future.
template <bool EDGE, bool PF>
EIK_HD bool row_march(bool act, float* R, int stride, int ke, float c, float c2, float* Wt, long wstride, bool* mono, int kedge = -1)
{
    const bool hw = c2 < c;
    bool slow = false, mn = true;
    float pn = 0.f, cn = 0.f;
    if (act) {
        pn = R[0];
        // 1-D transmission in front of the minimum (node 0: no cell before it; a one-node row at the edge: none after it either)
        const float est = pn + eik::fmin_ref(kInf, (EDGE && kedge == 0) ? kInf : c);
        cn = (est < kInf) ? est : kInf;
        R[0] = cn;
        if (Wt) Wt[0] = cn;
    }
    if (PF) {
        // The row lives in global memory: the past values of the NEXT eight nodes are requested before the current eight are
        // worked on, so a memory latency is paid once per sweep, not once per block.  (The sweep is in place, but a node is
        // only written after it and everything before it has been read.)
        constexpr int NB = 16;
        float nv[NB];
#pragma unroll
        for (int u = 0; u < NB; u++) nv[u] = (act && 1 + u <= ke) ? R[(long)(1 + u) * stride] : 0.f;
        for (int k0 = 1; EIKF_ANY(act && !slow && k0 <= ke); k0 += NB) {
            float pv[NB];
#pragma unroll
            for (int u = 0; u < NB; u++) {
                pv[u] = nv[u];
                nv[u] = (act && k0 + NB + u <= ke) ? R[(long)(k0 + NB + u) * stride] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < NB; u++) {
                const int k = k0 + u;
                if (act && !slow && k <= ke) {
                    const float pk = pv[u];
                    const float cv = node_update(kInf, pk, pn, cn, c, (EDGE && k == kedge) ? kInf : c, true);
                    if (hw && headwave_fires(cv, cn, c2)) slow = true;
                    R[(long)k * stride] = cv;
                    if (Wt) wt_store(Wt + (long)k * wstride, cv);
                    mn = mn && (cv - cn >= 0.f);
                    pn = pk; cn = cv;
                }
            }
        }
    } else {
        for (int k = 1; EIKF_ANY(act && !slow && k <= ke); k++) {
            if (act && !slow && k <= ke) {
                const float pk = R[(long)k * stride];
                const float cv = node_update(kInf, pk, pn, cn, c, (EDGE && k == kedge) ? kInf : c, true);
                if (hw && headwave_fires(cv, cn, c2)) slow = true;
                R[(long)k * stride] = cv;
                if (Wt) Wt[(long)k * wstride] = cv;
                mn = mn && (cv - cn >= 0.f);
                pn = pk; cn = cv;
            }
        }
    }
    *mono = mn && !slow;
#ifdef EIKF_STATS
    if (act) { g_stats[9]++; if (slow) g_stats[10]++; }
#endif
    return slow;
}

// ---- the march: one full-depth column from the previous one ------------------------------------------------
// The same walk as fast_sweep<false> in its ping-pong discipline, specialised for the steady state of a solve
// (kb = 0, ke = my, coarse medium, no write-through, no slow path).  P and C have indices -1 .. ke+1 whose two
// end slots hold kStop, so a walk ends at the array ends by the very test that ends it at a local maximum of the
// past column; S[-1] = S[ke] = INF stand for the cells the reference does not look at.
constexpr float kStop = -1.0e30f;


EIK_HD void march_sweep(bool act, const float* P, float* C, const float* S, int ke, int* hint)
{
    bool alive = act;
    int k = 0, kk = 0, d = -1, kmin = 0, vhi = -1, nseg = 0;
    float pmin = 0.f, cmin = 0.f, pn = 0.f, cn = 0.f, s0 = 0.f, sp = 0.f;
    // first local minimum of the new column, tracked as it is written; rv_lo = lowest node a later walk re-timed
    int ans_b = 0x7fffffff, ans_f = ke, rv_lo = 0x7fffffff;

    auto start_segment = [&](bool use_hint) {
        float pk;
        if (use_hint) {
            k = *hint;
            pk = P[(long)k * LS];
        } else {
            pk = P[(long)k * LS];
            while (k < ke) {
                const float pnx = P[(long)(k + 1) * LS];
                if (!(pnx < pk)) break;
                pk = pnx;
                k++;
            }
        }
        nseg++;
        kmin = k;
        pmin = pk;
        sp = S[(long)k * LS];
        const float sm = S[(long)(k - 1) * LS];
        cmin = fminf(kInf, pk + eik::fmin_ref(sm, sp));   // 1-D transmission in front of the minimum
        C[(long)k * LS] = cmin;
        kk = kmin - 1; d = -1; pn = pmin; cn = cmin; s0 = sm;
    };

    if (alive) start_segment(*hint >= 0);
    while (EIKF_ANY(alive)) {
        if (alive) {
            float pk2 = P[(long)kk * LS];
            float dt = pk2 - pn;
            bool go = true;
            if (dt < 0.f) {
                if (d < 0) {                         // the walk towards the top is over: turn round at the minimum
                    if (kmin == ke) { alive = false; go = false; }
                    else {
                        d = 1; kk = kmin + 1; pn = pmin; cn = cmin; s0 = sp; vhi = kmin;
                        pk2 = P[(long)kk * LS];
                        dt = pk2 - pn;
                    }
                }
                if (alive && dt < 0.f) {             // the segment is over: next local minimum, or done
                    go = false;
                    k = kk;
                    if (k > ke) alive = false;
                    else start_segment(false);
                }
            }
            if (go) {
                const float hs1 = S[(long)((d > 0) ? kk : kk - 1) * LS];
                float cold = kInf;
                if (kk <= vhi) {                     // a later walk comes back over nodes already timed
                    cold = C[(long)kk * LS];
                    rv_lo = (kk < rv_lo) ? kk : rv_lo;
                }
                const float lim = s0 * kRsqrt2;
                const float s0sq = s0 * s0;
                float est = pk2 + sqrt_pos(fmaf(-dt, dt, s0sq));
                float cv = fminf(cold, (dt < lim) ? est : kInf);
                const float dt2 = cn - pn;
                est = cn + sqrt_pos(fmaf(-dt2, dt2, s0sq));
                cv = fminf(cv, in_zero_lim(dt2, lim) ? est : kInf);
                cv = fminf(cv, pk2 + hs1);
                cv = fminf(cv, fmaf(s0, kSqrt2, pn));
                C[(long)kk * LS] = cv;
                if (d < 0) { if (cn >= cv && kk < ans_b) ans_b = kk; }
                else {
                    if (cv >= cn && kk - 1 < ans_f) ans_f = kk - 1;
                    vhi = kk;
                }
                pn = pk2; cn = cv; s0 = hs1;
                kk += d;
            }
        }
    }
    if (act) {
        const int first = (ans_b != 0x7fffffff) ? ans_b : ans_f;
        *hint = (first + 1 < rv_lo) ? first : -1;
#ifdef EIKF_STATS
        g_stats[0]++; if (*hint < 0) g_stats[1]++; if (nseg > 1) g_stats[2]++; if (nseg > 2) g_stats[3]++;
        if (*hint < 0 && rv_lo <= 1) g_stats[4]++;
#endif
    }
}

// ---- the march in lock-step --------------------------------------------------------------------------------------
// Without exact ties in the past column the order in which the reference visits the nodes of a line does not
// matter, only who is timed from whom: a node whose past time is not below its upper neighbour's is timed from
// that neighbour (the reference reaches it walking down from a local minimum), a node not below its lower
// neighbour's from below, a local maximum from both (the smaller wins), and a node strictly below its upper
// neighbour and not above its lower one is where the reference's search for a local minimum stops: a root, timed
// by 1-D transmission.  So a column is one pass down and one pass up, every lane at the same depth at the same
// time, with predicated results instead of per-lane walks: straight-line code, no votes, no state machine.  The
// passes also see an exact tie if there is one; such a column (4 in 100 000) is re-done by march_sweep, which
// follows the reference's order literally.
// P, C: indices -1 .. ke+1; P's end slots must hold kEdge (set by the caller); S[-1] = S[ke] = INF.
constexpr float kEdge = 1.0e30f;
// Sentinel of column node k outside the growing box of a solve with its source at depth node ys (GM mode): above every
// travel time and strictly increasing away from the source, so that no minimum, no tie and no chain can come from it.
EIK_HD float gm_sentinel(int k, int ys)
{
    const int d = (k > ys) ? k - ys : ys - k;
    return kEdge * (1.0f + (float)d * (1.0f / 8192.0f));
}

// ---- the march, both passes in one loop -----------------------------------------------------------------------
// The two passes never feed each other: a node timed from above cannot have a neighbour above it
// that is timed from below unless the two past times are equal (a tie, handled elsewhere), and a root's value only
// depends on the past column.  So the downward chain (A, k = i) and the upward chain (B, k = ke - i) run in the same
// loop iteration as two independent dependency chains -- twice the instruction-level parallelism per warp, which is
// what a kernel limited to ~2 warps per scheduler by its shared-memory footprint needs -- and meet in the middle;
// whoever reaches a node second merges with fminf.
// One node of each chain.  A: node k timed from k-1 (state: past times of k-1 and k, cell k-1, current time of k-1);
// B: node k timed from k+1.  Both also time the roots they pass (1-D transmission) so that neither waits for the other.
struct ChainA { float pprev, pk, sprev, cn; };
struct ChainB { float pnx, pk, s0, cn; };

EIK_HD float chain_a_node(ChainA& a, float pnext, float sk, bool& tie)
{
    const float dt = a.pk - a.pprev;
    const bool up = dt >= 0.f;
    const bool root = !up && (pnext >= a.pk);
    tie = tie || (dt == 0.f);
    const float lim = a.sprev * kRsqrt2;
    const float s0sq = a.sprev * a.sprev;
    float est = a.pk + sqrt_pos(fmaf(-dt, dt, s0sq));
    float cv = (dt < lim) ? est : kInf;
    const float dt2 = a.cn - a.pprev;
    est = a.cn + sqrt_pos(fmaf(-dt2, dt2, s0sq));
    cv = fminf(cv, in_zero_lim(dt2, lim) ? est : kInf);
    const float e3 = a.pk + sk;
    cv = fminf(cv, e3);
    cv = fminf(cv, fmaf(a.sprev, kSqrt2, a.pprev));
    // 1-D transmission in front of a minimum: pk + min(sprev, sk) = min(pk + sprev, pk + sk) (rounding is monotone), and
    // it stays below INF at every real node (one of its two cells is real), so the reference's clip against INF is void
    const float cmin = fminf(e3, a.pk + a.sprev);
    const float val = up ? cv : (root ? cmin : kInf);
    a.cn = val; a.pprev = a.pk; a.pk = pnext; a.sprev = sk;
    return val;
}

EIK_HD float chain_b_node(ChainB& b, float pprev, float hs1)
{
    const float dt = b.pk - b.pnx;
    const bool down = dt >= 0.f;
    const bool root = (b.pk < pprev) && !down;
    const float lim = b.s0 * kRsqrt2;
    const float s0sq = b.s0 * b.s0;
    float est = b.pk + sqrt_pos(fmaf(-dt, dt, s0sq));
    float cv = (dt < lim) ? est : kInf;
    const float dt2 = b.cn - b.pnx;
    est = b.cn + sqrt_pos(fmaf(-dt2, dt2, s0sq));
    cv = fminf(cv, in_zero_lim(dt2, lim) ? est : kInf);
    const float e3 = b.pk + hs1;
    cv = fminf(cv, e3);
    cv = fminf(cv, fmaf(b.s0, kSqrt2, b.pnx));
    const float cmin = fminf(e3, b.pk + b.s0);
    const float val = down ? cv : (root ? cmin : kInf);
    b.cn = val; b.pnx = b.pk; b.pk = pprev; b.s0 = hs1;
    return val;
}

template <bool PF = false>
EIK_HD bool march_sweep3(bool act, const float* P, float* C, const float* S, int ke, float* Wt = nullptr)
{
    bool tie = false;
    // the cells beyond the two ends of the range: S[-1] = S[ke] = INF for a full column (k = 0 .. my), real cells for a
    // column of the growing box (the reference's 1-D transmission at a minimum looks at them, src/time_2d.c:995-997)
    // (the nodes beyond the ends do not exist: a past time above every sentinel keeps the end nodes from being timed from them)
    ChainA a{4.0f * kEdge, P[0], S[-(long)LS], kInf};
    ChainB b{4.0f * kEdge, P[(long)ke * LS], S[(long)ke * LS], kInf};
    const float* pa = P + LS;                 // &P[ka + 1]
    const float* sa = S;                      // &S[ka]
    float* ca = C;                            // &C[ka]
    const float* pb = P + (long)(ke - 1) * LS;   // &P[kb - 1]
    const float* sb = S + (long)(ke - 1) * LS;   // &S[kb - 1]
    float* cb = C + (long)ke * LS;            // &C[kb]

    // one node of each chain; MERGE: the other chain has already been at these nodes
    // (only lanes that take part write: in the pipelined kernel a lane whose box phase is over has handed its column on
    //  in one of these buffers)
    auto step = [&](auto merge) {
        float a_val = chain_a_node(a, *pa, *sa, tie);
        if (decltype(merge)::value) a_val = fminf(a_val, *ca);
        if (act) *ca = a_val;
        float b_val = chain_b_node(b, *pb, *sb);
        if (decltype(merge)::value) b_val = fminf(b_val, *cb);
        if (act) *cb = b_val;
        pa += LS; sa += LS; ca += LS; pb -= LS; sb -= LS; cb -= LS;
    };
    // Columns in global memory (PF): four nodes of each chain per block, the operands of the NEXT block requested before
    // the current one is worked on.  In the second half a chain only reads back what the OTHER chain wrote in the first
    // half, so reading ahead of this half's stores is safe; every node gets its final value exactly once in the second
    // half, which is where the write-through to the window column Wt (element stride LS, or nullptr) happens.
    struct Blk { float pa[4], sa[4], pb[4], sb[4], ca[4], cb[4]; };
    auto load4 = [&](Blk& q, int off, bool merge) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            q.pa[u] = pa[(long)(off + u) * LS]; q.sa[u] = sa[(long)(off + u) * LS];
            q.pb[u] = pb[-(long)(off + u) * LS]; q.sb[u] = sb[-(long)(off + u) * LS];
            if (merge) { q.ca[u] = ca[(long)(off + u) * LS]; q.cb[u] = cb[-(long)(off + u) * LS]; }
        }
    };
    auto compute4 = [&](const Blk& q, bool merge) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            float a_val = chain_a_node(a, q.pa[u], q.sa[u], tie);
            float b_val = chain_b_node(b, q.pb[u], q.sb[u]);
            if (merge) {
                a_val = fminf(a_val, q.ca[u]);
                b_val = fminf(b_val, q.cb[u]);
                if (act && Wt) { wt_store(Wt + (ca - C) + (long)u * LS, a_val); wt_store(Wt + (cb - C) - (long)u * LS, b_val); }
            }
            if (act) { ca[(long)u * LS] = a_val; cb[-(long)u * LS] = b_val; }
        }
        pa += 4 * LS; sa += 4 * LS; ca += 4 * LS; pb -= 4 * LS; sb -= 4 * LS; cb -= 4 * LS;
    };
    auto run_blocks = [&](int n_blocks, bool merge) {      // n_blocks blocks of four nodes per chain
        if (n_blocks <= 0) return;
        Blk qa, qb;                    // ping-pong: one block is worked on while the other is in flight, no copies
        load4(qa, 0, merge);
        int blk = 0;
        for (; blk + 2 <= n_blocks; blk += 2) {
            load4(qb, 4, merge);
            compute4(qa, merge);
            if (blk + 2 < n_blocks) load4(qa, 4, merge);
            compute4(qb, merge);
        }
        if (blk < n_blocks) compute4(qa, merge);
    };
    struct No { enum { value = 0 }; };
    struct Yes { enum { value = 1 }; };

    const int n1 = (ke + 1) >> 1;            // iterations with ka < kb: nobody has been at either node
    if (PF) {
        auto step_wt = [&]() {     // one node of each chain in the second half, with write-through
            float a_val = fminf(chain_a_node(a, *pa, *sa, tie), *ca);
            if (act) *ca = a_val;
            if (act && Wt) wt_store(Wt + (ca - C), a_val);
            float b_val = fminf(chain_b_node(b, *pb, *sb), *cb);
            if (act) *cb = b_val;
            if (act && Wt) wt_store(Wt + (cb - C), b_val);
            pa += LS; sa += LS; ca += LS; pb -= LS; sb -= LS; cb -= LS;
        };
        int i = 0;
        run_blocks(n1 / 4, false);
        i = (n1 / 4) * 4;
        for (; i < n1; i++) step(No());
        if (!(ke & 1)) {   // odd node count: both chains meet on the middle node, the second to arrive merges with the first
            if (act) C[(long)(ke >> 1) * LS] = kInf;
            step_wt();
            i++;
        }
        const int nb2 = (ke + 1 - i) / 4;
        run_blocks(nb2, true);
        i += 4 * nb2;
        for (; i <= ke; i++) step_wt();
    } else {
#pragma unroll 2
        for (int i = 0; i < n1; i++) step(No());
        if (act && !(ke & 1)) C[(long)(ke >> 1) * LS] = kInf;   // odd node count: both chains meet on the middle node
#pragma unroll 2
        for (int i = n1; i <= ke; i++) step(Yes());
    }
    return act && tie;
}

// ---- perimeter <-> global window -------------------------------------------------------------------
// n elements from src (element stride ss) to dst (element stride ds), eight loads in flight before the first store:
// the copies between the global window and shared memory are latency-bound one-lane-wide loops otherwise.
EIK_HD void copy_strided(float* dst, long ds, const float* src, long ss, int n)
{
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = src[(long)(i + u) * ss];
#pragma unroll
        for (int u = 0; u < 8; u++) dst[(long)(i + u) * ds] = v[u];
    }
    for (; i < n; i++) dst[(long)i * ds] = src[(long)i * ss];
}

// col: the buffer that holds the lane's right column (L.COL, or whichever of COL / COL2 is current in GM mode)
template <class G>
EIK_HD void load_perimeter(const G& g, const Lane& L, int row_len, float* col)
{
    // a row is only kept while the box can still grow on that side; the two rows share one buffer
    // (top from the front, bottom from the back) whose length covers them exactly under that rule
    const long xs = (long)g.ny * g.ts;   // node (x,y) of the window lives at t[(x*ny + y)*ts]
    if (g.Y0 > 0) copy_strided(L.ROW, LS, &g.T(0, g.Y0), xs, g.X1 + 1);
    if (g.Y1 < g.my) copy_strided(L.ROW + (size_t)(row_len - 1) * LS, -LS, &g.T(0, g.Y1), xs, g.X1 + 1);
    copy_strided(col + (size_t)g.Y0 * LS, LS, &g.T(g.X1, g.Y0), g.ts, g.Y1 - g.Y0 + 1);
}

template <class Medium>
EIK_HD eik::Grid<Medium> make_grid(float* t, const Box& b, const Medium& m)
{
    eik::Grid<Medium> g;
    g.t = t; g.ts = LS; g.nx = b.nx; g.ny = b.ny; g.mx = b.mx; g.my = b.my; g.S = m; g.ys = b.ys;
    g.X1 = b.X1; g.Y0 = b.Y0; g.Y1 = b.Y1; g.side_limit = 0; g.status = eik::kOk; g.cnt = nullptr;
    return g;
}

// Slow path of one line: the generic sweep (with head waves and reverse propagation) on the global
// copy, then the perimeter is re-read.  `refill`: the fast sweep already wrote part of the line.
template <int AXIS, class Medium>
EIK_HD int slow_line(float* t, Box& b, const Medium& m, const Lane& L, int row_len, int line, int future,
                     int kb, int ke, bool refill, float* col)
{
    if (b.seed_pending) {   // the generic sweep may propagate back into the seed box: its interior has to be there
        for (int x = 0; x <= b.sX1; x++)
            for (int y = b.sY0; y <= b.sY1; y++) t[((size_t)x * b.ny + y) * LS] = eik::box_time(b.shs0, x, y - b.ys);
        b.seed_pending = 0;
    }
    eik::Grid<Medium> g = make_grid(t, b, m);
#ifdef EIKF_STATS
    g_stats[5]++; g_stats[6] += (ke - kb + 1); if (!refill) g_stats[7]++;
#endif
    if (refill)
        for (int k = kb; k <= ke; k++) eik::node<AXIS>(g, line, k) = kInf;
    eik::sweep_line<AXIS>(g, line, future, kb, ke);
    load_perimeter(g, L, row_len, col);
    return g.status;
}

// May the row sweep of this round take row_march?  Every lane that sweeps (act) must have a non-decreasing past row:
// known from the previous sweep of that row (then only the node the last column sweep appended is new), else checked.
EIK_HD bool row_ready(bool act, const float* R, int stride, int ke, bool known)
{
    bool ok = true;
    if (act) {
        if (known) ok = (ke < 1) || (R[(long)ke * stride] - R[(long)(ke - 1) * stride] >= 0.f);
    }
    const bool need_check = act && !known;
    if (EIKF_ANY(need_check)) {
        const bool m = row_is_monotone(need_check, R, stride, ke);
        if (need_check) ok = m;
    }
    return !EIKF_ANY(act && !ok);
}

// ---- expanding box + march on one grid, all lanes of the warp together --------------------------------
// FINE selects the medium type of the slow path.  out/rows: receiver-row output of the coarse grid
// (nullptr on the refined grid); out[r*out_rstride + x] receives t[x][rows[r]].
// GM: the lane's arrays live in global memory (eik_fine_kernel): sweeps that run in lock-step read ahead, and a column of
// the growing box is swept by the two-chain loop of the march (ping-pong between COL and COL2) whenever the lanes agree on
// its range, instead of the per-lane in-place walk whose every step waits for a global load.
template <bool FINE, bool GM, bool LCT>
// hand_col/hand_x1 (coarse grid only, may be nullptr): split mode.  A lane whose box spans the whole depth range
// writes its right column (hand_col[k*32], k = 0..my) and X1 there and stops; a separate kernel marches on from it.
EIK_HD int run_grid(Box& b, const Lane& L, const Dims& D, const eik::CoarseMedium& cm, int j0, int hy, float* out,
                    long out_rstride, const int* rows, int n_rows, float* full, int* xbox_end, float* hand_col = nullptr,
                    int* hand_x1 = nullptr)
{
    float* T = FINE ? L.WF : L.W;
    Med med;
    med.S = L.S; med.fine = FINE ? 1 : 0; med.j0 = j0; med.hy = hy;
    const eik::FineMedium fm{cm, j0, hy};
    int status = eik::kOk;
    const int RL = D.row_len;
    const int wx = FINE ? b.nx : D.wx;
    bool boxphase = b.active && (b.Y0 > 0 || b.Y1 < b.my);
    float* col = L.COL;      // the lane's current right column
    // second column buffer (indices -1..ny): the idle row buffer once the rows are no longer needed, or COL2 (GM)
    constexpr bool LC = GM || LCT;          // columns of the growing box by the two-chain loop (needs COL2)
    float* spare = LC ? L.COL2 : L.ROW + LS;
    int hint = -1;           // first local minimum of the current column (known on the march)
    bool mono_top = false, mono_bot = false;   // the top / bottom row is known not to decrease away from the axis
    if (LC && !FINE && b.active) {
        // both column buffers start as sentinels (see the column sweep below); the right column of the seed box is already
        // in COL (load_perimeter)
        for (int k = -1; k <= b.ny; k++) {
            if (k < b.Y0 || k > b.Y1) col[(long)k * LS] = gm_sentinel(k, b.ys);
            spare[(long)k * LS] = gm_sentinel(k, b.ys);
        }
    }
    if (xbox_end) *xbox_end = boxphase ? -1 : b.X1;
    // cell slowness of a row strip, with the masked dummy row of the coarse grid
    auto rowS = [&](int cy) -> float { return (!FINE && cy >= b.my) ? kInf : med.cell(cy); };
    const bool split = !FINE && hand_x1 != nullptr;
    auto hand_over = [&]() {     // this lane's box phase is over: the march kernel takes the column from here
        for (int k = 0; k <= b.my; k++) hand_col[(size_t)k * LS] = col[(size_t)k * LS];
        *hand_x1 = b.X1;
        b.active = 0;
    };
    if (split && b.active && !boxphase) hand_over();

    for (;;) {
        bool moved = false;
        if (!FINE && !EIKF_ANY(b.active && boxphase)) {
            // ---- every lane of the warp is on the march: one full-depth column per iteration
            int rr[8];
#pragma unroll
            for (int r = 0; r < 8; r++) rr[r] = (out && r < n_rows) ? rows[r] : 0;
            if (b.active) {
                col[-(long)LS] = kEdge; col[(size_t)b.ny * LS] = kEdge;
                spare[-(long)LS] = kEdge; spare[(size_t)b.ny * LS] = kEdge;
            }
            while (EIKF_ANY(b.active && b.X1 < b.mx)) {
                const bool need = b.active && b.X1 < b.mx;
                int line = 0;
                if (need) line = ++b.X1;
                const bool tie = march_sweep3<GM>(need, col, spare, L.S, b.my);
                if (EIKF_ANY(tie)) {   // an exact tie in the past column: follow the reference's order literally
                    if (tie) { col[-(long)LS] = kStop; col[(size_t)b.ny * LS] = kStop; }
                    int nohint = -1;
                    march_sweep(tie, col, spare, L.S, b.my, &nohint);
                    if (tie) { col[-(long)LS] = kEdge; col[(size_t)b.ny * LS] = kEdge; }
                }
                if (need) {
                    float* tmp = col; col = spare; spare = tmp;
                    if (out) {
#pragma unroll
                        for (int r = 0; r < 8; r++)
                            if (r < n_rows) out[(long)r * out_rstride + line] = col[(size_t)rr[r] * LS];
                        for (int r = 8; r < n_rows; r++) out[(long)r * out_rstride + line] = col[(size_t)rows[r] * LS];
                    }
                    if (full)
                        for (int y = 0; y < b.ny; y++) full[(size_t)line * b.ny + y] = col[(size_t)y * LS];
                }
            }
            break;
        }
        // ---- row above the box (reference x_side(Y0, -1), src/time_2d.c:934-939)
        {
            const bool need = b.active && b.Y0 > 0;
            if (EIKF_ANY(need)) {
                moved = true;
                int line = 0;
                bool slow = false, refill = true;
                if (need) {
                    line = --b.Y0;
                    if (b.preset_up) { slow = true; refill = false; }
                }
                // the coarse grid's last column of cells is masked (INF, src/time_2d.c:489-496): the last node of a row that
                // reaches it has no cell beyond it; the refined grid is not masked (:466), its rows are uniform up to the edge
                const bool edge = !FINE && b.X1 >= b.mx;
                // (The two slownesses are read through a volatile pointer on the coarse grid: with this function inlined into
                //  eik_pipe_kernel nvcc 12.9 / ptxas gave c the value of c2 -- S[line - 1] instead of S[line], INF on the top
                //  row -- at every optimisation level above -Xptxas -O1; found by printing one solve's box phase from both
                //  kernels (-DEIKF_TRACE), profiles/README.md r2c.  tests/test_pipe_gpu.py compares the kernels bit for bit.)
                float c = 1.f, c2 = kInf;      // c2 stays INF above the top row: no head wave there
                if (need) {
                    if (FINE) {
                        c = rowS(line);
                        if (line - 1 >= 0) c2 = rowS(line - 1);
                    } else {
                        const volatile float* sp = L.S + (long)line * LS;
                        c = (line >= b.my) ? kInf : sp[0];
                        if (line - 1 >= 0) c2 = sp[-(long)LS];
                    }
                }
                const bool fastlane = need && !slow;
                bool s2;
                if (D.row_march && row_ready(fastlane, L.ROW, LS, b.X1, mono_top)) {
                    float* wt = T + (size_t)line * LS;
                    s2 = EIKF_ANY(fastlane && edge)
                             ? row_march<true, GM>(fastlane, L.ROW, LS, b.X1, c, c2, wt, (long)b.ny * LS, &mono_top, edge ? b.X1 : -1)
                             : row_march<false, GM>(fastlane, L.ROW, LS, b.X1, c, c2, wt, (long)b.ny * LS, &mono_top);
                } else {
#ifdef EIKF_STATS
                    if (fastlane) g_stats[8]++;
#endif
                    s2 = fast_sweep<true>(fastlane, L.ROW, L.ROW, LS, true, 0, b.X1, med, c, c2,
                                          T + (size_t)line * LS, (long)b.ny * LS, nullptr, edge ? b.X1 : -1);
                    mono_top = false;
                }
#ifdef EIKF_TRACE
                if (!FINE && need && b.trace) printf("TR up line %d X1 %d slow %d s2 %d c %.6g c2 %.6g row[0] %.7g row[X1] %.7g win[0] %.7g\n", line, b.X1, (int)slow, (int)s2, c, c2, L.ROW[0], L.ROW[(size_t)b.X1 * LS], T[(size_t)line * LS]);
#endif
                if (need && (slow || s2)) {
                    const int rc = FINE ? slow_line<1>(T, b, fm, L, RL, line, -1, 0, b.X1, refill, col)
                                        : slow_line<1>(T, b, cm, L, RL, line, -1, 0, b.X1, refill, col);
                    if (rc != eik::kOk) status = rc;
                    b.preset_up = 0;
                    mono_top = false;
#ifdef EIKF_TRACE
                    if (!FINE && b.trace) printf("TR up line %d after slow: row[0] %.7g row[X1] %.7g win[0] %.7g\n", line, L.ROW[0], L.ROW[(size_t)b.X1 * LS], T[(size_t)line * LS]);
#endif
                }
                if (need) col[(size_t)line * LS] = L.ROW[(size_t)b.X1 * LS];
            }
        }
        // ---- column right of the box (reference y_side(X1, +1), src/time_2d.c:940-945)
        {
            const bool need = b.active && b.X1 < b.mx;
            if (EIKF_ANY(need)) {
                moved = true;
                int line = 0;
                if (need) line = ++b.X1;
                const bool wt = need && (FINE || boxphase) && line < wx;
                bool lockstep = false;
                if (LC && !FINE) {
                    // The two-chain loop of the march over the union of the lanes' ranges.  Outside its own range a lane's
                    // column buffers hold sentinels that grow away from the source depth (gm_sentinel), so its chains start
                    // and end at its own Y0 and Y1 exactly as they do at the ends of a full column.
                    const int kb = EIKF_MIN(need ? b.Y0 : 0x7fffffff), ke = EIKF_MAX(need ? b.Y1 : -1);
                    lockstep = true;
                    // (GM: a lane writes through over the whole union range: window nodes outside its own box are untimed
                    //  and are written again when a sweep of its own times them)
                    const bool tie = march_sweep3<GM>(need, col + (long)kb * LS, spare + (long)kb * LS, L.S + (long)kb * LS, ke - kb,
                                                      (GM && wt) ? T + ((size_t)line * b.ny + kb) * LS : nullptr);
                    if (need) {   // what the sweep left outside this lane's range is not a column value: sentinels again
                        for (int k = kb; k < b.Y0; k++) spare[(long)k * LS] = gm_sentinel(k, b.ys);
                        for (int k = b.Y1 + 1; k <= ke; k++) spare[(long)k * LS] = gm_sentinel(k, b.ys);
                    }
                    // An exact tie in the past column (4 in 100 000): the lane's walk (fast_sweep), from the intact past column
                    // into the other buffer.  In that discipline it reproduces every case of the reference's walk; a lane
                    // whose box is still growing in shared memory reports the ties the in-place walk gives up on and re-does the
                    // line on the window exactly where the in-place path does, so the two paths agree bit for bit on the device
                    // as well (the generic sweep rounds its square roots correctly, the fast sweeps use MUFU.SQRT).
                    bool s2 = false;
                    if (EIKF_ANY(tie))
                        s2 = fast_sweep<false>(tie, col, spare, LS, false, b.Y0, b.Y1, med, 0.f, 0.f,
                                               (tie && wt) ? T + (size_t)line * b.ny * LS : nullptr, LS, nullptr, -1, !GM && wt);
                    if (need) {
                        float* tmp = col; col = spare; spare = tmp;
                        if (!GM && wt && !tie) copy_strided(T + ((size_t)line * b.ny + b.Y0) * LS, LS, col + (long)b.Y0 * LS, LS, b.Y1 - b.Y0 + 1);
                    }
                    if (need && tie && s2) {
                        const int rc = slow_line<0>(T, b, cm, L, RL, line, 1, b.Y0, b.Y1, true, col);
                        if (rc != eik::kOk) status = rc;
                    }
                }
                if (!lockstep) {
                    // While the box is growing the column is swept in place (the row buffer is busy and the window
                    // is there to fall back on); on the march it ping-pongs between the column and the idle row buffer.
                    const bool inplace = boxphase;
                    float* dst = inplace ? col : spare;
                    const bool s2 = fast_sweep<false>(need, col, dst, LS, inplace, b.Y0, b.Y1, med, 0.f, 0.f,
                                                      wt ? T + (size_t)line * b.ny * LS : nullptr, LS, inplace ? nullptr : &hint);
                    if (need && !inplace) { spare = col; col = dst; }
                    if (need && s2) {   // an exact tie on an in-place column: re-done on the window
                        const int rc = FINE ? slow_line<0>(T, b, fm, L, RL, line, 1, b.Y0, b.Y1, true, col)
                                            : slow_line<0>(T, b, cm, L, RL, line, 1, b.Y0, b.Y1, true, col);
                        if (rc != eik::kOk) status = rc;
                    }
                }
                if (need) {
                    if (boxphase) {
                        if (b.Y0 > 0) L.ROW[(size_t)line * LS] = col[(size_t)b.Y0 * LS];
                        if (b.Y1 < b.my) L.ROW[(size_t)(RL - 1 - line) * LS] = col[(size_t)b.Y1 * LS];
                    } else {
                        if (out)
                            for (int r = 0; r < n_rows; r++) out[(long)r * out_rstride + line] = col[(size_t)rows[r] * LS];
                        if (full)
                            for (int y = 0; y < b.ny; y++) full[(size_t)line * b.ny + y] = col[(size_t)y * LS];
                    }
                }
            }
        }
        // ---- row below the box (reference x_side(Y1, +1), src/time_2d.c:946-951)
        {
            const bool need = b.active && b.Y1 < b.my;
            if (EIKF_ANY(need)) {
                moved = true;
                int line = 0;
                bool slow = false;
                if (need) line = ++b.Y1;
                const bool edge = !FINE && b.X1 >= b.mx;
                float c = 1.f, c2 = kInf;
                if (need) {        // volatile on the coarse grid: see the row above the box
                    if (FINE) {
                        c = rowS(line - 1);
                        c2 = rowS(line);
                    } else {
                        const volatile float* sp = L.S + (long)line * LS;
                        c = sp[-(long)LS];                       // line - 1 < my always
                        c2 = (line >= b.my) ? kInf : sp[0];
                    }
                }
                float* bot = L.ROW + (size_t)(RL - 1) * LS;
                const bool fastlane = need && !slow;
                bool s2;
                if (D.row_march && row_ready(fastlane, bot, -LS, b.X1, mono_bot)) {
                    float* wt = T + (size_t)line * LS;
                    s2 = EIKF_ANY(fastlane && edge)
                             ? row_march<true, GM>(fastlane, bot, -LS, b.X1, c, c2, wt, (long)b.ny * LS, &mono_bot, edge ? b.X1 : -1)
                             : row_march<false, GM>(fastlane, bot, -LS, b.X1, c, c2, wt, (long)b.ny * LS, &mono_bot);
                } else {
#ifdef EIKF_STATS
                    if (fastlane) g_stats[8]++;
#endif
                    s2 = fast_sweep<true>(fastlane, bot, bot, -LS, true, 0, b.X1, med, c, c2,
                                          T + (size_t)line * LS, (long)b.ny * LS, nullptr, edge ? b.X1 : -1);
                    mono_bot = false;
                }
#ifdef EIKF_TRACE
                if (!FINE && need && b.trace) printf("TR dn line %d X1 %d slow %d s2 %d c %.6g c2 %.6g win1[0] %.7g win2[0] %.7g\n", line, b.X1, (int)slow, (int)s2, c, c2, T[(size_t)1 * LS], T[(size_t)2 * LS]);
#endif
                if (need && (slow || s2)) {
                    const int rc = FINE ? slow_line<1>(T, b, fm, L, RL, line, 1, 0, b.X1, !slow, col)
                                        : slow_line<1>(T, b, cm, L, RL, line, 1, 0, b.X1, !slow, col);
                    if (rc != eik::kOk) status = rc;
                    mono_bot = false;
#ifdef EIKF_TRACE
                    if (!FINE && b.trace) printf("TR dn line %d after slow: win1[0] %.7g win2[0] %.7g\n", line, T[(size_t)1 * LS], T[(size_t)2 * LS]);
#endif
                }
                if (need) col[(size_t)line * LS] = L.ROW[(size_t)(RL - 1 - b.X1) * LS];
            }
        }
#ifdef EIKF_TRACE
        if (!FINE && b.active && b.trace) printf("TR round end X1 %d Y0 %d Y1 %d win1[0] %.7g win2[0] %.7g status %d\n", b.X1, b.Y0, b.Y1, T[(size_t)1 * LS], T[(size_t)2 * LS], status);
#endif
        if (b.active && boxphase && b.Y0 == 0 && b.Y1 == b.my) {
            boxphase = false;            // from here on the solve is a march over columns
            if (xbox_end) *xbox_end = b.X1;
            if (split) hand_over();
        }
        if (!EIKF_ANY(moved)) break;
    }
    return status;
}


// ---- one warp = up to 32 solves ----------------------------------------------------------------------
constexpr int kNeedGeneric = 1;   // status: the lane met a case only the generic solver handles

struct LaneTask {
    bool valid;
    int iz;
    const float* slow;   // global: h/v per depth cell of this lane's column, nz values
    float* out;          // receiver rows: out[r*out_rstride + x] = t[x][rows[r]]  (or nullptr)
    long out_rstride;
    float* full;         // whole field in the reference layout x*nz+y (or nullptr)
    float* hand_col;     // split mode: where the box phase leaves its last column (element k at [k*32]) ...
    int* hand_x1;        // ... and the column index it belongs to (-1: nothing left to march); nullptr = fused mode
#ifdef EIKF_TRACE
    int trace;
#endif
};

template <bool GM = false, bool LCT = false>
EIK_HD int solve_warp(const Dims& D, const Lane& L, const LaneTask& t, const int* rows, int n_rows)
{
    const int nx = D.nx, nz = D.nz, mx = nx - 1, my = nz - 1;
    int status = eik::kOk;
    // slowness column -> shared, dummy row masked the way the reference masks it (src/time_2d.c:489-492)
    if (t.valid) {
        for (int k = 0; k < my; k++) L.S[(size_t)k * LS] = t.slow[k];
        L.S[(size_t)my * LS] = kInf;
        L.S[-(long)LS] = kInf;     // "cell above the grid": lets the walk drop its k == 0 special cases
    }
    const eik::CoarseMedium cm{L.S, LS, mx, my};
    Box bc;   // coarse box
    bc.nx = nx; bc.ny = nz; bc.mx = mx; bc.my = my; bc.ys = t.iz; bc.X1 = 0; bc.Y0 = 0; bc.Y1 = 0; bc.preset_up = 0; bc.active = 0;
    bc.seed_pending = 0; bc.sX1 = 0; bc.sY0 = 0; bc.sY1 = 0; bc.shs0 = 0.f;
#ifdef EIKF_TRACE
    bc.trace = 0;
#endif
    Box bf = bc;   // refined box
#ifdef EIKF_TRACE
    bc.trace = t.valid ? t.trace : 0;
#endif
    int j0 = 0, hy = 0;
    bool whole = false;
    float hs0 = 0.f;
    if (t.valid) {
        eik::Grid<eik::CoarseMedium> g = make_grid(L.W, bc, cm);
        const int kind = eik::seed_search(g, true);
        bc.X1 = g.X1; bc.Y0 = g.Y0; bc.Y1 = g.Y1;
        hs0 = eik::source_slowness(g);
        whole = (kind == eik::kSeedBox && g.X1 == mx && g.Y0 == 0 && g.Y1 == my);
        if (!whole) {
            {
                // Only what can be read before a sweep writes it has to start as INF: the nodes round the source that the
                // minimal initialisation and the copy-back of the refined grid leave untimed.  Everything else in the window
                // is written through when it is timed and never read before (a window of the fine grid is 4.5 MB per lane;
                // on the Example plane clearing all of it was 2 GB of stores per launch of 1024 chains).  The host builds
                // start from windows full of NaN (tests/test_emu_cpu.py).
                const int xc = (D.wx - 1 < kInitMin + 2) ? D.wx - 1 : kInitMin + 2;
                const int ylo = (t.iz - kInitMin - 2 > 0) ? t.iz - kInitMin - 2 : 0, yhi = (t.iz + kInitMin + 2 < my) ? t.iz + kInitMin + 2 : my;
                for (int x = 0; x <= xc; x++)
                    for (int y = ylo; y <= yhi; y++) L.W[((size_t)x * nz + y) * LS] = kInf;
            }
            bc.active = 1;
        }
        if (kind == eik::kSeedRefine) {
            // geometry of the half-spacing grid (reference recursive_init, src/time_2d.c:844-864)
            int nxf, nyf, xsf, ysf, i0, hx;
            eik::refine_axis(0, nx, &nxf, &xsf, &i0, &hx);
            eik::refine_axis(t.iz, nz, &nyf, &ysf, &j0, &hy);
            bf.nx = nxf; bf.ny = nyf; bf.mx = nxf - 1; bf.my = nyf - 1; bf.ys = ysf;
            const eik::FineMedium fm{cm, j0, hy};
            for (int i = 0; i < nxf * nyf; i++) L.WF[(size_t)i * LS] = kInf;
            eik::Grid<eik::FineMedium> f = make_grid(L.WF, bf, fm);
            const int kf = eik::seed_search(f, false);
            eik::seed_fill(f, kf);
            bf.X1 = f.X1; bf.Y0 = f.Y0; bf.Y1 = f.Y1;
            bf.preset_up = (kf == eik::kSeedNearest && ysf > 0 && ysf < bf.my) ? 1 : 0;
            bf.active = 1;
            load_perimeter(f, L, D.row_len, L.COL);
        } else if (!whole) {
            if (GM && kind == eik::kSeedBox && !t.full) {
                // the outline of the homogeneous box (what the sweeps start from) and its receiver rows (what is output);
                // the interior follows if a slow path ever looks at it (slow_line)
                for (int y = g.Y0; y <= g.Y1; y++) g.T(g.X1, y) = eik::box_time(hs0, g.X1, y - t.iz);
                for (int x = 0; x <= g.X1; x++) { g.T(x, g.Y0) = eik::box_time(hs0, x, g.Y0 - t.iz); g.T(x, g.Y1) = eik::box_time(hs0, x, g.Y1 - t.iz); }
                if (t.out)
                    for (int r = 0; r < n_rows; r++)
                        if (rows[r] > g.Y0 && rows[r] < g.Y1)
                            for (int x = 0; x <= g.X1; x++) g.T(x, rows[r]) = eik::box_time(hs0, x, rows[r] - t.iz);
                bc.seed_pending = 1; bc.sX1 = g.X1; bc.sY0 = g.Y0; bc.sY1 = g.Y1; bc.shs0 = hs0;
            } else {
                eik::seed_fill(g, kind);
            }
            bc.preset_up = (kind == eik::kSeedNearest && t.iz > 0 && t.iz < my) ? 1 : 0;
        }
    }
    EIKF_SYNC();
    // ---- refined grids of the lanes that need one
    if (EIKF_ANY(bf.active)) {
        const int rc = run_grid<true, GM, LCT>(bf, L, D, cm, j0, hy, nullptr, 0, nullptr, 0, nullptr, nullptr);
        if (rc != eik::kOk) status = rc;
        if (bf.active) {
            // every second fine node is a coarse node (src/time_2d.c:887-890); the fine field is complete in WF
            for (int i = 0, ii = 0; ii < bf.nx; ii += 2, i++)
                copy_strided(L.W + ((size_t)i * nz + j0 + hy) * LS, LS, L.WF + ((size_t)ii * bf.ny + hy) * LS, 2 * LS,
                             (bf.ny - hy + 1) / 2);
            bc.X1 = (kInitMin < mx) ? kInitMin : mx;
            bc.Y0 = (t.iz - kInitMin > 0) ? t.iz - kInitMin : 0;
            bc.Y1 = (t.iz + kInitMin < my) ? t.iz + kInitMin : my;
        }
    }
    if (bc.active) {
        eik::Grid<eik::CoarseMedium> g = make_grid(L.W, bc, cm);
        load_perimeter(g, L, D.row_len, L.COL);
    }
    EIKF_SYNC();
    // ---- coarse grids
    int xbox_end = -1;
    {
        if (t.hand_x1) *t.hand_x1 = -1;
        const int rc = run_grid<false, GM, LCT>(bc, L, D, cm, 0, 0, t.out, t.out_rstride, rows, n_rows, t.full, &xbox_end, t.hand_col,
                                       t.hand_x1);
        if (rc != eik::kOk) status = rc;
    }
    // ---- the part of the output that was computed while the box was still growing (it may have been
    //      re-timed by reverse propagation until the very end of that phase) comes from the window
#ifdef EIKF_TRACE
    if (t.valid && t.trace) printf("TR out whole %d xbox_end %d rows %d %d win1[0] %.7g win2[0] %.7g\n", (int)whole, xbox_end, rows[0], n_rows > 1 ? rows[1] : -1, L.W[(size_t)1 * LS], L.W[(size_t)2 * LS]);
#endif
    if (t.valid && !whole) {
        if (t.out)
            for (int r = 0; r < n_rows; r++)
                copy_strided(t.out + (long)r * t.out_rstride, 1, L.W + (size_t)rows[r] * LS, (long)nz * LS, xbox_end + 1);
        if (t.full)
            for (int x = 0; x <= xbox_end; x++)
                for (int y = 0; y < nz; y++) t.full[(size_t)x * nz + y] = L.W[((size_t)x * nz + y) * LS];
    }
    if (t.valid && whole) {   // homogeneous model: the exact solution everywhere, nothing to propagate
        if (t.out)
            for (int r = 0; r < n_rows; r++)
                for (int x = 0; x < nx; x++) t.out[(long)r * t.out_rstride + x] = eik::box_time(hs0, x, rows[r] - t.iz);
        if (t.full)
            for (int x = 0; x < nx; x++)
                for (int y = 0; y < nz; y++) t.full[(size_t)x * nz + y] = eik::box_time(hs0, x, y - t.iz);
    }
    return status;
}


// lane pointers into a warp's shared-memory slice (base already offset by the lane)
EIK_HD void carve_shared(float* base, const Dims& D, Lane* L)
{
    L->S = base + (size_t)1 * LS;
    L->COL = L->S + (size_t)D.nz * LS + (size_t)1 * LS;
    L->ROW = L->COL + (size_t)(D.col_len + 1) * LS;
    L->COL2 = D.lock_cols ? L->ROW + (size_t)D.row_len * LS + (size_t)1 * LS : nullptr;
}
// the same for a slice in global memory, with the second column buffer behind the rows
EIK_HD void carve_global(float* base, const Dims& D, Lane* L)
{
    carve_shared(base, D, L);
    L->COL2 = L->ROW + (size_t)D.row_len * LS + (size_t)1 * LS;
}

}  // namespace eikf
