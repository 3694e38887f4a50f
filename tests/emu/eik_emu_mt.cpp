// tests/emu/eik_emu_mt.cpp -- TEST-ONLY host build of the warp-synchronous solver (csrc/eik_fast.cuh) with a warp of
// EIKF_HOST_LANES host threads (host_warp.h): one solve per lane, all lanes in lock-step, the collectives real.
// mode 0: shared-memory slice, per-lane in-place columns; 1: slice in global memory (eik_fine_kernel); 2: shared-memory slice
// with the second column buffer (Dims::lock_cols).  split != 0: the box phase hands its last column over (as it does to the
// tensor-memory march of eik_pipe_kernel) and the march is finished here by march_sweep3 on a copy of that column.
#include <atomic>
#include <thread>
#include <vector>
#define EIKF_HOST_WARP "../../tests/emu/host_warp.h"   // included by eik_fast.cuh inside namespace eikf
#include "../../mcmc_eq_b200/csrc/eik_fast.cuh"
#include <thread>
#include <vector>

using namespace eikf;   // LS and host_warp live in eikf (the header is included from inside it)

extern "C" int emu_mt_lanes() { return EIKF_HOST_LANES; }
// collectives that the lanes of the warp reached from different call sites since the library was loaded (must stay 0)
extern "C" int emu_mt_site_mismatches() { return host_warp::team().mismatches.load(); }

// what the scratch arrays hold before a solve (on the GPU: whatever the previous task left there)
static float g_fill = 0.f;
extern "C" void emu_mt_set_fill(float v) { g_fill = v; }

// slow: [lanes][ny] slowness columns, iz: [lanes] source depths (-1: lane without a solve), full: [lanes][nx*ny] fields,
// rows_out: [lanes][n_rows][nx], status: [lanes]
extern "C" void emu_mt_time_2d(const float* slow, int nx, int ny, const int* iz, float* full, const int* rows, int n_rows,
                               float* rows_out, int* status, int mode, int split)
{
    constexpr int W = EIKF_HOST_LANES;
    eikf::Dims D = eikf::make_dims(nx, ny);
    D.lock_cols = (mode == 2) ? 1 : 0;
    std::vector<float> SM((size_t)eikf::gmem_floats_per_lane(D) * W, g_fill), Wn(((size_t)D.wx * ny) * W, g_fill), WF((size_t)22 * 43 * W, g_fill);
    std::vector<float> hand((size_t)(ny + 2) * 3 * W, g_fill);
    auto body = [&](int lane) {
        eikf::host_warp::lane() = lane;
        eikf::Lane L;
        if (mode == 1) eikf::carve_global(SM.data() + lane, D, &L);
        else eikf::carve_shared(SM.data() + lane, D, &L);
        L.W = Wn.data() + lane; L.WF = WF.data() + lane;
        eikf::LaneTask t;
        t.valid = iz[lane] >= 0; t.iz = t.valid ? iz[lane] : 0; t.slow = slow + (size_t)lane * ny;
        t.out = rows_out + (size_t)lane * n_rows * nx; t.out_rstride = nx;
        t.full = split ? nullptr : full + (size_t)lane * nx * ny;
        int x1 = -1;
        t.hand_col = split ? L.COL : nullptr;
        t.hand_x1 = split ? &x1 : nullptr;
        int rc = (mode == 1) ? eikf::solve_warp<true, false>(D, L, t, rows, n_rows)
               : (mode == 2) ? eikf::solve_warp<false, true>(D, L, t, rows, n_rows) : eikf::solve_warp<false, false>(D, L, t, rows, n_rows);
        if (split) {
            // what eik_pipe_kernel does with the hand-over: past | current column with the end slots, slowness column
            const int ke = ny - 1, mx = nx - 1;
            float* P = hand.data() + lane + (size_t)1 * LS;                        // node k at P[k*LS], k = -1 .. ny
            float* C = P + (size_t)(ny + 2) * LS;
            float* S = C + (size_t)(ny + 2) * LS;
            const bool live = t.valid && x1 >= 0 && x1 < mx;
            for (int k = 0; k <= ke; k++) { P[(long)k * LS] = live ? L.COL[(size_t)k * LS] : 0.f; S[(long)k * LS] = (live && k < ke) ? L.S[(size_t)k * LS] : eikf::kInf; }
            S[-(long)LS] = eikf::kInf;
            const int xlo = EIKF_MIN(live ? x1 : 0x7fffffff);
            if (EIKF_ANY(live)) {
                for (int line = xlo + 1; line <= mx; line++) {
                    const bool need = live && line > x1;
                    P[-(long)LS] = eikf::kEdge; P[(long)ny * LS] = eikf::kEdge; C[-(long)LS] = eikf::kEdge; C[(long)ny * LS] = eikf::kEdge;
                    const bool tie = eikf::march_sweep3<false>(need, P, C, S, ke);
                    if (EIKF_ANY(tie)) {
                        if (tie) { P[-(long)LS] = eikf::kStop; P[(long)ny * LS] = eikf::kStop; }
                        int nohint = -1;
                        eikf::march_sweep(tie, P, C, S, ke, &nohint);
                    }
                    if (need) {
                        for (int r = 0; r < n_rows; r++) t.out[(long)r * nx + line] = C[(long)rows[r] * LS];
                        float* tmp = P; P = C; C = tmp;
                    }
                }
            }
        }
        status[lane] = rc;
    };
    std::vector<std::thread> th;
    for (int lane = 1; lane < W; lane++) th.emplace_back(body, lane);
    body(0);
    for (auto& x : th) x.join();
}
