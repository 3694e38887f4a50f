// forward.cuh -- launchers of the forward-model kernels (rasterise, misfit, totals).
#pragma once
#include "state.h"

namespace mq {

// items for "rebuild the tables of every chain" (mq_forward): phases selected by calct (1 P, 2 S, 3 both)
cudaError_t launch_build_items_all(Handle* h, const EvalView& v, int calct);
// slow[item][iz] = h / v(z0 + iz*h) from the model the view points at (src/misfit.c:205-214,256-266)
cudaError_t launch_rasterise(Handle* h, const EvalView& v, int max_items);
// eikonal for every item x source depth into the item's table (src/misfit.c:270-289)
cudaError_t launch_tables(Handle* h, int max_items);
// residuals, origin times and per-event class sums (src/misfit.c:83-153)
cudaError_t launch_misfit(Handle* h, const EvalView& v);
// one bilinear lookup with the kernel's own code (mq_traveltimet)
bool traveltimet_cell(float dist, float z, float hgrid, float z0, int nz, int nxmod, int* iz1, int* m1);
cudaError_t launch_traveltimet(float dist, float z, float hgrid, float z0, int nz, int nxmod, const float* d_corners, float* d_out,
                               cudaStream_t s);
// per-chain class sums mf_eval[c][8] from the per-event sums of the view
cudaError_t launch_totals(Handle* h, const EvalView& v);

// CUDA-event timing of the eikonal launches (mq_profile)
void profile_enable(Handle* h, bool on);
void profile_collect(Handle* h, double* ms, long* launches, bool reset);
void profile_destroy(Handle* h);
// time and launches per eikonal kernel (indices kEik* of eikonal.cuh) of the accumulation profile_collect last read
void profile_by_kernel(Handle* h, double* ms, long* launches);
// the same for the misfit kernel
void profile_misfit(Handle* h, double* ms, long* launches);

}  // namespace mq
