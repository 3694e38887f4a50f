// eikonal.cu -- batched Podvin-Lecomte eikonal kernels for sm_100a.
//
// Replaces the nz serial calls of time_2d() per phase and chain in the reference
// (src/misfit.c:270-289 -> src/time_2d.c:301) by one launch over every
// (chain, phase, source depth) triple.  One lane owns one solve: the update order of
// the expanding-box scheme is sequential inside a solve and must be kept for parity
// (SURVEY.md section 7 H1), so the parallelism is across solves.
#include "eikonal.cuh"
#include "eik_core.cuh"
#include "launch_count.h"

namespace mq {

size_t eik_scratch_floats_per_warp(int nxmod, int nz)
{
    return ((size_t)nxmod * nz + kFineNodes) * 32;
}

// Generic kernel: the whole time field of a lane lives in global memory, interleaved by
// lane (node i of lane l at scratch[i*32 + l]) so that lanes that sweep the same node
// -- the common case, all lanes of a warp share the source depth -- touch one 128-byte line.
__global__ void __launch_bounds__(128)
eik_generic_kernel(EikBatch b)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int warp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * warps_per_block;
    const int nodes = b.nxmod * b.nz;
    float* W = b.scratch + (size_t)warp * (((size_t)nodes + kFineNodes) * 32) + lane;
    float* WF = W + (size_t)nodes * 32;
    const int n_items = b.n_items_dev ? *b.n_items_dev : b.n_items;
    const int n_solves = b.src_iz ? b.n_solves : n_items * b.nz;
    const int n_tasks = (n_solves + 31) >> 5;

    for (int task = warp; task < n_tasks; task += n_warps) {
        const int g = task * 32 + lane;
        if (g < n_solves) {
            int iz, item;
            if (b.src_iz) { iz = b.src_iz[g]; item = g; }
            else { iz = g / n_items; item = g - iz * n_items; }
            const float* s = b.slow + (size_t)item * b.nz;
            const int rc = eik::solve(s, 1, b.nxmod, b.nz, iz, W, WF, 32, nullptr);
            if (b.status) b.status[g] = rc;
            if (b.status_min && rc < 0) atomicMin(b.status_min, rc);
            if (b.full_out) {
                float* o = b.full_out + (size_t)g * nodes;
                for (int i = 0; i < nodes; i++) o[i] = W[(size_t)i * 32];
            }
            if (b.n_rows > 0 && (b.row_out || b.row_out_base)) {
                float* tab = b.row_out ? b.row_out[item] : b.row_out_base + (size_t)item * b.row_item_stride;
                for (int r = 0; r < b.n_rows; r++) {
                    const int j = b.rows[r];
                    float* o = tab + ((size_t)r * b.nz + iz) * b.xpitch;
                    for (int x = 0; x < b.nxmod; x++) o[x] = W[((size_t)x * b.nz + j) * 32];
                }
            }
        }
        __syncwarp();
    }
}

cudaError_t eik_launch_generic(const EikBatch& b, cudaStream_t stream)
{
    const int n_solves = b.src_iz ? b.n_solves : b.n_items * b.nz;   // upper bound when n_items_dev is set
    if (n_solves <= 0) return cudaSuccess;
    const int n_tasks = (n_solves + 31) / 32;
    const int warps = n_tasks < b.max_warps ? n_tasks : b.max_warps;
    const int wpb = 4;
    const int blocks = (warps + wpb - 1) / wpb;
    eik_generic_kernel<<<blocks, wpb * 32, 0, stream>>>(b);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mq
