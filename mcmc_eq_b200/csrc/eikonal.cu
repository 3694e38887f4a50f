// eikonal.cu -- batched Podvin-Lecomte eikonal kernels for sm_100a.
//
// Replaces the nz serial calls of time_2d() per phase and chain in the reference
// (src/misfit.c:270-289 -> src/time_2d.c:301) by one launch over every
// (chain, phase, source depth) triple.  One lane owns one solve: the update order of
// the expanding-box scheme is sequential inside a solve and must be kept for parity
// (SURVEY.md section 7 H1), so the parallelism is across solves.
#include "eikonal.cuh"
#include "eik_core.cuh"
#include "eik_fast.cuh"
#include <stdlib.h>
#include <string.h>
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


// ---- warp-synchronous kernel ---------------------------------------------------------------------------
static eikf::Dims fast_dims(int nxmod, int nz) { return eikf::make_dims(nxmod, nz); }
static size_t fast_smem_floats_per_warp(const eikf::Dims& D) { return (size_t)eikf::smem_floats_per_lane(D) * 32; }
static size_t fast_scratch_floats_per_warp(const eikf::Dims& D) { return ((size_t)D.wx * D.nz + kFineNodes) * 32; }

bool eik_fast_supported(int nxmod, int nz)
{
    if (nxmod < 2 || nz < 2) return false;
    const eikf::Dims D = fast_dims(nxmod, nz);
    return fast_smem_floats_per_warp(D) * sizeof(float) <= 200 * 1024;
}

// One warp per block: a warp owns 32 solves and its slice of shared memory, nothing is shared between warps.
__global__ void __launch_bounds__(32, 9) eik_fast_kernel(EikBatch b, eikf::Dims D)
{
    extern __shared__ float smem[];
    const int lane = threadIdx.x;
    const int nodes = b.nxmod * b.nz;
    eikf::Lane L;
    eikf::carve_shared(smem + lane, D, &L);
    L.W = b.scratch + (size_t)blockIdx.x * (((size_t)D.wx * D.nz + kFineNodes) * 32) + lane;
    L.WF = L.W + (size_t)D.wx * D.nz * 32;
    const int n_items = b.n_items_dev ? *b.n_items_dev : b.n_items;
    const int n_solves = b.src_iz ? b.n_solves : n_items * b.nz;
    const int n_tasks = (n_solves + 31) >> 5;

    for (int task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        const int g = task * 32 + lane;
        eikf::LaneTask t;
        t.valid = g < n_solves;
        t.iz = 0; t.slow = nullptr; t.out = nullptr; t.out_rstride = 0; t.full = nullptr;
        if (t.valid) {
            int item;
            if (b.src_iz) { t.iz = b.src_iz[g]; item = g; }
            else { t.iz = g / n_items; item = g - t.iz * n_items; }
            t.slow = b.slow + (size_t)item * b.nz;
            if (b.n_rows > 0 && (b.row_out || b.row_out_base)) {
                float* tab = b.row_out ? b.row_out[item] : b.row_out_base + (size_t)item * b.row_item_stride;
                t.out = tab + (size_t)t.iz * b.xpitch;
                t.out_rstride = (long)b.nz * b.xpitch;
            }
            if (b.full_out) t.full = b.full_out + (size_t)g * nodes;
        }
        const int rc = eikf::solve_warp(D, L, t, b.rows, b.n_rows);
        if (t.valid) {
            if (b.status) b.status[g] = rc;
            if (b.status_min && rc < 0) atomicMin(b.status_min, rc);
        }
        __syncwarp();
    }
}

int eik_fast_max_warps(int nxmod, int nz, int device)
{
    const eikf::Dims D = fast_dims(nxmod, nz);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int per_sm = 0;
    const size_t smem = fast_smem_floats_per_warp(D) * sizeof(float);
    cudaFuncSetAttribute(eik_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(eik_fast_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, eik_fast_kernel, 32, smem);
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

cudaError_t eik_launch_fast(const EikBatch& b, cudaStream_t stream)
{
    const int n_solves = b.src_iz ? b.n_solves : b.n_items * b.nz;
    if (n_solves <= 0) return cudaSuccess;
    const eikf::Dims D = fast_dims(b.nxmod, b.nz);
    const size_t smem = fast_smem_floats_per_warp(D) * sizeof(float);
    static int configured_for = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    static int resident = 0;
    if (configured_for != dev * 100000 + b.nz) {
        resident = eik_fast_max_warps(b.nxmod, b.nz, dev);
        configured_for = dev * 100000 + b.nz;
    }
    const int n_tasks = (n_solves + 31) / 32;
    // scratch was sized for max_warps warps of the generic kernel, which is never less per warp
    const size_t have = (size_t)b.max_warps * eik_scratch_floats_per_warp(b.nxmod, b.nz);
    long warps = (long)(have / fast_scratch_floats_per_warp(D));
    if (warps > resident) warps = resident;
    if (warps > n_tasks) warps = n_tasks;
    if (warps < 1) warps = 1;
    eik_fast_kernel<<<(unsigned)warps, 32, smem, stream>>>(b, D);
    count_launch();
    return cudaGetLastError();
}

cudaError_t eik_launch(const EikBatch& b, cudaStream_t stream)
{
    static int force_generic = -1;
    if (force_generic < 0) {
        const char* e = getenv("MCMCEQ_EIKONAL");
        force_generic = (e && strcmp(e, "generic") == 0) ? 1 : 0;
    }
    if (!force_generic && eik_fast_supported(b.nxmod, b.nz)) return eik_launch_fast(b, stream);
    return eik_launch_generic(b, stream);
}

}  // namespace mq
