// tria.cuh -- the linear-gradient parameterisation (config line 29, TRIA = 1): velocity interpolated linearly between
// the depth-sorted nuclei instead of the Voronoi step function (src/misfit.c:217-250, src/analyse_eq.c:573-598).
//
// The reference sorts a copy of the model by depth (bubble sort, strict >, so stable) and, for a depth z, takes the
// segment i with zs[i] <= z < zs[i+1]; when no segment holds z (z at or below the deepest nucleus) its index variable
// keeps the value of the previous depth node.  Without the sort: the segment's upper node is the nucleus with the
// largest depth <= z -- among equal depths the one a stable sort places last, i.e. the highest index -- and its lower
// node the nucleus with the smallest depth > z, the lowest index among equals.
#pragma once

namespace tria {

// nuclei (lo, hi) of the segment that holds zq; false when there is none
__device__ __forceinline__ bool segment(const float* z, int dim, float zq, int* lo, int* hi)
{
    int a = -1, b = -1;
    float za = 0.f, zb = 0.f;
    for (int i = 0; i < dim; i++) {
        const float zi = z[i];
        if (zi <= zq) { if (a < 0 || zi >= za) { a = i; za = zi; } }
        else if (b < 0 || zi < zb) { b = i; zb = zi; }
    }
    *lo = a; *hi = b;
    return a >= 0 && b >= 0;
}

// segment used for depth node iz of the grid z0 + iz*h: the node's own, else that of the nearest node above that has one
__device__ __forceinline__ void segment_of_node(const float* z, int dim, int iz, float hgrid, float z0, int* lo, int* hi)
{
    for (int q = iz; q >= 0; q--) {
        const float zq = __fadd_rn(z0, __fmul_rn((float)q, hgrid));
        if (segment(z, dim, zq, lo, hi)) return;
    }
    // no node up to iz lies inside the model (the reference reads an uninitialised index here): first two nuclei
    *lo = 0; *hi = dim > 1 ? 1 : 0;
}

// a*z + b through (z0, v0), (z1, v1) in the reference's float arithmetic (src/misfit.c:243-245)
__device__ __forceinline__ float line_through(float zq, float z0, float v0, float z1, float v1)
{
    const float a = __fdiv_rn(__fsub_rn(v1, v0), __fsub_rn(z1, z0));
    const float b = __fsub_rn(v0, __fmul_rn(a, z0));
    return __fadd_rn(__fmul_rn(a, zq), b);
}

}  // namespace tria
