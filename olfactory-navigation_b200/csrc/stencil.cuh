// Grid-move stencil helpers shared by the belief-step and infotaxis kernels.
// Transition of exact_converter (reference agents/model_based_util/environment_converter.py:111-122):
// R_a(y,x) = (B_y(y+dy), B_x(x+dx)), B = clip ("stop") or wrap per axis (reference environment.py:711-769).
#pragma once
#include "common.cuh"

__device__ __forceinline__ float comp4(const float4& v, int k) {
    return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}
__device__ __forceinline__ void add4(float4& v, int k, float a) {
    if (k == 0) v.x += a; else if (k == 1) v.y += a; else if (k == 2) v.z += a; else v.w += a;
}
__device__ __forceinline__ void set4(float4& v, int k, float a) {
    if (k == 0) v.x = a; else if (k == 1) v.y = a; else if (k == 2) v.z = a; else v.w = a;
}

// Pushed values for destination columns [4*c4, 4*c4+3] coming from ONE source grid row `row`
// (global or shared memory, generic pointer), x-move dx in {-1,0,+1}.  Scalar neighbour loads:
// used on the rare paths (clip edge row, wrapped halo row).
__device__ __forceinline__ float4 xshift_row(const float* row, int c4, int dx, int W, int W4, int wrap_x) {
    const float4 v = *reinterpret_cast<const float4*>(row + 4 * c4);
    if (dx == 0) return v;
    float4 out;
    const int e = W - 1;
    if (dx > 0) {
        const float left = (c4 > 0) ? row[4 * c4 - 1] : (wrap_x ? row[e] : 0.f);
        out = make_float4(left, v.x, v.y, v.z);
        if (!wrap_x && (e >> 2) == c4) add4(out, e & 3, comp4(v, e & 3));      // clipped cell keeps its own mass
    } else {
        const float right = (c4 < W4 - 1) ? row[4 * c4 + 4] : 0.f;
        out = make_float4(v.y, v.z, v.w, right);
        if (wrap_x && (e >> 2) == c4) set4(out, e & 3, row[0]);
        if (!wrap_x && c4 == 0) out.x += v.x;
    }
    return out;
}


// Pushed belief b_a for destination grid row y, float4 column c4, read from a row-padded belief `bin`
// (global memory).  Handles every boundary case; scalar neighbour loads (L1/L2 hits).
__device__ __forceinline__ float4 pushed4(const float* bin, int y, int c4, int dy, int dx, const Geom& g) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int yp = y - dy;
    if (dy == 0 || g.wrap_y) {
        const int ys = (yp < 0) ? yp + g.H : (yp >= g.H ? yp - g.H : yp);
        return xshift_row(bin + (size_t)ys * g.Wp, c4, dx, g.W, g.W4, g.wrap_x);
    }
    if (yp >= 0 && yp < g.H) acc = xshift_row(bin + (size_t)yp * g.Wp, c4, dx, g.W, g.W4, g.wrap_x);
    if (y == (dy > 0 ? g.H - 1 : 0)) {     // clipped edge row keeps its own mass
        const float4 q = xshift_row(bin + (size_t)y * g.Wp, c4, dx, g.W, g.W4, g.wrap_x);
        acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
    }
    return acc;
}
