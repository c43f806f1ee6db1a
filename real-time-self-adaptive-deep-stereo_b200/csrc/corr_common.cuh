// Warp-coordinate helper shared by the correlation kernels (corr.cu, corr_tma.cu).
// Restates MadNet._linear_warping's tap/weight rule (reference Nets/MadNet.py:400-436).
#pragma once
#include "common.cuh"

namespace ms {

struct WarpTap { int i0, i1; float w0, w1; };

// warp coordinates for target column xp (already known to be inside [0,w)); uu = u[xp] or 0
__device__ __forceinline__ WarpTap warp_tap(int xp, float uu, int w, bool warped) {
    WarpTap t;
    if (!warped) { t.i0 = xp; t.i1 = xp; t.w0 = 1.f; t.w1 = 0.f; return t; }
    float cx = (float)xp + uu;
    float x0 = floorf(cx), x1 = x0 + 1.f;
    float x0s = fminf(fmaxf(x0, 0.f), (float)(w - 1));
    float x1s = fminf(fmaxf(x1, 0.f), (float)(w - 1));
    t.w0 = (x1 - cx) * (x0 == x0s ? 1.f : 0.f);
    t.w1 = (cx - x0) * (x1 == x1s ? 1.f : 0.f);
    t.i0 = (int)x0s; t.i1 = (int)x1s;
    return t;
}

struct Tap { int i0, i1; float w0, w1; };

}  // namespace ms
