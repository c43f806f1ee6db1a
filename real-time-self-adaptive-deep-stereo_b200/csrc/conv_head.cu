// Direct kernels for the single-channel disparity heads (bandwidth-bound, no tensor cores: N = 1 has no reuse).
//
// Replaces the cuDNN calls behind the linear 3x3 -> 1 convolutions of the reference: estimator `disp-6`
// (Nets/MadNet.py:113-118), `context-7` with the residual add (:160-168), DispNet `predict` / `prediction`
// (Nets/DispNet.py:49-50,143-146), and their input gradients (1 -> cin).  Round 1 ran these through the generic fp32
// gather GEMM (68 us per launch at 96x320x32); here a pixel's channel vector is one coalesced 128-bit load per lane.
#include <algorithm>

#include "common.cuh"

namespace ms {

// forward: y[p] = act(sum_{tap,c} x[p + tap][c] * w[tap][c] + b) (+ res[p]); 8 lanes per pixel, 4 channels per lane per step
__global__ void __launch_bounds__(256)
conv_head_fwd_kernel(ConvGemm g, size_t npix) {
    pdl_prologue();
    extern __shared__ float w_s[];                       // [taps][C]
    const int C = g.x.c, taps = g.kh * g.kw;
    for (int i = threadIdx.x; i < taps * C; i += blockDim.x) w_s[i] = g.wmat[i];
    __syncthreads();
    const int sub = threadIdx.x & 7;
    const size_t pix = (size_t)blockIdx.x * 32 + (threadIdx.x >> 3);
    const bool live = pix < npix;
    float acc = 0.f;
    if (live) {
        const int W = g.y.w, H = g.y.h;
        const int ox = (int)(pix % W);
        const size_t t = pix / W;
        const int oy = (int)(t % H);
        const size_t img = t / H;
        for (int r = 0; r < g.kh; ++r) {
            const int iy = oy + g.off_y + r * g.step;
            if (iy < 0 || iy >= g.x.h) continue;
            for (int s = 0; s < g.kw; ++s) {
                const int ix = ox + g.off_x + s * g.step;
                if (ix < 0 || ix >= g.x.w) continue;
                const float* xp = g.x.p + ((img * g.x.h + iy) * g.x.w + ix) * g.x.cs;
                const float* wp = w_s + (r * g.kw + s) * C;
                for (int c = sub * 4; c < C; c += 32) {
                    const float4 v = *reinterpret_cast<const float4*>(xp + c);
                    const float4 w = *reinterpret_cast<const float4*>(wp + c);
                    acc = fmaf(v.x, w.x, acc); acc = fmaf(v.y, w.y, acc); acc = fmaf(v.z, w.z, acc); acc = fmaf(v.w, w.w, acc);
                }
            }
        }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (live && sub == 0) {
        float t = acc + (g.bias ? g.bias[0] : 0.f);
        t = fmaxf(g.alpha * t, t);
        if (g.res) t += g.res[pix * g.res_cs];
        float* yp = g.y.p + pix * g.y.cs;
        if (g.accumulate) t += *yp;
        if (g.mask) t *= (g.mask[pix * g.mask_cs] > 0.f) ? 1.f : g.mask_alpha;
        *yp = t;
    }
}

// input gradient of a 1-channel conv: dx[p][c] = sum_tap dy[p + off + tap*step] * w[tap][c]   (then accumulate / mask)
__global__ void __launch_bounds__(256)
conv_head_dgrad_kernel(ConvGemm g, size_t npix) {
    pdl_prologue();
    extern __shared__ float w_s[];                       // [taps][C]
    const int C = g.y.c, taps = g.kh * g.kw, cq = C >> 2;
    for (int i = threadIdx.x; i < taps * C; i += blockDim.x) w_s[i] = g.wmat[i];
    __syncthreads();
    const size_t total = npix * cq;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t pix = i / cq;
        const int c = (int)(i - pix * cq) * 4;
        const int W = g.y.w, H = g.y.h;
        const int ox = (int)(pix % W);
        const size_t t = pix / W;
        const int oy = (int)(t % H);
        const size_t img = t / H;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < g.kh; ++r) {
            const int iy = oy + g.off_y + r * g.step;
            if (iy < 0 || iy >= g.x.h) continue;
            for (int s = 0; s < g.kw; ++s) {
                const int ix = ox + g.off_x + s * g.step;
                if (ix < 0 || ix >= g.x.w) continue;
                const float d = __ldg(g.x.p + ((img * g.x.h + iy) * g.x.w + ix) * g.x.cs);
                const float4 w = *reinterpret_cast<const float4*>(w_s + (r * g.kw + s) * C + c);
                a.x = fmaf(d, w.x, a.x); a.y = fmaf(d, w.y, a.y); a.z = fmaf(d, w.z, a.z); a.w = fmaf(d, w.w, a.w);
            }
        }
        float* yp = g.y.p + pix * g.y.cs + c;
        if (g.accumulate) { const float4 o = *reinterpret_cast<const float4*>(yp); a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w; }
        if (g.mask) {
            const float4 m = *reinterpret_cast<const float4*>(g.mask + pix * g.mask_cs + c);
            a.x *= m.x > 0.f ? 1.f : g.mask_alpha; a.y *= m.y > 0.f ? 1.f : g.mask_alpha;
            a.z *= m.z > 0.f ? 1.f : g.mask_alpha; a.w *= m.w > 0.f ? 1.f : g.mask_alpha;
        }
        *reinterpret_cast<float4*>(yp) = a;
    }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// 1 = forward head (y.c == 1), 2 = head dgrad (x.c == 1), 0 = not a head
// single-channel gather (1 -> 1 channel, any generic gather geometry): DispNet's `up_predict` 4x4 stride-2 conv_transpose of
// a disparity map (Nets/DispNet.py:51-53).  One thread per output pixel; the generic fp32 gather GEMM spent 166 us on the
// 192x640 instance of this 2 MFLOP operation.
__global__ void __launch_bounds__(256) conv_one_channel_kernel(ConvGemm g, size_t npix) {
    pdl_prologue();
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const int W = g.y.w, H = g.y.h;
    const int ox = (int)(pix % W);
    const size_t t = pix / W;
    const int oy = (int)(t % H);
    const size_t img = t / H;
    float acc = g.bias ? g.bias[0] : 0.f;
    for (int r = 0; r < g.kh; ++r) {
        int ty = oy * g.mul + g.off_y + r * g.step;
        if (g.div > 1) { if (ty % g.div) continue; ty /= g.div; }
        if (ty < 0 || ty >= g.x.h) continue;
        for (int s = 0; s < g.kw; ++s) {
            int tx = ox * g.mul + g.off_x + s * g.step;
            if (g.div > 1) { if (tx % g.div) continue; tx /= g.div; }
            if (tx < 0 || tx >= g.x.w) continue;
            acc = fmaf(g.x.p[((img * g.x.h + ty) * g.x.w + tx) * g.x.cs], g.wmat[r * g.kw + s], acc);
        }
    }
    acc = fmaxf(g.alpha * acc, acc);
    float* yp = g.y.p + pix * g.y.cs;
    if (g.res) acc += g.res[pix * g.res_cs];
    if (g.accumulate) acc += *yp;
    if (g.mask) acc *= (g.mask[pix * g.mask_cs] > 0.f) ? 1.f : g.mask_alpha;
    *yp = acc;
}
bool conv_one_channel_supported(const ConvGemm& g) { return g.x.c == 1 && g.y.c == 1 && g.x.n == g.y.n && g.alpha <= 1.f && g.alpha >= 0.f; }
int conv_one_channel(const ConvGemm& g, cudaStream_t st) {
    MS_REQUIRE(conv_one_channel_supported(g), "conv_one_channel: not a 1 -> 1 channel gather");
    const size_t npix = g.y.pixels();
    launch_k(conv_one_channel_kernel, dim3((unsigned)cdivz(npix, 256)), dim3(256), 0, st, g, npix);
    return check_launch("conv_one_channel");
}

int conv_head_kind(const ConvGemm& g) {
    if (g.mul != 1 || g.div != 1) return 0;
    if (g.x.h != g.y.h || g.x.w != g.y.w) return 0;
    const size_t wbytes = (size_t)g.kh * g.kw * std::max(g.x.c, g.y.c) * sizeof(float);
    if (wbytes > 96 * 1024) return 0;
    if (g.y.c == 1 && g.x.c >= 4 && (g.x.c & 3) == 0 && (g.x.cs & 3) == 0 && aligned16(g.x.p) && aligned16(g.wmat)) return 1;
    if (g.x.c == 1 && g.y.c >= 4 && (g.y.c & 3) == 0 && (g.y.cs & 3) == 0 && aligned16(g.y.p) && aligned16(g.wmat) && !g.res &&
        !g.bias && g.alpha == 1.f && (!g.mask || ((g.mask_cs & 3) == 0 && aligned16(g.mask))))
        return 2;
    return 0;
}

// g.wmat: [tap][cin] for the forward head, [tap][cin] (= [tap][1][cin]) for its dgrad -- the same memory either way
int conv_head(const ConvGemm& g, cudaStream_t st) {
    const int kind = conv_head_kind(g);
    MS_REQUIRE(kind != 0, "conv_head: not a single-channel head");
    static bool attr_done = false;
    if (!attr_done) {
        MS_CHECK_CUDA(cudaFuncSetAttribute(conv_head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        MS_CHECK_CUDA(cudaFuncSetAttribute(conv_head_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_done = true;
    }
    const size_t npix = g.y.pixels();
    const size_t smem = (size_t)g.kh * g.kw * std::max(g.x.c, g.y.c) * sizeof(float);
    if (kind == 1) {
        launch_k(conv_head_fwd_kernel, dim3((unsigned)cdivz(npix, 32)), dim3(256), smem, st, g, npix);
        return check_launch("conv_head_fwd");
    }
    const size_t total = npix * (g.y.c >> 2);
    const unsigned grid = (unsigned)std::min<size_t>(cdivz(total, 256), 148 * 8);
    launch_k(conv_head_dgrad_kernel, dim3(grid), dim3(256), smem, st, g, npix);
    return check_launch("conv_head_dgrad");
}

}  // namespace ms
