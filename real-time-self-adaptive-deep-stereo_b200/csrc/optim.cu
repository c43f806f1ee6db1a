// Momentum-SGD update over a contiguous parameter / gradient / momentum arena.
// Replaces tf.train.MomentumOptimizer(lr, 0.9) ApplyMomentum ops (reference Stereo_Online_Adaptation.py:85,
// :118, :128): accum = mu*accum + g ; var -= lr*accum (non-Nesterov).  `gscale` folds the 1/N of the
// data-parallel gradient mean into the same pass (the all-reduce delivers the sum).
#include "common.cuh"

namespace ms {

__global__ void momentum_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, size_t n,
                                float lr, float mu, float gscale) {
    pdl_prologue();
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 gv = *reinterpret_cast<const float4*>(g + i);
        float4 mv = *reinterpret_cast<float4*>(m + i);
        float4 wv = *reinterpret_cast<float4*>(w + i);
        mv.x = mu * mv.x + gv.x * gscale; mv.y = mu * mv.y + gv.y * gscale;
        mv.z = mu * mv.z + gv.z * gscale; mv.w = mu * mv.w + gv.w * gscale;
        wv.x -= lr * mv.x; wv.y -= lr * mv.y; wv.z -= lr * mv.z; wv.w -= lr * mv.w;
        *reinterpret_cast<float4*>(m + i) = mv;
        *reinterpret_cast<float4*>(w + i) = wv;
    } else {
        for (; i < n; ++i) {
            float mv = mu * m[i] + g[i] * gscale;
            m[i] = mv;
            w[i] -= lr * mv;
        }
    }
}

int momentum_update(float* w, const float* g, float* m, size_t n, float lr, float mu, float gscale, cudaStream_t st) {
    if (n == 0) return 0;
    MS_REQUIRE(((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m)) & 15) == 0,
               "momentum_update: arenas must be 16B aligned");
    size_t nthr = cdivz(n, 4);
    launch_k(momentum_kernel, dim3((unsigned)cdivz(nthr, 256)), dim3(256), 0, st, w, g, m, n, lr, mu, gscale);
    return check_launch("momentum_update");
}

}  // namespace ms
