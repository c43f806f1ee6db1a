// tcgen05 implicit-GEMM convolution (stride 1, any dilation): the tensor-core path for the conv stacks.
//
// Replaces the cuDNN calls behind tf.nn.conv2d / tf.nn.atrous_conv2d (reference Nets/sharedLayers.py:58,72) and
// their input-gradients for every stride-1 layer of MADNet/DispNet (estimators Nets/MadNet.py:73-120, context net
// :122-171, the stride-1 pyramid convs :173-249).
//
// GEMM view: M = 128 output pixels (a TH x TW patch of one image), N = output channels (<=256, padded to x16),
// K = taps x input channels in blocks of 32.  Per K-block one 4-D TMA box {32ch, TW, TH, 1} of the input at the
// tap-shifted coordinate lands in shared memory as a K-major SWIZZLE_128B tile (out-of-image taps and channels
// beyond C are zero-filled by TMA = the SAME padding), the weights arrive as a {32, N} K-major tile.
// fp32 fidelity on tf32 tensor cores uses the 3xTF32 split  a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi
// (hi = round-to-tf32, lo = exact remainder): the weight halves are prepared in global memory, the activation
// halves by 4 "splitter" warps in shared memory.  Accumulators live in TMEM (128 lanes x N fp32 columns);
// the same 4 warps run the epilogue (bias, leaky-ReLU, residual, dgrad mask/accumulate) from tcgen05.ld.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = splitter +
// epilogue.  mbarrier pipeline: full[s] (TMA -> splitter), ready[s] (splitter -> MMA), empty[s] (tcgen05.commit ->
// TMA), accum (tcgen05.commit -> epilogue).
#include <cuda.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ms {

// ------------------------------------------------------------------------------------------------
struct ConvTCParams {
    int kh, kw, off_y, off_x, step;   // gathered coordinate = out + off + tap*step
    int kblocks;                      // ceil(K channels / 32)
    int H, W, NB, TH, TW, tiles_x, tiles_y;
    int N, BN, tmem_cols, stages;
    int n_main, acc_stride;           // accumulators: [0]=cross terms, [1..n_main]=hi*hi partial sums (column stride)
    float* y; int ycs;
    const float* bias; float alpha;
    const float* mask; int mask_cs; float mask_alpha;
    const float* res; int res_cs;
    int accumulate;
    int ksplit;                       // split-K over channel blocks (small maps): partial sums -> `part`, then tc_splitk_reduce
    float* part;                      // [ksplit][pixels][BN]
    int dbg;                          // timing experiments only (MS_TC_DEBUG): 1 no split, 2 no proxy fence, 4 no MMA
};

constexpr int TC_THREADS = 320;           // warp 0 TMA, warp 1 MMA, warps 2-9 splitter + epilogue
constexpr int SPLIT_THREADS = 256;
constexpr int A_TILE_BYTES = 128 * 128;   // 128 pixels x 32 fp32

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const ConvTCParams p) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[4], ready_bar[4], empty_bar[4], accum_bar;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // 1024-byte aligned carve-up
    const uint32_t base = (s_addr(smem_dyn) + 1023u) & ~1023u;
    unsigned char* gbase = smem_dyn + (base - s_addr(smem_dyn));
    const uint32_t b_bytes = (uint32_t)p.BN * 128u;
    const uint32_t stage_bytes = 2u * A_TILE_BYTES + 2u * b_bytes;

    int bid = blockIdx.x;
    const int tx = bid % p.tiles_x; bid /= p.tiles_x;
    const int ty = bid % p.tiles_y;
    const int img = bid / p.tiles_y;
    const int x0 = tx * p.TW, y0 = ty * p.TH;
    const int taps = p.kh * p.kw;
    const int total = taps * p.kblocks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mb_init(&full_bar[s], 1); mb_init(&ready_bar[s], SPLIT_THREADS / 32); mb_init(&empty_bar[s], 1); }
        mb_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(&tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int tap = 0, kb = 0, s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < total; ++it) {
                mb_wait(&empty_bar[s], ph ^ 1u);
                unsigned char* st = gbase + (size_t)s * stage_bytes;
                mb_expect_tx(&full_bar[s], (uint32_t)A_TILE_BYTES + b_bytes);
                const int r = tap / p.kw, q = tap - r * p.kw;
                tma_load_4d(st, &mapA, &full_bar[s], kb * 32, x0 + p.off_x + q * p.step, y0 + p.off_y + r * p.step, img);
                tma_load_3d(st + 2 * A_TILE_BYTES, &mapB, &full_bar[s], kb * 32, 0, tap);
                if (++kb == p.kblocks) { kb = 0; ++tap; }
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // instruction descriptor: D=f32 (bit4), A=B=tf32 (2<<7, 2<<10), K-major both, N>>3 @17, M>>4 @24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < total; ++it) {
                mb_wait(&full_bar[s], ph);
                mb_wait(&ready_bar[s], ph);
                tc_fence_after();
                const uint32_t sa = base + (uint32_t)s * stage_bytes;
                const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + A_TILE_BYTES);
                const uint64_t b_hi = umma_desc_sw128(sa + 2 * A_TILE_BYTES), b_lo = umma_desc_sw128(sa + 2 * A_TILE_BYTES + b_bytes);
                if (!(p.dbg & 4))
#pragma unroll
                for (int j = 0; j < 4; ++j) {          // 4 x (K = 8 tf32 = 32 bytes) inside the 128-byte swizzle row
                    const uint64_t o = (uint64_t)(j * 2);
                    const int g = it * 4 + j;
                    // The tensor core adds into the fp32 accumulator with truncation, a bias that grows with the
                    // number of accumulation steps.  Cross terms (2^-11 of the main term) get their own accumulator
                    // and the hi*hi products rotate over n_main accumulators; the epilogue sums them in fp32 (RN).
                    tc_mma_tf32(tmem, a_lo + o, b_hi + o, idesc, g > 0 ? 1u : 0u);
                    tc_mma_tf32(tmem, a_hi + o, b_lo + o, idesc, 1u);
                    const uint32_t dmain = tmem + (uint32_t)((1 + g % p.n_main) * p.acc_stride);
                    tc_mma_tf32(dmain, a_hi + o, b_hi + o, idesc, g >= p.n_main ? 1u : 0u);
                }
                tc_commit(&empty_bar[s]);              // frees the stage once these MMAs have read it
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
            tc_commit(&accum_bar);
        }
    } else {
        // ================= splitter (warps 2..9): tf32 hi/lo halves of the activation AND weight tiles =============
        const int st_tid = threadIdx.x - 64;           // 0..255
        const int b_f4 = p.BN * 8;                     // float4 per weight tile
        {
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < total; ++it) {
                mb_wait(&full_bar[s], ph);
                unsigned char* stg = gbase + (size_t)s * stage_bytes;
                float4* __restrict__ ahi = reinterpret_cast<float4*>(stg);
                float4* __restrict__ alo = reinterpret_cast<float4*>(stg + A_TILE_BYTES);
                float4* __restrict__ bhi = reinterpret_cast<float4*>(stg + 2 * A_TILE_BYTES);
                float4* __restrict__ blo = reinterpret_cast<float4*>(stg + 2 * A_TILE_BYTES + b_bytes);
                if (!(p.dbg & 1)) {
                float4 v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = ahi[st_tid + e * SPLIT_THREADS];       // 1024 float4 per A tile
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int idx = st_tid + e * SPLIT_THREADS;
                    float4 h, l;
                    h.x = tf32_hi(v[e].x); h.y = tf32_hi(v[e].y); h.z = tf32_hi(v[e].z); h.w = tf32_hi(v[e].w);
                    l.x = v[e].x - h.x; l.y = v[e].y - h.y; l.z = v[e].z - h.z; l.w = v[e].w - h.w;
                    ahi[idx] = h; alo[idx] = l;
                }
                for (int i0 = st_tid; i0 < b_f4; i0 += 2 * SPLIT_THREADS) {                // BN*8 float4 per B tile
                    const int i1 = i0 + SPLIT_THREADS;
                    const bool two = i1 < b_f4;
                    float4 w0 = bhi[i0], w1 = two ? bhi[i1] : make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 h, l;
                    h.x = tf32_hi(w0.x); h.y = tf32_hi(w0.y); h.z = tf32_hi(w0.z); h.w = tf32_hi(w0.w);
                    l.x = w0.x - h.x; l.y = w0.y - h.y; l.z = w0.z - h.z; l.w = w0.w - h.w;
                    bhi[i0] = h; blo[i0] = l;
                    if (two) {
                        h.x = tf32_hi(w1.x); h.y = tf32_hi(w1.y); h.z = tf32_hi(w1.z); h.w = tf32_hi(w1.w);
                        l.x = w1.x - h.x; l.y = w1.y - h.y; l.z = w1.z - h.z; l.w = w1.w - h.w;
                        bhi[i1] = h; blo[i1] = l;
                    }
                }
                }
                if (!(p.dbg & 2)) fence_async_smem();   // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mb_arrive(&ready_bar[s]);
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
        // ================= epilogue =================
        mb_wait(&accum_bar, 0);
        tc_fence_after();
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;                   // accumulator row = pixel inside the patch
        const int py = y0 + m / p.TW, px = x0 + m % p.TW;
        const bool valid = (py < p.H) && (px < p.W);
        const size_t pix = ((size_t)img * p.H + py) * p.W + px;
        float* yrow = p.y + pix * p.ycs;
        const bool vec = ((p.ycs & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
        const int chunks = p.BN / 16, half = (chunks + 1) / 2;
        const int cbeg = (warp < 6 ? 0 : half) * 16, cend = (warp < 6 ? half : chunks) * 16;   // warps 2-5 | 6-9
        const bool plain = vec && !p.res && !p.accumulate && !p.mask;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        for (int c0 = cbeg; c0 < cend; c0 += 16) {
            uint32_t r0[16], r1[16], r2[16], r3[16];
            float bb[16];
            const bool fast = plain && c0 + 16 <= p.N;
            if (fast) {                                  // bias loads overlap the TMEM reads
#pragma unroll
                for (int j = 0; j < 16; ++j) bb[j] = p.bias ? __ldg(p.bias + c0 + j) : 0.f;
            }
            tc_ld16_nowait(tmem + lane_base + (uint32_t)c0, r0);
            tc_ld16_nowait(tmem + lane_base + (uint32_t)(p.acc_stride + c0), r1);
            if (p.n_main > 1) tc_ld16_nowait(tmem + lane_base + (uint32_t)(2 * p.acc_stride + c0), r2);
            if (p.n_main > 2) tc_ld16_nowait(tmem + lane_base + (uint32_t)(3 * p.acc_stride + c0), r3);
            tc_wait_ld();
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float t = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
                if (p.n_main > 1) t += __uint_as_float(r2[j]);
                if (p.n_main > 2) t += __uint_as_float(r3[j]);
                v[j] = t;
            }
            if (valid && fast) {
#pragma unroll
                for (int j = 0; j < 16; ++j) { const float t = v[j] + bb[j]; v[j] = fmaxf(p.alpha * t, t); }
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(yrow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                continue;
            }
            if (!valid) continue;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = c0 + j;
                if (n < p.N) {
                    float t = v[j];
                    if (p.bias) t += p.bias[n];
                    t = fmaxf(p.alpha * t, t);
                    if (p.res) t += p.res[pix * p.res_cs + n];
                    if (p.accumulate) t += yrow[n];
                    if (p.mask) t *= (p.mask[pix * p.mask_cs + n] > 0.f) ? 1.f : p.mask_alpha;
                    v[j] = t;
                }
            }
            if (vec && c0 + 16 <= p.N) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(yrow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < p.N) yrow[c0 + j] = v[j];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// v2 kernel for small tap extents (undilated 3x3, 1x1): the activation patch (tile + halo) is loaded ONCE per
// 32-channel block and the 9 tap-shifted A tiles are gathered from it in shared memory by the splitter warps while
// they compute the tf32 hi/lo halves (6x less activation traffic than one TMA box per tap); weights stream through a
// deeper raw ring; two operand stages decouple the splitter from the MMA issuer.
//   smem: op[2] x {A_hi, A_lo, B_hi, B_lo} | braw[NB] | patch[2]
//   barriers: pfull/pempty[2] (patch), bfull/bempty[NB] (weights), ready/free[2] (operands), accum
// ------------------------------------------------------------------------------------------------
__device__ unsigned long long g_tc_prof[32];
#define TCP_T0() const long long _t0 = clock64()
#define TCP_ADD(slot, cond) do { if (cond) g_tc_prof[slot] += (unsigned long long)(clock64() - _t0); } while (0)

struct ConvTCHaloParams {
    ConvTCParams c;
    int PW, PH;            // patch width / height in pixels
    int minx, miny;        // patch origin relative to the tile origin
    int nb_slots;          // weight ring depth
    uint32_t patch_bytes;
    int ns;                // operand stages (TS kernel): {A hi/lo in TMEM, B hi/lo in smem} in flight between splitter and MMA
};

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapB,
                    const ConvTCHaloParams hp) {
    const ConvTCParams& p = hp.c;
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t pfull[2], pempty[2], bfull[8], bempty[8], ready_bar[2], free_bar[2], accum_bar;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (s_addr(smem_dyn) + 1023u) & ~1023u;
    unsigned char* gbase = smem_dyn + (base - s_addr(smem_dyn));
    const uint32_t b_bytes = (uint32_t)p.BN * 128u;
    const uint32_t op_bytes = 2u * A_TILE_BYTES + 2u * b_bytes;
    const uint32_t braw_off = 2u * op_bytes;
    const uint32_t patch_stride = (hp.patch_bytes + 1023u) & ~1023u;
    const uint32_t patch_off = braw_off + (uint32_t)hp.nb_slots * b_bytes;
    const int NB = hp.nb_slots;

    int bid = blockIdx.x;
    const int tx = bid % p.tiles_x; bid /= p.tiles_x;
    const int ty = bid % p.tiles_y;
    const int img = bid / p.tiles_y;
    const int x0 = tx * p.TW, y0 = ty * p.TH;
    const int taps = p.kh * p.kw;
    const int total = taps * p.kblocks;
    const bool prof = (p.dbg & 8) && blockIdx.x == 0;
    const long long t_start = clock64();
    long long t_epi = 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mb_init(&pfull[i], 1); mb_init(&pempty[i], SPLIT_THREADS / 32);
            mb_init(&ready_bar[i], SPLIT_THREADS / 32); mb_init(&free_bar[i], 1);
        }
        for (int i = 0; i < NB; ++i) { mb_init(&bfull[i], 1); mb_init(&bempty[i], SPLIT_THREADS / 32); }
        mb_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(&tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        // ================= TMA producer: patches (per channel block) + weight tiles (per channel block x tap) ======
        if (lane == 0) {
            auto load_patch = [&](int kb) {
                const int pb = kb & 1;
                mb_wait(&pempty[pb], (((uint32_t)kb >> 1) & 1u) ^ 1u);
                mb_expect_tx(&pfull[pb], hp.patch_bytes);
                tma_load_4d(gbase + patch_off + (size_t)pb * patch_stride, &mapP, &pfull[pb], kb * 32, x0 + hp.minx, y0 + hp.miny, img);
            };
            load_patch(0);
            int slot = 0;
            uint32_t bph = 0;
            const int ahead = min(NB, taps) - 1;
            for (int kb = 0; kb < p.kblocks; ++kb) {
                for (int tap = 0; tap < taps; ++tap) {
                    { TCP_T0(); mb_wait(&bempty[slot], bph ^ 1u); TCP_ADD(0, prof); }
                    mb_expect_tx(&bfull[slot], b_bytes);
                    tma_load_3d(gbase + braw_off + (size_t)slot * b_bytes, &mapB, &bfull[slot], kb * 32, 0, tap);
                    if (++slot == NB) { slot = 0; bph ^= 1u; }
                    if (tap == ahead && kb + 1 < p.kblocks) load_patch(kb + 1);   // the splitter has left patch kb-1 by now
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
            for (int it = 0; it < total; ++it) {
                const int s = it & 1;
                { TCP_T0(); mb_wait(&ready_bar[s], ((uint32_t)it >> 1) & 1u); TCP_ADD(1, prof); }
                const long long t_issue = clock64();
                tc_fence_after();
                const uint32_t sa = base + (uint32_t)s * op_bytes;
                const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + A_TILE_BYTES);
                const uint64_t b_hi = umma_desc_sw128(sa + 2 * A_TILE_BYTES), b_lo = umma_desc_sw128(sa + 2 * A_TILE_BYTES + b_bytes);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint64_t o = (uint64_t)(j * 2);
                    const int g = it * 4 + j;
                    tc_mma_tf32(tmem, a_lo + o, b_hi + o, idesc, g > 0 ? 1u : 0u);
                    tc_mma_tf32(tmem, a_hi + o, b_lo + o, idesc, 1u);
                    const uint32_t dmain = tmem + (uint32_t)((1 + g % p.n_main) * p.acc_stride);
                    tc_mma_tf32(dmain, a_hi + o, b_hi + o, idesc, g >= p.n_main ? 1u : 0u);
                }
                tc_commit(&free_bar[s]);
                if (prof) g_tc_prof[2] += (unsigned long long)(clock64() - t_issue);
            }
            tc_commit(&accum_bar);
        }
    } else {
        // ================= splitter / gatherer (warps 2..9) =================
        const int st_tid = threadIdx.x - 64;           // 0..255
        const int b_f4 = p.BN * 8;
        {
            int slot = 0;
            uint32_t bph = 0;
            int it = 0;
            for (int kb = 0; kb < p.kblocks; ++kb) {
                const int pb = kb & 1;
                const bool sp = prof && threadIdx.x == 64;
                { TCP_T0(); mb_wait(&pfull[pb], ((uint32_t)kb >> 1) & 1u); TCP_ADD(3, sp); }
                const unsigned char* patch = gbase + patch_off + (size_t)pb * patch_stride;
                for (int tap = 0; tap < taps; ++tap, ++it) {
                    const int s = it & 1;
                    { TCP_T0(); mb_wait(&free_bar[s], (((uint32_t)it >> 1) & 1u) ^ 1u); TCP_ADD(4, sp); }   // operand stage released by the MMAs of it-2
                    { TCP_T0(); mb_wait(&bfull[slot], bph); TCP_ADD(5, sp); }
                    const long long t_work = clock64();
                    unsigned char* stg = gbase + (size_t)s * op_bytes;
                    float4* __restrict__ ahi = reinterpret_cast<float4*>(stg);
                    float4* __restrict__ alo = reinterpret_cast<float4*>(stg + A_TILE_BYTES);
                    float4* __restrict__ bhi = reinterpret_cast<float4*>(stg + 2 * A_TILE_BYTES);
                    float4* __restrict__ blo = reinterpret_cast<float4*>(stg + 2 * A_TILE_BYTES + b_bytes);
                    const float4* __restrict__ braw = reinterpret_cast<const float4*>(gbase + braw_off + (size_t)slot * b_bytes);
                    // ---- A: gather the tap-shifted tile out of the patch, split, write swizzled (SWIZZLE_128B)
                    const int tr = tap / p.kw, ts = tap - tr * p.kw;
                    const int dy = p.off_y + tr * p.step - hp.miny, dx = p.off_x + ts * p.step - hp.minx;
                    float4 v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int idx = st_tid + e * SPLIT_THREADS;        // 0..1023 = row*8 + chunk
                        const int row = idx >> 3, ch = idx & 7;
                        const int py = row / p.TW + dy, px = row % p.TW + dx;
                        v[e] = *reinterpret_cast<const float4*>(patch + ((size_t)(py * hp.PW + px) * 128 + ch * 16));
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int idx = st_tid + e * SPLIT_THREADS;
                        const int row = idx >> 3, ch = idx & 7;
                        const int d = row * 8 + (ch ^ (row & 7));
                        float4 h, l;
                        h.x = tf32_hi(v[e].x); h.y = tf32_hi(v[e].y); h.z = tf32_hi(v[e].z); h.w = tf32_hi(v[e].w);
                        l.x = v[e].x - h.x; l.y = v[e].y - h.y; l.z = v[e].z - h.z; l.w = v[e].w - h.w;
                        ahi[d] = h; alo[d] = l;
                    }
                    // ---- B: elementwise split (the TMA already wrote the swizzled layout)
                    for (int i0 = st_tid; i0 < b_f4; i0 += 2 * SPLIT_THREADS) {
                        const int i1 = i0 + SPLIT_THREADS;
                        const bool two = i1 < b_f4;
                        float4 w0 = braw[i0], w1 = two ? braw[i1] : make_float4(0.f, 0.f, 0.f, 0.f);
                        float4 h, l;
                        h.x = tf32_hi(w0.x); h.y = tf32_hi(w0.y); h.z = tf32_hi(w0.z); h.w = tf32_hi(w0.w);
                        l.x = w0.x - h.x; l.y = w0.y - h.y; l.z = w0.z - h.z; l.w = w0.w - h.w;
                        bhi[i0] = h; blo[i0] = l;
                        if (two) {
                            h.x = tf32_hi(w1.x); h.y = tf32_hi(w1.y); h.z = tf32_hi(w1.z); h.w = tf32_hi(w1.w);
                            l.x = w1.x - h.x; l.y = w1.y - h.y; l.z = w1.z - h.z; l.w = w1.w - h.w;
                            bhi[i1] = h; blo[i1] = l;
                        }
                    }
                    if (sp) g_tc_prof[6] += (unsigned long long)(clock64() - t_work);
                    { TCP_T0(); fence_async_smem(); TCP_ADD(7, sp); }
                    __syncwarp();
                    if (lane == 0) { mb_arrive(&ready_bar[s]); mb_arrive(&bempty[slot]); }
                    if (++slot == NB) { slot = 0; bph ^= 1u; }
                }
                __syncwarp();
                if (lane == 0) mb_arrive(&pempty[pb]);
            }
        }
        // ================= epilogue =================
        t_epi = clock64();
        mb_wait(&accum_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int py = y0 + m / p.TW, px = x0 + m % p.TW;
        const bool valid = (py < p.H) && (px < p.W);
        const size_t pix = ((size_t)img * p.H + py) * p.W + px;
        float* yrow = p.y + pix * p.ycs;
        const bool vec = ((p.ycs & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
        const int chunks = p.BN / 16, half = (chunks + 1) / 2;
        const int cbeg = (warp < 6 ? 0 : half) * 16, cend = (warp < 6 ? half : chunks) * 16;
        for (int c0 = cbeg; c0 < cend; c0 += 16) {
            float v[16];
            tc_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            for (int a = 1; a <= p.n_main; ++a) {
                float u[16];
                tc_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * p.acc_stride + c0), u);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += u[j];
            }
            if (!valid) continue;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = c0 + j;
                if (n < p.N) {
                    float t = v[j];
                    if (p.bias) t += p.bias[n];
                    t = fmaxf(p.alpha * t, t);
                    if (p.res) t += p.res[pix * p.res_cs + n];
                    if (p.accumulate) t += yrow[n];
                    if (p.mask) t *= (p.mask[pix * p.mask_cs + n] > 0.f) ? 1.f : p.mask_alpha;
                    v[j] = t;
                }
            }
            if (vec && c0 + 16 <= p.N) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(yrow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < p.N) yrow[c0 + j] = v[j];
            }
        }
    }
    if (prof && threadIdx.x == 64) { g_tc_prof[8] += (unsigned long long)(clock64() - t_epi); g_tc_prof[9] += (unsigned long long)(t_epi - t_start); g_tc_prof[10] += 1; }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// v3 kernel ("TS"): like the halo kernel, but the A operand lives in TENSOR MEMORY.  Profiling the v1/v2 kernels
// (per-role clock64 counters) showed the 3xTF32 scheme to be shared-memory-bandwidth bound: per K-block
// the splitter moved 96 KB and the three SS-mode MMAs re-read 96 KB of operands.  Here each splitter thread owns one
// accumulator row (= one output pixel): it reads its pixel's 32 channels from the (swizzled) halo patch, computes the
// tf32 hi/lo halves in registers and writes them straight to TMEM with tcgen05.st; the MMAs take A from TMEM
// (tcgen05.mma ... [a_tmem]) and only B from shared memory.  smem traffic per K-block: 16 KB (patch reads) + 48 KB
// (weight split) + 48 KB (B operand reads) instead of 192 KB, and the freed 64 KB deepen the weight ring.
//   TMEM columns: [0, (n_main+1)*acc_stride) accumulators | then 2 stages x {32 hi, 32 lo} columns of A
// ------------------------------------------------------------------------------------------------
#define TSP_T0() const long long _t0 = PROF ? clock64() : 0
#define TSP_ADD(slot, cond) do { if (PROF && (cond)) g_tc_prof[slot] += (unsigned long long)(clock64() - _t0); } while (0)
template <bool PROF>      // PROF: per-role clock64 counters + MS_TC_DEBUG kill switches (timing experiments only)
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_ts_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapB,
                  const ConvTCHaloParams hp) {
    const ConvTCParams& p = hp.c;
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t pfull[2], pempty[2], bfull[8], bempty[8], ready_bar[4], free_bar[4], accum_bar;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (s_addr(smem_dyn) + 1023u) & ~1023u;
    unsigned char* gbase = smem_dyn + (base - s_addr(smem_dyn));
    const uint32_t b_bytes = (uint32_t)p.BN * 128u;
    const uint32_t op_bytes = 2u * b_bytes;                       // B_hi, B_lo
    const int NS = hp.ns;
    const uint32_t braw_off = (uint32_t)NS * op_bytes;
    const uint32_t patch_stride = (hp.patch_bytes + 1023u) & ~1023u;
    const uint32_t patch_off = braw_off + (uint32_t)hp.nb_slots * b_bytes;
    const int NB = hp.nb_slots;
    const uint32_t a_col0 = (uint32_t)((p.n_main + 1) * p.acc_stride);

    int bid = blockIdx.x;
    const int tx = bid % p.tiles_x; bid /= p.tiles_x;
    const int ty = bid % p.tiles_y;
    const int img = bid / p.tiles_y;
    const int x0 = tx * p.TW, y0 = ty * p.TH;
    const int taps = p.kh * p.kw;
    // split-K unit = one (channel block, tap) iteration; this CTA owns global iterations [g0, g1), kb-major / tap-minor
    const int total_all = taps * p.kblocks;
    const int g0 = (int)(((long)blockIdx.y * total_all) / p.ksplit), g1 = (int)(((long)(blockIdx.y + 1) * total_all) / p.ksplit);
    const int total = g1 - g0;
    const bool prof = PROF && (p.dbg & 8) && blockIdx.x == 0;
    const long long t_start = PROF ? clock64() : 0;
    long long t_epi = 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mb_init(&pfull[i], 1); mb_init(&pempty[i], SPLIT_THREADS / 32); }
        for (int i = 0; i < NS; ++i) { mb_init(&ready_bar[i], SPLIT_THREADS / 32); mb_init(&free_bar[i], 1); }
        for (int i = 0; i < NB; ++i) { mb_init(&bfull[i], 1); mb_init(&bempty[i], SPLIT_THREADS / 32); }
        mb_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(&tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int kl = 0;                               // local patch counter (slot = kl & 1)
            auto load_patch = [&](int kb) {
                const int pb = kl & 1;
                mb_wait(&pempty[pb], (((uint32_t)kl >> 1) & 1u) ^ 1u);
                mb_expect_tx(&pfull[pb], hp.patch_bytes);
                tma_load_4d(gbase + patch_off + (size_t)pb * patch_stride, &mapP, &pfull[pb], kb * 32, x0 + hp.minx, y0 + hp.miny, img);
                ++kl;
            };
            int slot = 0;
            uint32_t bph = 0;
            const int ahead = min(NB, taps) - 1;
            int next_patch_g = g0;                    // first iteration that needs a patch not requested yet
            int kb = g0 / taps, tap = g0 - kb * taps;
            for (int g = g0; g < g1; ++g) {
                while (next_patch_g < g1 && next_patch_g <= g + ahead) {
                    const int kbn = next_patch_g / taps;
                    load_patch(kbn);
                    next_patch_g = (kbn + 1) * taps;
                }
                { TSP_T0(); mb_wait(&bempty[slot], bph ^ 1u); TSP_ADD(0, prof); }
                mb_expect_tx(&bfull[slot], b_bytes);
                tma_load_3d(gbase + braw_off + (size_t)slot * b_bytes, &mapB, &bfull[slot], kb * 32, 0, tap);
                if (++slot == NB) { slot = 0; bph ^= 1u; }
                if (++tap == taps) { tap = 0; ++kb; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
            int s = 0, rot = 0, gcount = 0;
            uint32_t ph = 0;
            for (int it = 0; it < total; ++it) {
                { TSP_T0(); mb_wait(&ready_bar[s], ph); TSP_ADD(1, prof); }
                const long long t_issue = PROF ? clock64() : 0;
                tc_fence_after();
                const uint32_t sb = base + (uint32_t)s * op_bytes;
                const uint64_t b_hi = umma_desc_sw128(sb), b_lo = umma_desc_sw128(sb + b_bytes);
                const uint32_t a_hi = tmem + a_col0 + (uint32_t)s * 64u, a_lo = a_hi + 32u;
                if (!(PROF && (p.dbg & 64)))        // timing experiments only (MS_TC_DEBUG): 64 = no MMAs, 16 = no A split, 32 = no B split
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint64_t o = (uint64_t)(j * 2);
                    const uint32_t ao = (uint32_t)(j * 8);
                    tc_mma_tf32_ts(tmem, a_lo + ao, b_hi + o, idesc, gcount > 0 ? 1u : 0u);
                    tc_mma_tf32_ts(tmem, a_hi + ao, b_lo + o, idesc, 1u);
                    const uint32_t dmain = tmem + (uint32_t)((1 + rot) * p.acc_stride);
                    tc_mma_tf32_ts(dmain, a_hi + ao, b_hi + o, idesc, gcount >= p.n_main ? 1u : 0u);
                    ++gcount;
                    if (++rot == p.n_main) rot = 0;
                }
                tc_commit(&free_bar[s]);
                if (PROF && prof) g_tc_prof[2] += (unsigned long long)(clock64() - t_issue);
                if (++s == NS) { s = 0; ph ^= 1u; }
            }
            tc_commit(&accum_bar);
        }
    } else {
        // ================= splitter (warps 2..9): thread <-> accumulator row =================
        const int st_tid = threadIdx.x - 64;
        const int b_f4 = p.BN * 8;
        const int q = warp & 3;                        // TMEM lane quarter of this warp
        const int hsel = (warp - 2) >> 2;              // which 16-channel half of the 32-channel block
        const int m = q * 32 + lane;
        const int my = m / p.TW, mx = m % p.TW;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        {
            int slot = 0;
            uint32_t bph = 0;
            int it = 0;
            int kl = -1, cur_kb = -1, pb = 0;
            int s = 0;
            uint32_t sph = 0;
            const unsigned char* patch = nullptr;
            const bool sp = prof && threadIdx.x == 64;
            int kb = g0 / taps, tap = g0 - kb * taps;
            int tr = tap / p.kw, ts = tap - tr * p.kw;
            const int pp_base = (my + p.off_y - hp.miny) * hp.PW + (mx + p.off_x - hp.minx);
            for (int g = g0; g < g1; ++g, ++it) {
                {
                    if (kb != cur_kb) {                 // entering a new channel block: release the old patch, wait for the new one
                        if (cur_kb >= 0) { __syncwarp(); if (lane == 0) mb_arrive(&pempty[pb]); }
                        cur_kb = kb; ++kl; pb = kl & 1;
                        { TSP_T0(); mb_wait(&pfull[pb], ((uint32_t)kl >> 1) & 1u); TSP_ADD(3, sp); }
                        patch = gbase + patch_off + (size_t)pb * patch_stride;
                    }
                    { TSP_T0(); mb_wait(&free_bar[s], sph ^ 1u); TSP_ADD(4, sp); }
                    { TSP_T0(); mb_wait(&bfull[slot], bph); TSP_ADD(5, sp); }
                    const long long t_work = PROF ? clock64() : 0;
                    // ---- A: this thread's pixel, 16 channels -> tf32 hi / lo -> TMEM
                    const int pp = pp_base + (tr * hp.PW + ts) * p.step;
                    const unsigned char* prow = patch + (size_t)pp * 128;
                    float hi[16], lo[16];
                    if (!(PROF && (p.dbg & 16))) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int ch = hsel * 4 + e;
                        const float4 v = *reinterpret_cast<const float4*>(prow + ((ch ^ (pp & 7)) * 16));
                        hi[4 * e] = tf32_hi(v.x); hi[4 * e + 1] = tf32_hi(v.y); hi[4 * e + 2] = tf32_hi(v.z); hi[4 * e + 3] = tf32_hi(v.w);
                        lo[4 * e] = v.x - hi[4 * e]; lo[4 * e + 1] = v.y - hi[4 * e + 1];
                        lo[4 * e + 2] = v.z - hi[4 * e + 2]; lo[4 * e + 3] = v.w - hi[4 * e + 3];
                    }
                    const uint32_t acol = tmem + lane_base + a_col0 + (uint32_t)s * 64u + (uint32_t)hsel * 16u;
                    tc_st16(acol, hi);
                    tc_st16(acol + 32u, lo);
                    }
                    // ---- B: elementwise split raw -> hi / lo (layout already swizzled by TMA)
                    unsigned char* stg = gbase + (size_t)s * op_bytes;
                    float4* __restrict__ bhi = reinterpret_cast<float4*>(stg);
                    float4* __restrict__ blo = reinterpret_cast<float4*>(stg + b_bytes);
                    const float4* __restrict__ braw = reinterpret_cast<const float4*>(gbase + braw_off + (size_t)slot * b_bytes);
                    if (!(PROF && (p.dbg & 32)))
                    for (int i0 = st_tid; i0 < b_f4; i0 += 2 * SPLIT_THREADS) {
                        const int i1 = i0 + SPLIT_THREADS;
                        const bool two = i1 < b_f4;
                        float4 w0 = braw[i0], w1 = two ? braw[i1] : make_float4(0.f, 0.f, 0.f, 0.f);
                        float4 h, l;
                        h.x = tf32_hi(w0.x); h.y = tf32_hi(w0.y); h.z = tf32_hi(w0.z); h.w = tf32_hi(w0.w);
                        l.x = w0.x - h.x; l.y = w0.y - h.y; l.z = w0.z - h.z; l.w = w0.w - h.w;
                        bhi[i0] = h; blo[i0] = l;
                        if (two) {
                            h.x = tf32_hi(w1.x); h.y = tf32_hi(w1.y); h.z = tf32_hi(w1.z); h.w = tf32_hi(w1.w);
                            l.x = w1.x - h.x; l.y = w1.y - h.y; l.z = w1.z - h.z; l.w = w1.w - h.w;
                            bhi[i1] = h; blo[i1] = l;
                        }
                    }
                    if (PROF && sp) g_tc_prof[6] += (unsigned long long)(clock64() - t_work);
                    { TSP_T0(); tc_wait_st(); fence_async_smem(); tc_fence_before(); TSP_ADD(7, sp); }
                    __syncwarp();
                    if (lane == 0) { mb_arrive(&ready_bar[s]); mb_arrive(&bempty[slot]); }
                    if (++slot == NB) { slot = 0; bph ^= 1u; }
                    if (++s == NS) { s = 0; sph ^= 1u; }
                    if (++ts == p.kw) { ts = 0; ++tr; }
                    if (++tap == taps) { tap = 0; tr = 0; ts = 0; ++kb; }
                }
            }
        }
        // ================= epilogue =================
        if (PROF) t_epi = clock64();
        mb_wait(&accum_bar, 0);
        if (PROF && prof && threadIdx.x == 64) g_tc_prof[11] += (unsigned long long)(clock64() - t_epi);
        tc_fence_after();
        const int py = y0 + my, px = x0 + mx;
        const bool valid = (py < p.H) && (px < p.W);
        const size_t pix = ((size_t)img * p.H + py) * p.W + px;
        float* yrow = p.y + pix * p.ycs;
        const bool vec = ((p.ycs & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
        const int chunks = p.BN / 16, half = (chunks + 1) / 2;
        const int cbeg = (warp < 6 ? 0 : half) * 16, cend = (warp < 6 ? half : chunks) * 16;
        const bool plain = vec && !p.res && !p.accumulate && !p.mask && p.ksplit == 1;
        for (int c0 = cbeg; c0 < cend; c0 += 16) {
            uint32_t r0[16], r1[16], r2[16], r3[16];
            float bb[16];
            const bool fast = plain && c0 + 16 <= p.N;
            if (fast) {                                  // bias loads overlap the TMEM reads
#pragma unroll
                for (int j = 0; j < 16; ++j) bb[j] = p.bias ? __ldg(p.bias + c0 + j) : 0.f;
            }
            tc_ld16_nowait(tmem + lane_base + (uint32_t)c0, r0);
            tc_ld16_nowait(tmem + lane_base + (uint32_t)(p.acc_stride + c0), r1);
            if (p.n_main > 1) tc_ld16_nowait(tmem + lane_base + (uint32_t)(2 * p.acc_stride + c0), r2);
            if (p.n_main > 2) tc_ld16_nowait(tmem + lane_base + (uint32_t)(3 * p.acc_stride + c0), r3);
            tc_wait_ld();
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float t = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
                if (p.n_main > 1) t += __uint_as_float(r2[j]);
                if (p.n_main > 2) t += __uint_as_float(r3[j]);
                v[j] = t;
            }
            if (!valid) continue;
            if (fast) {
#pragma unroll
                for (int j = 0; j < 16; ++j) { const float t = v[j] + bb[j]; v[j] = fmaxf(p.alpha * t, t); }
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(yrow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                continue;
            }
            if (p.ksplit > 1) {                     // raw partial sums; bias / activation happen in tc_splitk_reduce
                float* prow = p.part + ((size_t)blockIdx.y * ((size_t)p.NB * p.H * p.W) + pix) * p.BN + c0;
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(prow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                continue;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = c0 + j;
                if (n < p.N) {
                    float t = v[j];
                    if (p.bias) t += p.bias[n];
                    t = fmaxf(p.alpha * t, t);
                    if (p.res) t += p.res[pix * p.res_cs + n];
                    if (p.accumulate) t += yrow[n];
                    if (p.mask) t *= (p.mask[pix * p.mask_cs + n] > 0.f) ? 1.f : p.mask_alpha;
                    v[j] = t;
                }
            }
            if (vec && c0 + 16 <= p.N) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(yrow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < p.N) yrow[c0 + j] = v[j];
            }
        }
    }
    if (PROF && prof && threadIdx.x == 64) { g_tc_prof[8] += (unsigned long long)(clock64() - t_epi); g_tc_prof[9] += (unsigned long long)(t_epi - t_start); g_tc_prof[10] += 1; }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

__global__ void tc_splitk_reduce_kernel(ConvTCParams p, size_t npix) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix * p.N) return;
    const size_t pix = i / p.N;
    const int n = (int)(i - pix * p.N);
    float t = 0.f;
    for (int ks = 0; ks < p.ksplit; ++ks) t += p.part[((size_t)ks * npix + pix) * p.BN + n];
    if (p.bias) t += p.bias[n];
    t = fmaxf(p.alpha * t, t);
    if (p.res) t += p.res[pix * p.res_cs + n];
    float* yrow = p.y + pix * p.ycs;
    if (p.accumulate) t += yrow[n];
    if (p.mask) t *= (p.mask[pix * p.mask_cs + n] > 0.f) ? 1.f : p.mask_alpha;
    yrow[n] = t;
}

// ------------------------------------------------------------------------------------------------
// weight preparation: B[tap][n (BN rows)][k (Kpad)] (zero padded, K contiguous) from canonical HWIO, batched over layers
//   transposed_src = 1 : src is [tap][K][N]  (forward conv: K = cin, N = cout)
//   transposed_src = 0 : src is [tap][N][K]  (dgrad: N = cin, K = cout)
// ------------------------------------------------------------------------------------------------
__global__ void tc_prep_weights_kernel(const TcPrepJob* __restrict__ jobs) {
    const TcPrepJob j = jobs[blockIdx.y];
    const size_t total = (size_t)j.taps * j.BN * j.Kpad;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int k = (int)(i % j.Kpad);
        size_t q = i / j.Kpad;
        int n = (int)(q % j.BN);
        int t = (int)(q / j.BN);
        float v = 0.f;
        if (n < j.N && k < j.K)
            v = j.transposed_src ? j.src[((size_t)t * j.K + k) * j.N + n] : j.src[((size_t)t * j.N + n) * j.K + k];
        j.bh[i] = v;          // raw fp32; the tf32 hi/lo split happens in shared memory inside conv_tc_kernel
    }
}

int tc_prep_weights(const TcPrepJob* jobs_dev, int njobs, size_t max_total, cudaStream_t st) {
    if (njobs <= 0) return 0;
    unsigned gx = (unsigned)std::min<size_t>(cdivz(max_total, 256), 512);
    tc_prep_weights_kernel<<<dim3(gx, njobs), 256, 0, st>>>(jobs_dev);
    return check_launch("tc_prep_weights");
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct MapKey {
    uintptr_t addr; int rank; int swz; uint64_t d[4]; uint64_t s[3]; uint32_t b[4];
    bool operator<(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) < 0; }
};
static std::map<MapKey, CUtensorMap>& map_cache() { static std::map<MapKey, CUtensorMap> c; return c; }

// cached cuTensorMapEncodeTiled (fp32, SWIZZLE_128B, zero OOB fill)
static int get_map(const CUtensorMap** out, void* addr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box, bool swizzle128 = true) {
    MapKey k;
    memset(&k, 0, sizeof k);
    k.addr = reinterpret_cast<uintptr_t>(addr); k.rank = rank; k.swz = swizzle128 ? 1 : 0;
    for (int i = 0; i < rank; ++i) { k.d[i] = dims[i]; k.b[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) k.s[i] = strides_bytes[i];
    auto& c = map_cache();
    auto it = c.find(k);
    if (it == c.end()) {
        EncodeTiledFn enc = get_encode();
        MS_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUtensorMap m;
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, addr, dims, strides_bytes, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r)); return -1; }
        if (c.size() >= 4096) {
            // cache full (callers that keep passing fresh pointers): do not evict -- pointers handed out earlier in the same
            // launch sequence must stay valid -- but serve this descriptor from a small ring of uncached slots
            static thread_local CUtensorMap spill[16];
            static thread_local unsigned spill_i = 0;
            CUtensorMap* slot = &spill[spill_i++ & 15u];
            *slot = m;
            *out = slot;
            return 0;
        }
        it = c.emplace(k, m).first;
    }
    *out = &it->second;
    return 0;
}

int tc_get_map(const CUtensorMap** out, void* addr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, bool swizzle128) {
    return get_map(out, addr, rank, dims, strides_bytes, box, swizzle128);
}

bool conv_tc_supported(const ConvGemm& g) {
    if (g.mul != 1 || g.div != 1) return false;                       // stride-1 gathers only
    if (g.x.h != g.y.h || g.x.w != g.y.w) return false;
    if (g.x.c < 8 || g.y.c < 8 || g.y.c > 256) return false;
    if ((g.x.cs & 3) || (reinterpret_cast<uintptr_t>(g.x.p) & 15)) return false;
    if (g.y.h * g.y.w < 64) return false;
    return true;
}

// measured on B200 (scripts/tc_bench.py): below ~64x64 channels the fp32 CUDA-core kernel is as fast or faster
bool conv_tc_profitable(const ConvGemm& g) {
    return conv_tc_supported(g) && (long)g.x.c * g.y.c >= 1024;
}

int conv_tc_read_prof(unsigned long long* out32, int reset) {
    MS_CHECK_CUDA(cudaMemcpyFromSymbol(out32, g_tc_prof, sizeof(unsigned long long) * 32));
    if (reset) { unsigned long long z[32] = {0}; MS_CHECK_CUDA(cudaMemcpyToSymbol(g_tc_prof, z, sizeof z)); }
    return 0;
}

void conv_tc_weight_dims(int N, int K, int& BN, int& Kpad) { BN = (N + 15) / 16 * 16; Kpad = (K + 31) / 32 * 32; }

size_t conv_tc_scratch_floats(int taps, int N, int K) {
    int BN, Kpad; conv_tc_weight_dims(N, K, BN, Kpad);
    return 2 * (size_t)taps * BN * Kpad + 64;
}

int conv_tc_init() {
    static bool done = false;
    if (done) return 0;
    MS_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    MS_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    MS_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    MS_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    done = true;
    return 0;
}

// bw: prepared weights [taps][BN][Kpad] (see tc_prep_weights) for this GEMM orientation.
size_t conv_tc_part_floats() { return (size_t)8 << 20; }      // 32 MB split-K partial-sum scratch

int conv_tc(const ConvGemm& g, const float* bw, cudaStream_t st, float* part) {
    MS_REQUIRE(conv_tc_supported(g), "conv_tc: unsupported geometry");
    if (conv_tc_init()) return -1;
    const int taps = g.kh * g.kw, K = g.x.c, N = g.y.c;
    int BN, Kpad; conv_tc_weight_dims(N, K, BN, Kpad);
    const int kblocks = Kpad / 32;

    ConvTCParams p{};
    p.kh = g.kh; p.kw = g.kw; p.off_y = g.off_y; p.off_x = g.off_x; p.step = g.step;
    p.kblocks = kblocks; p.H = g.y.h; p.W = g.y.w; p.NB = g.y.n;
    if (g.y.w >= 16 || g.y.h < 16) { p.TW = 16; p.TH = 8; } else { p.TW = 8; p.TH = 16; }
    p.tiles_x = cdiv(p.W, p.TW); p.tiles_y = cdiv(p.H, p.TH);
    p.N = N; p.BN = BN;
    p.acc_stride = (BN + 31) / 32 * 32;
    p.n_main = std::max(1, std::min(3, 512 / p.acc_stride - 1));
    {
        int need = (p.n_main + 1) * p.acc_stride;
        p.tmem_cols = need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : (need <= 256 ? 256 : 512)));
    }
    const size_t stage_bytes = 2 * (size_t)A_TILE_BYTES + 2 * (size_t)BN * 128;
    int stages = (int)std::min<size_t>(4, (200 * 1024) / stage_bytes);
    MS_REQUIRE(stages >= 2, "conv_tc: tile does not fit shared memory");
    p.stages = stages;
    p.y = g.y.p; p.ycs = g.y.cs; p.bias = g.bias; p.alpha = g.alpha;
    p.mask = g.mask; p.mask_cs = g.mask_cs; p.mask_alpha = g.mask_alpha;
    p.res = g.res; p.res_cs = g.res_cs; p.accumulate = g.accumulate;
    { const char* e = getenv("MS_TC_DEBUG"); p.dbg = e ? atoi(e) : 0; }
    p.ksplit = 1; p.part = nullptr;

    const CUtensorMap *mapA, *mapB;
    {
        cuuint64_t dims[3] = {(cuuint64_t)Kpad, (cuuint64_t)BN, (cuuint64_t)taps};
        cuuint64_t strides[2] = {(cuuint64_t)Kpad * 4, (cuuint64_t)BN * Kpad * 4};
        cuuint32_t box[3] = {32, (cuuint32_t)BN, 1};
        if (get_map(&mapB, const_cast<float*>(bw), 3, dims, strides, box)) return -1;
    }
    const int grid = p.NB * p.tiles_x * p.tiles_y;
    // ---- v2 (halo patch) when the tap extent is small: undilated 3x3 / 1x1
    {
        const int ext_x = (g.kw - 1) * std::abs(g.step), ext_y = (g.kh - 1) * std::abs(g.step);
        static int halo_enabled = -1;
        if (halo_enabled < 0) { const char* e = getenv("MS_TC_HALO"); halo_enabled = (e && e[0] == '0') ? 0 : 1; }
        static int ts_enabled = -1;
        if (ts_enabled < 0) { const char* e = getenv("MS_TC_TS"); ts_enabled = (e && e[0] == '0') ? 0 : 1; }
        if (ts_enabled && halo_enabled && ext_x <= 2 && ext_y <= 2) {
            // ---- v3: A operand in tensor memory
            ConvTCHaloParams hp{};
            ConvTCParams pt = p;
            // operand stages between splitter and MMA (64 TMEM columns + 2*BN*128 bytes each).  Measured
            // (profiles/r1_tc_stage_depth.log): 2 / 3 / 4 stages run the 128->128 layer in 77.8 / 77.8 / 75.8 us, and a
            // third stage costs cout=128 layers their second hi*hi accumulator (truncation bias -8e-6 instead of -4e-6),
            // so 2 is the default; MS_TC_NS overrides.
            static int ns_env = -1;
            if (ns_env < 0) { const char* e = getenv("MS_TC_NS"); ns_env = e ? atoi(e) : 2; }
            int NS = std::max(2, std::min(4, ns_env));
            while (NS > 2 && 2 * pt.acc_stride + NS * 64 > 512) --NS;
            pt.n_main = std::max(1, std::min(3, (512 - NS * 64) / pt.acc_stride - 1));
            const int need = (pt.n_main + 1) * pt.acc_stride + NS * 64;
            if (need <= 512) {
                pt.tmem_cols = need <= 256 ? 256 : 512;
                hp.c = pt;
                hp.PW = p.TW + ext_x; hp.PH = p.TH + ext_y;
                hp.minx = g.off_x + (g.step < 0 ? (g.kw - 1) * g.step : 0);
                hp.miny = g.off_y + (g.step < 0 ? (g.kh - 1) * g.step : 0);
                hp.patch_bytes = (uint32_t)hp.PW * hp.PH * 128u;
                const size_t b_bytes = (size_t)BN * 128, op_bytes = 2 * b_bytes;
                const size_t patch_stride = (hp.patch_bytes + 1023) & ~(size_t)1023;
                size_t fixed = NS * op_bytes + 2 * patch_stride + 1024;
                while (NS > 2 && fixed + 3 * b_bytes > 224 * 1024) { --NS; fixed = NS * op_bytes + 2 * patch_stride + 1024; }
                hp.ns = NS;
                int nb = fixed + 2 * b_bytes <= 224 * 1024 ? (int)std::min<size_t>(8, (224 * 1024 - fixed) / b_bytes) : 0;
                if (nb >= 2) {
                    hp.nb_slots = nb;
                    const CUtensorMap* mapP;
                    cuuint64_t dims[4] = {(cuuint64_t)g.x.c, (cuuint64_t)g.x.w, (cuuint64_t)g.x.h, (cuuint64_t)g.x.n};
                    cuuint64_t strides[3] = {(cuuint64_t)g.x.cs * 4, (cuuint64_t)g.x.w * g.x.cs * 4, (cuuint64_t)g.x.h * g.x.w * g.x.cs * 4};
                    cuuint32_t box[4] = {32, (cuuint32_t)hp.PW, (cuuint32_t)hp.PH, 1};
                    if (get_map(&mapP, g.x.p, 4, dims, strides, box, true)) return -1;
                    const size_t smem2 = fixed + (size_t)nb * b_bytes;
                    // small maps leave most SMs idle while each CTA walks the whole K loop: split K over channel blocks
                    const size_t npix = (size_t)p.NB * p.H * p.W;
                    int ksplit = 1;
                    static int tapsplit = -1;
                    if (tapsplit < 0) { const char* e = getenv("MS_TC_TAPSPLIT"); tapsplit = (e && e[0] == '0') ? 0 : 1; }
                    const int units = tapsplit ? taps * kblocks : kblocks;       // split-K granularity: (kb, tap) iterations
                    if (part && grid <= 74 && units > 1) {
                        ksplit = std::min(units, std::max(1, 148 / grid));
                        if (!tapsplit) ksplit = std::min(ksplit, kblocks);
                        if ((size_t)ksplit * npix * BN > conv_tc_part_floats()) ksplit = 1;
                    }
                    hp.c.ksplit = ksplit; hp.c.part = part;
                    if (p.dbg & 8) conv_tc_ts_kernel<true><<<dim3(grid, ksplit), TC_THREADS, smem2, st>>>(*mapP, *mapB, hp);
                    else conv_tc_ts_kernel<false><<<dim3(grid, ksplit), TC_THREADS, smem2, st>>>(*mapP, *mapB, hp);
                    if (ksplit > 1) {
                        tc_splitk_reduce_kernel<<<(unsigned)cdivz(npix * p.N, 256), 256, 0, st>>>(hp.c, npix);
                        return check_launch("conv_tc_ts+reduce", 2);
                    }
                    return check_launch("conv_tc_ts");
                }
            }
        }
        if (halo_enabled && ext_x <= 2 && ext_y <= 2) {
            ConvTCHaloParams hp{};
            hp.c = p;
            hp.PW = p.TW + ext_x; hp.PH = p.TH + ext_y;
            hp.minx = g.off_x + (g.step < 0 ? (g.kw - 1) * g.step : 0);
            hp.miny = g.off_y + (g.step < 0 ? (g.kh - 1) * g.step : 0);
            hp.patch_bytes = (uint32_t)hp.PW * hp.PH * 128u;
            const size_t b_bytes = (size_t)BN * 128, op_bytes = 2 * (size_t)A_TILE_BYTES + 2 * b_bytes;
            const size_t patch_stride = (hp.patch_bytes + 1023) & ~(size_t)1023;
            const size_t fixed = 2 * op_bytes + 2 * patch_stride + 1024;
            int nb = fixed + 2 * b_bytes <= 224 * 1024 ? (int)std::min<size_t>(8, (224 * 1024 - fixed) / b_bytes) : 0;
            if (nb >= 2) {
                hp.nb_slots = nb;
                const CUtensorMap* mapP;
                cuuint64_t dims[4] = {(cuuint64_t)g.x.c, (cuuint64_t)g.x.w, (cuuint64_t)g.x.h, (cuuint64_t)g.x.n};
                cuuint64_t strides[3] = {(cuuint64_t)g.x.cs * 4, (cuuint64_t)g.x.w * g.x.cs * 4, (cuuint64_t)g.x.h * g.x.w * g.x.cs * 4};
                cuuint32_t box[4] = {32, (cuuint32_t)hp.PW, (cuuint32_t)hp.PH, 1};
                if (get_map(&mapP, g.x.p, 4, dims, strides, box, false)) return -1;
                const size_t smem2 = fixed + (size_t)nb * b_bytes;
                conv_tc_halo_kernel<<<grid, TC_THREADS, smem2, st>>>(*mapP, *mapB, hp);
                return check_launch("conv_tc_halo");
            }
        }
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)g.x.c, (cuuint64_t)g.x.w, (cuuint64_t)g.x.h, (cuuint64_t)g.x.n};
        cuuint64_t strides[3] = {(cuuint64_t)g.x.cs * 4, (cuuint64_t)g.x.w * g.x.cs * 4, (cuuint64_t)g.x.h * g.x.w * g.x.cs * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
        if (get_map(&mapA, g.x.p, 4, dims, strides, box)) return -1;
    }
    const size_t smem = (size_t)stages * stage_bytes + 1024;
    conv_tc_kernel<<<grid, TC_THREADS, smem, st>>>(*mapA, *mapB, p);
    return check_launch("conv_tc");
}

// one-shot convenience (operator-level C ABI): prepares the weight halves into `scratch` first.
//   g.wmat must be the CANONICAL weights: [tap][x.c][y.c] if !wmat_is_nk else [tap][y.c][x.c].
int conv_tc_oneshot(const ConvGemm& g, int wmat_is_nk, float* scratch, size_t scratch_floats, cudaStream_t st) {
    MS_REQUIRE(conv_tc_supported(g), "conv_tc: unsupported geometry");
    const int taps = g.kh * g.kw, K = g.x.c, N = g.y.c;
    int BN, Kpad; conv_tc_weight_dims(N, K, BN, Kpad);
    const size_t per = (size_t)taps * BN * Kpad;
    MS_REQUIRE(scratch_floats >= 2 * per + 64, "conv_tc: scratch too small");
    MS_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 15) == 0, "conv_tc: scratch must be 16B aligned");
    TcPrepJob job{g.wmat, scratch, scratch + per, taps, N, K, BN, Kpad, wmat_is_nk ? 0 : 1};
    TcPrepJob* jd = reinterpret_cast<TcPrepJob*>(scratch + 2 * per);   // 64 spare floats hold the job descriptor
    MS_CHECK_CUDA(cudaMemcpyAsync(jd, &job, sizeof job, cudaMemcpyHostToDevice, st));
    if (tc_prep_weights(jd, 1, per, st)) return -1;
    return conv_tc(g, scratch, st, nullptr);
}

}  // namespace ms
