// wgrad_bf: tcgen05 weight (+ bias) gradient of a convolution on split-bf16 operands, stride 1 or 2, any dilation.
//
// Replaces the filter- and bias-gradient sub-graphs tf.gradients derives for tf.nn.conv2d / tf.nn.atrous_conv2d /
// tf.nn.bias_add (reference Nets/sharedLayers.py:58-59,72-73) inside the train ops of
// Stereo_Online_Adaptation.py:118,128.
//
//   dW[r][s][ci][co] = sum over output pixels p of  X[p * stride + offset(r, s)][ci] * dY[p][co]
//   db[co]           = sum over output pixels p of  dY[p][co]
//
// Both operands already exist as bf16 hi / lo planes in NHWC (the forward activation planes conv_bf reads, and the
// gradient planes the dgrad epilogue writes), i.e. with the GEMM's reduction index (pixels) as the SLOW index.  The
// UMMA descriptors take that layout directly as "MN-major" operands: a TMA box {64 channels, 8 pixels, rows}
// with SWIZZLE_128B is exactly the canonical MN-major SW128 tile ((8,n),(8,k)):((1,LBO),(8,SBO)) [uint128 units] with
// SBO = 1024 B (8 pixels) and LBO = the distance between 64-channel blocks -- no transposes, no in-kernel splitting
// (the round-1 kernel, wgrad_tc.cu, needs an NHWC->NCHW copy of dY and splits fp32 into tf32 halves in the main loop).
//
// GEMM per CTA: one filter COLUMN s (all kh taps of it: one TMEM accumulator per tap), one 128-row block of ci,
// one block of BN output channels, one slice of the pixel tiles (split-K):
//   M = ci (128 TMEM lanes), N = BN co, K = pixels in tiles of 8 x TH.
//   For stride 1 and small dilation the kh taps read ONE halo patch of X (TH + (kh-1)*dil rows; a tap is a row offset =
//   a whole number of 1024-byte swizzle atoms = a different descriptor start address); otherwise one box per tap
//   (stride 2 through TMA element strides).
//   Three kind::f16 MMAs per K step: X_lo*dY_hi + X_hi*dY_lo + X_hi*dY_hi (~2^-16 relative product error).
//   Bias gradient: CTAs of column 0 / block 0 run two more MMAs per K step with an all-ones A tile -> db in one more
//   accumulator (no separate reduction kernel over dY).
// Partial sums [split][tap][ci][co] (+ [split][co]) go to the workspace; a fixed-order reduce finishes (deterministic).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ms {

constexpr int WB_THREADS = 320;
constexpr int WB_MAX_KH = 7;
constexpr uint32_t WB_ONES_BYTES = 4096;       // 2 channel blocks x 2 k-groups x 1024 B of bf16 1.0
constexpr uint32_t WB_ATOM = 1024;             // one tile row: 8 pixels x 64 channels x 2 B

struct WgradBfParams {
    int kh, kw;
    int tiles_x, tiles_y, ntiles, splits;
    int TH;                        // pixel tile = 8 wide x TH high (TH even)
    int sx;                        // forward stride: input pixel = out * sx + offset
    int nbox, box_rows, x_rows;    // X boxes per 64-channel block and plane; tile rows per box; rows of the X region
    short box_dy[WB_MAX_KH];       // input-row origin of box b relative to y0 * sx
    short tap_row[WB_MAX_KH];      // first X-region row of tap r
    int pad_l, dil;
    int ci, co, mblocks, nblocks, BN, xblk, dblk;
    int nstages;
    uint32_t stage_bytes, x_plane_bytes, d_plane_bytes;
    int tmem_cols;
    int with_bias;
    int xfmt, dfmt;                // plane formats of X and dY (0 = bf16, 1 = fp16 of value / 16)
    float* part;                   // [split][tap][ci][co]
    float* bpart;                  // [split][co]
    int tap0, taps_total;          // this launch covers filter rows [tap0 / kw, tap0 / kw + kh) of a taps_total-tap filter
    int debug;                     // MS_WB_DEBUG bit mask (diagnosis): 1 no MMAs, 2 no TMA loads, 4 main product only, 8 product-major issue order
};

__device__ __forceinline__ void wb_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// MN-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO (distance
// between 64-element blocks along M/N) | SBO (distance between 8-row groups along K) | version 1 | layout_type 2.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_byte_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_byte_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(WB_THREADS, 1)
wgrad_bf_kernel(const __grid_constant__ CUtensorMap mapXh, const __grid_constant__ CUtensorMap mapXl,
                const __grid_constant__ CUtensorMap mapDh, const __grid_constant__ CUtensorMap mapDl,
                const __grid_constant__ WgradBfParams p) {
    pdl_prologue();
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[4], empty_bar[4], accum_bar;
    __shared__ uint32_t tmem_slot;

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const uint32_t base = (s_addr(smem_dyn) + 1023u) & ~1023u;
    unsigned char* gbase = smem_dyn + (base - s_addr(smem_dyn));
    const uint32_t stage0 = WB_ONES_BYTES;

    int bx = blockIdx.x;
    const int nb = bx % p.nblocks; bx /= p.nblocks;
    const int mb = bx % p.mblocks;
    const int s = bx / p.mblocks;                       // filter column
    const int split = blockIdx.y;
    const int t0 = (int)(((long)split * p.ntiles) / p.splits), t1 = (int)(((long)(split + 1) * p.ntiles) / p.splits);
    const int total = t1 - t0;
    const bool do_bias = p.with_bias && s == 0 && mb == 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.nstages; ++i) { mb_init(&full_bar[i], 1); mb_init(&empty_bar[i], 1); }
        mb_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(&tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2) {                                    // the all-ones A tile of the bias-gradient MMAs
        uint32_t* ones = reinterpret_cast<uint32_t*>(gbase);
        const uint32_t one2 = p.dfmt == 0 ? 0x3F803F80u : 0x3C003C00u;      // 1.0 in the format of the dY planes
        for (int i = threadIdx.x - 64; i < (int)(WB_ONES_BYTES / 4); i += WB_THREADS - 64) ones[i] = one2;
        fence_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        // ================= TMA producer: warp-uniform loop, elected lane issues =================
        const bool leader = elect_one();
        if (total > 0) {
            int st = 0; uint32_t ph = 0;
            const int tiles_img = p.tiles_x * p.tiles_y;
            const int offx = s * p.dil - p.pad_l;
            for (int t = t0; t < t1; ++t) {
                const int img = t / tiles_img;
                const int rem = t - img * tiles_img;
                const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
                mb_wait(&empty_bar[st], ph ^ 1u);
                if (p.debug & 2) { if (leader) mb_arrive(&full_bar[st]); if (++st == p.nstages) { st = 0; ph ^= 1u; } continue; }
                if (leader) mb_expect_tx(&full_bar[st], p.stage_bytes);
                unsigned char* dst = gbase + stage0 + (size_t)st * p.stage_bytes;
                for (int pl = 0; pl < 2 && leader; ++pl) {
                    const CUtensorMap* mx = pl ? &mapXl : &mapXh;
                    unsigned char* xd = dst + (size_t)pl * p.x_plane_bytes;
                    for (int b = 0; b < p.xblk; ++b)
                        for (int q = 0; q < p.nbox; ++q)
                            tma_load_4d(xd + (size_t)(b * p.x_rows + q * p.box_rows) * WB_ATOM, mx, &full_bar[st],
                                        mb * 128 + b * 64, tx * 8 * p.sx + offx, ty * p.TH * p.sx + p.box_dy[q], img);
                    const CUtensorMap* md = pl ? &mapDl : &mapDh;
                    unsigned char* dd = dst + 2 * (size_t)p.x_plane_bytes + (size_t)pl * p.d_plane_bytes;
                    for (int b = 0; b < p.dblk; ++b)
                        tma_load_4d(dd + (size_t)(b * p.TH) * WB_ATOM, md, &full_bar[st], nb * p.BN + b * 64, tx * 8, ty * p.TH, img);
                }
                if (++st == p.nstages) { st = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: warp-uniform loop, one elected lane issues (tc_ptx.cuh:elect_one) =================
        const bool leader = elect_one();
        const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
        if (total > 0) {
            // D = f32, A = B = bf16, both MN-major (bits 15, 16), N >> 3 at bit 17, M >> 4 at bit 24
            // A (= X) / B (= dY) element formats follow the planes: 0 = f16, 1 = bf16 in the descriptor
            const uint32_t fa = p.xfmt == 0 ? 1u : 0u, fb = p.dfmt == 0 ? 1u : 0u;
            const uint32_t idesc = (1u << 4) | (fa << 7) | (fb << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t idesc_ones = (1u << 4) | (fb << 7) | (fb << 10) | (1u << 15) | (1u << 16) |
                                        ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t lbo_x = p.xblk > 1 ? (uint32_t)p.x_rows * WB_ATOM : 0u;
            const uint32_t lbo_d = (uint32_t)p.TH * WB_ATOM;
            const uint64_t ones = umma_desc_mn_sw128(base, 2048u, 1024u);
            const uint32_t acc_bias = tmem + (uint32_t)(p.kh * p.BN);
            int st = 0; uint32_t ph = 0;
            uint32_t started = 0;
            const int ksteps = p.TH >> 1;
            for (int t = t0; t < t1; ++t) {
                mb_wait(&full_bar[st], ph);
                tc_fence_after();
                const uint32_t sb = base + stage0 + (uint32_t)st * p.stage_bytes;
                const uint32_t xh = sb, xl = sb + p.x_plane_bytes;
                const uint32_t dh = sb + 2u * p.x_plane_bytes, dl = dh + p.d_plane_bytes;
                for (int j = 0; j < ((p.debug & 1) ? 0 : ksteps); ++j) {
                    const uint64_t bh = umma_desc_mn_sw128(dh + (uint32_t)(2 * j) * WB_ATOM, lbo_d, 1024u);
                    const uint64_t bl = umma_desc_mn_sw128(dl + (uint32_t)(2 * j) * WB_ATOM, lbo_d, 1024u);
                    if (p.debug & 12) {                 // diagnosis variants: main product only (4) / product-major order (8)
                        for (int pr = (p.debug & 4) ? 2 : 0; pr < 3; ++pr)
                            for (int r = 0; r < p.kh; ++r) {
                                const uint32_t ro = (uint32_t)(p.tap_row[r] + 2 * j) * WB_ATOM;
                                const uint64_t ah = umma_desc_mn_sw128(xh + ro, lbo_x, 1024u);
                                const uint64_t al = umma_desc_mn_sw128(xl + ro, lbo_x, 1024u);
                                const uint32_t acc = tmem + (uint32_t)(r * p.BN);
                                const bool first = pr == ((p.debug & 4) ? 2 : 0);
                                if (leader) wb_mma_f16(acc, pr == 0 ? al : ah, pr == 1 ? bl : bh, idesc, first ? started : 1u);
                            }
                        started = 1u;
                        continue;
                    }
                    for (int r = 0; r < p.kh; ++r) {
                        const uint32_t ro = (uint32_t)(p.tap_row[r] + 2 * j) * WB_ATOM;
                        const uint64_t ah = umma_desc_mn_sw128(xh + ro, lbo_x, 1024u);
                        const uint64_t al = umma_desc_mn_sw128(xl + ro, lbo_x, 1024u);
                        const uint32_t acc = tmem + (uint32_t)(r * p.BN);
                        if (leader) {
                            wb_mma_f16(acc, al, bh, idesc, started);
                            wb_mma_f16(acc, ah, bl, idesc, 1u);
                            wb_mma_f16(acc, ah, bh, idesc, 1u);
                        }
                    }
                    if (do_bias && leader) {
                        wb_mma_f16(acc_bias, ones, bl, idesc_ones, started);
                        wb_mma_f16(acc_bias, ones, bh, idesc_ones, 1u);
                    }
                    started = 1u;
                }
                if (leader) tc_commit(&empty_bar[st]);
                if (++st == p.nstages) { st = 0; ph ^= 1u; }
            }
            if (leader) tc_commit(&accum_bar);
            __syncwarp();
        }
    } else {
        // ================= epilogue (warps 2..9): thread <-> ci row, columns <-> (tap, co) =================
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;
        const int m = q * 32 + lane;
        const int ci_g = mb * 128 + m;
        const bool valid = ci_g < p.ci;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        if (total > 0) {
            mb_wait(&accum_bar, 0);
            tc_fence_after();
        }
        const int cpt = p.BN >> 4;                       // 16-column chunks per tap
        const int chunks = p.kh * cpt;
        const int cb = half ? (chunks + 1) / 2 : 0, ce = half ? chunks : (chunks + 1) / 2;
        const bool vec = (p.co & 3) == 0;
        const int taps = p.taps_total;
        for (int c = cb; c < ce; ++c) {
            const int r = c / cpt, c0 = (c - r * cpt) * 16;
            uint32_t v[16];
            if (total > 0) {
                tc_ld16_nowait(tmem + lane_base + (uint32_t)(r * p.BN + c0), v);
                tc_wait_ld();
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0u;
            }
            if (!valid) continue;
            const int tap = p.tap0 + r * p.kw + s;
            const int co0 = nb * p.BN + c0;
            float* prow = p.part + (((size_t)split * taps + tap) * p.ci + ci_g) * p.co + co0;
            if (vec && co0 + 16 <= p.co) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(prow + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                        __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (co0 + j < p.co) prow[j] = __uint_as_float(v[j]);
            }
        }
        if (do_bias && warp == 4) {                      // warp 4: lane quarter 0, row 0 holds sum_p dY[p][co]
            for (int c0 = 0; c0 < p.BN; c0 += 16) {
                uint32_t v[16];
                if (total > 0) {
                    tc_ld16_nowait(tmem + (uint32_t)(p.kh * p.BN + c0), v);
                    tc_wait_ld();
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0u;
                }
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int co = nb * p.BN + c0 + j;
                        if (co < p.co) p.bpart[(size_t)split * p.co + co] = __uint_as_float(v[j]);
                    }
                }
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

// dw[i] = sum_k part[k][i] (fixed order), db[c] = sum_k bpart[k][c]
__global__ void wgrad_bf_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, size_t n4, int split,
                                       const float* __restrict__ bpart, float* __restrict__ db, int co, int accumulate,
                                       float wscale, float bscale) {
    pdl_prologue();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        const float4* src = reinterpret_cast<const float4*>(part) + i;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < split; ++k) {
            const float4 v = __ldcs(src + (size_t)k * n4);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        a.x *= wscale; a.y *= wscale; a.z *= wscale; a.w *= wscale;      // undoes the 1/16 pre-scale of fp16 planes (exact)
        float4* d = reinterpret_cast<float4*>(dw) + i;
        if (accumulate) { const float4 o = *d; a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w; }
        *d = a;
    } else if (db && i < n4 + (size_t)co) {
        const int c = (int)(i - n4);
        float a = 0.f;
        for (int k = 0; k < split; ++k) a += bpart[(size_t)k * co + c];
        a *= bscale;
        if (accumulate) a += db[c];
        db[c] = a;
    }
}

struct WbPlan {
    int TH, nbox, box_rows, x_rows, nstages, nblocks, BN, mblocks, xblk, dblk, splits, tiles_x, tiles_y, ntiles, tmem_cols;
    uint32_t stage_bytes, x_plane_bytes, d_plane_bytes;
    bool ok;
};

static WbPlan wb_plan(const ConvWgrad& q) {
    WbPlan P{};
    const int ci = q.x.c, co = q.dy.c, kh = q.kh, kw = q.kw;
    P.ok = false;
    if (q.stride != 1 && q.stride != 2) return P;
    if (kh > WB_MAX_KH || kh * kw > 49) return P;
    if (ci < 3 || co < 16 || (co & 3)) return P;
    if ((size_t)q.dy.h * q.dy.w < 32) return P;
    // output-channel blocks: (kh taps + bias) accumulators of BN columns in 512 TMEM columns
    int nbk = 1, BN = 0;
    for (;; ++nbk) {
        BN = ((co + nbk - 1) / nbk + 15) / 16 * 16;
        if ((kh + 1) * BN <= 512 && BN <= 256) break;
        if (nbk > 16) return P;
    }
    P.nblocks = nbk; P.BN = BN;
    P.mblocks = (ci + 127) / 128;
    P.xblk = ci > 64 ? 2 : 1;
    P.dblk = (BN + 63) / 64;
    const int need = (kh + 1) * BN;
    P.tmem_cols = need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : (need <= 256 ? 256 : 512)));
    const size_t budget = 222 * 1024 - WB_ONES_BYTES;
    const int cand[4] = {16, 8, 4, 2};
    for (int i = 0; i < 4; ++i) {
        const int TH = cand[i];
        if (TH > 2 && TH >= 2 * q.dy.h) continue;                      // do not pad tiny maps to tall tiles
        const int halo = (kh - 1) * q.dil;
        const bool shared = q.stride == 1 && halo < (kh - 1) * TH && (TH + halo) <= 256;
        // stride 2, dilation 1: taps of equal row parity read the same strided patch (tap r = row r/2 of box r%2)
        const bool parity = q.stride == 2 && q.dil == 1 && kh > 2;
        const int nbox = shared ? 1 : (parity ? 2 : kh);
        const int box_rows = shared ? TH + halo : (parity ? TH + (kh - 1) / 2 : TH);
        const int x_rows = nbox * box_rows;
        const size_t xpb = (size_t)P.xblk * x_rows * WB_ATOM, dpb = (size_t)P.dblk * TH * WB_ATOM;
        const size_t stage = 2 * (xpb + dpb);
        const int ns = (int)std::min<size_t>(4, budget / stage);
        if (ns < 2 || (ns < 3 && TH > 2)) continue;
        if (box_rows * q.stride > 256) continue;
        P.TH = TH; P.nbox = nbox; P.box_rows = box_rows; P.x_rows = x_rows; P.nstages = ns;
        P.stage_bytes = (uint32_t)stage; P.x_plane_bytes = (uint32_t)xpb; P.d_plane_bytes = (uint32_t)dpb;
        P.ok = true;
        break;
    }
    if (!P.ok) return P;
    P.tiles_x = cdiv(q.dy.w, 8); P.tiles_y = cdiv(q.dy.h, P.TH);
    P.ntiles = q.dy.n * P.tiles_x * P.tiles_y;
    const int cols = kw * P.mblocks * P.nblocks;
    int splits = std::max(1, (148 + cols / 2) / cols);
    splits = std::min(splits, std::min(P.ntiles, 64));
    // bound the number of accumulation steps per CTA (the tensor core adds into fp32 with truncation)
    const int max_tiles = std::max(1, 8192 / (8 * P.TH));
    splits = std::max(splits, std::min(64, cdiv(P.ntiles, max_tiles)));
    P.splits = std::max(1, std::min(splits, P.ntiles));
    return P;
}

bool wgrad_bf_supported(const ConvWgrad& q) { return wb_plan(q).ok; }

size_t wgrad_bf_workspace_floats(int kh, int kw, int ci, int co) {
    return 64 * ((size_t)kh * kw * ci * co + (size_t)co) + 64;
}

int wgrad_bf_init() {
    static bool done = false;
    if (done) return 0;
    MS_CHECK_CUDA(cudaFuncSetAttribute(wgrad_bf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    done = true;
    return 0;
}

// One kernel launch over the filter rows [r0, r0 + q.kh) of a kh_total-row filter (q.pad_t already shifted to row r0).
// splits_force > 0: use that split factor (all row groups of one filter share the partial-sum layout).  Returns the split
// factor used through *splits_out.
static int wgrad_bf_launch(const ConvWgrad& q, const ActPlanes& xp, const ActPlanes& dp, int r0, int kh_total, int splits_force,
                           int* splits_out, cudaStream_t st) {
    WbPlan P = wb_plan(q);
    MS_REQUIRE(P.ok, "wgrad_bf: unsupported geometry");
    {   // the split factor never outgrows the caller's workspace
        const size_t per = (size_t)kh_total * q.kw * q.x.c * q.dy.c + (size_t)q.dy.c;
        P.splits = (int)std::max<size_t>(1, std::min<size_t>((size_t)P.splits, q.workspace_floats / std::max<size_t>(per, 1)));
        if (splits_force > 0) P.splits = std::max(1, std::min(splits_force, P.ntiles));
        MS_REQUIRE(splits_force <= 0 || P.splits == splits_force, "wgrad_bf: row groups disagree on the split factor");
    }
    MS_REQUIRE(xp.fmt == dp.fmt, "wgrad_bf: tcgen05 kind::f16 rejects mixed f16 x bf16 operands (probed: illegal instruction); both plane sets must share a format");
    MS_REQUIRE(xp.hi && xp.lo && dp.hi && dp.lo && (xp.cs & 7) == 0 && (dp.cs & 7) == 0 && xp.cs >= q.x.c && dp.cs >= q.dy.c,
               "wgrad_bf: operand planes missing");
    if (wgrad_bf_init()) return -1;
    const int ci = q.x.c, co = q.dy.c, taps = kh_total * q.kw;
    const size_t wn = (size_t)taps * ci * co;
    MS_REQUIRE((wn & 3) == 0, "wgrad_bf: taps*ci*co must be a multiple of 4");
    MS_REQUIRE(q.workspace_floats >= wn + co, "wgrad_bf: workspace too small");
    MS_REQUIRE((reinterpret_cast<uintptr_t>(q.workspace) & 15) == 0 && (reinterpret_cast<uintptr_t>(q.dw) & 15) == 0,
               "wgrad_bf: workspace / dw must be 16-byte aligned");
    static WgradBfParams p;
    memset(&p, 0, sizeof p);
    p.kh = q.kh; p.kw = q.kw;
    p.tiles_x = P.tiles_x; p.tiles_y = P.tiles_y; p.ntiles = P.ntiles; p.splits = P.splits;
    p.TH = P.TH; p.sx = q.stride; p.nbox = P.nbox; p.box_rows = P.box_rows; p.x_rows = P.x_rows;
    const bool parity = q.stride == 2 && q.dil == 1 && q.kh > 2 && P.nbox == 2;
    for (int r = 0; r < q.kh; ++r) {
        if (P.nbox == 1) { p.tap_row[r] = (short)(r * q.dil); }
        else if (parity) { p.tap_row[r] = (short)((r & 1) * P.box_rows + (r >> 1)); }
        else { p.tap_row[r] = (short)(r * P.TH); p.box_dy[r] = (short)(r * q.dil - q.pad_t); }
    }
    if (P.nbox == 1) p.box_dy[0] = (short)(-q.pad_t);
    if (parity) { p.box_dy[0] = (short)(-q.pad_t); p.box_dy[1] = (short)(1 - q.pad_t); }
    p.pad_l = q.pad_l; p.dil = q.dil;
    p.ci = ci; p.co = co; p.mblocks = P.mblocks; p.nblocks = P.nblocks; p.BN = P.BN; p.xblk = P.xblk; p.dblk = P.dblk;
    p.nstages = P.nstages; p.stage_bytes = P.stage_bytes; p.x_plane_bytes = P.x_plane_bytes; p.d_plane_bytes = P.d_plane_bytes;
    p.tmem_cols = P.tmem_cols;
    p.with_bias = (q.db && r0 == 0) ? 1 : 0;
    p.tap0 = r0 * q.kw; p.taps_total = taps;
    p.xfmt = xp.fmt; p.dfmt = dp.fmt;
    p.part = q.workspace;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("MS_WB_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
    p.bpart = q.workspace + (size_t)P.splits * wn;

    const CUtensorMap *mXh, *mXl, *mDh, *mDl;
    {
        cuuint64_t dims[4] = {(cuuint64_t)ci, (cuuint64_t)q.x.w, (cuuint64_t)q.x.h, (cuuint64_t)q.x.n};
        cuuint64_t strides[3] = {(cuuint64_t)xp.cs * 2, (cuuint64_t)q.x.w * xp.cs * 2, (cuuint64_t)q.x.h * q.x.w * xp.cs * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(8 * q.stride), (cuuint32_t)(P.box_rows * q.stride), 1};
        cuuint32_t es[4] = {1, (cuuint32_t)q.stride, (cuuint32_t)q.stride, 1};
        if (bf_get_map(&mXh, xp.hi, 4, dims, strides, box, es, 128)) return -1;
        if (bf_get_map(&mXl, xp.lo, 4, dims, strides, box, es, 128)) return -1;
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)co, (cuuint64_t)q.dy.w, (cuuint64_t)q.dy.h, (cuuint64_t)q.dy.n};
        cuuint64_t strides[3] = {(cuuint64_t)dp.cs * 2, (cuuint64_t)q.dy.w * dp.cs * 2, (cuuint64_t)q.dy.h * q.dy.w * dp.cs * 2};
        cuuint32_t box[4] = {64, 8, (cuuint32_t)P.TH, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        if (bf_get_map(&mDh, dp.hi, 4, dims, strides, box, es, 128)) return -1;
        if (bf_get_map(&mDl, dp.lo, 4, dims, strides, box, es, 128)) return -1;
    }
    const size_t smem = WB_ONES_BYTES + (size_t)P.nstages * P.stage_bytes + 1024;
    launch_k(wgrad_bf_kernel, dim3(dim3(q.kw * P.mblocks * P.nblocks, P.splits)), dim3(WB_THREADS), smem, st, *mXh, *mXl, *mDh, *mDl, p);
    if (splits_out) *splits_out = P.splits;
    return check_launch("wgrad_bf");
}

// Filter rows per launch.  Every filter row keeps its own accumulator (BN TMEM columns each, plus the bias row), so a tall
// filter narrows BN: 5 rows leave 64 columns.  Measured (profiles/r2_wgrad_bf_killswitch.log, MMAs only, no loads):
// 5 rows x BN 64 cost 200 cycles per MMA, 3 rows x BN 128 cost 119 -- per column three times dearer -- so filters with
// 5 or more rows run as groups of <= 3 rows, each group its own launch with its own, shorter halo.
static int wb_group_rows(const ConvWgrad& q) {
    static int grp = -1;
    if (grp < 0) { const char* e = getenv("MS_WB_GROUPS"); grp = (e && e[0] == '0') ? 0 : 1; }
    // in the DispNet step graph (profiles/r2_layers_in_graph_cfg4.json vs r2_layers_cfg4_groups.json): 5 x 5 layers gain
    // (conv2 462 -> 349 us, conv3 238 -> 152 us), the 4 x 4 transposed-conv gradients as 2 + 2 rows lose (36 -> 73 us)
    if (!grp || q.kh <= 4) return q.kh;
    return 3;
}

// xp / dp: bf16 planes of q.x / q.dy
int wgrad_bf(const ConvWgrad& q, const ActPlanes& xp, const ActPlanes& dp, cudaStream_t st) {
    const int khg = wb_group_rows(q);
    int splits = 0;
    for (int r0 = 0; r0 < q.kh; r0 += khg) {
        ConvWgrad g = q;
        g.kh = std::min(khg, q.kh - r0);
        g.pad_t = q.pad_t - r0 * q.dil;
        if (r0) g.db = q.db;                       // (bias only in the first group; the launch looks at r0)
        MS_REQUIRE(wb_plan(g).ok, "wgrad_bf: unsupported geometry");
        if (r0 == 0 && khg < q.kh) {               // the shortest group has the same tile count: plans agree on ntiles, take the first group's split
            WbPlan P0 = wb_plan(g);
            const size_t per = (size_t)q.kh * q.kw * q.x.c * q.dy.c + (size_t)q.dy.c;
            splits = (int)std::max<size_t>(1, std::min<size_t>((size_t)P0.splits, q.workspace_floats / std::max<size_t>(per, 1)));
            for (int r1 = khg; r1 < q.kh; r1 += khg) {      // ... clamped to what every group can honour
                ConvWgrad h = q; h.kh = std::min(khg, q.kh - r1); h.pad_t = q.pad_t - r1 * q.dil;
                WbPlan P1 = wb_plan(h);
                MS_REQUIRE(P1.ok, "wgrad_bf: unsupported geometry");
                splits = std::min(splits, P1.ntiles);
            }
            splits = std::min(splits, P0.ntiles);
        }
        int used = 0;
        if (wgrad_bf_launch(g, xp, dp, r0, q.kh, splits, &used, st)) return -1;
        splits = used;
    }
    const int ci = q.x.c, co = q.dy.c, taps = q.kh * q.kw;
    const size_t wn = (size_t)taps * ci * co;
    float* part = q.workspace;
    float* bpart = q.workspace + (size_t)splits * wn;
    struct { int splits; } P{splits};
    struct { float* part; float* bpart; } p{part, bpart};
    const size_t n4 = wn / 4;
    const size_t work = n4 + (q.db ? (size_t)co : 0);
    const float sx16 = xp.fmt == 1 ? 1.f / xp.scale : 1.f, sd16 = dp.fmt == 1 ? 1.f / dp.scale : 1.f;
    launch_k(wgrad_bf_reduce_kernel, dim3((unsigned)cdivz(work, 256)), dim3(256), 0, st, p.part, q.dw, n4, P.splits, p.bpart, q.db, co, q.accumulate,
                                                                      sx16 * sd16, sd16);
    return check_launch("wgrad_bf_reduce", 1);
}

// ------------------------------------------------------------------------------------------------
// tcgen05.mma cost probe (diagnosis, scripts/mma_probe.py): one CTA per SM, one thread issues `iters` back-to-back
// kind::f16 MMAs (M = 128, K = 16, bf16) on zero-filled shared memory -- no loads, no epilogue -- and the cycles between the
// first issue and the completion of the last one are reported per CTA.
//   a_mn / b_mn : operand layout (0 = K-major SW128 as in conv_bf, 1 = MN-major SW128 as in wgrad_bf)
//   n           : MMA N;   n_acc : accumulators visited round-robin (each n columns);  rot : 1 = rotate the operand
//   addresses over 4 atoms like the real K loop, 0 = the same operands every time;  uni : issue scheme (see the kernel)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t probe_desc_k_sw128(uint32_t smem_byte_addr) {
    return (uint64_t)((smem_byte_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int a_mn, int b_mn, int n, int n_acc, int rot, int iters, int uni, long long* out) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (s_addr(smem_dyn) + 1023u) & ~1023u;
    unsigned char* gbase = smem_dyn + (base - s_addr(smem_dyn));
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gbase)[i] = 0u;
    fence_async_smem();
    if (threadIdx.x == 0) { mb_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int warp = uniform_warp_idx();
    // uni = 0: the loop runs on lane 0 only (round-2 kernels before this probe); uni = 1: on the whole warp, uniform control
    // flow, the MMA itself predicated on an elected lane
    if ((uni && warp == 0) || (!uni && threadIdx.x == 0)) {
        const bool leader = uni ? elect_one() : true;
        const uint32_t tmem = uni ? __shfl_sync(0xffffffffu, tmem_slot, 0) : tmem_slot;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
                               ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = base, b0 = base + 48 * 1024;
        // operands and accumulators of 4 consecutive K steps, fixed before the loop: the loop body is 4 MMAs and a counter
        // (the first version of this probe selected layouts, rotated addresses and took `i % n_acc` INSIDE the loop and
        // measured its own 150 - 240 cycles of integer work per iteration, profiles/r2_mma_probe_lane0_loop.log)
        uint64_t ad[4], bd[4];
        uint32_t acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t step = rot ? (uint32_t)k : 0u;
            ad[k] = a_mn ? umma_desc_mn_sw128(a0 + step * 2048u, 8u * 1024u, 1024u) : probe_desc_k_sw128(a0) + (uint64_t)(step * 2u);
            bd[k] = b_mn ? umma_desc_mn_sw128(b0 + step * 2048u, 8u * 1024u, 1024u) : probe_desc_k_sw128(b0) + (uint64_t)(step * 2u);
            acc[k] = tmem + (uint32_t)((k % n_acc) * n);
        }
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 4) {
            const uint32_t on = i ? 1u : 0u;
            if (leader) {
                wb_mma_f16(acc[0], ad[0], bd[0], idesc, on);
                wb_mma_f16(acc[1], ad[1], bd[1], idesc, n_acc > 1 ? on : 1u);
                wb_mma_f16(acc[2], ad[2], bd[2], idesc, n_acc > 2 ? on : 1u);
                wb_mma_f16(acc[3], ad[3], bd[3], idesc, n_acc > 3 ? on : 1u);
            }
        }
        const long long t1 = clock64();
        if (leader) tc_commit(&bar);
        mb_wait(&bar, 0);
        const long long t2 = clock64();
        if (leader) {
            out[2 * blockIdx.x] = t1 - t0;          // issue loop
            out[2 * blockIdx.x + 1] = t2 - t0;      // until the last MMA retired
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    }
}
int mma_probe(int a_mn, int b_mn, int n, int n_acc, int rot, int iters, int uni, int ctas, long long* out_dev, cudaStream_t st) {
    MS_REQUIRE(n >= 16 && n <= 256 && (n & 15) == 0 && (n_acc == 1 || n_acc == 2 || n_acc == 4) && n_acc * n <= 512 && ctas >= 1 && (iters & 3) == 0,
               "mma_probe: bad arguments");
    static bool init = false;
    if (!init) {
        MS_CHECK_CUDA(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        init = true;
    }
    mma_probe_kernel<<<ctas, 128, 97 * 1024 + 1024, st>>>(a_mn, b_mn, n, n_acc, rot, iters, uni, out_dev);
    return check_launch("mma_probe");
}

// one-shot convenience (operator-level C ABI / tests): splits x and dy into planes, runs the kernel.
//   scratch layout (bytes): [x hi | x lo | dy hi | dy lo | partial sums]
size_t wgrad_bf_oneshot_scratch_bytes(const ConvWgrad& q) {
    const size_t xe = q.x.pixels() * ((q.x.c + 7) / 8 * 8), de = q.dy.pixels() * ((q.dy.c + 7) / 8 * 8);
    return 2 * (xe * 2 + 256) + 2 * (de * 2 + 256) + wgrad_bf_workspace_floats(q.kh, q.kw, q.x.c, q.dy.c) * 4 + 1024;
}

int wgrad_bf_oneshot(const ConvWgrad& q0, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    MS_REQUIRE(wgrad_bf_supported(q0), "wgrad_bf: unsupported geometry");
    MS_REQUIRE(scratch_bytes >= wgrad_bf_oneshot_scratch_bytes(q0), "wgrad_bf: scratch too small");
    MS_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "wgrad_bf: scratch must be 256B aligned");
    unsigned char* b = reinterpret_cast<unsigned char*>(scratch);
    auto take = [&](size_t bytes) { unsigned char* r = b; b += (bytes + 255) / 256 * 256; return r; };
    ActPlanes xp, dp;
    xp.cs = (q0.x.c + 7) / 8 * 8; dp.cs = (q0.dy.c + 7) / 8 * 8; xp.fmt = 0; dp.fmt = 0; xp.scale = dp.scale = 1.f;
    const size_t xe = q0.x.pixels() * xp.cs, de = q0.dy.pixels() * dp.cs;
    xp.hi = take(xe * 2); xp.lo = take(xe * 2); dp.hi = take(de * 2); dp.lo = take(de * 2);
    ConvWgrad q = q0;
    q.workspace_floats = wgrad_bf_workspace_floats(q.kh, q.kw, q.x.c, q.dy.c);
    q.workspace = reinterpret_cast<float*>(take(q.workspace_floats * 4));
    if (split_planes(q.x, xp, st)) return -1;
    if (split_planes(q.dy, dp, st)) return -1;
    return wgrad_bf(q, xp, dp, st);
}

}  // namespace ms
