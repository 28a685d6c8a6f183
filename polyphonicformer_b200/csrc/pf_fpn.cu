// SemanticFPNWrapper.forward up to feature_add_all_level (polyphonic/funcs/semantic_fpn.py:198-219 of the reference, in the
// shipped configuration configs/_base_/models/polyphonic_former.py:78-96): seven [3x3 conv + GroupNorm32 + ReLU] modules
// over the four FPN levels, bilinear x2 steps between them, the sine positional encoding added to the coarsest level, and
// the sum of the four level outputs.  (conv_pred / aux_convs, :221-229, are pf_fpn_pred.)
//
//   level 0 (stride 4):   conv0 with stride 2                                    -> H x W
//   level 1 (stride 8):   conv0                                                  -> H x W
//   level 2 (stride 16):  conv0, x2, conv1                                       -> H x W
//   level 3 (stride 32):  + positional encoding, conv0, x2, conv1, x2, conv2     -> H x W        (H x W = the decoder map)
//
// Activations are channels-last bf16 hi / lo planes over a zero-padded grid, one image = R = round_up((h+2)(w+2), 128) rows of
// 256 channels, so a 3x3 tap is a row shift of dy (w + 2) + dx and the convolution is nine shifted GEMMs on the tensor cores
// (pf_sgemm.cuh, no im2col buffer); the stride-2 convolution reads the four parity phases of its input as four planes.
// GroupNorm spans the whole image, so the conv epilogue writes the raw fp32 output plus per-tile group statistics, a tiny
// kernel merges them in fp64 into one (scale, shift) per channel, and normalisation + ReLU happen inside the kernel that
// consumes the map: the x2 up-sampler that writes the next convolution's planes, or the final four-level sum.
#include <math.h>
#include <stdlib.h>

#include "pf_internal.h"
#include "pf_sgemm.cuh"

#ifndef PF_CONV_CLUSTER_DEFAULT
#define PF_CONV_CLUSTER_DEFAULT 0
#endif
#include "pf_sm100.cuh"

namespace pf {

constexpr int FPN_CONVS = 7;
enum { CV_L0 = 0, CV_L1, CV_L2A, CV_L2B, CV_L3A, CV_L3B, CV_L3C };

static inline int fpn_rows(int h, int w) { return ((h + 2) * (w + 2) + 127) / 128 * 128; }

__device__ __forceinline__ void store_split(uint16_t* hi, uint16_t* lo, size_t idx, float v) {
    const float h = bf16_round(v);
    hi[idx] = (uint16_t)(__float_as_uint(h) >> 16);
    lo[idx] = (uint16_t)(__float_as_uint(bf16_round(v - h)) >> 16);
}

// mmdet SinePositionalEncoding(num_feats = 128, normalize = True, temperature 1e4, scale 2 pi, eps 1e-6, offset 0) for an
// all-valid mask (mmdet/models/utils/positional_encoding.py:57-92): channels [0, 128) encode y, [128, 256) encode x
__device__ __forceinline__ float sine_posenc(int c, int y, int x, int h, int w) {
    const int i = c & 127;
    const float embed = c < 128 ? (float)(y + 1) / ((float)h + 1e-6f) : (float)(x + 1) / ((float)w + 1e-6f);
    const float dim_t = powf(10000.f, (float)(2 * (i >> 1)) / 128.f);
    const float v = embed * 6.283185307179586f / dim_t;
    return (i & 1) ? cosf(v) : sinf(v);
}

// NCHW fp32 [B][256][hs][ws] -> channels-last padded bf16 planes.  PHASES = 1: grid (h + 2) x (w + 2) with h = hs, w = ws.
// PHASES = 4: the four parity phases (py, px) of the map as four planes of grid (hs/2 + 2) x (ws/2 + 2).
// Block = one padded row of one plane group; 256 threads = channels on the store side.
template <int PHASES>
__global__ void __launch_bounds__(256) fpn_pack_kernel(const float* __restrict__ src, int hs, int ws, int R, int add_posenc,
                                                       uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    constexpr int S = PHASES == 4 ? 2 : 1;
    constexpr int SC = 128, CPP = SC / S;              // source columns / padded columns per pass
    const int h = hs / S, w = ws / S, pitch = w + 2;
    // blockIdx.z = channel quarter (64 channels) [+ 4 * row parity for PHASES = 4]
    const int y1 = blockIdx.x, b = blockIdx.y, cq = blockIdx.z & 3, py = blockIdx.z >> 2;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    __shared__ float tile[SC][65];                     // [source column][channel of the quarter]
    const bool ring_row = y1 == 0 || y1 == h + 1;
    const int ysrc = (y1 - 1) * S + py;
    for (int x0 = 0; x0 < pitch; x0 += CPP) {
        __syncthreads();
        if (!ring_row) {
            // load: warp = channel (8 per round), lane = 4 source columns 32 apart: 4 x 128 contiguous bytes per channel row
            const int xs0 = (x0 - 1) * S + lane;       // source column of padded column x0, phase 0, + lane
#pragma unroll 2
            for (int cc = warp; cc < 64; cc += 8) {
                const int c = cq * 64 + cc;
                const float* p = src + (((size_t)b * 256 + c) * hs + ysrc) * ws;
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int xs = xs0 + 32 * k;
                    v[k] = (xs >= 0 && xs < ws) ? __ldg(p + xs) : 0.f;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int xs = xs0 + 32 * k;
                    if (add_posenc && xs >= 0 && xs < ws) v[k] += sine_posenc(c, ysrc, xs, hs, ws);
                    tile[lane + 32 * k][cc] = v[k];
                }
            }
        }
        __syncthreads();
        // store: warp = padded column (x S phases), lane = 2 channels -> 128 contiguous bytes per warp store and plane
        for (int jj = warp; jj < SC; jj += 8) {
            const int j = jj / S, px = jj - j * S, x1 = x0 + j;
            if (x1 >= pitch) continue;
            const bool ring = ring_row || x1 == 0 || x1 == w + 1;
            const float a0 = ring ? 0.f : tile[jj][2 * lane], a1 = ring ? 0.f : tile[jj][2 * lane + 1];
            const float h0 = bf16_round(a0), h1 = bf16_round(a1);
            const size_t plane = (size_t)b * PHASES + (PHASES == 4 ? py * 2 + px : 0);
            const size_t idx = ((plane * R + (size_t)y1 * pitch + x1) * 256 + cq * 64 + 2 * lane) * 2;      // bytes
            *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(hi) + idx) = pack_bf16x2(h0, h1);
            *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(lo) + idx) = pack_bf16x2(a0 - h0, a1 - h1);
        }
    }
}

// per-tile (sum, sum of squares) of every group -> (scale, shift) per channel of every image.  One CTA per image.
__global__ void __launch_bounds__(256) fpn_gn_finalize_kernel(const float2* __restrict__ stats, int tiles_per_img, int count,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float eps, float2* __restrict__ affine) {
    pdl_wait();
    __shared__ double s_s[8][32], s_ss[8][32];
    __shared__ float s_mean[32], s_rstd[32];
    const int b = blockIdx.x, t = threadIdx.x, g = t & 31, part = t >> 5;
    double s = 0, ss = 0;
    for (int i = part; i < tiles_per_img; i += 8) {     // fixed assignment and order: deterministic
        const float2 v = __ldg(stats + ((size_t)b * tiles_per_img + i) * 32 + g);
        s += v.x, ss += v.y;
    }
    s_s[part][g] = s, s_ss[part][g] = ss;
    __syncthreads();
    if (t < 32) {
        s = ss = 0;
        for (int k = 0; k < 8; ++k) s += s_s[k][t], ss += s_ss[k][t];
        const double mean = s / count, var = fmax(ss / count - mean * mean, 0.0);
        s_mean[t] = (float)mean, s_rstd[t] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    const float sc = __ldg(gamma + t) * s_rstd[t >> 3];
    affine[b * 256 + t] = make_float2(sc, __ldg(beta + t) - s_mean[t >> 3] * sc);
}

__device__ __forceinline__ float4 affine_relu4(float4 v, const float2* af) {
    return make_float4(fmaxf(fmaf(v.x, af[0].x, af[0].y), 0.f), fmaxf(fmaf(v.y, af[1].x, af[1].y), 0.f),
                       fmaxf(fmaf(v.z, af[2].x, af[2].y), 0.f), fmaxf(fmaf(v.w, af[3].x, af[3].y), 0.f));
}

// ReLU(GroupNorm(raw)) at grid (h, w), bilinear x2 (align_corners = False, semantic_fpn.py:125-129) -> the padded planes of
// the next convolution at grid (2h, 2w).  Block = one padded output row of one image; warp = output pixel, lane = 8 channels.
__global__ void __launch_bounds__(256) fpn_apply_up2_kernel(const float* __restrict__ raw, const float2* __restrict__ affine, int h,
                                                            int w, int Rin, int Rout, uint16_t* __restrict__ hi,
                                                            uint16_t* __restrict__ lo) {
    pdl_wait();
    const int Y1 = blockIdx.x, b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H2 = 2 * h, W2 = 2 * w, pin = w + 2, pout = W2 + 2;
    float2 af[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) af[i] = __ldg(affine + b * 256 + lane * 8 + i);
    const bool ring_row = Y1 == 0 || Y1 == H2 + 1;
    const int Y = Y1 - 1, y = Y >> 1;
    const int ya = (Y & 1) ? y : max(y - 1, 0), yb = (Y & 1) ? min(y + 1, h - 1) : y;
    const float wya = (Y & 1) ? 0.75f : 0.25f, wyb = 1.f - wya;      // rows ya, yb (at Y = 0 / 2h - 1 they coincide)
    for (int X1 = warp; X1 < pout; X1 += 8) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 0.f;
        if (!ring_row && X1 >= 1 && X1 <= W2) {
            const int X = X1 - 1, x = X >> 1;
            const int xa = (X & 1) ? x : max(x - 1, 0), xb = (X & 1) ? min(x + 1, w - 1) : x;
            const float wxa = (X & 1) ? 0.75f : 0.25f, wxb = 1.f - wxa;
            const float* base = raw + ((size_t)b * Rin) * 256 + lane * 8;
            const float* paa = base + ((size_t)(ya + 1) * pin + xa + 1) * 256;
            const float* pab = base + ((size_t)(ya + 1) * pin + xb + 1) * 256;
            const float* pba = base + ((size_t)(yb + 1) * pin + xa + 1) * 256;
            const float* pbb = base + ((size_t)(yb + 1) * pin + xb + 1) * 256;
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
                const float4 vaa = affine_relu4(__ldg(reinterpret_cast<const float4*>(paa) + hlf), af + 4 * hlf);
                const float4 vab = affine_relu4(__ldg(reinterpret_cast<const float4*>(pab) + hlf), af + 4 * hlf);
                const float4 vba = affine_relu4(__ldg(reinterpret_cast<const float4*>(pba) + hlf), af + 4 * hlf);
                const float4 vbb = affine_relu4(__ldg(reinterpret_cast<const float4*>(pbb) + hlf), af + 4 * hlf);
                o[4 * hlf + 0] = wya * (wxa * vaa.x + wxb * vab.x) + wyb * (wxa * vba.x + wxb * vbb.x);
                o[4 * hlf + 1] = wya * (wxa * vaa.y + wxb * vab.y) + wyb * (wxa * vba.y + wxb * vbb.y);
                o[4 * hlf + 2] = wya * (wxa * vaa.z + wxb * vab.z) + wyb * (wxa * vba.z + wxb * vbb.z);
                o[4 * hlf + 3] = wya * (wxa * vaa.w + wxb * vab.w) + wyb * (wxa * vba.w + wxb * vbb.w);
            }
        }
        uint32_t ph[4], pl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float a0 = o[2 * i], a1 = o[2 * i + 1], h0 = bf16_round(a0), h1 = bf16_round(a1);
            ph[i] = pack_bf16x2(h0, h1), pl[i] = pack_bf16x2(a0 - h0, a1 - h1);
        }
        const size_t idx = (((size_t)b * Rout + (size_t)Y1 * pout + X1) * 256 + lane * 8) * 2;   // bytes
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(hi) + idx) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(lo) + idx) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
}

// feature_add_all_level = l0 + l1 + l2 + l3 (semantic_fpn.py:216-219), each ReLU(GroupNorm(raw_l)), written channel-major:
// bf16 [B][256][HWp] (the input of pf_fpn_pred) and optionally fp32 [B][256][HW].  Block = 32 pixels of one image.
struct FpnSumArgs {
    const float* raw[4];
    const float2* affine[4];
    int h, w, R, HWp;
    uint16_t* fused;
    float* fused32;
};
__global__ void __launch_bounds__(256) fpn_sum_kernel(const __grid_constant__ FpnSumArgs a) {
    pdl_wait();
    __shared__ float tile[32][257];
    const int b = blockIdx.y, p0 = blockIdx.x * 32, t = threadIdx.x, HW = a.h * a.w, pitch = a.w + 2;
    float2 af[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) af[l] = __ldg(a.affine[l] + b * 256 + t);
#pragma unroll 1
    for (int j0 = 0; j0 < 32; j0 += 8) {
        float r[8][4];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {               // 32 independent loads in flight per thread
            const int p = p0 + j0 + jj;
            const int y = p / a.w, x = p - y * a.w;
            const size_t idx = ((size_t)b * a.R + (size_t)(y + 1) * pitch + x + 1) * 256 + t;
#pragma unroll
            for (int l = 0; l < 4; ++l) r[jj][l] = p < HW ? __ldg(a.raw[l] + idx) : 0.f;
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            float v = 0.f;
#pragma unroll
            for (int l = 0; l < 4; ++l) v += fmaxf(fmaf(r[jj][l], af[l].x, af[l].y), 0.f);   // ((l0 + l1) + l2) + l3
            tile[j0 + jj][t] = p0 + j0 + jj < HW ? v : 0.f;
        }
    }
    __syncthreads();
    const int lane = t & 31, warp = t >> 5;
    for (int c = warp; c < 256; c += 8) {
        const int p = p0 + lane;
        const float v = tile[lane][c];
        if (p < HW) {
            a.fused[((size_t)b * 256 + c) * a.HWp + p] = (uint16_t)(__float_as_uint(bf16_round(v)) >> 16);
            if (a.fused32) a.fused32[((size_t)b * 256 + c) * HW + p] = v;
        } else if (p < a.HWp) {
            a.fused[((size_t)b * 256 + c) * a.HWp + p] = 0;
        }
    }
}

struct FpnScratch {
    uint16_t* planes[2];      // [hi, lo] of the largest activation (level 0 phases); reused by every convolution input
    float* raw[FPN_CONVS];
    float2* stats;            // [B * tiles][32], reused
    float2* affine[FPN_CONVS];
    size_t total;
};
static size_t al256f(size_t v) { return (v + 255) / 256 * 256; }
static FpnScratch carve_fpn(void* base, int B, int H, int W) {
    FpnScratch s;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? static_cast<char*>(base) + off : nullptr;
        off = al256f(off + bytes);
        return p;
    };
    const size_t Rfull = fpn_rows(H, W), Rhalf = fpn_rows(H / 2, W / 2), Rq = fpn_rows(H / 4, W / 4);
    for (int i = 0; i < 2; ++i) s.planes[i] = static_cast<uint16_t*>(take((size_t)B * 4 * Rfull * 256 * 2));
    const size_t rr[FPN_CONVS] = {Rfull, Rfull, Rhalf, Rfull, Rq, Rhalf, Rfull};
    for (int i = 0; i < FPN_CONVS; ++i) {
        s.raw[i] = static_cast<float*>(take((size_t)B * rr[i] * 256 * 4));
        s.affine[i] = static_cast<float2*>(take((size_t)B * 256 * sizeof(float2)));
    }
    s.stats = static_cast<float2*>(take((size_t)B * (Rfull / 128) * 32 * sizeof(float2)));
    s.total = off;
    return s;
}

// weight-multicast cluster size of the convolution kernel: 0 = plain persistent kernel; PF_CONV_CLUSTER overrides
static int conv_cluster() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PF_CONV_CLUSTER");
        v = e ? atoi(e) : PF_CONV_CLUSTER_DEFAULT;
        if (v != 2 && v != 4) v = 0;
    }
    return v;
}

template <int CL>
static int launch_conv_mc(const CUtensorMap& ah, const CUtensorMap& al, const uint16_t* wbase, const SgArgs& a, int n_mtiles,
                          cudaStream_t st) {
    CUtensorMap ws;
    if (int e = make_tmap_bf16_2d(&ws, wbase, 2 * 9 * 256, 256, 256, SC_TN / CL, SG_KC)) return e;
    auto kern = sgemm_conv256_mc_kernel<CL>;
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM);
    if (ce != cudaSuccess) return set_error(PF_ERR_CUDA, "sgemm smem attribute: %s", cudaGetErrorString(ce));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.blockDim = dim3(SG_THREADS), cfg.dynamicSmemBytes = SC_SMEM, cfg.stream = st, cfg.attrs = attr, cfg.numAttrs = 2;
    cfg.gridDim = dim3(CL);
    static int max_clusters[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && max_clusters[dev] == 0) {
        int n = 0;
        cfg.gridDim = dim3(num_sms() / CL * CL);
        ce = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
        if (ce != cudaSuccess || n <= 0) return set_error(PF_ERR_CUDA, "cudaOccupancyMaxActiveClusters: %s", cudaGetErrorString(ce));
        max_clusters[dev] = n;
    }
    int n_clusters = (n_mtiles + CL - 1) / CL;
    if (dev >= 0 && dev < 64 && n_clusters > max_clusters[dev]) n_clusters = max_clusters[dev];
    cfg.gridDim = dim3(n_clusters * CL);
    ce = cudaLaunchKernelEx(&cfg, kern, ah, al, ws, a, n_mtiles);
    if (ce != cudaSuccess) return set_error(PF_ERR_CUDA, "sgemm_conv256_mc_kernel launch: %s", cudaGetErrorString(ce));
    count_launch();
    return PF_OK;
}

// one [3x3 conv + GN statistics] over planes of grid (h, w): raw fp32 + (scale, shift)
static int fpn_conv(const pf_fpn_weights* w, int cv, const FpnScratch& sc, int B, int h, int wd, bool stride2, cudaStream_t st) {
    const int R = fpn_rows(h, wd), pitch = wd + 2, planes = stride2 ? 4 : 1;
    CUtensorMap ah, al, wm;
    const uint64_t adims[3] = {256, (uint64_t)R, (uint64_t)B * planes}, astr[2] = {512, (uint64_t)R * 512};
    const uint32_t abox[3] = {SG_KC, 128, 1};
    if (int e = make_tmap_bf16_nd(&ah, sc.planes[0], 3, adims, astr, abox)) return e;
    if (int e = make_tmap_bf16_nd(&al, sc.planes[1], 3, adims, astr, abox)) return e;
    if (int e = make_tmap_bf16_2d(&wm, w->conv_w + (size_t)cv * 2 * 9 * 256 * 256, 2 * 9 * 256, 256, 256, SC_TN, SG_KC)) return e;
    SgArgs a = {};
    a.mode = SG_CONV, a.n_kb = 9 * 4, a.cin_blocks = 4, a.w_tap_rows = 256, a.w_lo = 9 * 256;
    a.rows_per_img = R, a.planes_per_img = planes;
    for (int tp = 0; tp < 9; ++tp) {
        const int dy = tp / 3 - 1, dx = tp % 3 - 1;
        if (stride2) {   // source (2y + dy, 2x + dx): phase (dy & 1, dx & 1) at (y + (dy < 0 ? -1 : 0), x + (dx < 0 ? -1 : 0))
            a.plane[tp] = (dy & 1) * 2 + (dx & 1);
            a.shift[tp] = (dy < 0 ? -pitch : 0) + (dx < 0 ? -1 : 0);
        } else {
            a.plane[tp] = 0, a.shift[tp] = dy * pitch + dx;
        }
    }
    a.raw = sc.raw[cv], a.stats = sc.stats, a.grid_h = h, a.grid_w = wd;
    const int tiles = R / 128;
    cudaError_t ce = cudaFuncSetAttribute(sgemm_conv256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM);
    if (ce != cudaSuccess) return set_error(PF_ERR_CUDA, "sgemm smem attribute: %s", cudaGetErrorString(ce));
    const int n_mtiles = B * tiles;
    const uint16_t* wbase = w->conv_w + (size_t)cv * 2 * 9 * 256 * 256;
    const int cl = n_mtiles >= 16 ? conv_cluster() : 0;
    if (cl == 4) {
        if (int e = launch_conv_mc<4>(ah, al, wbase, a, n_mtiles, st)) return e;
    } else if (cl == 2) {
        if (int e = launch_conv_mc<2>(ah, al, wbase, a, n_mtiles, st)) return e;
    } else if (int e = launch_pdl("sgemm_conv256_kernel", sgemm_conv256_kernel, dim3(n_mtiles < num_sms() ? n_mtiles : num_sms()),
                                  dim3(SG_THREADS), SC_SMEM, st, ah, al, wm, a, n_mtiles))
        return e;
    return launch_pdl("fpn_gn_finalize_kernel", fpn_gn_finalize_kernel, dim3(B), dim3(256), 0, st, (const float2*)sc.stats, tiles,
                      h * wd * 8, w->gn_gamma + cv * 256, w->gn_beta + cv * 256, w->gn_eps, sc.affine[cv]);
}

static int fpn_up2(const FpnScratch& sc, int cv, int B, int h, int wd, cudaStream_t st) {
    return launch_pdl("fpn_apply_up2_kernel", fpn_apply_up2_kernel, dim3(2 * h + 2, B), dim3(256), 0, st, (const float*)sc.raw[cv],
                      (const float2*)sc.affine[cv], h, wd, fpn_rows(h, wd), fpn_rows(2 * h, 2 * wd), sc.planes[0], sc.planes[1]);
}

}  // namespace pf

extern "C" size_t pf_semantic_fpn_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0 || H % 4 || W % 4) return 0;
    return pf::carve_fpn(nullptr, B, H, W).total;
}

extern "C" int pf_semantic_fpn(const pf_fpn_weights* w, const float* p0, const float* p1, const float* p2, const float* p3,
                               uint16_t* fused, float* fused32, void* workspace, size_t workspace_bytes, int B, int H, int W,
                               int HWp, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    reset_launch_count();
    PF_REQUIRE(w && p0 && p1 && p2 && p3 && fused && workspace, PF_ERR_ARG, "pf_semantic_fpn: null pointer");
    PF_REQUIRE(B > 0 && H > 0 && W > 0 && H % 4 == 0 && W % 4 == 0, PF_ERR_ARG,
               "pf_semantic_fpn: the decoder map %dx%d must be a multiple of 4 (levels at x2, x1, /2, /4)", H, W);
    PF_REQUIRE(HWp >= H * W && HWp % 8 == 0, PF_ERR_ALIGN, "pf_semantic_fpn: HWp=%d", HWp);
    PF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PF_ERR_ALIGN, "pf_semantic_fpn: workspace not 256-byte aligned");
    const FpnScratch sc = carve_fpn(workspace, B, H, W);
    PF_REQUIRE(workspace_bytes >= sc.total, PF_ERR_WORKSPACE, "pf_semantic_fpn: workspace %zu < %zu", workspace_bytes, sc.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int h2 = H / 2, w2 = W / 2, h4 = H / 4, w4 = W / 4;
    auto pack1 = [&](const float* src, int h, int wd, int posenc) {
        fpn_pack_kernel<1><<<dim3(h + 2, B, 4), 256, 0, st>>>(src, h, wd, fpn_rows(h, wd), posenc, sc.planes[0], sc.planes[1]);
    };
    // level 0: one stride-2 convolution on the four parity phases (semantic_fpn.py:92-104)
    fpn_pack_kernel<4><<<dim3(H + 2, B, 8), 256, 0, st>>>(p0, 2 * H, 2 * W, fpn_rows(H, W), 0, sc.planes[0], sc.planes[1]);
    PF_CHECK_LAUNCH("fpn_pack_kernel<4>");
    if (int e = fpn_conv(w, CV_L0, sc, B, H, W, true, st)) return e;
    // level 1
    pack1(p1, H, W, 0);
    PF_CHECK_LAUNCH("fpn_pack_kernel<1>");
    if (int e = fpn_conv(w, CV_L1, sc, B, H, W, false, st)) return e;
    // level 2: conv, x2, conv
    pack1(p2, h2, w2, 0);
    PF_CHECK_LAUNCH("fpn_pack_kernel<1>");
    if (int e = fpn_conv(w, CV_L2A, sc, B, h2, w2, false, st)) return e;
    if (int e = fpn_up2(sc, CV_L2A, B, h2, w2, st)) return e;
    if (int e = fpn_conv(w, CV_L2B, sc, B, H, W, false, st)) return e;
    // level 3: + positional encoding, conv, x2, conv, x2, conv
    pack1(p3, h4, w4, 1);
    PF_CHECK_LAUNCH("fpn_pack_kernel<1>");
    if (int e = fpn_conv(w, CV_L3A, sc, B, h4, w4, false, st)) return e;
    if (int e = fpn_up2(sc, CV_L3A, B, h4, w4, st)) return e;
    if (int e = fpn_conv(w, CV_L3B, sc, B, h2, w2, false, st)) return e;
    if (int e = fpn_up2(sc, CV_L3B, B, h2, w2, st)) return e;
    if (int e = fpn_conv(w, CV_L3C, sc, B, H, W, false, st)) return e;
    FpnSumArgs sa;
    const int last[4] = {CV_L0, CV_L1, CV_L2B, CV_L3C};
    for (int l = 0; l < 4; ++l) sa.raw[l] = sc.raw[last[l]], sa.affine[l] = sc.affine[last[l]];
    sa.h = H, sa.w = W, sa.R = fpn_rows(H, W), sa.HWp = HWp, sa.fused = fused, sa.fused32 = fused32;
    return launch_pdl("fpn_sum_kernel", fpn_sum_kernel, dim3((HWp + 31) / 32, B), dim3(256), 0, st, sa);
}
