// "Shifted-row" split-bf16 GEMM on the 5th-generation tensor cores: the building block of every convolution / wide FC
// outside the decoder's stage loop (the tracking head's 3x3 convs and FC, SemanticFPN's 3x3 convs).
//
//   D[128 rows][128 cols] (fp32, tensor memory) = sum over k-blocks of  A_kb[128][64] * W_kb[128][64]^T
//   with fp32-level accuracy from bf16 operands:  A = Ah + Al, W = Wh + Wl,  D += Al*Wh + Ah*Wl + Ah*Wh
//
// A 3x3 convolution over a zero-padded row-major grid is nine such GEMMs whose A tiles are the SAME activation rows
// shifted by dy * pitch + dx: the producer warp just adds the tap's row shift to the TMA coordinate (rows outside the
// tensor, negative ones included, read as zero), so there is no im2col buffer.  SG_FC walks (position, channel block)
// pairs of a [item][position][channel] activation instead (the flatten + Linear after the convs).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (TMEM lane quarter = warp % 4).
#pragma once
#include "pf_internal.h"
#include "pf_sm100.cuh"

namespace pf {

constexpr int SG_THREADS = 192;
constexpr int SG_TN = 128;                     // output columns per CTA
constexpr int SG_KC = 64;                      // K per ring stage
constexpr int SG_NSTG = 3;
constexpr int SG_PLANE = 128 * SG_KC * 2;      // one [128][64] bf16 box
constexpr int SG_STAGE = 4 * SG_PLANE;         // A hi | A lo | W hi | W lo
constexpr int SG_BAR_OFF = SG_NSTG * SG_STAGE;
constexpr int SG_STAT_OFF = SG_BAR_OFF + 128;  // [2 passes][4 warps][16 groups] floats
constexpr int SG_SMEM_USED = SG_STAT_OFF + 2 * 4 * 16 * 4;
constexpr int SG_SMEM = SG_SMEM_USED + 1024;   // slack for the 1024-byte alignment of the ring

enum { SG_CONV = 0, SG_FC = 1 };
enum { SG_EPI_GN64 = 0, SG_EPI_PARTIAL = 1, SG_EPI_RAWSTATS = 2 };

struct SgArgs {
    int mode;              // SG_CONV / SG_FC
    int n_kb;              // k-blocks per CTA
    int cin_blocks;        // SG_CONV: input channels / 64 (k-block kb -> tap kb / cin_blocks, channel block kb % cin_blocks)
    int w_tap_rows;        // SG_CONV: rows of one tap in the weight map (= padded output channels)
    int w_lo;              // rows from the hi plane to the lo plane in the weight map
    int fc_pos_per_split;  // SG_FC: grid positions per split (blockIdx.z)
    int fc_grid, fc_pitch; // SG_FC: valid positions are (y, x) with y, x < fc_grid at row y * fc_pitch + x of an item
    int shift[9];          // SG_CONV: row shift of tap t
    int plane[9];          // SG_CONV: plane (third tensor coordinate) of tap t inside an image: the four parity phases of a
                           //          stride-2 convolution; 0 otherwise
    int rows_per_img;      // SG_CONV: rows of one image (a multiple of 128): tile row m0 -> (image m0 / rows_per_img, row m0 %
    int planes_per_img;    //          rows_per_img); the A tensor is [images * planes_per_img][rows_per_img][channels]
    // ---- epilogue
    int n_items;           // GN64: items (RoIs) that exist; rows of later items are written as zeros
    const float *gamma, *beta;   // GN64: [256] of this layer
    float eps;
    uint16_t *out_hi, *out_lo;   // GN64: bf16 planes [rows][256]
    float* partial;              // PARTIAL: fp32 [split][part_rows][part_ld]
    int part_rows, part_ld;
    // RAWSTATS: a convolution over a zero-padded (grid_h + 2) x (grid_w + 2) grid per image whose GroupNorm spans the whole
    // image: raw fp32 output rows [row][256] for the interior positions + per-tile (sum, sum of squares) of every group of 8
    // channels over them
    float* raw;
    float2* stats;               // [gridDim.y][32 groups]
    int grid_h, grid_w;
};

__device__ __forceinline__ void sg_named_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ float sg_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void sg_ld32(uint32_t taddr, float (&y)[32]) {
    uint32_t v[32];
    tmem_ld32(taddr, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(v[i]);
}

// 32 fp32 values -> bf16 hi / lo (x = hi + lo to 2^-17) -> two 64-byte row segments
__device__ __forceinline__ void sg_store_split32(uint16_t* hi, uint16_t* lo, const float (&y)[32]) {
    uint32_t h[16], l[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float a = y[2 * i], b = y[2 * i + 1];
        const float ah = bf16_round(a), bh = bf16_round(b);
        h[i] = pack_bf16x2(ah, bh);
        l[i] = pack_bf16x2(a - ah, b - bh);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        reinterpret_cast<uint4*>(hi)[i] = make_uint4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
        reinterpret_cast<uint4*>(lo)[i] = make_uint4(l[4 * i], l[4 * i + 1], l[4 * i + 2], l[4 * i + 3]);
    }
}

template <int EPI>
__global__ void __launch_bounds__(SG_THREADS, 1)
sgemm_kernel(const __grid_constant__ CUtensorMap tmap_ah, const __grid_constant__ CUtensorMap tmap_al,
             const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ SgArgs a) {
    extern __shared__ uint8_t sg_smem_raw[];
    uint8_t* smem = sg_smem_raw + ((1024u - (smem_u32(sg_smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SG_BAR_OFF);
    uint64_t* full = bars;
    uint64_t* empty = bars + SG_NSTG;
    uint64_t* accfull = bars + 2 * SG_NSTG;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SG_NSTG + 1);
    float* s_stat = reinterpret_cast<float*>(smem + SG_STAT_OFF);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * SG_TN, m0 = blockIdx.y * 128, split = blockIdx.z;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_ah);
        tma_prefetch_desc(&tmap_al);
        tma_prefetch_desc(&tmap_w);
        for (int i = 0; i < SG_NSTG; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(accfull, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<128>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);
    pdl_wait();                 // activations written by the previous kernel of the stream are visible from here on
    pdl_launch_dependents();

    if (warp == 0) {
        // ================= TMA producer (whole warp, one elected lane issues) =================
        for (int it = 0; it < a.n_kb; ++it) {
            const int s = it % SG_NSTG;
            if (it >= SG_NSTG) mbar_wait(&empty[s], ((it / SG_NSTG) & 1) ^ 1);
            int ac0, ac1, ac2, wc0, wr;
            if (a.mode == SG_CONV) {
                const int tap = it / a.cin_blocks, cb = it - tap * a.cin_blocks;
                const int img = m0 / a.rows_per_img;
                ac0 = cb * SG_KC, ac1 = m0 - img * a.rows_per_img + a.shift[tap], ac2 = img * a.planes_per_img + a.plane[tap];
                wc0 = cb * SG_KC, wr = tap * a.w_tap_rows + n0;
            } else {
                const int pi = split * a.fc_pos_per_split + (it >> 2), cb = it & 3;
                ac0 = cb * SG_KC, ac1 = (pi / a.fc_grid) * a.fc_pitch + pi % a.fc_grid, ac2 = m0;
                wc0 = pi * 256 + cb * SG_KC, wr = n0;
            }
            uint8_t* st = smem + s * SG_STAGE;
            mbar_arrive_expect_tx_warp(&full[s], SG_STAGE);
            tma_load_3d_warp(st, &tmap_ah, &full[s], ac0, ac1, ac2, kEvictNormal);
            tma_load_3d_warp(st + SG_PLANE, &tmap_al, &full[s], ac0, ac1, ac2, kEvictNormal);
            tma_load_2d_warp(st + 2 * SG_PLANE, &tmap_w, &full[s], wc0, wr, kEvictNormal);
            tma_load_2d_warp(st + 3 * SG_PLANE, &tmap_w, &full[s], wc0, wr + a.w_lo, kEvictNormal);
        }
    } else if (warp == 1) {
        // ================= MMA issuer: D = Al*Wh + Ah*Wl + Ah*Wh =================
        constexpr uint32_t idesc = make_idesc_bf16(128, SG_TN, 0, 0);
        for (int it = 0; it < a.n_kb; ++it) {
            const int s = it % SG_NSTG;
            mbar_wait(&full[s], (it / SG_NSTG) & 1);
            tc_fence_after();
            const uint32_t st = smem_u32(smem + s * SG_STAGE);
            const uint64_t dah = make_smem_desc_sw128(st, 16, 1024), dal = make_smem_desc_sw128(st + SG_PLANE, 16, 1024);
            const uint64_t dwh = make_smem_desc_sw128(st + 2 * SG_PLANE, 16, 1024);
            const uint64_t dwl = make_smem_desc_sw128(st + 3 * SG_PLANE, 16, 1024);
#pragma unroll
            for (int k16 = 0; k16 < SG_KC / 16; ++k16) {
                const uint64_t o = (uint64_t)(k16 * 2);   // K-major: +32 bytes (>> 4) per K = 16 step
                umma_bf16_ss_warp(tmem_base, dal + o, dwh + o, idesc, (it | k16) != 0);
                umma_bf16_ss_warp(tmem_base, dah + o, dwl + o, idesc, 1);
                umma_bf16_ss_warp(tmem_base, dah + o, dwh + o, idesc, 1);
            }
            umma_commit_warp(&empty[s]);
        }
        umma_commit_warp(accfull);
    } else {
        // ================= epilogue: thread = output row (TMEM lane) =================
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        mbar_wait(accfull, 0);
        tc_fence_after();
        if (EPI == SG_EPI_GN64) {
            // GroupNorm(32 groups of 8 channels) over the 7 x 7 valid positions of one item = 64 consecutive rows
            // (7 grid rows of pitch 8, column 7 and rows >= 56 are zero padding), then ReLU, then bf16 hi / lo planes.
            const int item = blockIdx.y * 2 + (r >> 6), p = r & 63;
            const bool valid = item < a.n_items && p < 56 && (p & 7) < 7;
            const float inv_n = 1.f / 392.f;
            float mean[16], rstd[16];
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    float y[32];
                    sg_ld32(trow + ch * 32, y);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float s = 0.f;
                        if (valid) {
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const float d = pass ? y[g * 8 + c] - mean[ch * 4 + g] : y[g * 8 + c];
                                s += pass ? d * d : d;
                            }
                        }
                        s = sg_warp_sum(s);
                        if (lane == 0) s_stat[(pass * 4 + q) * 16 + ch * 4 + g] = s;
                    }
                }
                sg_named_bar();
#pragma unroll
                for (int g = 0; g < 16; ++g) {
                    const float t = (s_stat[(pass * 4 + q) * 16 + g] + s_stat[(pass * 4 + (q ^ 1)) * 16 + g]) * inv_n;
                    if (pass) rstd[g] = rsqrtf(t + a.eps);
                    else mean[g] = t;
                }
            }
            const size_t row = (size_t)blockIdx.y * 128 + r;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float y[32];
                sg_ld32(trow + ch * 32, y);
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int col = n0 + ch * 32 + c;
                    const float v = (y[c] - mean[ch * 4 + (c >> 3)]) * rstd[ch * 4 + (c >> 3)] * __ldg(a.gamma + col) + __ldg(a.beta + col);
                    y[c] = valid ? fmaxf(v, 0.f) : 0.f;
                }
                sg_store_split32(a.out_hi + row * 256 + n0 + ch * 32, a.out_lo + row * 256 + n0 + ch * 32, y);
            }
        } else if (EPI == SG_EPI_RAWSTATS) {
            const int img = m0 / a.rows_per_img, qrow = m0 - img * a.rows_per_img + r, pitch = a.grid_w + 2;
            const int y1 = qrow / pitch, x1 = qrow - y1 * pitch;
            const bool valid = y1 >= 1 && y1 <= a.grid_h && x1 >= 1 && x1 <= a.grid_w;
            float* dst = a.raw + ((size_t)m0 + r) * 256 + n0;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float y[32];
                sg_ld32(trow + ch * 32, y);
                if (valid) {
#pragma unroll
                    for (int c = 0; c < 32; c += 4)
                        *reinterpret_cast<float4*>(dst + ch * 32 + c) = make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]);
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float s = 0.f, ss = 0.f;
                    if (valid) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) s += y[g * 8 + c], ss += y[g * 8 + c] * y[g * 8 + c];
                    }
                    s = sg_warp_sum(s), ss = sg_warp_sum(ss);
                    if (lane == 0) s_stat[q * 16 + ch * 4 + g] = s, s_stat[64 + q * 16 + ch * 4 + g] = ss;
                }
            }
            sg_named_bar();
            if (r < 16) {
                const float s = (s_stat[r] + s_stat[16 + r]) + (s_stat[32 + r] + s_stat[48 + r]);
                const float ss = (s_stat[64 + r] + s_stat[80 + r]) + (s_stat[96 + r] + s_stat[112 + r]);
                a.stats[(size_t)blockIdx.y * 32 + blockIdx.x * 16 + r] = make_float2(s, ss);
            }
        } else if (EPI == SG_EPI_PARTIAL) {
            const bool ok = m0 + r < a.part_rows;
            float* dst = a.partial + ((size_t)split * a.part_rows + m0 + r) * a.part_ld + n0;
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                float y[32];
                sg_ld32(trow + ch * 32, y);     // warp-collective: every lane loads, only rows that exist store
                if (ok) {
#pragma unroll
                    for (int c = 0; c < 32; c += 4)
                        *reinterpret_cast<float4*>(dst + ch * 32 + c) = make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]);
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<128>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// The same shifted-row convolution for image-sized problems (SemanticFPN): PERSISTENT CTAs over [128 rows][256 cols] tiles
// (all output channels of 128 padded positions), two accumulators in tensor memory (2 x 256 columns) so that the epilogue
// of tile i -- 128 KB of raw fp32 rows + the GroupNorm partial statistics -- runs under the main loop of tile i + 1.
// What bounds it (ncu, profiles/r2_ncu_full_neck.raw.csv: 0.33 ms per 128x256-map convolution at B = 4): the TENSOR PIPE,
// 86 % active -- 477 GFLOP of issued MMAs (1052 tiles x 36 k-blocks x 12 MMAs of M128 N256 K16) in 334 us = 1.43 PFLOP/s,
// the measured sustained bf16 rate of the part (MEASURED_PEAKS.json: 1.40).  The algorithmic rate is a third of that: the
// 3-MMA split is the price of fp32-level accuracy from bf16 tensor cores.  So cutting operand traffic cannot help -- the
// clustered variant below (weight tiles TMA-multicast across 2 or 4 CTAs) runs at exactly the same speed -- only fewer MMAs
// per product would, and those give up the 5e-6 parity of the pyramid.
constexpr int SC_TN = 256, SC_NSTG = 2;
constexpr int SC_APLANE = 128 * SG_KC * 2, SC_WPLANE = SC_TN * SG_KC * 2;
constexpr int SC_STAGE = 2 * SC_APLANE + 2 * SC_WPLANE;          // 96 KB
constexpr int SC_BAR_OFF = SC_NSTG * SC_STAGE;
constexpr int SC_STAT_OFF = SC_BAR_OFF + 128;                     // [2 (sum, sum of squares)][4 warps][32 groups] floats
constexpr int SC_SMEM = SC_STAT_OFF + 2 * 4 * 32 * 4 + 1024;

static __global__ void __launch_bounds__(SG_THREADS, 1)
sgemm_conv256_kernel(const __grid_constant__ CUtensorMap tmap_ah, const __grid_constant__ CUtensorMap tmap_al,
                     const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ SgArgs a, int n_mtiles) {
    extern __shared__ uint8_t sg_smem_raw[];
    uint8_t* smem = sg_smem_raw + ((1024u - (smem_u32(sg_smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SC_BAR_OFF);
    uint64_t* full = bars;
    uint64_t* empty = bars + SC_NSTG;
    uint64_t* accfull = bars + 2 * SC_NSTG;
    uint64_t* accempty = bars + 2 * SC_NSTG + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SC_NSTG + 4);
    float* s_stat = reinterpret_cast<float*>(smem + SC_STAT_OFF);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_ah);
        tma_prefetch_desc(&tmap_al);
        tma_prefetch_desc(&tmap_w);
        for (int i = 0; i < SC_NSTG; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&accfull[i], 1);
            mbar_init(&accempty[i], 128);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        // ================= TMA producer =================
        int g = 0;
        for (int tile = blockIdx.x; tile < n_mtiles; tile += gridDim.x) {
            const int m0 = tile * 128, img = m0 / a.rows_per_img, q0 = m0 - img * a.rows_per_img;
            for (int it = 0; it < a.n_kb; ++it, ++g) {
                const int s = g % SC_NSTG;
                if (g >= SC_NSTG) mbar_wait(&empty[s], ((g / SC_NSTG) & 1) ^ 1);
                const int tap = it / a.cin_blocks, cb = it - tap * a.cin_blocks;
                const int ac0 = cb * SG_KC, ac1 = q0 + a.shift[tap], ac2 = img * a.planes_per_img + a.plane[tap];
                const int wr = tap * a.w_tap_rows;
                uint8_t* st = smem + s * SC_STAGE;
                mbar_arrive_expect_tx_warp(&full[s], SC_STAGE);
                tma_load_3d_warp(st, &tmap_ah, &full[s], ac0, ac1, ac2, kEvictNormal);
                tma_load_3d_warp(st + SC_APLANE, &tmap_al, &full[s], ac0, ac1, ac2, kEvictNormal);
                tma_load_2d_warp(st + 2 * SC_APLANE, &tmap_w, &full[s], ac0, wr, kEvictLast);
                tma_load_2d_warp(st + 2 * SC_APLANE + SC_WPLANE, &tmap_w, &full[s], ac0, wr + a.w_lo, kEvictLast);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: D = Al*Wh + Ah*Wl + Ah*Wh =================
        constexpr uint32_t idesc = make_idesc_bf16(128, SC_TN, 0, 0);
        int g = 0, li = 0;
        for (int tile = blockIdx.x; tile < n_mtiles; tile += gridDim.x, ++li) {
            const int buf = li & 1;
            mbar_wait(&accempty[buf], ((li >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator (passes at first use)
            tc_fence_after();
            const uint32_t d = tmem_base + (uint32_t)buf * SC_TN;
            for (int it = 0; it < a.n_kb; ++it, ++g) {
                const int s = g % SC_NSTG;
                mbar_wait(&full[s], (g / SC_NSTG) & 1);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + s * SC_STAGE);
                const uint64_t dah = make_smem_desc_sw128(st, 16, 1024), dal = make_smem_desc_sw128(st + SC_APLANE, 16, 1024);
                const uint64_t dwh = make_smem_desc_sw128(st + 2 * SC_APLANE, 16, 1024);
                const uint64_t dwl = make_smem_desc_sw128(st + 2 * SC_APLANE + SC_WPLANE, 16, 1024);
#pragma unroll
                for (int k16 = 0; k16 < SG_KC / 16; ++k16) {
                    const uint64_t o = (uint64_t)(k16 * 2);
                    umma_bf16_ss_warp(d, dal + o, dwh + o, idesc, (it | k16) != 0);
                    umma_bf16_ss_warp(d, dah + o, dwl + o, idesc, 1);
                    umma_bf16_ss_warp(d, dah + o, dwh + o, idesc, 1);
                }
                umma_commit_warp(&empty[s]);
            }
            umma_commit_warp(&accfull[buf]);
        }
    } else {
        // ================= epilogue: thread = output row (TMEM lane) =================
        const int q = warp & 3, r = q * 32 + lane, pitch = a.grid_w + 2;
        int li = 0;
        for (int tile = blockIdx.x; tile < n_mtiles; tile += gridDim.x, ++li) {
            const int buf = li & 1, m0 = tile * 128;
            const int img = m0 / a.rows_per_img, qrow = m0 - img * a.rows_per_img + r;
            const int y1 = qrow / pitch, x1 = qrow - y1 * pitch;
            const bool valid = y1 >= 1 && y1 <= a.grid_h && x1 >= 1 && x1 <= a.grid_w;
            float* dst = a.raw + ((size_t)m0 + r) * 256;
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * SC_TN;
            mbar_wait(&accfull[buf], (li >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < SC_TN / 32; ++ch) {
                float y[32];
                sg_ld32(trow + ch * 32, y);
                if (valid) {
#pragma unroll
                    for (int c = 0; c < 32; c += 4)
                        *reinterpret_cast<float4*>(dst + ch * 32 + c) = make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]);
                }
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                    float s = 0.f, ss = 0.f;
                    if (valid) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) s += y[gi * 8 + c], ss += y[gi * 8 + c] * y[gi * 8 + c];
                    }
                    s = sg_warp_sum(s), ss = sg_warp_sum(ss);
                    if (lane == 0) s_stat[q * 32 + ch * 4 + gi] = s, s_stat[128 + q * 32 + ch * 4 + gi] = ss;
                }
            }
            tc_fence_before();
            mbar_arrive(&accempty[buf]);          // all 128 epilogue threads: the accumulator may be overwritten
            sg_named_bar();
            if (r < 32) {
                const float s = (s_stat[r] + s_stat[32 + r]) + (s_stat[64 + r] + s_stat[96 + r]);
                const float ss = (s_stat[128 + r] + s_stat[160 + r]) + (s_stat[192 + r] + s_stat[224 + r]);
                a.stats[(size_t)tile * 32 + r] = make_float2(s, ss);
            }
            sg_named_bar();                       // s_stat is free again
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// The clustered variant (opt-in: PF_CONV_CLUSTER=2|4; parity-tested, NOT faster: the kernel is tensor-bound, see above): CL consecutive tiles form a
// thread-block cluster; the weight tile of a k-block is the same for
// every tile, so each CTA fetches 1 / CL of it and TMA-multicasts the slice into all CL shared memories (L2 -> SM operand
// traffic per tile and k-block: 32 KB of activations + 64 / CL KB of weights instead of 96 KB).  A ring stage is refilled
// only after ALL CTAs of the cluster have consumed it: the MMA warp's commit arrives on the `empty` barrier of every peer.
template <int CL>
static __global__ void __launch_bounds__(SG_THREADS, 1)
sgemm_conv256_mc_kernel(const __grid_constant__ CUtensorMap tmap_ah, const __grid_constant__ CUtensorMap tmap_al,
                        const __grid_constant__ CUtensorMap tmap_ws, const __grid_constant__ SgArgs a, int n_mtiles) {
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
    constexpr int WSLICE = SC_WPLANE / CL;          // bytes of this CTA's slice of one weight plane (256 / CL rows)
    uint32_t crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int n_clusters = gridDim.x / CL, cid = blockIdx.x / CL;
    const int rounds = (n_mtiles + n_clusters * CL - 1) / (n_clusters * CL);
    extern __shared__ uint8_t sg_smem_raw[];
    uint8_t* smem = sg_smem_raw + ((1024u - (smem_u32(sg_smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SC_BAR_OFF);
    uint64_t* full = bars;
    uint64_t* empty = bars + SC_NSTG;
    uint64_t* accfull = bars + 2 * SC_NSTG;
    uint64_t* accempty = bars + 2 * SC_NSTG + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SC_NSTG + 4);
    float* s_stat = reinterpret_cast<float*>(smem + SC_STAT_OFF);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_ah);
        tma_prefetch_desc(&tmap_al);
        tma_prefetch_desc(&tmap_ws);
        for (int i = 0; i < SC_NSTG; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], CL);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&accfull[i], 1);
            mbar_init(&accempty[i], 128);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);
    // every CTA's barriers are initialised before a peer may multicast into it / arrive on it
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        // ================= TMA producer =================
        int g = 0;
        for (int rd = 0; rd < rounds; ++rd) {
            const int tile = (rd * n_clusters + cid) * CL + (int)crank;   // tiles past the end: all-zero operands (out of range)
            const int m0 = tile * 128, img = m0 / a.rows_per_img, q0 = m0 - img * a.rows_per_img;
            for (int it = 0; it < a.n_kb; ++it, ++g) {
                const int s = g % SC_NSTG;
                if (g >= SC_NSTG) mbar_wait(&empty[s], ((g / SC_NSTG) & 1) ^ 1);
                const int tap = it / a.cin_blocks, cb = it - tap * a.cin_blocks;
                const int ac0 = cb * SG_KC, ac1 = q0 + a.shift[tap], ac2 = img * a.planes_per_img + a.plane[tap];
                const int wr = tap * a.w_tap_rows;
                uint8_t* st = smem + s * SC_STAGE;
                mbar_arrive_expect_tx_warp(&full[s], SC_STAGE);
                tma_load_3d_warp(st, &tmap_ah, &full[s], ac0, ac1, ac2, kEvictNormal);
                tma_load_3d_warp(st + SC_APLANE, &tmap_al, &full[s], ac0, ac1, ac2, kEvictNormal);
                const int wrow = wr + (int)crank * (SC_TN / CL);
                tma_load_2d_mc_warp(st + 2 * SC_APLANE + crank * WSLICE, &tmap_ws, &full[s], ac0, wrow, kMask, kEvictLast);
                tma_load_2d_mc_warp(st + 2 * SC_APLANE + SC_WPLANE + crank * WSLICE, &tmap_ws, &full[s], ac0, wrow + a.w_lo, kMask,
                                    kEvictLast);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: D = Al*Wh + Ah*Wl + Ah*Wh =================
        constexpr uint32_t idesc = make_idesc_bf16(128, SC_TN, 0, 0);
        int g = 0, li = 0;
        for (int rd = 0; rd < rounds; ++rd, ++li) {
            const int buf = li & 1;
            mbar_wait(&accempty[buf], ((li >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator (passes at first use)
            tc_fence_after();
            const uint32_t d = tmem_base + (uint32_t)buf * SC_TN;
            for (int it = 0; it < a.n_kb; ++it, ++g) {
                const int s = g % SC_NSTG;
                mbar_wait(&full[s], (g / SC_NSTG) & 1);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + s * SC_STAGE);
                const uint64_t dah = make_smem_desc_sw128(st, 16, 1024), dal = make_smem_desc_sw128(st + SC_APLANE, 16, 1024);
                const uint64_t dwh = make_smem_desc_sw128(st + 2 * SC_APLANE, 16, 1024);
                const uint64_t dwl = make_smem_desc_sw128(st + 2 * SC_APLANE + SC_WPLANE, 16, 1024);
#pragma unroll
                for (int k16 = 0; k16 < SG_KC / 16; ++k16) {
                    const uint64_t o = (uint64_t)(k16 * 2);
                    umma_bf16_ss_warp(d, dal + o, dwh + o, idesc, (it | k16) != 0);
                    umma_bf16_ss_warp(d, dah + o, dwl + o, idesc, 1);
                    umma_bf16_ss_warp(d, dah + o, dwh + o, idesc, 1);
                }
                umma_commit_mc_warp(&empty[s], kMask);
            }
            umma_commit_warp(&accfull[buf]);
        }
    } else {
        // ================= epilogue: thread = output row (TMEM lane) =================
        const int q = warp & 3, r = q * 32 + lane, pitch = a.grid_w + 2;
        int li = 0;
        for (int rd = 0; rd < rounds; ++rd, ++li) {
            const int tile = (rd * n_clusters + cid) * CL + (int)crank;
            const bool live = tile < n_mtiles;
            const int buf = li & 1, m0 = tile * 128;
            const int img = m0 / a.rows_per_img, qrow = m0 - img * a.rows_per_img + r;
            const int y1 = qrow / pitch, x1 = qrow - y1 * pitch;
            const bool valid = live && y1 >= 1 && y1 <= a.grid_h && x1 >= 1 && x1 <= a.grid_w;
            float* dst = a.raw + ((size_t)m0 + r) * 256;
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * SC_TN;
            mbar_wait(&accfull[buf], (li >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < SC_TN / 32; ++ch) {
                float y[32];
                sg_ld32(trow + ch * 32, y);
                if (valid) {
#pragma unroll
                    for (int c = 0; c < 32; c += 4)
                        *reinterpret_cast<float4*>(dst + ch * 32 + c) = make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]);
                }
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                    float s = 0.f, ss = 0.f;
                    if (valid) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) s += y[gi * 8 + c], ss += y[gi * 8 + c] * y[gi * 8 + c];
                    }
                    s = sg_warp_sum(s), ss = sg_warp_sum(ss);
                    if (lane == 0) s_stat[q * 32 + ch * 4 + gi] = s, s_stat[128 + q * 32 + ch * 4 + gi] = ss;
                }
            }
            tc_fence_before();
            mbar_arrive(&accempty[buf]);          // all 128 epilogue threads: the accumulator may be overwritten
            sg_named_bar();
            if (r < 32 && live) {
                const float s = (s_stat[r] + s_stat[32 + r]) + (s_stat[64 + r] + s_stat[96 + r]);
                const float ss = (s_stat[128 + r] + s_stat[160 + r]) + (s_stat[192 + r] + s_stat[224 + r]);
                a.stats[(size_t)tile * 32 + r] = make_float2(s, ss);
            }
            sg_named_bar();                       // s_stat is free again
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    // no CTA leaves while a peer may still multicast into its shared memory or arrive on its barriers
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}


}  // namespace pf
