// K5 -- the tail of KernelHead._decode_init_proposals (polyphonic/kernel_head.py:250-336 of the reference): the step
// that produces the decoder's inputs from the three SemanticFPN maps (localization, semantic, depth):
//
//     loc   = ReLU(GN32(W_loc . maps[0]))      sem = ReLU(GN32(W_seg . maps[1]))      dep = ReLU(GN32(W_dep . maps[2]))
//     mask_preds  = init_kernels . loc          seg_preds = conv_seg . sem + b         depth_pred = conv_direct_depth . dep + b
//     x_feats     = sem + loc                   mask_preds := [mask_preds ; seg_preds[num_things:]]   (cat_stuff_mask)
//
// Three launches (then pf_mask_pool + pf_init_proposals, which already exist, finish kernel_head.py:313-336):
//   1. conv1x1_maps (pf_einsum.cu): the three 256x256 1x1 convolutions as one launch of the einsum kernel -> Y fp32.
//      Its epilogue also accumulates every row's sum and sum of squares: the GroupNorm statistics cost no extra pass.
//   2. gn_finalize_kernel: merges those partials (32 groups of 8 channels over the whole map, biased variance, fp64)
//      and folds gamma / beta into one (scale, shift) pair per channel.
//   3. head_apply_kernel (below): one streaming pass over Y.  Thread = pixel: normalise + ReLU the three maps 16
//      channels at a time (the next 16 already in flight), write x_feats / depth_feats in the decoder's bf16 layout, and feed the SAME registers to the
//      tensor cores as the A operand of the three prediction heads: the activations are split into bf16 hi + lo, packed
//      two channels per 32-bit word and stored with tcgen05.st into tensor memory (row = pixel in lane, K = channel
//      along the columns), the head weights (bf16 hi + lo, K-major, 128-byte swizzle) sit in shared memory for the
//      whole kernel, two A buffers per group, three TS-mode MMAs per K step (Al.Wh + Ah.Wl + Ah.Wh, fp32 accumulate) keep fp32-level accuracy.
//      D = [128 pixels][112 | 32 | 16 columns] per worker group; the group reads it back with tcgen05.ld, adds the
//      biases, writes mask_preds / seg_preds / depth_pred coalesced along the pixels and ballots the sign bits of the
//      initial masks (kernel_head.py:314-317) in the layout pf_mask_pool consumes.
//   Warp roles (576 threads, one persistent CTA per SM): warps 0..7 / 8..15 = two worker groups (each owns 256 TMEM
//   columns and its own tiles); inside a group two sets of 4 warps (TMEM lane quadrant = warp % 4) share the tile's
//   pixels and split its channel chunks (even / odd = A buffer 0 / 1); warps 16 / 17 = the groups' MMA issuers.
//   (With 8 worker warps the kernel was issue-latency-bound: 121 of 175 us remained with every global access off.)
// HBM-bound: 3 KB of Y read, 1 KB of bf16 features and ~0.5 KB of predictions written per pixel.
#include "pf_internal.h"
#include "pf_sm100.cuh"

namespace pf {

constexpr int H_THREADS = 576;
constexpr int H_ROWS = 160;                 // head rows: init_kernels 0..111 | conv_seg 112..143 | conv_direct_depth 144..159
constexpr int H_ROW_SEG = 112, H_ROW_DEP = 144;
constexpr int H_KBLK = H_ROWS * 128;        // bytes of one 64-channel block of one weight plane
constexpr int H_WBYTES = 2 * 4 * H_KBLK;    // hi / lo planes x 4 channel blocks = 163840
constexpr int H_CH = 16;                    // channels per chunk = one K step of the MMAs
constexpr int H_NCH = 256 / H_CH;
constexpr int H_PF = 2;                     // L2 prefetch distance in a set's own chunks
constexpr int H_DCOLS = 160;                // accumulator columns of a group
constexpr int H_ACOLS = 48;                 // one A chunk: 3 maps x (hi, lo) x 8 columns; two buffers per group
constexpr int H_GCOLS = 256;                // TMEM columns per group: 160 + 2 * 48
constexpr int H_AFF_OFF = H_WBYTES;                         // float2 [2 groups][3][256]
constexpr int H_BIAS_OFF = H_AFF_OFF + 2 * 3 * 256 * 8;     // float [160]
constexpr int H_BAR_OFF = H_BIAS_OFF + H_ROWS * 4;
constexpr int H_SMEM = H_BAR_OFF + 128 + 1024;
static_assert(H_DCOLS + 2 * H_ACOLS == H_GCOLS, "tensor memory budget of a worker group");

struct HeadParams {
    const float* Y;          // [6B][nblk][128][32]: unit = half * 3B + map * B + b, 32-pixel blocks
    const float2* affine;    // [3B][256] (scale, shift): unit = map * B + b
    const float* head_b;     // [160]
    uint16_t* feats;         // out bf16 [2][B][256][HWp]
    float* x32;              // optional fp32 [B][256][HW]
    float* d32;              // optional fp32 [B][256][HW]
    float* mask_preds;       // [B][P + n_stuff][HW]
    float* seg_preds;        // [B][num_classes][HW]
    float* depth_pred;       // [B][HW]
    uint32_t* bits;          // optional [B][WORDS][128]: sign bits of the P initial masks
    int B, HW, HWp, P, num_classes, num_things, words;
    int tiles_per_img, n_tiles, nblk;
};

__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory"); }   // the 8 warps of a group

// 16 channels of one map (already normalised + ReLU) -> bf16 hi / lo, two channels per 32-bit column
__device__ __forceinline__ void st_split8(uint32_t taddr, const float* v) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float a = v[2 * j], b = v[2 * j + 1];
        hi[j] = pack_bf16x2(a, b);
        lo[j] = pack_bf16x2(a - __uint_as_float(hi[j] << 16), b - __uint_as_float(hi[j] & 0xFFFF0000u));
    }
    tmem_st8(taddr, hi);
    tmem_st8(taddr + 8, lo);
}

__global__ void __launch_bounds__(H_THREADS, 1)
head_apply_kernel(const __grid_constant__ CUtensorMap tmap_w, const HeadParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float2* s_aff = reinterpret_cast<float2*>(smem + H_AFF_OFF);
    float* s_bias = reinterpret_cast<float*>(smem + H_BIAS_OFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + H_BAR_OFF);
    uint64_t* wfull = bars;          // head weights landed
    uint64_t* afull = bars + 1;      // [group][buffer] the A chunk is in tensor memory (128 arrivals)
    uint64_t* afree = bars + 5;      // [group][buffer] the MMAs that read it have completed
    uint64_t* dfull = bars + 9;      // [group] the tile's accumulators are complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    // warp index through a shuffle: provably warp-uniform, so the role branches below are not treated as divergent
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_w);
        mbar_init(wfull, 1);
        for (int i = 0; i < 4; ++i) {
            mbar_init(&afull[i], 128);
            mbar_init(&afree[i], 1);
        }
        mbar_init(&dfull[0], 1);
        mbar_init(&dfull[1], 1);
        mbar_fence_init();
    }
    if (warp == 16) tmem_alloc<512>(tmem_slot);
    for (int i = threadIdx.x; i < H_ROWS; i += H_THREADS) s_bias[i] = __ldg(p.head_b + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);
    pdl_launch_dependents();

    if (warp >= 16) {
        // ================= MMA issuer of group g (the weights are static: their TMA does not wait for the grid) ====
        const int g = warp - 16;
        if (g == 0 && lane == 0) {
            mbar_arrive_expect_tx(wfull, H_WBYTES);
            for (int i = 0; i < 8; ++i)   // (plane, channel block): a [160 rows][64 channels] box
                tma_load_2d(smem + i * H_KBLK, &tmap_w, wfull, (i & 3) * 64, (i >> 2) * H_ROWS, kEvictLast);
        }
        mbar_wait(wfull, 0);
        const uint32_t tD = tmem_base + g * H_GCOLS, tA = tD + H_DCOLS;
        const uint32_t wbase = smem_u32(smem);
        constexpr uint32_t idesc[3] = {make_idesc_bf16(128, 112, 0, 0), make_idesc_bf16(128, 32, 0, 0),
                                       make_idesc_bf16(128, 16, 0, 0)};
        constexpr int rowoff[3] = {0, H_ROW_SEG, H_ROW_DEP};
        uint32_t n = 0;
        for (int t = blockIdx.x * 2 + g; t < p.n_tiles; t += gridDim.x * 2) {
#pragma unroll 1
            for (int k = 0; k < H_NCH; ++k, ++n) {
                const uint32_t buf = n & 1;
                mbar_wait_backoff(&afull[g * 2 + buf], (n >> 1) & 1, 32);   // shares a scheduler with two worker warps
                tc_fence_after();
                const int c0 = k * H_CH;
                const uint64_t koff = (uint64_t)(((c0 & 63) * 2) >> 4);
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    const uint32_t wa = wbase + (c0 >> 6) * H_KBLK + rowoff[m] * 128;
                    const uint64_t bh = make_smem_desc_sw128(wa, 16, 1024) + koff;
                    const uint64_t bl = make_smem_desc_sw128(wa + 4 * H_KBLK, 16, 1024) + koff;
                    const uint32_t ah = tA + buf * H_ACOLS + m * 16, al = ah + 8;
                    const uint32_t d = tD + rowoff[m];
                    umma_bf16_ts_warp(d, al, bh, idesc[m], k != 0);
                    umma_bf16_ts_warp(d, ah, bl, idesc[m], 1);
                    umma_bf16_ts_warp(d, ah, bh, idesc[m], 1);
                }
                umma_commit_warp(&afree[g * 2 + buf]);
                if (k == H_NCH - 1) umma_commit_warp(&dfull[g]);
            }
        }
    } else {
        // ================= worker: group g (its own tiles), set (even / odd chunks -> A buffer `set`), quadrant q ====
        // thread = pixel; the two sets of a group work on the SAME pixels (TMEM lanes) and different channels
        const int g = warp >> 3, set = (warp >> 2) & 1, q = warp & 3;
        const uint32_t tD = tmem_base + ((uint32_t)(q * 32) << 16) + g * H_GCOLS;
        const uint32_t ta = tD + H_DCOLS + set * H_ACOLS;
        uint64_t* my_full = &afull[g * 2 + set];
        uint64_t* my_free = &afree[g * 2 + set];
        float2* aff = s_aff + g * 3 * 256;
        const int tg = threadIdx.x & 255;
        const int NM = p.P + p.num_classes - p.num_things;   // channels of mask_preds
        // Y is pixel-blocked ([unit][32-px block][128 channels][32 px]): a warp reads ONE block, its channel stride is a
        // compile-time 128 bytes, so the 48 loads of a chunk are immediates off three pointers (the first version spent
        // 45 % of its instructions on 64-bit index arithmetic, ncu source page).
        const size_t hw = (size_t)p.HW, hwp = (size_t)p.HWp;
        const size_t map_stride = (size_t)p.B * p.nblk * 4096;            // Y: from one map's unit to the next map's
        const size_t half_jump = ((size_t)3 * p.B * p.nblk - 1) * 4096;   // channel 128 of half 0 -> channel 0 of half 1
        const size_t feat_branch = (size_t)p.B * 256 * p.HWp;             // feats: x_feats -> depth_feats
        pdl_wait();   // Y and the affine table come from the previous kernels
        int cur_b = -1;
        uint32_t use = 0, tile_i = 0;   // uses of this set's A buffer so far
        bool pending = false;           // the last chunk is in tensor memory but not yet handed to the MMA warp
        auto flush_a = [&]() {
            if (pending) {
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(my_full);
                pending = false;
            }
        };
        for (int t = blockIdx.x * 2 + g; t < p.n_tiles; t += gridDim.x * 2, ++tile_i) {
            const int b = t / p.tiles_per_img;
            const int px = (t - b * p.tiles_per_img) * 128 + q * 32 + lane;
            const bool ok = px < p.HW, okp = px < p.HWp;
            if (b != cur_b) {   // (scale, shift) of this image's three maps -> shared memory
                group_bar(g);
                for (int i = tg; i < 3 * 256; i += 256)
                    aff[i] = __ldg(p.affine + ((size_t)(i >> 8) * p.B + b) * 256 + (i & 255));
                group_bar(g);
                cur_b = b;
            }
            int blk = (px - lane) >> 5;   // this warp's 32-pixel block (clamped: rows of A / D beyond HW are never stored)
            if (blk >= p.nblk) blk = p.nblk - 1;
            const float* yb = p.Y + ((size_t)b * p.nblk + blk) * 4096 + lane;
            uint16_t* xb = p.feats + (size_t)b * 256 * hwp + (okp ? px : 0);
#pragma unroll 1
            for (int k = set; k < H_NCH; k += 2) {
                const float* y0 = yb + k * 512 + (k >= 8 ? half_jump : 0);
                const float* y1 = y0 + map_stride;
                const float* y2 = y1 + map_stride;
                float v[48];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __ldcs(y0 + j * 32), v[16 + j] = __ldcs(y1 + j * 32), v[32 + j] = __ldcs(y2 + j * 32);
                if (lane == 0 && k + 2 * H_PF < H_NCH) {   // L2 prefetch of this set's chunk H_PF turns ahead (3 x 2 KB)
                    const float* yc = y0 + 2 * H_PF * 512 + ((k < 8 && k + 2 * H_PF >= 8) ? half_jump : 0);
                    prefetch_l2_bulk(yc, 2048);
                    prefetch_l2_bulk(yc + map_stride, 2048);
                    prefetch_l2_bulk(yc + 2 * map_stride, 2048);
                }
                const float2* a0 = aff + k * H_CH;
#pragma unroll
                for (int m = 0; m < 3; ++m)
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float2 a = a0[m * 256 + j];
                        v[m * 16 + j] = fmaxf(fmaf(v[m * 16 + j], a.x, a.y), 0.f);
                    }
                if (use >= 1) {   // the MMAs that read this buffer on its previous use have completed
                    mbar_wait(my_free, (use - 1) & 1);
                    tc_fence_after();
                }
                st_split8(ta, v);
                st_split8(ta + 16, v + 16);
                st_split8(ta + 32, v + 32);
                pending = true;
                ++use;
                // x_feats = sem + loc (kernel_head.py:303), depth_feats: the decoder's bf16 maps (+ optional fp32 copies)
                if (okp) {
                    uint16_t* xp = xb + (size_t)k * 16 * hwp;
                    uint16_t* dp = xp + feat_branch;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const uint32_t pk = ok ? pack_bf16x2(v[j] + v[16 + j], v[32 + j]) : 0u;   // pad columns: zero
                        *xp = (uint16_t)(pk & 0xFFFFu), *dp = (uint16_t)(pk >> 16);
                        xp += hwp, dp += hwp;
                    }
                }
                if (p.x32 && ok) {
                    float* o = p.x32 + ((size_t)b * 256 + k * 16) * hw + px;
                    float* d = p.d32 + ((size_t)b * 256 + k * 16) * hw + px;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        *o = v[j] + v[16 + j], *d = v[32 + j];
                        o += hw, d += hw;
                    }
                }
                flush_a();   // the tcgen05.st above had the global stores to land: hand the chunk to the MMA warp
            }
            // ---- the three heads of this tile: accumulators -> predictions (set 0: columns 0..95, set 1: 96..159)
            flush_a();
            mbar_wait(&dfull[g], tile_i & 1);
            tc_fence_after();
            float* mp = p.mask_preds + (size_t)b * NM * p.HW + px;
            float* sp = p.seg_preds + (size_t)b * p.num_classes * p.HW + px;
            uint32_t* bw = p.bits ? p.bits + ((size_t)b * p.words + (px >> 5)) * 128 : nullptr;
#pragma unroll 1
            for (int cb = set * 3; cb < 3 + set * 2; ++cb) {   // columns [32 cb, 32 cb + 32)
                uint32_t v[32];
                tmem_ld32(tD + cb * 32, v);
                tmem_ld_wait();
                uint32_t word = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = cb * 32 + j;
                    const float f = __uint_as_float(v[j]) + s_bias[col];
                    if (col < H_ROW_SEG) {
                        const uint32_t bal = __ballot_sync(0xffffffffu, ok && col < p.P && f > 0.f);
                        if (lane == j) word = bal;
                        if (ok && col < p.P) mp[(size_t)col * p.HW] = f;
                    } else if (col < H_ROW_DEP) {
                        const int cls = col - H_ROW_SEG;
                        if (ok && cls < p.num_classes) {
                            sp[(size_t)cls * p.HW] = f;
                            if (cls >= p.num_things) mp[(size_t)(p.P + cls - p.num_things) * p.HW] = f;
                        }
                    } else if (col == H_ROW_DEP) {
                        if (ok) p.depth_pred[(size_t)b * p.HW + px] = f;
                    }
                }
                if (bw && cb < 4 && (px - lane) < p.HW) bw[cb * 32 + lane] = word;   // rows >= P: zero
            }
            tc_fence_before();
            group_bar(g);   // both sets have read D: the next tile's first MMA may overwrite it
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc<512>(tmem_base);
}

// GroupNorm finalisation of one (map-image unit u, group of 8 channels): merge the per-row partial sums the conv
// epilogue left (fp64), then per channel (scale, shift) = (gamma * rstd, beta - mean * gamma * rstd).
// torch.nn.GroupNorm: biased variance, eps inside the sqrt.
__global__ void __launch_bounds__(32)
gn_finalize_kernel(const float2* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float2* __restrict__ affine, int B, int HW, int cpu, float eps) {
    pdl_wait();
    pdl_launch_dependents();
    const int grp = blockIdx.x;            // 0..31
    const int u = blockIdx.y;              // map * B + b
    const int c0 = grp * 8;
    const int unit = (c0 >> 7) * 3 * B + u;
    double ds = 0.0, dq = 0.0;
    for (int i = threadIdx.x; i < 8 * cpu; i += 32) {
        const float2 v = __ldg(stats + ((size_t)unit * cpu + (i >> 3)) * 128 + (c0 & 127) + (i & 7));
        ds += (double)v.x, dq += (double)v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
        dq += __shfl_xor_sync(0xffffffffu, dq, o);
    }
    if (threadIdx.x < 8) {
        const double total = 8.0 * (double)HW;
        const double mean = ds / total;
        double var = dq / total - mean * mean;
        if (var < 0.0) var = 0.0;
        const double rstd = 1.0 / sqrt(var + (double)eps);
        const int map = u / B, c = c0 + threadIdx.x;
        const double sc = (double)gamma[map * 256 + c] * rstd;
        affine[(size_t)u * 256 + c] = make_float2((float)sc, (float)((double)beta[map * 256 + c] - mean * sc));
    }
}

// GroupNorm + ReLU of the blocked conv output -> bf16 maps [3][B][256][HWp] (+ optional fp32 [3][B][256][HW]): the
// elementwise tail of pf_fpn_pred.  One warp per 16 KB block [128 channels][32 px]: a lane owns 4 consecutive pixels
// (one float4 in, 8 bytes of bf16 out), 8 lanes a channel row, 4 rows per warp instruction.
__global__ void __launch_bounds__(256)
gn_apply_kernel(const float* __restrict__ Y, const float2* __restrict__ affine, uint16_t* __restrict__ maps,
                float* __restrict__ maps32, int B, int HW, int HWp, int nblk) {
    pdl_wait();
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long nblocks = (long long)6 * B * nblk;
    for (long long blkid = (long long)blockIdx.x * 8 + warp; blkid < nblocks; blkid += (long long)gridDim.x * 8) {
        const int unit = (int)(blkid / nblk), blk = (int)(blkid - (long long)unit * nblk);
        const int half = unit / (3 * B), u = unit - half * 3 * B;   // u = map * B + b
        const float4* src = reinterpret_cast<const float4*>(Y + (size_t)blkid * 4096);
        const float2* af = affine + (size_t)u * 256 + half * 128;
        const int px = blk * 32 + (lane & 7) * 4;
        uint16_t* dst = maps + ((size_t)u * 256 + half * 128) * HWp + px;
        float* dst32 = maps32 ? maps32 + ((size_t)u * 256 + half * 128) * HW + px : nullptr;
#pragma unroll 4
        for (int r0 = 0; r0 < 128; r0 += 4) {
            const int r = r0 + (lane >> 3);
            const float4 v = __ldcs(src + r * 8 + (lane & 7));
            const float2 a = __ldg(af + r);
            float o[4] = {fmaxf(fmaf(v.x, a.x, a.y), 0.f), fmaxf(fmaf(v.y, a.x, a.y), 0.f), fmaxf(fmaf(v.z, a.x, a.y), 0.f),
                          fmaxf(fmaf(v.w, a.x, a.y), 0.f)};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (px + i >= HW) o[i] = 0.f;   // pad columns [HW, HWp) are zero
            if (px + 3 < HWp)                   // HWp % 8 == 0 and px % 4 == 0: all four or none
                *reinterpret_cast<uint2*>(dst + (size_t)r * HWp) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
            if (dst32) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (px + i < HW) dst32[(size_t)r * HW + i] = o[i];
            }
        }
    }
}

}  // namespace pf

extern "C" size_t pf_kernel_head_workspace_bytes(int B, int HW) {
    if (B <= 0 || HW <= 0) return 0;
    // Y fp32 [6B][ceil(HW/32)][128][32] | affine float2 [3B][256] | statistics partials float2 [6B][<= SMs][128]
    return (size_t)6 * B * ((HW + 31) / 32) * 16384 + (size_t)3 * B * 256 * 8 + (size_t)6 * B * 148 * 128 * 8;
}

extern "C" int pf_kernel_head(const pf_head_weights* w, const uint16_t* maps, uint16_t* feats, float* x32, float* d32,
                              float* mask_preds, float* seg_preds, float* depth_pred, uint32_t* bits, void* workspace,
                              size_t workspace_bytes, int B, int HW, int HWp, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(w && maps && feats && mask_preds && seg_preds && depth_pred && workspace, PF_ERR_ARG, "pf_kernel_head: null pointer");
    PF_REQUIRE(w->conv_split && w->gn_gamma && w->gn_beta && w->head_w && w->head_b, PF_ERR_ARG, "pf_kernel_head: null weight pointer");
    PF_REQUIRE(B > 0 && HW > 0 && HWp >= HW && HWp % 8 == 0, PF_ERR_ARG, "pf_kernel_head: bad shape B=%d HW=%d HWp=%d", B, HW, HWp);
    PF_REQUIRE(w->num_proposals > 0 && w->num_proposals <= H_ROW_SEG && w->num_classes > 0 && w->num_classes <= H_ROW_DEP - H_ROW_SEG &&
                   w->num_thing_classes >= 0 && w->num_thing_classes <= w->num_classes,
               PF_ERR_ARG, "pf_kernel_head: unsupported head sizes P=%d classes=%d things=%d", w->num_proposals, w->num_classes,
               w->num_thing_classes);
    PF_REQUIRE((x32 == nullptr) == (d32 == nullptr), PF_ERR_ARG, "pf_kernel_head: x32 and d32 go together");
    PF_REQUIRE(workspace_bytes >= pf_kernel_head_workspace_bytes(B, HW), PF_ERR_WORKSPACE, "pf_kernel_head: workspace too small");
    PF_REQUIRE(((reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(maps) | reinterpret_cast<uintptr_t>(feats) |
                 reinterpret_cast<uintptr_t>(w->conv_split) | reinterpret_cast<uintptr_t>(w->head_w)) & 15) == 0,
               PF_ERR_ALIGN, "pf_kernel_head: pointers must be 16-byte aligned");
    reset_launch_count();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* Y = static_cast<float*>(workspace);
    float2* affine = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(workspace) + (size_t)6 * B * ((HW + 31) / 32) * 16384);

    float2* stats = affine + (size_t)3 * B * 256;
    const int cpu = conv1x1_ctas_per_unit(B, HW);
    PF_REQUIRE(cpu <= 148, PF_ERR_WORKSPACE, "pf_kernel_head: %d CTAs per unit exceed the statistics scratch", cpu);
    if (int e = conv1x1_maps(maps, 3, w->conv_split, Y, stats, B, HW, HWp, stream)) return e;
    if (int e = launch_pdl("gn_finalize_kernel", gn_finalize_kernel, dim3(32, 3 * B), dim3(32), 0, st, (const float2*)stats,
                           w->gn_gamma, w->gn_beta, affine, B, HW, cpu, w->gn_eps))
        return e;

    HeadParams p;
    p.Y = Y, p.affine = affine, p.head_b = w->head_b, p.feats = feats, p.x32 = x32, p.d32 = d32;
    p.mask_preds = mask_preds, p.seg_preds = seg_preds, p.depth_pred = depth_pred, p.bits = bits;
    p.B = B, p.HW = HW, p.HWp = HWp, p.P = w->num_proposals, p.num_classes = w->num_classes, p.num_things = w->num_thing_classes;
    p.words = (HW + 31) / 32;
    p.tiles_per_img = (HWp + 127) / 128, p.n_tiles = B * p.tiles_per_img, p.nblk = (HW + 31) / 32;
    CUtensorMap tmap_w;
    if (int e = make_tmap_bf16_2d(&tmap_w, w->head_w, 2 * H_ROWS, 256, 256, H_ROWS, 64)) return e;
    cudaError_t ea = cudaFuncSetAttribute(head_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM);
    if (ea != cudaSuccess) return set_error(PF_ERR_CUDA, "head_apply smem attribute: %s", cudaGetErrorString(ea));
    int grid = num_sms();
    if (grid * 2 > p.n_tiles) grid = (p.n_tiles + 1) / 2;
    return launch_pdl("head_apply_kernel", head_apply_kernel, dim3(grid), dim3(H_THREADS), H_SMEM, st, tmap_w, p);
}

extern "C" int pf_fpn_pred(const uint16_t* conv_split, const float* gn_gamma, const float* gn_beta, float gn_eps,
                           const uint16_t* fused, uint16_t* maps, float* maps32, void* workspace, size_t workspace_bytes,
                           int B, int HW, int HWp, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(conv_split && gn_gamma && gn_beta && fused && maps && workspace, PF_ERR_ARG, "pf_fpn_pred: null pointer");
    PF_REQUIRE(B > 0 && HW > 0 && HWp >= HW && HWp % 8 == 0, PF_ERR_ARG, "pf_fpn_pred: bad shape B=%d HW=%d HWp=%d", B, HW, HWp);
    PF_REQUIRE(workspace_bytes >= pf_kernel_head_workspace_bytes(B, HW), PF_ERR_WORKSPACE, "pf_fpn_pred: workspace too small");
    PF_REQUIRE(((reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(fused) | reinterpret_cast<uintptr_t>(maps) |
                 reinterpret_cast<uintptr_t>(conv_split)) & 15) == 0,
               PF_ERR_ALIGN, "pf_fpn_pred: pointers must be 16-byte aligned");
    reset_launch_count();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nblk = (HW + 31) / 32;
    float* Y = static_cast<float*>(workspace);
    float2* affine = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(workspace) + (size_t)6 * B * nblk * 16384);
    float2* stats = affine + (size_t)3 * B * 256;
    const int cpu = conv1x1_ctas_per_unit(B, HW);
    PF_REQUIRE(cpu <= 148, PF_ERR_WORKSPACE, "pf_fpn_pred: %d CTAs per unit exceed the statistics scratch", cpu);
    (void)Y;   // two passes over the bf16 input instead of an fp32 intermediate: statistics, then conv + GN + ReLU -> maps
    if (int e = conv1x1_fused(fused, 1, conv_split, stats, nullptr, nullptr, nullptr, B, HW, HWp, stream)) return e;
    if (int e = launch_pdl("gn_finalize_kernel", gn_finalize_kernel, dim3(32, 3 * B), dim3(32), 0, st, (const float2*)stats,
                           gn_gamma, gn_beta, affine, B, HW, cpu, gn_eps))
        return e;
    return conv1x1_fused(fused, 1, conv_split, nullptr, affine, maps, maps32, B, HW, HWp, stream);
}
