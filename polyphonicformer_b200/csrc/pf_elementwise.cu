// Streaming (HBM-bound) helper kernels of the decoder: storage cast, mask binarisation, x2 bilinear upsampling.
#include "pf_internal.h"
#include "pf_sm100.cuh"
#include "pf_up2.cuh"

namespace pf {

// ---------------------------------------------------------------------------------------------------------------
// fp32 [B][256][HW] (x2 tensors) -> bf16 [2][B][256][HWp]; pad columns [HW, HWp) are written as zero.
__global__ void cast_feats_kernel(const float* __restrict__ x, const float* __restrict__ d,
                                  uint16_t* __restrict__ out, int rows_per_branch, int HW, int HWp) {
    const int row = blockIdx.y;  // 0 .. 2*rows_per_branch
    const float* src = (row < rows_per_branch ? x + (size_t)row * HW : d + (size_t)(row - rows_per_branch) * HW);
    uint16_t* dst = out + (size_t)row * HWp;
    const bool vec = (HW & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    for (int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4; c < HWp; c += gridDim.x * blockDim.x * 4) {
        float v[4];
        if (vec && c + 3 < HW) {
            const float4 f = __ldcs(reinterpret_cast<const float4*>(src + c));
            v[0] = f.x, v[1] = f.y, v[2] = f.z, v[3] = f.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = (c + i < HW) ? src[c + i] : 0.f;
        }
        // HWp % 8 == 0 so c + 3 < HWp whenever c < HWp
        *reinterpret_cast<uint2*>(dst + c) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// mask logits fp32 [B][N][HW] -> bits u32 [B][WORDS][128]: one block per (b, group of BIN_WORDS words); warp w ballots the
// 32 pixels of a word for rows n = w, w+8, ...  (kernel_update_head.py:236-238: sigmoid(x) > 0.5  <=>  x > 0)
constexpr int BIN_WORDS = 4;
constexpr int BIN_WARPS = 16;
__global__ void __launch_bounds__(BIN_WARPS * 32) binarise_kernel(const float* __restrict__ logits,
                                                                  uint32_t* __restrict__ bits, int N, int HW, int words) {
    __shared__ uint32_t s_bits[BIN_WORDS][128];
    // wait FIRST, then release the dependents: the pooling kernel that follows starts streaming the feature maps
    // without waiting, which is only safe once everything before this kernel (their producer) has completed
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.y;
    const int w0 = blockIdx.x * BIN_WORDS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < BIN_WORDS * 128; i += BIN_WARPS * 32) (&s_bits[0][0])[i] = 0u;
    __syncthreads();
    const float* base = logits + (size_t)b * N * HW;
    // two rows x four words per iteration: 8 independent 128-byte loads in flight per warp
    for (int n = warp; n < N; n += 2 * BIN_WARPS) {
        float v[2][BIN_WORDS];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int nn = n + r * BIN_WARPS;
#pragma unroll
            for (int k = 0; k < BIN_WORDS; ++k) {
                const int hw = (w0 + k) * 32 + lane;
                v[r][k] = (nn < N && hw < HW) ? __ldcs(base + (size_t)nn * HW + hw) : 0.f;
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int nn = n + r * BIN_WARPS;
#pragma unroll
            for (int k = 0; k < BIN_WORDS; ++k) {
                const uint32_t m = __ballot_sync(0xffffffffu, v[r][k] > 0.f);
                if (lane == 0 && nn < N) s_bits[k][nn] = m;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BIN_WORDS * 128; i += BIN_WARPS * 32) {
        const int k = i >> 7, n = i & 127;
        if (w0 + k < words) bits[((size_t)b * words + w0 + k) * 128 + n] = s_bits[k][n];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) on [maps][H][W] planes
// (polyphonic/kernel_update.py:133-143).  src = (dst + 0.5)/2 - 0.5 clamped at 0; weights follow ATen's
// upsample_bilinear2d: h0l*(w0l*a + w1l*b) + h1l*(w0l*c + w1l*d).
// Thread = 1 input pixel column pair (x, x+1) -> 4 output columns, walking down a strip of UP_ROWS input rows with a
// rolling window of horizontally interpolated rows: every input row is interpolated horizontally once per strip and
// used by the (up to) three output rows that depend on it.
constexpr int UP_ROWS = 8;

// horizontal pass: output columns 2x .. 2x+3 of one input row
__device__ __forceinline__ void up_hrow(const float* __restrict__ row, int x, int W, float (&h)[4]) {
    const int c0 = max(x - 1, 0), c2 = min(x + 1, W - 1), c3 = min(x + 2, W - 1);
    const float a0 = __ldg(row + c0), a1 = __ldg(row + x), a2 = __ldg(row + c2), a3 = __ldg(row + c3);
    // out col 2x   : src = x - 0.25 -> (x-1, x) weights (0.25, 0.75); at x == 0 the source clamps to column 0
    // (up2_mix: the rounding order pf_panoptic's on-the-fly sampling uses too, pf_up2.cuh)
    h[0] = (x == 0) ? a1 : up2_mix(0.25f, a0, 0.75f, a1);
    h[1] = up2_mix(0.75f, a1, 0.25f, a2);   // src = x + 0.25
    h[2] = up2_mix(0.25f, a1, 0.75f, a2);   // src = x + 0.75
    h[3] = up2_mix(0.75f, a2, 0.25f, a3);   // src = x + 1.25 -> (x+1, x+2)
}

__global__ void __launch_bounds__(128) upsample2x_kernel(const float* __restrict__ in, float* __restrict__ out, int H,
                                                         int W) {
    pdl_launch_dependents();
    pdl_wait();
    const int map = blockIdx.z;
    const int y0 = blockIdx.y * UP_ROWS;
    const int y1 = min(y0 + UP_ROWS, H);
    const float* src = in + (size_t)map * H * W;
    float* dst = out + (size_t)map * (4 * (size_t)H * W);
    const int W2 = 2 * W;
    const int xp = blockIdx.x * blockDim.x + threadIdx.x;
    const int x = xp * 2;
    if (x >= W) return;
    const bool vec = x + 1 < W && (W2 & 3) == 0;
    float hA[4], hB[4], hC[4];
    up_hrow(src + (size_t)max(y0 - 1, 0) * W, x, W, hA);
    up_hrow(src + (size_t)y0 * W, x, W, hB);
#pragma unroll 2
    for (int y = y0; y < y1; ++y) {
        up_hrow(src + (size_t)min(y + 1, H - 1) * W, x, W, hC);
        // output row 2y   : src_y = y - 0.25 -> rows (y-1, y) weights (0.25, 0.75); at y == 0 it clamps to row 0
        // output row 2y+1 : src_y = y + 0.25 -> rows (y, y+1) weights (0.75, 0.25); at y == H-1 both rows are H-1
        float o0[4], o1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o0[k] = (y == 0) ? hB[k] : up2_mix(0.25f, hA[k], 0.75f, hB[k]);
            o1[k] = up2_mix(0.75f, hB[k], 0.25f, hC[k]);
        }
        float* d0 = dst + (size_t)(2 * y) * W2 + 2 * x;
        float* d1 = d0 + W2;
        if (vec) {
            __stcs(reinterpret_cast<float4*>(d0), make_float4(o0[0], o0[1], o0[2], o0[3]));
            __stcs(reinterpret_cast<float4*>(d1), make_float4(o1[0], o1[1], o1[2], o1[3]));
        } else {
            const int nout = (x + 1 < W) ? 4 : 2;
            for (int k = 0; k < nout; ++k) d0[k] = o0[k], d1[k] = o1[k];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) hA[k] = hB[k], hB[k] = hC[k];
    }
}

}  // namespace pf

extern "C" int pf_cast_feats(const float* x_feats, const float* depth_feats, uint16_t* feats, int B, int HW, int HWp,
                             void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(x_feats && depth_feats && feats, PF_ERR_ARG, "pf_cast_feats: null pointer");
    PF_REQUIRE(B > 0 && HW > 0 && HWp >= HW && HWp % 8 == 0, PF_ERR_ARG, "pf_cast_feats: bad shape B=%d HW=%d HWp=%d", B, HW, HWp);
    PF_REQUIRE((reinterpret_cast<uintptr_t>(feats) & 15) == 0, PF_ERR_ALIGN, "pf_cast_feats: feats not 16-byte aligned");
    const int rows = B * PF_C;
    int gx = (HWp / 4 + 255) / 256;
    if (gx > 32) gx = 32;
    dim3 grid(gx, 2 * rows);
    cast_feats_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x_feats, depth_feats, feats, rows, HW, HWp);
    PF_CHECK_LAUNCH("cast_feats_kernel");
    return PF_OK;
}

extern "C" int pf_cast_maps(const float* maps, uint16_t* out, int rows, int HW, int HWp, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(maps && out, PF_ERR_ARG, "pf_cast_maps: null pointer");
    PF_REQUIRE(rows > 0 && HW > 0 && HWp >= HW && HWp % 8 == 0, PF_ERR_ARG, "pf_cast_maps: bad shape rows=%d HW=%d HWp=%d", rows, HW, HWp);
    PF_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, PF_ERR_ALIGN, "pf_cast_maps: out not 16-byte aligned");
    int gx = (HWp / 4 + 255) / 256;
    if (gx > 32) gx = 32;
    cast_feats_kernel<<<dim3(gx, rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(maps, maps, out, rows, HW, HWp);
    PF_CHECK_LAUNCH("cast_feats_kernel");
    return PF_OK;
}

extern "C" int pf_binarise(const float* mask_logits, uint32_t* bits, int B, int N, int HW, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(mask_logits && bits, PF_ERR_ARG, "pf_binarise: null pointer");
    PF_REQUIRE(B > 0 && N > 0 && N <= PF_MAX_N && HW > 0, PF_ERR_ARG, "pf_binarise: bad shape B=%d N=%d HW=%d", B, N, HW);
    const int words = (HW + 31) / 32;
    dim3 grid((words + BIN_WORDS - 1) / BIN_WORDS, B);
    return launch_pdl("binarise_kernel", binarise_kernel, grid, dim3(BIN_WARPS * 32), 0, static_cast<cudaStream_t>(stream),
                      mask_logits, bits, N, HW, words);
}

extern "C" int pf_upsample2x(const float* in, float* out, int maps, int H, int W, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(in && out, PF_ERR_ARG, "pf_upsample2x: null pointer");
    PF_REQUIRE(maps > 0 && H > 0 && W > 0 && maps <= 65535 && H <= 65535, PF_ERR_ARG, "pf_upsample2x: bad shape maps=%d H=%d W=%d", maps, H, W);
    PF_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, PF_ERR_ALIGN, "pf_upsample2x: out not 16-byte aligned");
    const int pairs = (W + 1) / 2;
    int threads = 128;
    while (threads > 32 && threads / 2 >= pairs) threads /= 2;
    dim3 grid((pairs + threads - 1) / threads, (H + UP_ROWS - 1) / UP_ROWS, maps);
    return launch_pdl("upsample2x_kernel", upsample2x_kernel, grid, dim3(threads), 0, static_cast<cudaStream_t>(stream), in,
                      out, H, W);
}
