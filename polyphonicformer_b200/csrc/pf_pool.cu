// K1 -- mask pooling (polyphonic/kernel_update_head.py:236-242, polyphonic/kernel_head.py:313-320 of the reference):
//     pooled[g][n][c] = sum_hw 1[mask_logit[b][n][hw] > 0] * feats[g][c][hw]
// as a split-K [128 x HW] x [HW x 256] GEMM on tcgen05.  The mask operand is {0,1}, so every product is exact and
// only the fp32 summation order differs from the reference.
//
//   A  = mask tile [128 n][64 hw] bf16, expanded on chip from the packed sign bits (1 KB per tile instead of the
//        28 KB of fp32 logits), written K-major / 128-byte-swizzled by the 4 expander warps;
//   B  = feats[g][:, hw0:hw0+64] bf16, TMA box [256 c][64 hw] = K-major operand with N = 256;
//   D  = [128 n][256 c] fp32 in TMEM, accumulated over the CTA's slab of HW, then written as one split-K partial.
//   Deterministic: partials are reduced in a fixed order by the consumer (pf_kernel_update / pf_pool_reduce).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = bit expanders, then epilogue.  HBM-bound: 2*C*HW*2 bytes of features per image dominate.
#include "pf_internal.h"
#include "pf_sm100.cuh"
#include "pf_debug.cuh"

namespace pf {

constexpr int P_C = PF_C;
constexpr int P_BHW = 64;  // K per pipeline stage
constexpr int P_STAGES = 4;
constexpr int P_A_BYTES = 128 * P_BHW * 2;  // 16384
constexpr int P_B_BYTES = P_C * P_BHW * 2;  // 32768
constexpr int P_STAGE_BYTES = P_A_BYTES + P_B_BYTES;
constexpr int P_TMEM_COLS = 256;
constexpr int P_THREADS = 192;
constexpr int P_SMEM = P_STAGES * P_STAGE_BYTES + 256 + 1024;

struct PoolParams {
    const uint32_t* bits;  // [B][WORDS][128]
    float* partial;        // [G][S][64 column groups][N][4]: column-group-major so that thread = row stores / loads coalesce
    float* cntp;           // [G][S][N]
    int N, HW, words, B, S, tiles_per_unit;
    int Btot, b0;          // batch window: this launch covers images b0 .. b0+B-1 of a [n_branch][Btot] feature tensor
    int early_feats;       // the feature maps were complete before the PREVIOUS kernel started: stream them before pdl_wait
};

// 8 mask bits -> 8 bf16 {0,1} packed in 4 u32 (element 2i in the low half)
__device__ __forceinline__ uint4 expand8(uint32_t b) {
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        r[i] = ((b >> (2 * i)) & 1u) * 0x3F80u | ((b >> (2 * i + 1)) & 1u) * 0x3F800000u;
    return make_uint4(r[0], r[1], r[2], r[3]);
}

__global__ void __launch_bounds__(P_THREADS, 1)
pool_kernel(const __grid_constant__ CUtensorMap tmap_feats, const PoolParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_STAGES * P_STAGE_BYTES);
    uint64_t* fullB = bars;
    uint64_t* fullA = bars + P_STAGES;
    uint64_t* empty = bars + 2 * P_STAGES;
    uint64_t* accfull = bars + 3 * P_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * P_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x / p.S;  // branch * B + b
    const int split = blockIdx.x % p.S;
    const int b = unit % p.B;
    const int tile_begin = (int)((long long)split * p.tiles_per_unit / p.S);
    const int tile_end = (int)((long long)(split + 1) * p.tiles_per_unit / p.S);
    const int ntiles = tile_end - tile_begin;
    long long* dbg = dbg_claim_all(10);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_feats);
        for (int i = 0; i < P_STAGES; ++i) {
            mbar_init(&fullB[i], 1);
            mbar_init(&fullA[i], 128);
            mbar_init(&empty[i], 1);
        }
        mbar_init(accfull, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<P_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);   // warp-uniform (REDUX -> uniform register): tcgen05 operands need no per-instruction R2UR
    DBG(1);
    pdl_launch_dependents();
    // Inside the decode loop the feature tiles are long-complete inputs (early_feats): the producer streams them
    // without waiting for the previous kernel.  Only the mask bits come from it, so the expander warps (which also do
    // every global write of this CTA) wait on the grid dependency.

    if (warp == 0) {
        if (lane == 0) {
            if (!p.early_feats) pdl_wait();
            for (int i = 0; i < ntiles; ++i) {
                const int s = i % P_STAGES;
                mbar_wait(&empty[s], ((i / P_STAGES) & 1) ^ 1);
                if (i == 4) DBG(2);
                if (i == ntiles - 1) DBG(3);
                mbar_arrive_expect_tx(&fullB[s], P_B_BYTES);
                // Inside the decode loop x_feats (branch 0) is read by every kernel of every stage: keep it in the 126 MB L2
                // (evict-last); depth_feats is only read here (and by the last einsum): stream it (evict-first).  A
                // stand-alone call (pf_mask_pool) cannot assume a reader right behind it: pure streaming.
                tma_load_2d(smem + s * P_STAGE_BYTES + P_A_BYTES, &tmap_feats, &fullB[s], (tile_begin + i) * P_BHW,
                            ((unit / p.B) * p.Btot + p.b0 + b) * P_C, (p.early_feats && unit < p.B) ? kEvictLast : kEvictFirst);
            }
        }
    } else if (warp == 1) {
        // MMA issuer: the whole warp runs the loop, one lane elected inside the PTX block issues
        constexpr uint32_t idesc = make_idesc_bf16(128, P_C, 0, 0);
        for (int i = 0; i < ntiles; ++i) {
            const int s = i % P_STAGES;
            const uint32_t ph = (i / P_STAGES) & 1;
            mbar_wait(&fullA[s], ph);
            mbar_wait(&fullB[s], ph);
            tc_fence_after();
            const uint32_t a_base = smem_u32(smem + s * P_STAGE_BYTES);
            const uint64_t da = make_smem_desc_sw128(a_base, 16, 1024), db = make_smem_desc_sw128(a_base + P_A_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < P_BHW / 16; ++k)
                umma_bf16_ss_warp(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (i | k) != 0);
            umma_commit_warp(&empty[s]);
        }
        umma_commit_warp(accfull);
    } else {
        // ---- expanders: thread = mask row r (TMEM lane r later in the epilogue)
        pdl_wait();
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t* brow = p.bits + (size_t)b * p.words * 128 + r;
        uint32_t count = 0;
        for (int i = 0; i < ntiles; ++i) {
            const int s = i % P_STAGES;
            const int w0 = (tile_begin + i) * 2;
            const uint32_t m0 = (w0 < p.words) ? __ldg(brow + (size_t)w0 * 128) : 0u;
            const uint32_t m1 = (w0 + 1 < p.words) ? __ldg(brow + (size_t)(w0 + 1) * 128) : 0u;
            count += __popc(m0) + __popc(m1);
            mbar_wait(&empty[s], ((i / P_STAGES) & 1) ^ 1);
            uint8_t* sA = smem + s * P_STAGE_BYTES;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                *reinterpret_cast<uint4*>(sA + sw128_offset(r, c)) = expand8((m0 >> (8 * c)) & 0xFFu);
                *reinterpret_cast<uint4*>(sA + sw128_offset(r, c + 4)) = expand8((m1 >> (8 * c)) & 0xFFu);
            }
            fence_proxy_async_smem();
            mbar_arrive(&fullA[s]);
        }
        // ---- epilogue: one split-K partial per CTA
        const size_t slab = (size_t)unit * p.S + split;
        if (r < p.N) p.cntp[slab * p.N + r] = (float)count;
        if (ntiles > 0) {
            mbar_wait(accfull, 0);
            tc_fence_after();
        }
        // partial[slab][column group of 4][row][4]: the 32 rows of a warp write 512 contiguous bytes per store instruction
        // (row-major rows of 1 KB made every store touch 32 different 128-byte lines: the epilogue was LSU-bound)
        float* obase = p.partial + slab * (size_t)p.N * P_C + (size_t)(r < p.N ? r : 0) * 4;
#pragma unroll 1
        for (int c0 = 0; c0 < P_C; c0 += 32) {
            uint32_t v[32];
            if (ntiles > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = 0u;
            }
            if (r < p.N) {
#pragma unroll
                for (int c = 0; c < 32; c += 4)
                    *reinterpret_cast<uint4*>(obase + (size_t)((c0 + c) >> 2) * p.N * 4) = make_uint4(v[c], v[c + 1], v[c + 2], v[c + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    DBG(13);
    if (warp == 1) tmem_dealloc<P_TMEM_COLS>(tmem_base);
}
}  // namespace pf
PF_DEFINE_DBG_SETTER(set_dbg_pool)
namespace pf {

// sum over splits in a fixed order: pooled[g][n][c], count[b][n].  Thread = (unit, column group, row), row fastest: the
// partials are column-group-major ([G][S][64][N][4]).
__global__ void __launch_bounds__(256) pool_reduce_kernel(const float* __restrict__ partial,
                                                          const float* __restrict__ cntp, float* __restrict__ pooled,
                                                          float* __restrict__ count, int G, int B, int N, int S,
                                                          const float* __restrict__ addend, int out_rows) {
    pdl_wait();
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= G * 64 * N) return;
    const int n = idx % N, cg = (idx / N) & 63, g = idx / (64 * N);
    const float4* src = reinterpret_cast<const float4*>(partial + ((size_t)g * S) * N * P_C) + (size_t)cg * N + n;
    const size_t stride4 = (size_t)N * P_C / 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = 0;
    for (; s + 6 <= S; s += 6) {
        float4 v[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) v[i] = __ldg(src + (size_t)(s + i) * stride4);
#pragma unroll
        for (int i = 0; i < 6; ++i) acc.x += v[i].x, acc.y += v[i].y, acc.z += v[i].z, acc.w += v[i].w;
    }
    for (; s < S; ++s) {
        const float4 v = __ldg(src + (size_t)s * stride4);
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    if (addend) {   // KernelHead: proposal_feats = init_kernels.weight + obj_feats (kernel_head.py:324-326)
        const float4 a = __ldg(reinterpret_cast<const float4*>(addend + (size_t)n * P_C) + cg);
        acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
    }
    reinterpret_cast<float4*>(pooled + ((size_t)g * out_rows + n) * P_C)[cg] = acc;
    if (cg == 0 && g < B && count) {
        float k = 0.f;
        for (int s2 = 0; s2 < S; ++s2) k += cntp[((size_t)g * S + s2) * N + n];
        count[g * N + n] = k;
    }
}

}  // namespace pf

extern "C" int pf_pool_splits(int B, int n_branch, int HW) {
    if (B <= 0 || n_branch <= 0 || HW <= 0) return 0;
    int sms = 148;
    if (pf::check_device() == PF_OK) sms = pf::num_sms();
    const int tiles = (HW + pf::P_BHW - 1) / pf::P_BHW;
    int S = sms / (B * n_branch);
    if (S < 1) S = 1;
    if (S > tiles) S = tiles;
    return S;
}

namespace pf {
int mask_pool_window(const uint16_t* feats, const uint32_t* bits, float* partial, float* cntp, int Btot, int b0, int B,
                     int N, int HW, int HWp, int n_branch, int S, int early_feats, void* stream);
}
extern "C" int pf_mask_pool(const uint16_t* feats, const uint32_t* bits, float* partial, float* cntp, int B, int N,
                            int HW, int HWp, int n_branch, int S, void* stream) {
    return pf::mask_pool_window(feats, bits, partial, cntp, B, 0, B, N, HW, HWp, n_branch, S, 0, stream);
}

// feats: the FULL [n_branch][Btot][256][HWp] tensor; bits / partial / cntp: buffers of the window's B images
int pf::mask_pool_window(const uint16_t* feats, const uint32_t* bits, float* partial, float* cntp, int Btot, int b0, int B,
                         int N, int HW, int HWp, int n_branch, int S, int early_feats, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(Btot >= B && b0 >= 0 && b0 + B <= Btot, PF_ERR_ARG, "pf_mask_pool: bad batch window %d+%d of %d", b0, B, Btot);
    PF_REQUIRE(feats && bits && partial && cntp, PF_ERR_ARG, "pf_mask_pool: null pointer");
    PF_REQUIRE(B > 0 && N > 0 && N <= PF_MAX_N && HW > 0 && (n_branch == 1 || n_branch == 2), PF_ERR_ARG,
               "pf_mask_pool: bad shape B=%d N=%d HW=%d n_branch=%d", B, N, HW, n_branch);
    PF_REQUIRE(HWp >= HW && HWp % 8 == 0, PF_ERR_ALIGN, "pf_mask_pool: HWp=%d must be >= HW and a multiple of 8", HWp);
    const int tiles = (HW + P_BHW - 1) / P_BHW;
    PF_REQUIRE(S >= 1 && S <= tiles, PF_ERR_ARG, "pf_mask_pool: S=%d out of range [1,%d]", S, tiles);
    PF_REQUIRE((reinterpret_cast<uintptr_t>(partial) & 15) == 0, PF_ERR_ALIGN, "pf_mask_pool: partial not 16-byte aligned");

    CUtensorMap tmap;
    if (int e = make_tmap_bf16_2d(&tmap, feats, (uint64_t)n_branch * Btot * P_C, (uint64_t)HW, (uint64_t)HWp, P_C, P_BHW)) return e;
    PoolParams p;
    p.Btot = Btot, p.b0 = b0, p.early_feats = early_feats;
    p.bits = bits, p.partial = partial, p.cntp = cntp;
    p.N = N, p.HW = HW, p.words = (HW + 31) / 32, p.B = B, p.S = S, p.tiles_per_unit = tiles;
    cudaError_t e = cudaFuncSetAttribute(pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM);
    if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "pool smem attribute: %s", cudaGetErrorString(e));
    return launch_pdl("pool_kernel", pool_kernel, dim3(n_branch * B * S), dim3(P_THREADS), P_SMEM,
                      static_cast<cudaStream_t>(stream), tmap, p);
}

extern "C" int pf_pool_reduce(const float* partial, const float* cntp, float* pooled, float* count, int B, int N,
                              int n_branch, int S, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(partial && cntp && pooled, PF_ERR_ARG, "pf_pool_reduce: null pointer");
    return launch_pdl("pool_reduce_kernel", pool_reduce_kernel, dim3((n_branch * B * N * 64 + 255) / 256), dim3(256), 0,
                      static_cast<cudaStream_t>(stream), partial, cntp, pooled, count, n_branch * B, B, N, S,
                      (const float*)nullptr, N);
}

// KernelHead._decode_init_proposals, the part that is "K1 with N = num_proposals" (polyphonic/kernel_head.py:313-336):
//   proposal_feats[b][n] = init_kernels.weight[n] + sum_hw 1[mask_preds[b][n][hw] > 0] * x_feats[b][:, hw]   n < P
//   proposal_feats[b][P + j] = conv_seg.weight[num_thing_classes + j]                                        (stuff kernels)
// `partial` / `cntp` come from pf_mask_pool(n_branch = 1) over the P proposal masks.
extern "C" int pf_init_proposals(const float* partial, const float* cntp, const float* init_kernels,
                                 const float* stuff_kernels, float* proposal_feats, int B, int P, int n_stuff, int S,
                                 void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(partial && cntp && init_kernels && proposal_feats && (n_stuff == 0 || stuff_kernels), PF_ERR_ARG,
               "pf_init_proposals: null pointer");
    PF_REQUIRE(B > 0 && P > 0 && n_stuff >= 0 && P + n_stuff <= PF_MAX_N && S > 0, PF_ERR_ARG, "pf_init_proposals: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int N = P + n_stuff;
    if (int e = launch_pdl("pool_reduce_kernel", pool_reduce_kernel, dim3((B * P * 64 + 255) / 256), dim3(256), 0, st, partial, cntp,
                           proposal_feats, (float*)nullptr, B, B, P, S, init_kernels, N))
        return e;
    for (int b = 0; b < B && n_stuff > 0; ++b) {   // the same stuff kernels for every image
        cudaError_t e = cudaMemcpyAsync(proposal_feats + ((size_t)b * N + P) * P_C, stuff_kernels,
                                        (size_t)n_stuff * P_C * sizeof(float), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "pf_init_proposals: stuff kernel copy: %s", cudaGetErrorString(e));
    }
    return PF_OK;
}
