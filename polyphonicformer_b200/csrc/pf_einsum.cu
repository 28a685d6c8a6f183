// K3 -- kernel x feature-map einsum (polyphonic/kernel_update_head.py:308-334 of the reference):
//     logits[g][n][hw] = sum_c kern[g][n][c] * feats[g][c][hw] + kbias[g][n]
// as a batched [128 x 256] x [256 x HW] GEMM on the tcgen05 tensor cores.
//
//   A  = kern[g] as bf16 hi + bf16 lo (the fp32 kernel split in two by its producer; two MMAs per K step, fp32
//        accumulate), resident in TENSOR MEMORY for the whole CTA (2 x 128 columns: row n in lane n, K packed two
//        bf16 per 32-bit column), written once by the epilogue warps with tcgen05.st.  With A in shared memory
//        (SS mode) every MMA re-read 4 KB of A and the kernel was paced by those reads (~87 cycles per
//        128x64x16 MMA, in-kernel timeline), not by HBM; TS mode reads only the 2 KB feature slice per MMA;
//   B  = feats[g][:, hw0:hw0+128] bf16, TMA-loaded as two [256 c][64 hw] boxes = MN-major operand, 3-stage ring
//        (192 KB in flight per SM);
//   D  = [128 lanes][128 columns] fp32 in TMEM, 2 accumulator buffers so the epilogue overlaps the next tile.  The
//        128-pixel tile = MMA N 128 halves the MMA issues per pixel against 64-pixel tiles (an issued MMA costs ~48
//        cycles whatever its N: with N = 64 the logits variant was issue-bound at 0.70 of the HBM peak);
//   epilogue: tcgen05.ld -> + bias -> the packed sign bits consumed by the next stage's pooling (thread = kernel row
//             n, 32 consecutive pixels = one u32 word) and/or fp32 logits.  Logits leave through shared memory:
//             each thread writes its 32 pixels into a 128-byte-swizzled [128 rows][32 px] tile, one thread issues a
//             TMA tensor store of the tile (rows >= N are clipped by the tensor map), two tiles in flight.  (Direct
//             per-thread stores touch 32 different 128-byte lines per warp instruction and were LSU-bound.)
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
// The kernel is HBM-bound (59 FLOP/B against a ridge of ~213): see DESIGN.md for the roofline.
#include "pf_internal.h"
#include "pf_sm100.cuh"
#include "pf_debug.cuh"

namespace pf {

constexpr int E_C = PF_C;            // 256 = K of the GEMM
constexpr int E_BHW = 64;            // pixels per tile (= N of the MMA, one 128-byte swizzle atom of bf16)
constexpr int E_STAGES = 6;          // feature ring depth in 64-pixel boxes (3 stages of 128-pixel tiles)
constexpr int E_OUT_BYTES = 128 * 32 * 4;   // one staged half tile: [128 rows][32 px] fp32
constexpr int E_ACC = 2;             // 2 * E_ACC * 64 = 256 TMEM columns of accumulators: 4 buffers of 64-pixel tiles
                                     // or 2 buffers of 128-pixel tiles
constexpr int E_TMEM_A = 2 * E_ACC * E_BHW; // first TMEM column of A hi; A lo follows 128 columns later
constexpr int E_TMEM_COLS = 512;            // 256 accumulator + 2 x 128 A columns: the whole tensor memory
constexpr int E_B_BYTES = E_C * E_BHW * 2;  // 32768 per stage
constexpr int E_THREADS = 192;
constexpr int E_SMEM = E_STAGES * E_B_BYTES + 2 * E_OUT_BYTES + 256 /*barriers*/ + 1024 /*alignment slack*/;
static_assert(E_SMEM <= 232448, "shared memory budget of one CTA");

struct EinsumParams {
    const uint16_t* kern;   // [G][2][N][256] bf16 hi / lo planes
    const float* kbias;  // [G][N]
    float* logits;       // [G][N][HW] or null
    uint32_t* bits;      // [B][WORDS][128] or null
    int N, HW, words, B;
    int tiles_per_unit, ctas_per_unit;
    int Btot, b0;        // batch window inside the [2][Btot] feature / logits tensors
    int unit0;           // first unit of this launch (B: depth branch only)
    int early_feats;     // the feature maps were complete before the PREVIOUS kernel started: prefetch them before pdl_wait
    int kdiv;            // kernel set of unit u = u / kdiv (1: one set per unit; B: one set shared by the B images of a map)
    int fmod;            // feature map of unit u = u % fmod (pf_kernel_head: the two 128-row halves of a conv share a map)
    int out_blocked;     // 0, or the number of 32-pixel blocks per unit: logits leave as [unit][block][128 rows][32 px]
    float2* stats;       // STATS: [units][ctas_per_unit][128] per-row (sum, sum of squares) over the CTA's pixels
    // apply (pf_fpn_pred's second pass; TMA_OUT = false, logits = bits = null): row n of unit u = half * 3B + v is channel
    // half * 128 + n of map-image v; y = ReLU(acc * scale + shift) leaves as bf16 [3B][256][HWp] (+ optional fp32
    // [3B][256][HW]): GroupNorm + ReLU inside the convolution's epilogue, no fp32 intermediate in memory
    const float2* apply_affine;   // [3B][256] (scale, shift), or null
    uint16_t* apply_out;
    float* apply_out32;
    int apply_units, HWp;         // 3B; row pitch of apply_out
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2,
                                             uint64_t hint) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
                 :
                 : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
                 : "memory");
}
__device__ __forceinline__ void epi_bar4() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps

// TMA_OUT: fp32 logits leave as TMA tensor stores of staged tiles (tmap_out over [units][N][HW]).
// Tile width: 128 pixels (two TMA boxes, MMA N = 128); the two logits staging tiles sit behind the ring.
// STATS (pf_kernel_head's convolutions): every epilogue thread also accumulates the sum and the sum of squares of its
// row over the CTA's pixels -- the GroupNorm statistics, without a second pass over the output.
template <bool TMA_OUT, bool STATS = false>
__global__ void __launch_bounds__(E_THREADS, 1)
einsum_kernel(const __grid_constant__ CUtensorMap tmap_feats, const __grid_constant__ CUtensorMap tmap_out,
              const EinsumParams p) {
    constexpr int TILE = 128;                                        // pixels per tile = MMA N = accumulator columns
    constexpr int NBOX = TILE / E_BHW;
    constexpr int STAGE_BYTES = NBOX * E_B_BYTES;
    constexpr int STAGES = E_STAGES / NBOX;
    constexpr int NACC = (2 * E_ACC * E_BHW) / TILE;                 // accumulator buffers in 256 TMEM columns
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sB = smem;
    uint8_t* sOut = sB + E_STAGES * E_B_BYTES;       // TMA_OUT only: 2 x [128][32] fp32, 1024-byte aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + 2 * E_OUT_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + E_STAGES;
    uint64_t* tfull = bars + 2 * E_STAGES;
    uint64_t* tempty = bars + 2 * E_STAGES + 4;
    uint64_t* abar = bars + 2 * E_STAGES + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * E_STAGES + 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = p.unit0 + blockIdx.x / p.ctas_per_unit;
    const int j = blockIdx.x % p.ctas_per_unit;
    const int tiles_per_unit = (p.HW + TILE - 1) / TILE;
    const int tile_begin = (int)((long long)j * tiles_per_unit / p.ctas_per_unit);
    const int tile_end = (int)((long long)(j + 1) * tiles_per_unit / p.ctas_per_unit);
    const int ntiles = tile_end - tile_begin;
    const int gunit = (unit / p.B) * p.Btot + p.b0 + unit % p.B;   // unit inside the full-batch feature / logits tensors
    const int funit = gunit % p.fmod;                              // its feature map
    const int kunit = unit / p.kdiv;                               // its kernel set
    long long* dbg = dbg_claim_all(TMA_OUT ? 21 : 20);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_feats);
        if (TMA_OUT) tma_prefetch_desc(&tmap_out);
        mbar_init(abar, 128);   // every epilogue thread arrives once its row of A is in tensor memory
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < NACC; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<E_TMEM_COLS>(tmem_slot);

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);   // warp-uniform (REDUX -> uniform register)
    DBG(1);
    pdl_launch_dependents();

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            // inside the decode loop the feature tiles are long-complete inputs: fill the ring before the grid
            // dependency resolves
            if (!p.early_feats) pdl_wait();
            // inside the decode loop x_feats stays L2-resident across kernels; a stand-alone call streams
            const uint64_t fhint = (p.early_feats && unit < p.B) ? kEvictLast : kEvictFirst;
            for (int i = 0; i < ntiles; ++i) {
                const int s = i % STAGES;
                if (i >= STAGES) mbar_wait(&empty[s], ((i / STAGES) & 1) ^ 1);
                if (i == ntiles - 1) DBG(3);
                mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
#pragma unroll
                for (int bx = 0; bx < NBOX; ++bx)   // a box that starts beyond HW is zero-filled by the TMA unit
                    tma_load_2d(sB + s * STAGE_BYTES + bx * E_B_BYTES, &tmap_feats, &full[s],
                                (tile_begin + i) * TILE + bx * E_BHW, funit * E_C, fhint);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: the whole warp runs the loop, one elected lane issues =================
        constexpr uint32_t idesc = make_idesc_bf16(128, TILE, /*a_mn=*/0, /*b_mn=*/1);
        const uint32_t a_hi = tmem_base + E_TMEM_A, a_lo = a_hi + 128;
        mbar_wait(abar, 0);
        tc_fence_after();
        for (int i = 0; i < ntiles; ++i) {
            const int s = i % STAGES, a = i % NACC;
            mbar_wait(&tempty[a], ((i / NACC) & 1) ^ 1);
            mbar_wait(&full[s], (i / STAGES) & 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + a * TILE;
            // B (MN-major): 16 K-rows of 128 bytes per K step -> the start-address field advances by 2048 >> 4;
            // successive 64-pixel chunks of N are one box (32 KB) apart (leading byte offset)
            const uint64_t db0 = make_smem_desc_sw128(smem_u32(sB + s * STAGE_BYTES), E_B_BYTES, 1024);
#pragma unroll
            for (int kb = 0; kb < E_C / 16; ++kb) {
                // A (tensor memory): 16 K values = 8 columns per step; hi and lo planes into the same accumulator
                umma_bf16_ts_warp(d_tmem, a_hi + kb * 8, db0 + (uint64_t)(kb * 128), idesc, kb > 0);
                umma_bf16_ts_warp(d_tmem, a_lo + kb * 8, db0 + (uint64_t)(kb * 128), idesc, 1);
            }
            umma_commit_warp(&empty[s]);
            umma_commit_warp(&tfull[a]);
        }
    } else {
        // ================= epilogue (warps 2..5 -> TMEM lane quadrants 2,3,0,1) =================
        pdl_wait();   // kern / kbias are read and bits / logits written only after the previous kernels have completed
        const int q = warp & 3;
        const int n = q * 32 + lane;
        const bool row_ok = n < p.N;
        {   // A operand -> tensor memory: this thread's kernel row, hi plane then lo plane (rows >= N: zeros)
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const uint4* src = reinterpret_cast<const uint4*>(p.kern + (((size_t)kunit * 2 + h) * p.N + (row_ok ? n : 0)) * E_C);
#pragma unroll 1
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    uint32_t v[32];
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const uint4 t = row_ok ? __ldg(src + c0 / 4 + jj) : make_uint4(0u, 0u, 0u, 0u);
                        v[4 * jj] = t.x, v[4 * jj + 1] = t.y, v[4 * jj + 2] = t.z, v[4 * jj + 3] = t.w;
                    }
                    tmem_st32(tmem_base + ((uint32_t)(q * 32) << 16) + E_TMEM_A + h * 128 + c0, v);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(abar);
        }
        const float bias = (row_ok && p.kbias) ? __ldg(p.kbias + (size_t)kunit * p.N + n) : 0.f;
        float* orow = p.logits ? p.logits + ((size_t)gunit * p.N + (row_ok ? n : 0)) * p.HW : nullptr;
        const bool vec_ok = (p.HW & 3) == 0;
        uint32_t* brow = (p.bits && unit < p.B) ? p.bits + (size_t)unit * p.words * 128 + n : nullptr;
        float st_s[2] = {0.f, 0.f}, st_q[2] = {0.f, 0.f};
        const int half = unit / p.apply_units, v3 = unit - half * p.apply_units, ch = half * 128 + n;   // apply mode only
        const float2 af = p.apply_affine ? __ldg(p.apply_affine + (size_t)v3 * 256 + ch) : make_float2(1.f, 0.f);
        for (int i = 0; i < ntiles; ++i) {
            const int a = i % NACC;
            mbar_wait(&tfull[a], (i / NACC) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * TILE;
            const int hw0 = (tile_begin + i) * TILE;
#pragma unroll
            for (int h = 0; h < TILE / 32; ++h) {
                uint32_t v[32];
                tmem_ld32(taddr + h * 32, v);
                tmem_ld_wait();
                if (h == TILE / 32 - 1) {   // the accumulator buffer is free once its last chunk is in registers
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[a]);
                }
                const int hwb = hw0 + h * 32;
                if (hwb >= p.HW) continue;
                uint32_t word = 0;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float f = __uint_as_float(v[c]) + bias;
                    v[c] = __float_as_uint(f);
                    word |= (f > 0.f ? 1u : 0u) << c;
                    if (STATS) st_s[c & 1] += f, st_q[c & 1] = fmaf(f, f, st_q[c & 1]);   // pixels beyond HW are exact zeros
                }
                const int valid = p.HW - hwb;  // > 0
                if (valid < 32) word &= (1u << valid) - 1u;
                if (brow) brow[(size_t)(hwb >> 5) * 128] = word;
                if (TMA_OUT) {
                    // staged tile (2*i + h) & 1: wait until the store issued two half tiles ago has read it
                    uint8_t* tile = sOut + (h & 1) * E_OUT_BYTES;   // TILE / 32 is even: the two tiles alternate across tiles too
                    if (threadIdx.x == 64) tma_store_wait_read<1>();
                    epi_bar4();
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<uint4*>(tile + sw128_offset(n, c)) =
                            make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                    fence_proxy_async_smem();
                    epi_bar4();
                    if (threadIdx.x == 64) {
                        // evict-last: the x2 up-sampling that follows reads these logits back, ideally from L2
                        if (p.out_blocked) tma_store_3d(&tmap_out, tile, 0, 0, gunit * p.out_blocked + (hwb >> 5), kEvictNormal);
                        else tma_store_3d(&tmap_out, tile, hwb, 0, gunit, kEvictLast);
                        tma_store_commit();
                    }
                } else if (p.apply_affine) {
                    uint32_t pk[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        float a0 = fmaxf(fmaf(__uint_as_float(v[2 * c]), af.x, af.y), 0.f);
                        float a1 = fmaxf(fmaf(__uint_as_float(v[2 * c + 1]), af.x, af.y), 0.f);
                        if (2 * c >= valid) a0 = 0.f;          // pad columns [HW, HWp) are written as zero
                        if (2 * c + 1 >= valid) a1 = 0.f;
                        v[2 * c] = __float_as_uint(a0), v[2 * c + 1] = __float_as_uint(a1);
                        pk[c] = pack_bf16x2(a0, a1);
                    }
                    uint16_t* dst = p.apply_out + ((size_t)v3 * 256 + ch) * p.HWp + hwb;
                    const int validp = p.HWp - hwb;            // > 0, a multiple of 8
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (8 * c < validp) reinterpret_cast<uint4*>(dst)[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                    if (p.apply_out32) {
                        float* d32 = p.apply_out32 + ((size_t)v3 * 256 + ch) * p.HW + hwb;
                        if (vec_ok && valid >= 32) {
#pragma unroll
                            for (int c = 0; c < 32; c += 4)
                                __stcs(reinterpret_cast<float4*>(d32 + c), make_float4(__uint_as_float(v[c]), __uint_as_float(v[c + 1]),
                                                                                       __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])));
                        } else {
#pragma unroll
                            for (int c = 0; c < 32; ++c)
                                if (c < valid) d32[c] = __uint_as_float(v[c]);
                        }
                    }
                } else if (orow && row_ok) {
                    float* dst = orow + hwb;
                    if (vec_ok && valid >= 32) {
#pragma unroll
                        for (int c = 0; c < 32; c += 4)
                            __stcs(reinterpret_cast<float4*>(dst + c),
                                   make_float4(__uint_as_float(v[c]), __uint_as_float(v[c + 1]),
                                               __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])));
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            if (c < valid) dst[c] = __uint_as_float(v[c]);
                    }
                }
            }
        }
        if (STATS)
            p.stats[((size_t)unit * p.ctas_per_unit + j) * 128 + n] = make_float2(st_s[0] + st_s[1], st_q[0] + st_q[1]);
    }
    if (TMA_OUT && threadIdx.x == 64) tma_store_wait_read<0>();   // shared memory must outlive the last store's read
    tc_fence_before();
    __syncthreads();
    DBG(13);
    if (warp == 1) tmem_dealloc<E_TMEM_COLS>(tmem_base);
}

}  // namespace pf
PF_DEFINE_DBG_SETTER(set_dbg_einsum)

// fp32 kernels [G][N][256] -> bf16 hi / lo [G][2][N][256] (hi = bf16(x), lo = bf16(x - hi))
namespace pf {
__global__ void split_kernels_kernel(const float* __restrict__ kern, uint16_t* __restrict__ out, int N, int total4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // one float4
    if (i >= total4) return;
    const int per_unit4 = N * E_C / 4;
    const int unit = i / per_unit4, r = i % per_unit4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(kern) + i);
    const float h0 = bf16_round(v.x), h1 = bf16_round(v.y), h2 = bf16_round(v.z), h3 = bf16_round(v.w);
    uint2* dst_hi = reinterpret_cast<uint2*>(out) + (size_t)(unit * 2) * per_unit4 + r;
    uint2* dst_lo = dst_hi + per_unit4;
    *dst_hi = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
    *dst_lo = make_uint2(pack_bf16x2(v.x - h0, v.y - h1), pack_bf16x2(v.z - h2, v.w - h3));
}
}  // namespace pf

extern "C" int pf_split_kernels(const float* kern, uint16_t* kern_split, int n_units, int N, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(kern && kern_split && n_units > 0 && N > 0 && N <= PF_MAX_N, PF_ERR_ARG, "pf_split_kernels: bad argument");
    PF_REQUIRE(((reinterpret_cast<uintptr_t>(kern) | reinterpret_cast<uintptr_t>(kern_split)) & 15) == 0, PF_ERR_ALIGN,
               "pf_split_kernels: pointers must be 16-byte aligned");
    const int total4 = n_units * N * E_C / 4;
    split_kernels_kernel<<<(total4 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(kern, kern_split, N, total4);
    PF_CHECK_LAUNCH("split_kernels_kernel");
    return PF_OK;
}

namespace pf {
int mask_einsum_window(const uint16_t* feats, const uint16_t* kern, const float* kbias, float* logits, uint32_t* bits_out,
                       int Btot, int b0, int B, int N, int HW, int HWp, int n_units, int branch0, int early_feats,
                       void* stream);
}
extern "C" int pf_mask_einsum(const uint16_t* feats, const uint16_t* kern, const float* kbias, float* logits,
                              uint32_t* bits_out, int B, int N, int HW, int HWp, int n_units, void* stream) {
    return pf::mask_einsum_window(feats, kern, kbias, logits, bits_out, B, 0, B, N, HW, HWp, n_units, 0, 0, stream);
}

// feats / logits: the FULL [2][Btot][..] tensors; kern / kbias / bits_out: buffers of the window's B images.
// branch0 = 1 with n_units = B runs the depth branch alone (units B .. 2B-1).
int pf::mask_einsum_window(const uint16_t* feats, const uint16_t* kern, const float* kbias, float* logits,
                           uint32_t* bits_out, int Btot, int b0, int B, int N, int HW, int HWp, int n_units, int branch0,
                           int early_feats, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(Btot >= B && b0 >= 0 && b0 + B <= Btot, PF_ERR_ARG, "pf_mask_einsum: bad batch window %d+%d of %d", b0, B, Btot);
    PF_REQUIRE(feats && kern && kbias, PF_ERR_ARG, "pf_mask_einsum: null input");
    PF_REQUIRE(logits || bits_out, PF_ERR_ARG, "pf_mask_einsum: no output requested");
    PF_REQUIRE(B > 0 && N > 0 && N <= PF_MAX_N && HW > 0, PF_ERR_ARG, "pf_mask_einsum: bad shape B=%d N=%d HW=%d", B, N, HW);
    PF_REQUIRE(n_units == B || n_units == 2 * B, PF_ERR_ARG, "pf_mask_einsum: n_units must be B or 2B");
    PF_REQUIRE(branch0 == 0 || (branch0 == 1 && n_units == B && !bits_out), PF_ERR_ARG, "pf_mask_einsum: bad branch selection");
    PF_REQUIRE(HWp >= HW && HWp % 8 == 0, PF_ERR_ALIGN, "pf_mask_einsum: HWp=%d must be >= HW and a multiple of 8", HWp);
    PF_REQUIRE((reinterpret_cast<uintptr_t>(kern) & 15) == 0, PF_ERR_ALIGN, "pf_mask_einsum: kern_split not 16-byte aligned");
    PF_REQUIRE(!logits || (reinterpret_cast<uintptr_t>(logits) & 15) == 0, PF_ERR_ALIGN, "pf_mask_einsum: logits not 16-byte aligned");

    CUtensorMap tmap;
    if (int e = make_tmap_bf16_2d(&tmap, feats, (uint64_t)(branch0 + n_units / B) * Btot * E_C, (uint64_t)HW, (uint64_t)HWp, E_C, E_BHW)) return e;

    EinsumParams p;
    p.kern = kern, p.kbias = kbias, p.logits = logits, p.bits = bits_out;
    p.N = N, p.HW = HW, p.words = (HW + 31) / 32, p.B = B;
    p.Btot = Btot, p.b0 = b0, p.early_feats = early_feats, p.unit0 = branch0 * B;
    p.kdiv = 1, p.fmod = 0x7fffffff, p.stats = nullptr, p.out_blocked = 0;
    p.apply_affine = nullptr, p.apply_out = nullptr, p.apply_out32 = nullptr, p.apply_units = 1, p.HWp = HWp;
    // fp32 logits through TMA stores when the row pitch allows it (HW * 4 bytes must be a 16-byte multiple)
    const bool tma_out = logits && (HW % 4) == 0;
    const int tile = 128;   // must match einsum_kernel::TILE
    p.tiles_per_unit = (HW + tile - 1) / tile;
    int cpu = num_sms() / n_units;
    if (cpu < 1) cpu = 1;
    if (cpu > p.tiles_per_unit) cpu = p.tiles_per_unit;
    p.ctas_per_unit = cpu;

    CUtensorMap tmap_o = tmap;
    if (tma_out)
        if (int e = make_tmap_f32_3d(&tmap_o, logits, (uint64_t)(branch0 + n_units / B) * Btot, (uint64_t)N, (uint64_t)HW, 128, 32)) return e;
    auto kern_fn = tma_out ? einsum_kernel<true, false> : einsum_kernel<false, false>;
    cudaError_t ea = cudaFuncSetAttribute(kern_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, E_SMEM);
    if (ea != cudaSuccess) return set_error(PF_ERR_CUDA, "einsum smem attribute: %s", cudaGetErrorString(ea));
    return launch_pdl("einsum_kernel", kern_fn, dim3(n_units * cpu), dim3(E_THREADS), E_SMEM,
                      static_cast<cudaStream_t>(stream), tmap, tmap_o, p);
}

// pf_kernel_head's three 1x1 convolutions (kernel_head.py:250-251, 264-265, 277-278; ConvModule without bias) as ONE
// launch of the einsum kernel: unit u = half * 3B + map * B + b computes rows [128 * half, 128 * half + 128) of
// W_map . maps[map % n_inputs][b]; conv_split holds the six (half, map) row blocks as bf16 hi / lo planes [6][2][128][256].
// Y: fp32 [6B][ceil(HW/32)][128][32] (pixel-blocked).
int pf::conv1x1_ctas_per_unit(int B, int HW) {
    const int tile = 128;
    int cpu = num_sms() / (6 * B);
    const int tiles = (HW + tile - 1) / tile;
    return cpu < 1 ? 1 : (cpu > tiles ? tiles : cpu);
}

int pf::conv1x1_maps(const uint16_t* maps, int n_inputs, const uint16_t* conv_split, float* Y, float2* stats, int B, int HW,
                     int HWp, void* stream) {
    using namespace pf;
    const int n_units = 6 * B;
    CUtensorMap tmap;
    if (int e = make_tmap_bf16_2d(&tmap, maps, (uint64_t)n_inputs * B * E_C, (uint64_t)HW, (uint64_t)HWp, E_C, E_BHW)) return e;
    EinsumParams p;
    p.kern = conv_split, p.kbias = nullptr, p.logits = Y, p.bits = nullptr;
    p.N = 128, p.HW = HW, p.words = (HW + 31) / 32, p.B = n_units;
    p.Btot = n_units, p.b0 = 0, p.early_feats = 0, p.unit0 = 0;
    p.kdiv = B, p.fmod = n_inputs * B, p.stats = stats;   // n_inputs = 1: the three convolutions read the same map
    p.apply_affine = nullptr, p.apply_out = nullptr, p.apply_out32 = nullptr, p.apply_units = 3 * B, p.HWp = HWp;
    // blocked output: 16 KB [128 rows][32 px] blocks, written whole by the TMA stores (no row-pitch constraint) and
    // read back by head_apply_kernel with a compile-time channel stride
    const int nblk = (HW + 31) / 32;
    p.out_blocked = nblk;
    const int tile = 128;
    p.tiles_per_unit = (HW + tile - 1) / tile;
    const int cpu = conv1x1_ctas_per_unit(B, HW);
    p.ctas_per_unit = cpu;
    CUtensorMap tmap_o;
    if (int e = make_tmap_f32_3d(&tmap_o, Y, (uint64_t)n_units * nblk, 128, 32, 128, 32)) return e;
    auto kern_fn = einsum_kernel<true, true>;
    cudaError_t ea = cudaFuncSetAttribute(kern_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, E_SMEM);
    if (ea != cudaSuccess) return set_error(PF_ERR_CUDA, "einsum smem attribute: %s", cudaGetErrorString(ea));
    return launch_pdl("einsum_kernel(conv1x1)", kern_fn, dim3(n_units * cpu), dim3(E_THREADS), E_SMEM,
                      static_cast<cudaStream_t>(stream), tmap, tmap_o, p);
}

// The same three convolutions WITHOUT the fp32 intermediate (pf_fpn_pred): pass 1 (affine == null) only accumulates the
// GroupNorm statistics; pass 2 recomputes the convolution and applies (scale, shift) + ReLU in its epilogue, writing the
// bf16 maps [3][B][256][HWp] (+ optional fp32 [3][B][256][HW]) directly.  The input map is read twice (2 x 67 MB at B = 4)
// instead of writing and re-reading 2 x 403 MB of fp32.
int pf::conv1x1_fused(const uint16_t* maps, int n_inputs, const uint16_t* conv_split, float2* stats, const float2* affine,
                      uint16_t* out, float* out32, int B, int HW, int HWp, void* stream) {
    using namespace pf;
    const int n_units = 6 * B;
    CUtensorMap tmap;
    if (int e = make_tmap_bf16_2d(&tmap, maps, (uint64_t)n_inputs * B * E_C, (uint64_t)HW, (uint64_t)HWp, E_C, E_BHW)) return e;
    EinsumParams p;
    p.kern = conv_split, p.kbias = nullptr, p.logits = nullptr, p.bits = nullptr;
    p.N = 128, p.HW = HW, p.words = (HW + 31) / 32, p.B = n_units;
    p.Btot = n_units, p.b0 = 0, p.early_feats = 0, p.unit0 = 0;
    p.kdiv = B, p.fmod = n_inputs * B, p.stats = stats, p.out_blocked = 0;
    p.apply_affine = affine, p.apply_out = out, p.apply_out32 = out32, p.apply_units = 3 * B, p.HWp = HWp;
    p.tiles_per_unit = (HW + 127) / 128;
    p.ctas_per_unit = conv1x1_ctas_per_unit(B, HW);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!affine) {
        auto kern_fn = einsum_kernel<false, true>;
        cudaError_t ea = cudaFuncSetAttribute(kern_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, E_SMEM);
        if (ea != cudaSuccess) return set_error(PF_ERR_CUDA, "einsum smem attribute: %s", cudaGetErrorString(ea));
        return launch_pdl("einsum_kernel(conv1x1 statistics)", kern_fn, dim3(n_units * p.ctas_per_unit), dim3(E_THREADS), E_SMEM, st,
                          tmap, tmap, p);
    }
    auto kern_fn = einsum_kernel<false, false>;
    cudaError_t ea = cudaFuncSetAttribute(kern_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, E_SMEM);
    if (ea != cudaSuccess) return set_error(PF_ERR_CUDA, "einsum smem attribute: %s", cudaGetErrorString(ea));
    return launch_pdl("einsum_kernel(conv1x1 + GN + ReLU)", kern_fn, dim3(n_units * p.ctas_per_unit), dim3(E_THREADS), E_SMEM, st, tmap,
                      tmap, p);
}
