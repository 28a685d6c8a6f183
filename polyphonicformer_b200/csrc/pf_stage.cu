// K2, fused -- the small-N block of one decoder stage as ONE launch:
//   KernelUpdator (polyphonic/funcs/kernel_updator.py:55-93), inter-kernel multi-head attention + LN, FFN + LN and the
//   cls / mask / depth FC heads (polyphonic/kernel_update_head.py:245-288), feat_transform folded (include/pf_decoder.h).
//
// One 8-CTA thread-block cluster per unit (= one image of one branch: 128 kernel rows x 256 features).  CTA `rank` of
// the cluster owns feature columns [32 rank, 32 rank + 32) of every 256-wide activation, hidden columns
// [256 rank, 256 rank + 256) of the FFN, and attention head `rank`.  A layer = one phase:
//     TMA (A = the full-width activation of the previous phase, bf16 hi / lo planes in the L2-resident arena;
//          W = the 32-row weight blocks of this CTA's output columns, boxes of [32][64] picked straight out of the
//          weight stacks)  ->  tcgen05.mma  D = Al*Wh + Ah*Wl + Ah*Wh, fp32 in TMEM  ->  epilogue, thread = kernel row
// and the phases are separated by cluster barriers (release / acquire at cluster scope: the arena slice a CTA wrote with
// ordinary stores is visible to its peers' TMA loads), not by kernel boundaries.  What never leaves the SM:
//   * LayerNorm statistics: (mean, M2) of every 32-column piece go to all 8 CTAs through DSMEM mailboxes, one cluster
//     barrier, Chan merge (as in the per-layer kernels of pf_update.cu);
//   * the residuals obj0 / obj1, LN(param_out), LN(input_out): column-local, parked in spare TENSOR MEMORY columns
//     (lane = kernel row) between the phases that produce and consume them;
//   * q / k / v of head `rank` and the whole attention of that head (shared memory).
// The weights of the next phase are prefetched into the ring while the current phase's epilogue and barrier run.
// Phases: prep | dyn+inp | gates | fc | qkv + attention | out-proj | ffn1 | ffn2 (split-K over the cluster) | reduce + LN
// | heads | kernels + kbias + cls.   Warp roles (576 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..17
// = 16 worker warps (TMEM lane quarter = warp % 4, 32-column chunk = (warp - 2) / 4).
#include <string.h>

#include <mutex>

#include "pf_internal.h"
#include "pf_sm100.cuh"
#include "pf_update.cuh"

namespace pf {

constexpr int G_THREADS = 576;
constexpr int G_CL = 8;                          // CTAs per cluster = column slices of 32
constexpr int G_NSTG = 3;
constexpr int G_PLANE = 128 * 64 * 2;            // 16384: one [128][64] bf16 box (A plane); W planes use <= 4 boxes of [32][64]
constexpr int G_WBOX = 32 * 64 * 2;              // 4096
constexpr int G_STAGE = 4 * G_PLANE;             // A hi | A lo | W hi | W lo
constexpr int G_BAR_OFF = G_NSTG * G_STAGE;      // 196608
constexpr int G_MAIL_OFF = G_BAR_OFF + 256;
constexpr int G_MAIL_BYTES = 2 * 128 * 9 * 8;
constexpr int G_ST4_OFF = G_MAIL_OFF + G_MAIL_BYTES;   // reduce phase: [4 chunks][128 rows] (mean, M2) of 8 columns
constexpr int G_SMEM_USED = G_ST4_OFF + 4 * 128 * 8;
constexpr int G_SMEM = G_SMEM_USED + 1024;
static_assert(G_SMEM <= 232448, "shared memory budget of one CTA");
// attention scratch aliases ring stage 2 (idle between the qkv MMAs and the out-proj loads)
constexpr int G_ATT_OFF = 2 * G_STAGE;
constexpr int G_KT_LD = 132;
static_assert((128 * 32 + 32 * G_KT_LD + 128 * 32) * 4 <= G_STAGE, "attention scratch must fit one ring stage");
// tensor memory columns: [0, 256) accumulators; spare columns hold row-private fp32 state between phases
constexpr int TC_PON = 256, TC_ION = 288, TC_RES = 320;

struct StageArgs {
    pf_stage_weights w;
    const float *partial, *cntp;     // pooling partials [2B][S][N][256], counts [2B][S][N]
    const float *obj_in, *dep_in;    // [B][N][256]
    float *obj_out, *dep_out;        // [B][N][256]  (must not alias the inputs)
    float* cls_out;                  // [B][N][num_classes] or null
    float* kern;                     // [2B][N][256] or null
    uint16_t* kern_split;            // [2B][2][N][256]
    float* kbias;                    // [2B][N]
    uint16_t* arena;
    float* part;                     // ffn2 split-K partials [2B][8][128][256]
    int B, N, S, cls_sigmoid;
};

struct GPass {
    int a_slot, nbox, ffn, k0, tcol;
    int whi[4], wlo[4];              // rows of the 32-row weight boxes (hi / lo plane) in the stack
};
struct GPhase {
    int npass;
    GPass p[2];
};
enum { PH_DUAL = 0, PH_GATE, PH_FC, PH_QKV, PH_OUT, PH_FFN1, PH_FFN2, PH_HEADS, PH_KERN };

__device__ __forceinline__ void set_pass(GPass& p, int a_slot, int tcol, int nbox, int lo_off, int r0, int r1 = 0, int r2 = 0,
                                         int r3 = 0) {
    p.a_slot = a_slot, p.nbox = nbox, p.ffn = 0, p.k0 = 0, p.tcol = tcol;
    p.whi[0] = r0, p.whi[1] = r1, p.whi[2] = r2, p.whi[3] = r3;
#pragma unroll
    for (int j = 0; j < 4; ++j) p.wlo[j] = p.whi[j] + lo_off;
}

// the GEMM of phase `ph` for CTA `r` of a cluster of branch `br` (row numbers: include/pf_decoder.h, struct pf_branch_weights)
__device__ __forceinline__ void get_phase(int ph, int r, int br, const pf_branch_weights& bw, int ffn, GPhase& g) {
    g.npass = 1;
    switch (ph) {
    case PH_DUAL:   // [param_in | param_out] of this CTA's 32 features from pooled', [input_in | input_out] from the kernel
        g.npass = 2;
        set_pass(g.p[0], SLOT_POOLED, 0, 2, 512, bw.dyn_w + 32 * r, bw.dyn_w + 256 + 32 * r);
        set_pass(g.p[1], SLOT_INP, 64, 2, 512, bw.inp_w + 32 * r, bw.inp_w + 256 + 32 * r);
        break;
    case PH_GATE: {  // gate rows are interleaved in blocks of 64: [input_gate 64 t .. | update_gate 64 t ..]
        const int base = bw.gate_w + 128 * (r >> 1) + 32 * (r & 1);
        set_pass(g.p[0], SLOT_GATEIN, 0, 2, 512, base, base + 64);
        break;
    }
    case PH_FC:
        set_pass(g.p[0], SLOT_MIX, 0, 1, 256, bw.fc_w + 32 * r);
        break;
    case PH_QKV:    // head r: its q, k and v rows of in_proj
        set_pass(g.p[0], SLOT_OBJ0, 0, 3, 768, bw.qkv_w + 32 * r, bw.qkv_w + 256 + 32 * r, bw.qkv_w + 512 + 32 * r);
        break;
    case PH_OUT:
        set_pass(g.p[0], SLOT_ATT, 0, 1, 256, bw.out_w + 32 * r);
        break;
    case PH_FFN1:   // hidden columns [256 r, 256 r + 256) as two 128-column tiles
        g.npass = 2;
        for (int t = 0; t < 2; ++t) {
            const int w0 = bw.ffn1_w + 256 * r + 128 * t;
            set_pass(g.p[t], SLOT_OBJ1, 128 * t, 4, ffn, w0, w0 + 32, w0 + 64, w0 + 96);
        }
        break;
    case PH_FFN2:   // all 256 output columns over this CTA's own 256 hidden channels (split-K over the cluster)
        g.npass = 2;
        for (int t = 0; t < 2; ++t) {
            const int w0 = bw.ffn2_w + 128 * t;
            set_pass(g.p[t], SLOT_HID0 + r, 128 * t, 4, 256, w0, w0 + 32, w0 + 64, w0 + 96);
            g.p[t].ffn = 1, g.p[t].k0 = 256 * r;
        }
        break;
    case PH_HEADS:
        if (br == 0) set_pass(g.p[0], SLOT_OBJ2, 0, 2, 512, bw.head_w + 32 * r, bw.head_w + 256 + 32 * r);   // cls_fcs | mask_fcs
        else set_pass(g.p[0], SLOT_OBJ2, 0, 1, 256, bw.head_w + 32 * r);                                    // depth_regs
        break;
    default:        // PH_KERN: fc_mask / fc_depth (folded); CTA 0 adds the logit-bias row, CTA 1 of the mask branch fc_cls
        set_pass(g.p[0], SLOT_HEAD1, 0, 1, 256, bw.kern_w + 32 * r);
        if (r == 0) {
            g.p[0].nbox = 2, g.p[0].whi[1] = bw.kbrow_w, g.p[0].wlo[1] = bw.kbrow_w + 128;
        }
        if (br == 0 && r == 1) {
            g.npass = 2;
            set_pass(g.p[1], SLOT_HEAD0, 64, 1, 128, bw.cls_w);
        }
        break;
    }
}

__device__ __forceinline__ void cluster_sync_all() {
    cluster_arrive();
    cluster_wait();
}
// bounded wait: a protocol bug becomes a trap (and an error code on the host) instead of a hung GPU
__device__ __forceinline__ void gwait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {
            printf("pf stage_kernel: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }   // the 16 worker warps
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld8f(uint32_t taddr, float (&y)[8]) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(v[i]);
}
__device__ __forceinline__ void tmem_st32f(uint32_t taddr, const float (&y)[32]) {
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(y[i]);
    tmem_st32(taddr, v);
    tmem_st_wait();
}
// v: global memory, same address in all lanes (one broadcast transaction per float4)
__device__ __forceinline__ void add_gvec32(float (&y)[32], const float* v) {
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
        const float4 t = ld4(v + c);
        y[c] += t.x, y[c + 1] += t.y, y[c + 2] += t.z, y[c + 3] += t.w;
    }
}
__device__ __forceinline__ void fma_gvec32(float (&y)[32], float s, const float* v) {
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
        const float4 t = ld4(v + c);
        y[c] += s * t.x, y[c + 1] += s * t.y, y[c + 2] += s * t.z, y[c + 3] += s * t.w;
    }
}
// y = (y - mean) * rstd * gamma + beta with ln = {gamma[256], beta[256]}, columns col0 .. col0 + 31
__device__ __forceinline__ void ln_apply32(float (&y)[32], float mean, float rstd, const float* ln, int col0) {
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
        const float4 ga = ld4(ln + col0 + c), be = ld4(ln + 256 + col0 + c);
        y[c] = (y[c] - mean) * rstd * ga.x + be.x, y[c + 1] = (y[c + 1] - mean) * rstd * ga.y + be.y;
        y[c + 2] = (y[c + 2] - mean) * rstd * ga.z + be.z, y[c + 3] = (y[c + 3] - mean) * rstd * ga.w + be.w;
    }
}
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// 32 / 8 values of one row -> bf16 hi / lo planes of arena slot `slot`, columns col0 ..; rows >= N are written as zeros
__device__ __forceinline__ void planes32(const float (&y)[32], bool rok, uint16_t* arena, int unit, int slot, int row, int col0) {
    uint16_t* hi = arena_row(arena, unit, slot, 0, row) + col0;
    uint16_t* lo = arena_row(arena, unit, slot, 1, row) + col0;
#pragma unroll
    for (int c = 0; c < 32; c += 4)
        store_planes4(rok ? make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]) : make_float4(0.f, 0.f, 0.f, 0.f), hi + c, lo + c);
}
__device__ __forceinline__ void planes8(const float (&y)[8], bool rok, uint16_t* arena, int unit, int slot, int row, int col0) {
    uint16_t* hi = arena_row(arena, unit, slot, 0, row) + col0;
    uint16_t* lo = arena_row(arena, unit, slot, 1, row) + col0;
#pragma unroll
    for (int c = 0; c < 8; c += 4)
        store_planes4(rok ? make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]) : make_float4(0.f, 0.f, 0.f, 0.f), hi + c, lo + c);
}

__global__ void __launch_bounds__(G_THREADS, 1)
stage_kernel(const __grid_constant__ CUtensorMap tmap_w256, const __grid_constant__ CUtensorMap tmap_wffn,
             const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ StageArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_BAR_OFF);
    uint64_t* full = bars;
    uint64_t* empty = bars + G_NSTG;
    uint64_t* accfull = bars + 2 * G_NSTG;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * G_NSTG + 2);
    Mail mail{reinterpret_cast<float2(*)[128][9]>(smem + G_MAIL_OFF)};
    float2 (*st4)[128] = reinterpret_cast<float2(*)[128]>(smem + G_ST4_OFF);
    if (smem + G_SMEM_USED > smem_raw + G_SMEM) __trap();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();     // == blockIdx.x (cluster = the 8 CTAs of one unit)
    const int unit = blockIdx.y;                 // branch * B + b
    const int br = unit / a.B, b = unit % a.B;
    const int N = a.N;
    const pf_branch_weights& bw = a.w.br[br];
    const int ffn = a.w.ffn_channels;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_w256);
        tma_prefetch_desc(&tmap_wffn);
        tma_prefetch_desc(&tmap_a);
        for (int i = 0; i < G_NSTG; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(&accfull[0], 1);
        mbar_init(&accfull[1], 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);

    // ------------------------------------------------------------------ role state
    uint32_t itg = 0;        // producer / MMA: ring uses so far (stage = itg % 3, phase parity = (itg / 3) & 1)
    int pre = 0;             // producer: weight boxes of the upcoming phase already in flight (first `pre` ring uses)
    uint32_t par0 = 0, par1 = 0;   // workers: parity of the next completion of accfull[0] / accfull[1]

    auto stage_bytes = [](const GPass& p) { return (uint32_t)(2 * G_PLANE + p.nbox * 2 * G_WBOX); };
    // (all three run by the WHOLE producer warp; one lane elected inside each PTX block issues)
    auto issue_w = [&](const GPass& p, int kb, int s) {
        uint8_t* st = smem + s * G_STAGE;
        const CUtensorMap* wm = p.ffn ? &tmap_wffn : &tmap_w256;
        const int k = p.k0 + kb * 64;
        mbar_arrive_expect_tx_warp(&full[s], stage_bytes(p));
        for (int j = 0; j < p.nbox; ++j) {
            tma_load_2d_warp(st + 2 * G_PLANE + j * G_WBOX, wm, &full[s], k, p.whi[j], kEvictNormal);
            tma_load_2d_warp(st + 3 * G_PLANE + j * G_WBOX, wm, &full[s], k, p.wlo[j], kEvictNormal);
        }
    };
    auto issue_a = [&](const GPass& p, int kb, int s) {
        uint8_t* st = smem + s * G_STAGE;
        const int row = ((unit * NSLOT + p.a_slot) * 2) * 128;
        tma_load_2d_warp(st, &tmap_a, &full[s], kb * 64, row, kEvictNormal);
        tma_load_2d_warp(st + G_PLANE, &tmap_a, &full[s], kb * 64, row + 128, kEvictNormal);
    };
    auto prefetch_w = [&](int ph) {          // weights do not depend on anything: start the next phase's ring early
        GPhase g;
        get_phase(ph, rank, br, bw, ffn, g);
        const int total = g.npass * 4;
        pre = total < G_NSTG ? total : G_NSTG;
        for (int it = 0; it < pre; ++it) {
            const uint32_t n = itg + it;
            gwait(&empty[n % G_NSTG], ((n / G_NSTG) & 1) ^ 1);
            issue_w(g.p[it >> 2], it & 3, n % G_NSTG);
        }
    };
    auto produce = [&](int ph, int next_ph) {
        GPhase g;
        get_phase(ph, rank, br, bw, ffn, g);
        fence_proxy_async_all();             // the peers' generic-proxy stores to the arena (acquired at the cluster barrier) -> TMA
        const int total = g.npass * 4;
        for (int it = 0; it < total; ++it) {
            const uint32_t n = itg + it;
            const int s = n % G_NSTG;
            if (it >= pre) {
                gwait(&empty[s], ((n / G_NSTG) & 1) ^ 1);
                issue_w(g.p[it >> 2], it & 3, s);
            }
            issue_a(g.p[it >> 2], it & 3, s);
        }
        itg += total, pre = 0;
        if (next_ph >= 0) prefetch_w(next_ph);
    };
    auto mma = [&](int ph) {
        GPhase g;
        get_phase(ph, rank, br, bw, ffn, g);
        tc_fence_after();
        for (int ps = 0; ps < g.npass; ++ps) {
            const GPass& p = g.p[ps];
            const uint32_t idesc = make_idesc_bf16(128, 32 * p.nbox, 0, 0);
            const uint32_t d = tmem_base + (uint32_t)p.tcol;
            for (int kb = 0; kb < 4; ++kb) {
                const uint32_t n = itg++;
                const int s = n % G_NSTG;
                gwait(&full[s], (n / G_NSTG) & 1);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + s * G_STAGE);
                const uint64_t dah = make_smem_desc_sw128(st, 16, 1024), dal = make_smem_desc_sw128(st + G_PLANE, 16, 1024);
                const uint64_t dwh = make_smem_desc_sw128(st + 2 * G_PLANE, 16, 1024);
                const uint64_t dwl = make_smem_desc_sw128(st + 3 * G_PLANE, 16, 1024);
#pragma unroll
                for (int k16 = 0; k16 < 4; ++k16) {
                    const uint64_t o = (uint64_t)(k16 * 2);
                    umma_bf16_ss_warp(d, dal + o, dwh + o, idesc, (kb | k16) != 0);
                    umma_bf16_ss_warp(d, dah + o, dwl + o, idesc, 1);
                    umma_bf16_ss_warp(d, dah + o, dwh + o, idesc, 1);
                }
                umma_commit_warp(&empty[s]);
            }
            umma_commit_warp(&accfull[ps & 1]);
        }
    };
    // every thread of the cluster passes the same sequence of cluster barriers
    auto csync = [&]() {
        tc_fence_before();
        cluster_sync_all();
        tc_fence_after();
    };

    // ------------------------------------------------------------------ worker identity
    const bool worker = warp >= 2;
    const int ew = warp - 2, q = warp & 3, chunk = ew >> 2;
    const int row = q * 32 + lane;
    const bool rok = row < N;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int c32 = 32 * rank;                   // this CTA's first feature column
    const size_t grow = (size_t)b * N + row;     // row of [B][N][..] tensors
    float y[32];
    float cnt = 0.f;

    auto wait_acc0 = [&]() { gwait(&accfull[0], par0), par0 ^= 1u, tc_fence_after(); };
    auto wait_acc1 = [&]() { gwait(&accfull[1], par1), par1 ^= 1u, tc_fence_after(); };
    auto publish = [&](int arr) {
        float mu, m2;
        stats32(y, mu, m2);
        mail.publish(arr, rank, row, mu, m2, G_CL);
    };

    // =================================================================== prep
    if (warp == 0) prefetch_w(PH_DUAL);
    pdl_wait();                  // the pooling partials / previous stage's kernels are visible from here on
    pdl_launch_dependents();
    if (worker) {
        // pooled' = sum of the split-K pooling partials (fixed order), kernel operand (+ mask kernel for the depth
        // branch, kernel_update_head.py:250), mask pixel count; thread = (row, 8 columns)
        const int c0 = c32 + 8 * chunk;
        float p8[8], k8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) p8[i] = 0.f, k8[i] = 0.f;
        if (rok) {
            const int S = a.S;
            const float* src = a.partial + (((size_t)unit * S) * N + row) * 256 + c0;
            const size_t stride = (size_t)N * 256;
            int s = 0;
            for (; s + 4 <= S; s += 4) {
                float4 u[4], v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) u[i] = ld4(src + (size_t)(s + i) * stride), v[i] = ld4(src + (size_t)(s + i) * stride + 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    p8[0] += u[i].x, p8[1] += u[i].y, p8[2] += u[i].z, p8[3] += u[i].w;
                    p8[4] += v[i].x, p8[5] += v[i].y, p8[6] += v[i].z, p8[7] += v[i].w;
                }
            }
            for (; s < S; ++s) {
                const float4 u = ld4(src + (size_t)s * stride), v = ld4(src + (size_t)s * stride + 4);
                p8[0] += u.x, p8[1] += u.y, p8[2] += u.z, p8[3] += u.w, p8[4] += v.x, p8[5] += v.y, p8[6] += v.z, p8[7] += v.w;
            }
            for (int s2 = 0; s2 < S; ++s2) cnt += __ldg(a.cntp + ((size_t)b * S + s2) * N + row);   // the mask unit's counts
            const float4 u = ld4(a.obj_in + grow * 256 + c0), v = ld4(a.obj_in + grow * 256 + c0 + 4);
            k8[0] = u.x, k8[1] = u.y, k8[2] = u.z, k8[3] = u.w, k8[4] = v.x, k8[5] = v.y, k8[6] = v.z, k8[7] = v.w;
            if (br == 1) {
                const float4 d0 = ld4(a.dep_in + grow * 256 + c0), d1 = ld4(a.dep_in + grow * 256 + c0 + 4);
                // depth_proposal + proposal_feat (same operand order as the per-layer prep kernel)
                k8[0] = d0.x + k8[0], k8[1] = d0.y + k8[1], k8[2] = d0.z + k8[2], k8[3] = d0.w + k8[3];
                k8[4] = d1.x + k8[4], k8[5] = d1.y + k8[5], k8[6] = d1.z + k8[6], k8[7] = d1.w + k8[7];
            }
        }
        planes8(p8, rok, a.arena, unit, SLOT_POOLED, row, c0);
        planes8(k8, rok, a.arena, unit, SLOT_INP, row, c0);
        fence_proxy_async_all();
    }
    csync();

    // =================================================================== dyn + inp   (kernel_updator.py:58-69, 78-79)
    if (warp == 0) produce(PH_DUAL, PH_GATE);
    else if (warp == 1) mma(PH_DUAL);
    else {
        wait_acc0();
        wait_acc1();
        if (chunk == 0) {           // gate_feats = input_in * param_in -> A operand of the gates
            float z[32];
            tmem_ld32f(tlane + 0, y);
            tmem_ld32f(tlane + 64, z);
            add_gvec32(y, bw.dyn_b + c32);
            if (bw.dyn_cb) fma_gvec32(y, cnt, bw.dyn_cb + c32);
            add_gvec32(z, bw.inp_b + c32);
#pragma unroll
            for (int c = 0; c < 32; ++c) y[c] *= z[c];
            planes32(y, rok, a.arena, unit, SLOT_GATEIN, row, c32);
            fence_proxy_async_all();
        } else if (chunk == 1) {    // param_out
            tmem_ld32f(tlane + 32, y);
            add_gvec32(y, bw.dyn_b + 256 + c32);
            if (bw.dyn_cb) fma_gvec32(y, cnt, bw.dyn_cb + 256 + c32);
            publish(0);
        } else if (chunk == 2) {    // input_out
            tmem_ld32f(tlane + 96, y);
            add_gvec32(y, bw.inp_b + 256 + c32);
            publish(1);
        }
    }
    csync();                        // LayerNorm pieces exchanged
    if (worker && (chunk == 1 || chunk == 2)) {
        float mean, rstd;
        mail.combine(chunk - 1, row, mean, rstd);
        ln_apply32(y, mean, rstd, chunk == 1 ? bw.ln_norm_out : bw.ln_input_norm_out, c32);
        tmem_st32f(tlane + (chunk == 1 ? TC_PON : TC_ION), y);
    }
    csync();

    // =================================================================== gates + mixing   (:69-88)
    if (warp == 0) produce(PH_GATE, PH_FC);
    else if (warp == 1) mma(PH_GATE);
    else {
        wait_acc0();
        if (chunk < 2) {            // chunk 0: input_gate, chunk 1: update_gate (pre-LN) of this CTA's 32 features
            const int gpos = 128 * (rank >> 1) + 32 * (rank & 1) + 64 * chunk;   // position in the interleaved gate_b
            tmem_ld32f(tlane + 32 * chunk, y);
            add_gvec32(y, bw.gate_b + gpos);
            publish(chunk);
        }
    }
    csync();
    if (worker && chunk < 2) {
        float mean, rstd;
        mail.combine(chunk, row, mean, rstd);
        ln_apply32(y, mean, rstd, chunk == 0 ? bw.ln_input_norm_in : bw.ln_norm_in, c32);
        float t[32];
        tmem_ld32f(tlane + (chunk == 0 ? TC_ION : TC_PON), t);
#pragma unroll
        for (int c = 0; c < 32; ++c) y[c] = sigmoid_fast(y[c]) * t[c];   // input_gate * input_out | update_gate * param_out
        if (chunk == 1) tmem_st32f(tlane + TC_PON, y);                   // hand update_gate * param_out to the chunk-0 warp
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (worker && chunk == 0) {
        float t[32];
        tmem_ld32f(tlane + TC_PON, t);
#pragma unroll
        for (int c = 0; c < 32; ++c) y[c] = t[c] + y[c];                 // features = update_gate * param_out + input_gate * input_out
        planes32(y, rok, a.arena, unit, SLOT_MIX, row, c32);
        fence_proxy_async_all();
    }
    csync();

    // =================================================================== fc_layer + fc_norm + ReLU   (:89-92)
    if (warp == 0) produce(PH_FC, PH_QKV);
    else if (warp == 1) mma(PH_FC);
    else {
        wait_acc0();
        if (chunk == 0) {
            tmem_ld32f(tlane + 0, y);
            add_gvec32(y, bw.fc_b + c32);
            publish(0);
        }
    }
    csync();
    if (worker && chunk == 0) {
        float mean, rstd;
        mail.combine(0, row, mean, rstd);
        ln_apply32(y, mean, rstd, bw.ln_fc_norm, c32);
#pragma unroll
        for (int c = 0; c < 32; ++c) y[c] = fmaxf(y[c], 0.f);
        planes32(y, rok, a.arena, unit, SLOT_OBJ0, row, c32);
        tmem_st32f(tlane + TC_RES, y);                                   // residual of the attention block
        fence_proxy_async_all();
    }
    csync();

    // =================================================================== in-proj of head `rank` + its attention
    // (mmcv MultiheadAttention -> nn.MultiheadAttention, seq-first; kernel_update_head.py:259-260)
    if (warp == 0) produce(PH_QKV, -1);     // no weight prefetch: the attention scratch aliases ring stage 2
    else if (warp == 1) mma(PH_QKV);
    else {
        float* s_q = reinterpret_cast<float*>(smem + G_ATT_OFF);   // [128][32], 16-byte chunk c of row n at chunk c ^ (n & 7)
        float* s_kt = s_q + 128 * 32;                                // K transposed: [32 d][G_KT_LD]
        float* s_v = s_kt + 32 * G_KT_LD;                            // [128][32], chunk c of key j at chunk c ^ ((j >> 2) & 7)
        wait_acc0();
        if (chunk == 0) {
            const float scale = 0.17677669529663687f;               // 1/sqrt(32), applied to q before q k^T as torch does
            tmem_ld32f(tlane + 0, y);
            add_gvec32(y, bw.qkv_b + c32);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(s_q + row * 32 + ((c ^ (row & 7)) << 2)) =
                    make_float4(y[4 * c] * scale, y[4 * c + 1] * scale, y[4 * c + 2] * scale, y[4 * c + 3] * scale);
        } else if (chunk == 1) {
            tmem_ld32f(tlane + 32, y);
            add_gvec32(y, bw.qkv_b + 256 + c32);
#pragma unroll
            for (int d = 0; d < 32; ++d) s_kt[d * G_KT_LD + row] = rok ? y[d] : 0.f;
        } else if (chunk == 2) {
            tmem_ld32f(tlane + 64, y);
            add_gvec32(y, bw.qkv_b + 512 + c32);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(s_v + row * 32 + ((c ^ ((row >> 2) & 7)) << 2)) =
                    rok ? make_float4(y[4 * c], y[4 * c + 1], y[4 * c + 2], y[4 * c + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        worker_bar();
        // softmax(q k^T) v: 8 threads per query pair, keys dealt to the 8 threads in blocks of 4 (as attention_kernel)
        const int tw = ew * 32 + lane;                  // 0 .. 511
        const int part = tw & 7;
        const int qbase = (tw >> 8) * 64 + ((tw & 255) >> 3);
        float o[2][32], den[2], mx[2];
        float sc[2][4][4];
        int nq[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) nq[t] = qbase + 32 * t, mx[t] = -INFINITY;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j0 = 4 * (part + 8 * i);
            float4 acc[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int d4 = 0; d4 < 8; ++d4) {
                float4 qv[2];
#pragma unroll
                for (int t = 0; t < 2; ++t) qv[t] = *reinterpret_cast<const float4*>(s_q + nq[t] * 32 + ((d4 ^ (nq[t] & 7)) << 2));
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 kk = *reinterpret_cast<const float4*>(s_kt + (4 * d4 + e) * G_KT_LD + j0);
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const float qd = e == 0 ? qv[t].x : (e == 1 ? qv[t].y : (e == 2 ? qv[t].z : qv[t].w));
                        acc[t].x += qd * kk.x, acc[t].y += qd * kk.y, acc[t].z += qd * kk.z, acc[t].w += qd * kk.w;
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                sc[t][i][0] = j0 < N ? acc[t].x : -INFINITY, sc[t][i][1] = j0 + 1 < N ? acc[t].y : -INFINITY;
                sc[t][i][2] = j0 + 2 < N ? acc[t].z : -INFINITY, sc[t][i][3] = j0 + 3 < N ? acc[t].w : -INFINITY;
                mx[t] = fmaxf(fmaxf(mx[t], fmaxf(sc[t][i][0], sc[t][i][1])), fmaxf(sc[t][i][2], sc[t][i][3]));
            }
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
#pragma unroll
            for (int s2 = 1; s2 < 8; s2 <<= 1) mx[t] = fmaxf(mx[t], __shfl_xor_sync(0xffffffffu, mx[t], s2));
            den[t] = 0.f;
#pragma unroll
            for (int d = 0; d < 32; ++d) o[t][d] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j0 = 4 * (part + 8 * i);          // (j0 >> 2) & 7 == part
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float pj[2];
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    pj[t] = __expf(sc[t][i][e] - mx[t]);   // exp(-inf) = 0 for padded keys (their V rows are zero)
                    den[t] += pj[t];
                }
#pragma unroll
                for (int d4 = 0; d4 < 8; ++d4) {
                    const float4 vv = *reinterpret_cast<const float4*>(s_v + (j0 + e) * 32 + ((d4 ^ part) << 2));
#pragma unroll
                    for (int t = 0; t < 2; ++t)
                        o[t][d4 * 4] += pj[t] * vv.x, o[t][d4 * 4 + 1] += pj[t] * vv.y, o[t][d4 * 4 + 2] += pj[t] * vv.z,
                            o[t][d4 * 4 + 3] += pj[t] * vv.w;
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            float dn = den[t];
#pragma unroll
            for (int s2 = 1; s2 < 8; s2 <<= 1) dn += __shfl_xor_sync(0xffffffffu, dn, s2);
            const float inv = 1.f / dn;
            // octet reduce-scatter: after 3 halving steps lane `part` holds the full sums of dims [4 part, 4 part + 4)
            float r16[16], r8[8], r4[4];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float send = (part & 4) ? o[t][k] : o[t][16 + k];
                const float keep = (part & 4) ? o[t][16 + k] : o[t][k];
                r16[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float send = (part & 2) ? r16[k] : r16[8 + k];
                const float keep = (part & 2) ? r16[8 + k] : r16[k];
                r8[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float send = (part & 1) ? r8[k] : r8[4 + k];
                const float keep = (part & 1) ? r8[4 + k] : r8[k];
                r4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
            const int n = nq[t];
            store_planes4(n < N ? make_float4(r4[0] * inv, r4[1] * inv, r4[2] * inv, r4[3] * inv) : make_float4(0.f, 0.f, 0.f, 0.f),
                          arena_row(a.arena, unit, SLOT_ATT, 0, n) + c32 + part * 4,
                          arena_row(a.arena, unit, SLOT_ATT, 1, n) + c32 + part * 4);
        }
        fence_proxy_async_all();
    }
    csync();

    // =================================================================== attention_norm(x + out_proj(attn))
    if (warp == 0) produce(PH_OUT, PH_FFN1);
    else if (warp == 1) mma(PH_OUT);
    else {
        wait_acc0();
        if (chunk == 0) {
            float t[32];
            tmem_ld32f(tlane + 0, y);
            tmem_ld32f(tlane + TC_RES, t);
            add_gvec32(y, bw.out_b + c32);
#pragma unroll
            for (int c = 0; c < 32; ++c) y[c] += t[c];
            publish(0);
        }
    }
    csync();
    if (worker && chunk == 0) {
        float mean, rstd;
        mail.combine(0, row, mean, rstd);
        ln_apply32(y, mean, rstd, bw.ln_attn, c32);
        planes32(y, rok, a.arena, unit, SLOT_OBJ1, row, c32);
        tmem_st32f(tlane + TC_RES, y);                                   // residual of the FFN block
        fence_proxy_async_all();
    }
    csync();

    // =================================================================== FFN layer 1 + ReLU: hidden columns [256 rank, +256)
    if (warp == 0) produce(PH_FFN1, PH_FFN2);
    else if (warp == 1) mma(PH_FFN1);
    else {
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
            if (t == 0) wait_acc0();
            else wait_acc1();
            const int hc = 128 * t + 32 * chunk;                         // column inside this CTA's 256 hidden channels
            tmem_ld32f(tlane + hc, y);
            add_gvec32(y, bw.ffn1_b + 256 * rank + hc);
#pragma unroll
            for (int c = 0; c < 32; ++c) y[c] = fmaxf(y[c], 0.f);
            planes32(y, rok, a.arena, unit, SLOT_HID0 + rank, row, hc);
        }
        fence_proxy_async_all();
    }
    csync();

    // =================================================================== FFN layer 2 over this CTA's hidden channels -> partial
    if (warp == 0) produce(PH_FFN2, PH_HEADS);
    else if (warp == 1) mma(PH_FFN2);
    else {
        float* prow = a.part + (((size_t)unit * G_CL + rank) * 128 + row) * 256;
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
            if (t == 0) wait_acc0();
            else wait_acc1();
            const int oc = 128 * t + 32 * chunk;
            tmem_ld32f(tlane + oc, y);
            store32(prow + oc, y);
        }
    }
    csync();

    // =================================================================== ffn_norm(x + sum of partials + b2)   (:271-272)
    {
        const int c0 = c32 + 8 * chunk;
        float x[8];
        if (worker) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = 0.f;
            const float* pp = a.part + (((size_t)unit * G_CL) * 128 + row) * 256 + c0;
            float4 u[G_CL], v[G_CL];
#pragma unroll
            for (int j = 0; j < G_CL; ++j) {   // written by the peers during THIS launch: coherent (L2) loads, not the read-only path
                u[j] = __ldcg(reinterpret_cast<const float4*>(pp + (size_t)j * 128 * 256));
                v[j] = __ldcg(reinterpret_cast<const float4*>(pp + (size_t)j * 128 * 256 + 4));
            }
#pragma unroll
            for (int j = 0; j < G_CL; ++j) {
                x[0] += u[j].x, x[1] += u[j].y, x[2] += u[j].z, x[3] += u[j].w;
                x[4] += v[j].x, x[5] += v[j].y, x[6] += v[j].z, x[7] += v[j].w;
            }
            float res[8];
            tmem_ld8f(tlane + TC_RES + 8 * chunk, res);
            const float4 bu = ld4(bw.ffn2_b + c0), bv = ld4(bw.ffn2_b + c0 + 4);
            x[0] += bu.x + res[0], x[1] += bu.y + res[1], x[2] += bu.z + res[2], x[3] += bu.w + res[3];
            x[4] += bv.x + res[4], x[5] += bv.y + res[5], x[6] += bv.z + res[6], x[7] += bv.w + res[7];
            float mu = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) mu += x[i];
            mu *= 0.125f;
            float m2 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) m2 += (x[i] - mu) * (x[i] - mu);
            st4[chunk][row] = make_float2(mu, m2);
            worker_bar();
            if (chunk == 0) {       // merge the 4 pieces of 8 columns into this CTA's piece of 32 (Chan et al.)
                float2 p[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) p[k] = st4[k][row];
                const float m32 = ((p[0].x + p[1].x) + (p[2].x + p[3].x)) * 0.25f;
                float q2 = 0.f, dv = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) q2 += p[k].y, dv += (p[k].x - m32) * (p[k].x - m32);
                mail.publish(0, rank, row, m32, q2 + 8.f * dv, G_CL);
            }
        }
        csync();
        if (worker) {
            float mean, rstd;
            mail.combine(0, row, mean, rstd);
            const float4 gu = ld4(bw.ln_ffn + c0), gv = ld4(bw.ln_ffn + c0 + 4);
            const float4 eu = ld4(bw.ln_ffn + 256 + c0), ev = ld4(bw.ln_ffn + 256 + c0 + 4);
            const float gam[8] = {gu.x, gu.y, gu.z, gu.w, gv.x, gv.y, gv.z, gv.w};
            const float bet[8] = {eu.x, eu.y, eu.z, eu.w, ev.x, ev.y, ev.z, ev.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = (x[i] - mean) * rstd * gam[i] + bet[i];
            if (rok) {              // obj_feat / depth_feat_new: the stage's outputs
                float* dst = (br == 0 ? a.obj_out : a.dep_out) + grow * 256 + c0;
                *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(x[4], x[5], x[6], x[7]);
            }
            planes8(x, rok, a.arena, unit, SLOT_OBJ2, row, c0);
            fence_proxy_async_all();
        }
        csync();
    }

    // =================================================================== cls_fcs | mask_fcs | depth_regs: Linear + LN (+ ReLU)
    if (warp == 0) produce(PH_HEADS, PH_KERN);
    else if (warp == 1) mma(PH_HEADS);
    else {
        wait_acc0();
        if (chunk == 0 || (chunk == 1 && br == 0)) {
            tmem_ld32f(tlane + 32 * chunk, y);
            publish(chunk);
        }
    }
    csync();
    if (worker && (chunk == 0 || (chunk == 1 && br == 0))) {
        float mean, rstd;
        mail.combine(chunk, row, mean, rstd);
        ln_apply32(y, mean, rstd, chunk == 0 ? bw.ln_head_a : bw.ln_head_b, c32);
        if (bw.head_relu) {
#pragma unroll
            for (int c = 0; c < 32; ++c) y[c] = fmaxf(y[c], 0.f);
        }
        // mask branch: chunk 0 = cls feature -> HEAD0, chunk 1 = mask feature -> HEAD1; depth branch: chunk 0 -> HEAD1
        planes32(y, rok, a.arena, unit, (br == 0 && chunk == 0) ? SLOT_HEAD0 : SLOT_HEAD1, row, c32);
        fence_proxy_async_all();
    }
    csync();

    // =================================================================== fc_mask / fc_depth (folded) -> dynamic kernels, kbias, fc_cls
    if (warp == 0) produce(PH_KERN, -1);
    else if (warp == 1) mma(PH_KERN);
    else {
        wait_acc0();
        if (chunk == 0) {
            tmem_ld32f(tlane + 0, y);
            add_gvec32(y, bw.kern_b + c32);
            if (rok) {
                uint16_t* hi = a.kern_split + (((size_t)unit * 2) * N + row) * 256 + c32;
                uint16_t* lo = hi + (size_t)N * 256;
#pragma unroll
                for (int c = 0; c < 32; c += 4) store_planes4(make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]), hi + c, lo + c);
                if (a.kern) store32(a.kern + ((size_t)unit * N + row) * 256 + c32, y);
            }
        } else if (chunk == 1 && rank == 0) {
            float t[8];
            tmem_ld8f(tlane + 32, t);                                    // column 32 = the logit-bias row
            if (rok) a.kbias[(size_t)unit * N + row] = t[0] + __ldg(bw.kbrow_b);
        }
        if (br == 0 && rank == 1) {
            wait_acc1();
            if (chunk == 2 && a.cls_out) {
                tmem_ld32f(tlane + 64, y);
                const int ncls = a.w.num_classes;
                if (rok) {
                    float* dst = a.cls_out + grow * ncls;
#pragma unroll
                    for (int c = 0; c < PF_MAX_CLASSES; ++c)
                        if (c < ncls) {
                            const float v = y[c] + __ldg(bw.cls_b + c);
                            dst[c] = a.cls_sigmoid ? sigmoid_fast(v) : v;
                        }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

static int g_fused_update = 1;

}  // namespace pf

extern "C" int pf_set_fused_update(int on) {
    const int old = pf::g_fused_update;
    pf::g_fused_update = on ? 1 : 0;
    return old;
}

namespace pf {

bool fused_update_enabled() { return g_fused_update != 0; }

size_t stage_fused_ws_bytes(int B) { return (size_t)2 * B * G_CL * 128 * 256 * sizeof(float); }

// one launch: grid (8, 2B), clusters of 8 along x
int launch_stage_fused(const pf_stage_weights* w, const float* partial, const float* cntp, int S, const float* obj_in,
                       const float* dep_in, float* obj_out, float* dep_out, float* cls_out, float* kern, uint16_t* kern_split,
                       float* kbias, uint16_t* arena, float* part, int B, int N, int cls_sigmoid, cudaStream_t st) {
    CUtensorMap mw, mf, ma;
    if (int e = cached_tmap_2d(&mw, w->wstack256, (uint64_t)w->wstack256_rows, 256, 32)) return e;
    if (int e = cached_tmap_2d(&mf, w->wstack_ffn, (uint64_t)w->wstack_ffn_rows, (uint64_t)w->ffn_channels, 32)) return e;
    if (int e = cached_tmap_2d(&ma, arena, (uint64_t)2 * B * NSLOT * 256, 256, 128)) return e;
    StageArgs a;
    memset(&a, 0, sizeof(a));
    a.w = *w;
    a.partial = partial, a.cntp = cntp, a.S = S, a.obj_in = obj_in, a.dep_in = dep_in, a.obj_out = obj_out, a.dep_out = dep_out;
    a.cls_out = cls_out, a.kern = kern, a.kern_split = kern_split, a.kbias = kbias, a.arena = arena, a.part = part;
    a.B = B, a.N = N, a.cls_sigmoid = cls_sigmoid;

    static std::mutex attr_mu;
    static bool attr_done[64] = {};
    {
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lock(attr_mu);
        if (dev < 0 || dev >= 64 || !attr_done[dev]) {
            cudaError_t e = cudaFuncSetAttribute(stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
            if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "stage_kernel smem attribute: %s", cudaGetErrorString(e));
            if (dev >= 0 && dev < 64) attr_done[dev] = true;
        }
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(G_CL, 2 * B, 1);
    cfg.blockDim = dim3(G_THREADS);
    cfg.dynamicSmemBytes = G_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    attrs[1].id = cudaLaunchAttributeClusterDimension;
    attrs[1].val.clusterDim.x = G_CL, attrs[1].val.clusterDim.y = 1, attrs[1].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, stage_kernel, mw, mf, ma, a);
    if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "stage_kernel launch: %s", cudaGetErrorString(e));
    count_launch();
    return PF_OK;
}

}  // namespace pf
