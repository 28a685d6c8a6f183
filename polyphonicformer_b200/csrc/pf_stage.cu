// K2, fused -- the small-N block of one decoder stage as ONE launch:
//   KernelUpdator (polyphonic/funcs/kernel_updator.py:55-93), inter-kernel multi-head attention + LN, FFN + LN and the
//   cls / mask / depth FC heads (polyphonic/kernel_update_head.py:245-288), feat_transform folded (include/pf_decoder.h).
//
// One 8-CTA thread-block cluster per unit (= one image of one branch: 128 kernel rows x 256 features).  CTA `rank` of
// the cluster owns feature columns [32 rank, 32 rank + 32) of every 256-wide activation, hidden columns
// [256 rank, 256 rank + 256) of the FFN, and attention head `rank`.  A layer = one step:
//     TMA (A = the full-width activation of the previous step, bf16 hi / lo planes in the L2-resident arena;
//          W = the 32-row weight blocks of this CTA's output columns, boxes of [32][64] picked straight out of the
//          weight stacks)  ->  tcgen05.mma  D = Al*Wh + Ah*Wl + Ah*Wh, fp32 in TMEM  ->  epilogue, thread = kernel row
// and the steps are separated by cluster barriers (release / acquire at cluster scope: the arena slice a CTA wrote with
// ordinary stores is visible to its peers' TMA loads), not by kernel boundaries.  What never leaves the SM:
//   * LayerNorm statistics: (mean, M2) of every 32-column piece go to all 8 CTAs through DSMEM mailboxes, one cluster
//     barrier, Chan merge (as in the per-layer kernels of pf_update.cu);
//   * the residuals obj0 / obj1, LN(param_out), LN(input_out): column-local, parked in spare TENSOR MEMORY columns
//     (lane = kernel row) between the steps that produce and consume them;
//   * q / k / v of head `rank` and the whole attention of that head, itself on the tensor cores: S = Q K^T and O = P V as
//     3-MMA bf16 hi / lo products from shared-memory operand planes, softmax between them from tensor memory.
// The weights of the next step are prefetched into the ring while the current step's epilogue and barriers run.
//
// The kernel body is ONE loop over a 12-entry step table per warp role (not straight-line code): the first version was
// unrolled per layer, 36 k instructions with its tables in local memory, and ran 2x slower than the 12 separate launches
// (instruction fetch + stack traffic after every L1-flushing cluster barrier; profiles/README.md).
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..17 = 16 worker warps (TMEM lane quarter =
// warp % 4, 32-column chunk = (warp - 2) / 4).
#include <string.h>

#include <mutex>

#include "pf_internal.h"
#include "pf_sm100.cuh"
#include "pf_update.cuh"
#include "pf_debug.cuh"

namespace pf {

constexpr int G_THREADS = 576;
constexpr int G_CL = 8;                          // CTAs per cluster = column slices of 32
constexpr int G_NSTG = 3;
constexpr int G_PLANE = 128 * 64 * 2;            // 16384: one [128][64] bf16 box (A plane); W planes use <= 4 boxes of [32][64]
constexpr int G_WBOX = 32 * 64 * 2;              // 4096
constexpr int G_STAGE = 4 * G_PLANE;             // A hi | A lo | W hi | W lo

// per-CTA slices of the stage's fp32 vectors (biases, LayerNorm gamma | beta), offsets in floats; gathered once per set
// of weights by pf_pack_vec_slices into [2 branches][8 ranks][VS_TOTAL] and bulk-copied into shared memory at kernel start
enum {
    VS_DYNB_IN = 0, VS_DYNB_OUT = 32, VS_DYNCB_IN = 64, VS_DYNCB_OUT = 96, VS_INPB_IN = 128, VS_INPB_OUT = 160,
    VS_GATEB_IG = 192, VS_GATEB_UG = 224, VS_FCB = 256, VS_QKVB = 288 /* q | k | v */, VS_OUTB = 384, VS_FFN1B = 416 /* 256 */,
    VS_FFN2B = 672, VS_KERNB = 704, VS_CLSB = 736, VS_KBROWB = 768 /* 1 (+31 pad) */,
    VS_LN_NORM_OUT = 800 /* gamma 32 | beta 32 */, VS_LN_INORM_OUT = 864, VS_LN_INORM_IN = 928, VS_LN_NORM_IN = 992,
    VS_LN_FC = 1056, VS_LN_ATTN = 1120, VS_LN_FFN = 1184, VS_LN_HEAD_A = 1248, VS_LN_HEAD_B = 1312, VS_TOTAL = 1376
};

constexpr int G_BAR_OFF = G_NSTG * G_STAGE;      // 196608
constexpr int G_MAIL_OFF = G_BAR_OFF + 256;
constexpr int G_MAIL_BYTES = 2 * 128 * 9 * 8;
constexpr int G_ST4_OFF = G_MAIL_OFF + G_MAIL_BYTES;   // [4 chunks][128 rows] float2: reduce-step statistics / softmax (max, sum)
constexpr int G_VEC_OFF = G_ST4_OFF + 4 * 128 * 8;
constexpr int G_STEP_OFF = G_VEC_OFF + VS_TOTAL * 4;
// tensor memory columns: [0, 256) accumulators; spare columns hold row-private fp32 state between steps
constexpr int TC_PON = 256, TC_ION = 288, TC_RES = 320;
// attention operand planes alias the (idle) ring: rows of 128 bytes, 128-byte swizzle, 16-byte chunks 0..3 of a row in use
constexpr int AT_Q = 0, AT_K = 2 * G_PLANE, AT_V = 4 * G_PLANE, AT_P = 6 * G_PLANE;   // Q, K, V: hi | lo (16 KB each); P: hi (32 KB) | lo
constexpr int AT_S_TCOL = 128, AT_O_TCOL = 0;
static_assert(AT_P + 4 * G_PLANE <= G_NSTG * G_STAGE, "attention planes must fit the ring");

struct StageArgs {
    pf_stage_weights w;
    const float* vec_slices;         // [2][8][VS_TOTAL]
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

// ---------------------------------------------------------------------------------------------------------------
// step table (shared memory, built once per CTA)
struct GPass {
    int a_slot, nbox, ffn, k0, tcol, acc;   // acc: which accumulator barrier the pass commits to
    int whi[4], wlo[4];                     // rows of the 32-row weight boxes (hi / lo plane) in the stack
};
struct GStep {
    int npass, nsync, noprefetch, pad;
    GPass p[2];
};
enum { ST_DUAL = 0, ST_GATE, ST_FC, ST_QKV, ST_OUT, ST_FFN1A, ST_FFN1B, ST_FFN2A, ST_FFN2B, ST_REDUCE, ST_HEADS, ST_KERN, NSTEP };
constexpr int G_SMEM_USED = G_STEP_OFF + NSTEP * (int)sizeof(GStep);
constexpr int G_SMEM = G_SMEM_USED + 1024;
static_assert(G_SMEM <= 232448, "shared memory budget of one CTA");

__device__ __forceinline__ void set_pass(GPass& p, int a_slot, int tcol, int acc, int nbox, int lo_off, int r0, int r1 = 0,
                                         int r2 = 0, int r3 = 0) {
    p.a_slot = a_slot, p.nbox = nbox, p.ffn = 0, p.k0 = 0, p.tcol = tcol, p.acc = acc;
    p.whi[0] = r0, p.whi[1] = r1, p.whi[2] = r2, p.whi[3] = r3;
#pragma unroll
    for (int j = 0; j < 4; ++j) p.wlo[j] = p.whi[j] + lo_off;
}

// step `st` for CTA `r` of a cluster of branch `br` (row numbers: include/pf_decoder.h, struct pf_branch_weights);
// nsync = cluster barriers that follow the step's GEMM (LayerNorm exchange + end of step)
__device__ void build_step(int st, int r, int br, const pf_branch_weights& bw, int ffn, GStep& g) {
    g.npass = 1, g.nsync = 2, g.noprefetch = 0, g.pad = 0;
    set_pass(g.p[1], 0, 0, 1, 1, 0, 0);
    switch (st) {
    case ST_DUAL:   // [param_in | param_out] of this CTA's 32 features from pooled', [input_in | input_out] from the kernel
        g.npass = 2;
        set_pass(g.p[0], SLOT_POOLED, 0, 0, 2, 512, bw.dyn_w + 32 * r, bw.dyn_w + 256 + 32 * r);
        set_pass(g.p[1], SLOT_INP, 64, 1, 2, 512, bw.inp_w + 32 * r, bw.inp_w + 256 + 32 * r);
        break;
    case ST_GATE: {  // gate rows are interleaved in blocks of 64: [input_gate 64 t .. | update_gate 64 t ..]
        const int base = bw.gate_w + 128 * (r >> 1) + 32 * (r & 1);
        set_pass(g.p[0], SLOT_GATEIN, 0, 0, 2, 512, base, base + 64);
        break;
    }
    case ST_FC:
        set_pass(g.p[0], SLOT_MIX, 0, 0, 1, 256, bw.fc_w + 32 * r);
        break;
    case ST_QKV:    // head r: its q, k and v rows of in_proj; the attention operand planes then take over the ring
        set_pass(g.p[0], SLOT_OBJ0, 0, 0, 3, 768, bw.qkv_w + 32 * r, bw.qkv_w + 256 + 32 * r, bw.qkv_w + 512 + 32 * r);
        g.nsync = 1, g.noprefetch = 1;
        break;
    case ST_OUT:
        set_pass(g.p[0], SLOT_ATT, 0, 0, 1, 256, bw.out_w + 32 * r);
        break;
    case ST_FFN1A:
    case ST_FFN1B: {   // hidden columns [256 r, 256 r + 256) as two 128-column tiles
        const int t = st - ST_FFN1A, w0 = bw.ffn1_w + 256 * r + 128 * t;
        set_pass(g.p[0], SLOT_OBJ1, 128 * t, t, 4, ffn, w0, w0 + 32, w0 + 64, w0 + 96);
        g.nsync = t;
        break;
    }
    case ST_FFN2A:
    case ST_FFN2B: {   // all 256 output columns over this CTA's own 256 hidden channels (split-K over the cluster)
        const int t = st - ST_FFN2A, w0 = bw.ffn2_w + 128 * t;
        set_pass(g.p[0], SLOT_HID0 + r, 128 * t, t, 4, 256, w0, w0 + 32, w0 + 64, w0 + 96);
        g.p[0].ffn = 1, g.p[0].k0 = 256 * r;
        g.nsync = t;
        break;
    }
    case ST_REDUCE:
        g.npass = 0;
        break;
    case ST_HEADS:
        if (br == 0) set_pass(g.p[0], SLOT_OBJ2, 0, 0, 2, 512, bw.head_w + 32 * r, bw.head_w + 256 + 32 * r);   // cls_fcs | mask_fcs
        else set_pass(g.p[0], SLOT_OBJ2, 0, 0, 1, 256, bw.head_w + 32 * r);                                    // depth_regs
        break;
    default:        // ST_KERN: fc_mask / fc_depth (folded); CTA 0 adds the logit-bias row, CTA 1 of the mask branch fc_cls
        set_pass(g.p[0], SLOT_HEAD1, 0, 0, 1, 256, bw.kern_w + 32 * r);
        g.nsync = 0;
        if (r == 0) g.p[0].nbox = 2, g.p[0].whi[1] = bw.kbrow_w, g.p[0].wlo[1] = bw.kbrow_w + 128;
        if (br == 0 && r == 1) {
            g.npass = 2;
            set_pass(g.p[1], SLOT_HEAD0, 64, 1, 1, 128, bw.cls_w);
        }
        break;
    }
}

// what a worker thread of 32-column chunk `c` does in a generic step (all fields warp-uniform scalars)
struct Epi {
    int tcol;       // accumulator columns of y, or -1: nothing to do in this step
    int vb, vcb;    // + bias, + count * count-bias (vector slice offsets, -1: none)
    int res_tcol;   // + 32 parked columns of tensor memory (-1)
    int z_tcol, vbz;   // * (tmem[z_tcol] + vec[vbz]) (-1)
    int pub;        // LayerNorm: mailbox array (-1: no LayerNorm)
    int vln;        // gamma | beta slice
    int act;        // 0 none, 1 ReLU, 2 sigmoid
    int mul_tcol;   // * 32 parked columns after the activation (-1)
    int hand;       // gates: 1 = park y for the chunk-0 warp, 2 = add what chunk 1 parked
    int out_slot, out_col;   // bf16 hi / lo planes into the arena (-1)
    int out_tcol;   // park y in tensor memory (-1)
};
__device__ __forceinline__ Epi epi_none() {
    Epi e;
    e.tcol = e.vb = e.vcb = e.res_tcol = e.z_tcol = e.vbz = e.pub = e.vln = e.mul_tcol = e.out_slot = e.out_tcol = -1;
    e.act = e.hand = e.out_col = 0;
    return e;
}
__device__ Epi get_epi(int st, int c, int rank, int br, int head_relu) {
    Epi e = epi_none();
    const int c32 = 32 * rank;
    switch (st) {
    case ST_DUAL:   // kernel_updator.py:58-69, 78-79
        if (c == 0) e.tcol = 0, e.vb = VS_DYNB_IN, e.vcb = VS_DYNCB_IN, e.z_tcol = 64, e.vbz = VS_INPB_IN, e.out_slot = SLOT_GATEIN, e.out_col = c32;
        else if (c == 1) e.tcol = 32, e.vb = VS_DYNB_OUT, e.vcb = VS_DYNCB_OUT, e.pub = 0, e.vln = VS_LN_NORM_OUT, e.out_tcol = TC_PON;
        else if (c == 2) e.tcol = 96, e.vb = VS_INPB_OUT, e.pub = 1, e.vln = VS_LN_INORM_OUT, e.out_tcol = TC_ION;
        break;
    case ST_GATE:   // :69-88: chunk 0 input_gate * input_out, chunk 1 update_gate * param_out, summed by chunk 0
        if (c == 0) e.tcol = 0, e.vb = VS_GATEB_IG, e.pub = 0, e.vln = VS_LN_INORM_IN, e.act = 2, e.mul_tcol = TC_ION, e.hand = 2, e.out_slot = SLOT_MIX, e.out_col = c32;
        else if (c == 1) e.tcol = 32, e.vb = VS_GATEB_UG, e.pub = 1, e.vln = VS_LN_NORM_IN, e.act = 2, e.mul_tcol = TC_PON, e.hand = 1;
        break;
    case ST_FC:     // :89-92
        if (c == 0) e.tcol = 0, e.vb = VS_FCB, e.pub = 0, e.vln = VS_LN_FC, e.act = 1, e.out_slot = SLOT_OBJ0, e.out_col = c32, e.out_tcol = TC_RES;
        break;
    case ST_OUT:    // attention_norm(x + out_proj(attn)), kernel_update_head.py:259-260
        if (c == 0) e.tcol = 0, e.vb = VS_OUTB, e.res_tcol = TC_RES, e.pub = 0, e.vln = VS_LN_ATTN, e.out_slot = SLOT_OBJ1, e.out_col = c32, e.out_tcol = TC_RES;
        break;
    case ST_FFN1A:
    case ST_FFN1B: {   // hidden = ReLU(W1 x + b1), :271-272
        const int hc = 128 * (st - ST_FFN1A) + 32 * c;
        e.tcol = hc, e.vb = VS_FFN1B + hc, e.act = 1, e.out_slot = SLOT_HID0 + rank, e.out_col = hc;
        break;
    }
    case ST_HEADS:  // cls_fcs | mask_fcs (ReLU) or depth_regs (no activation): Linear(no bias) + LN, :278-283
        if (c == 0) e.tcol = 0, e.pub = 0, e.vln = VS_LN_HEAD_A, e.act = head_relu, e.out_slot = br == 0 ? SLOT_HEAD0 : SLOT_HEAD1, e.out_col = c32;
        else if (c == 1 && br == 0) e.tcol = 32, e.pub = 1, e.vln = VS_LN_HEAD_B, e.act = head_relu, e.out_slot = SLOT_HEAD1, e.out_col = c32;
        break;
    default:
        break;
    }
    return e;
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
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }      // the 16 worker warps
__device__ __forceinline__ void att_arrive() { asm volatile("bar.arrive 2, 544;" ::: "memory"); }    // workers -> MMA warp
__device__ __forceinline__ void att_sync() { asm volatile("bar.sync 2, 544;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void bulk_g2s_warp(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {   // whole warp, one lane issues
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}\n" ::"r"(smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
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
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// 8 fp32 -> 8 bf16 hi (one 16-byte chunk) and 8 bf16 lo
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        const float h0 = __uint_as_float(h[i] << 16), h1 = __uint_as_float(h[i] & 0xFFFF0000u);
        l[i] = pack_bf16x2(v[2 * i] - h0, v[2 * i + 1] - h1);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]), lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// The fused kernel's arena: per (unit, slot, plane) a block of [32 column groups][128 rows][8] bf16 (column-group-major),
// so that the 32 rows of a warp store 512 contiguous bytes per instruction and a TMA box [8 groups][128 rows][8] lands in
// shared memory as the canonical NO-swizzle K-major UMMA layout (core matrix = 8 rows x 16 bytes, contiguous).
__device__ __forceinline__ uint16_t* arena_blk(uint16_t* arena, int unit, int slot, int plane, int kg, int row) {
    return arena + ((((size_t)unit * NSLOT + slot) * 2 + plane) * 32 + kg) * 1024 + row * 8;
}
// 32 / 8 values of one row -> bf16 hi / lo planes of arena slot `slot`, columns col0 ..; rows >= N are written as zeros
__device__ __forceinline__ void planes32(const float (&y)[32], bool rok, uint16_t* arena, int unit, int slot, int row, int col0) {
#pragma unroll
    for (int c = 0; c < 32; c += 8) {
        uint4 h = make_uint4(0u, 0u, 0u, 0u), l = h;
        if (rok) split8(&y[c], h, l);
        *reinterpret_cast<uint4*>(arena_blk(arena, unit, slot, 0, (col0 + c) >> 3, row)) = h;
        *reinterpret_cast<uint4*>(arena_blk(arena, unit, slot, 1, (col0 + c) >> 3, row)) = l;
    }
}
__device__ __forceinline__ void planes8(const float (&y)[8], bool rok, uint16_t* arena, int unit, int slot, int row, int col0) {
    uint4 h = make_uint4(0u, 0u, 0u, 0u), l = h;
    if (rok) split8(&y[0], h, l);
    *reinterpret_cast<uint4*>(arena_blk(arena, unit, slot, 0, col0 >> 3, row)) = h;
    *reinterpret_cast<uint4*>(arena_blk(arena, unit, slot, 1, col0 >> 3, row)) = l;
}
// shared-memory matrix descriptor without swizzle (layout type 0): core matrices of 8 rows x 16 bytes; lbo = bytes between
// the two core matrices of one K = 16 step, sbo = bytes between 8-row groups
__device__ __forceinline__ uint64_t make_smem_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}
// 32 values of row `row` -> hi / lo operand planes in shared memory (rows of 128 bytes, 128-byte swizzle), 16-byte chunks
// chunk0 .. chunk0 + 3 of the row
__device__ __forceinline__ void smem_planes32(const float (&y)[32], uint8_t* hi_plane, uint8_t* lo_plane, int row, int chunk0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint4 h, l;
        split8(&y[8 * c], h, l);
        *reinterpret_cast<uint4*>(hi_plane + sw128_offset(row, chunk0 + c)) = h;
        *reinterpret_cast<uint4*>(lo_plane + sw128_offset(row, chunk0 + c)) = l;
    }
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
    uint64_t* vecbar = bars + 2 * G_NSTG + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * G_NSTG + 3);
    Mail mail{reinterpret_cast<float2(*)[128][9]>(smem + G_MAIL_OFF)};
    float2 (*st4)[128] = reinterpret_cast<float2(*)[128]>(smem + G_ST4_OFF);
    const float* vec = reinterpret_cast<const float*>(smem + G_VEC_OFF);
    GStep* steps = reinterpret_cast<GStep*>(smem + G_STEP_OFF);
    if (smem + G_SMEM_USED > smem_raw + G_SMEM) __trap();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();     // == blockIdx.x (cluster = the 8 CTAs of one unit)
    const int unit = blockIdx.y;                 // branch * B + b
    const int br = unit / a.B, b = unit % a.B;
    const int N = a.N;

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
        mbar_init(vecbar, 1);
        mbar_fence_init();
        // this CTA's slices of the stage's vectors (weights: independent of the previous kernel)
        mbar_arrive_expect_tx(vecbar, VS_TOTAL * 4);
        bulk_g2s(smem + G_VEC_OFF, a.vec_slices + ((size_t)br * G_CL + rank) * VS_TOTAL, VS_TOTAL * 4, vecbar);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + NSTEP) build_step(threadIdx.x - 64, rank, br, a.w.br[br], a.w.ffn_channels, steps[threadIdx.x - 64]);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);
    // optional in-kernel timeline (pf_debug_timeline): the CTAs of unit 0 record %globaltimer at the step boundaries
    // (first worker thread); record tags 500 + rank / 600 + rank
    __shared__ long long* s_dbg;
    long long* dbg = nullptr;
    if (g_dbg && blockIdx.y == 0) {
        if (threadIdx.x == 64) {
            const unsigned long long slot = atomicAdd(reinterpret_cast<unsigned long long*>(g_dbg), 6ull);
            s_dbg = g_dbg + 16 + slot * 16;
            for (int i = 0; i < 6; ++i) s_dbg[16 * i + 15] = 500 + 100 * i + blockIdx.x;
        }
        __syncthreads();
        if (threadIdx.x == 64) dbg = s_dbg;
    }
    long long* dbg3 = (g_dbg && blockIdx.y == 0 && lane == 0 && warp < 2) ? s_dbg + 80 : nullptr;   // producer / MMA warp probes
    auto stamp = [&](int k) {        // slot k of the worker timeline (slots 15, 31, ... hold the record tags)
        if (dbg) dbg[k + k / 15] = gtime();
    };
    stamp(0);

    auto stage_bytes = [](const GPass& p) { return (uint32_t)(2 * G_PLANE + p.nbox * 2 * G_WBOX); };
    // (both run by the WHOLE producer warp; one lane elected inside each PTX block issues)
    auto issue_w = [&](const GPass& p, int kb, int s) {
        uint8_t* st = smem + s * G_STAGE;
        const CUtensorMap* wm = p.ffn ? &tmap_wffn : &tmap_w256;
        const int k = p.k0 + kb * 64;
        mbar_arrive_expect_tx_warp(&full[s], stage_bytes(p));
#pragma unroll 1
        for (int j = 0; j < p.nbox; ++j) {
            tma_load_2d_warp(st + 2 * G_PLANE + j * G_WBOX, wm, &full[s], k, p.whi[j], kEvictNormal);
            tma_load_2d_warp(st + 3 * G_PLANE + j * G_WBOX, wm, &full[s], k, p.wlo[j], kEvictNormal);
        }
    };
    auto issue_a = [&](const GPass& p, int kb, int s) {
        uint8_t* st = smem + s * G_STAGE;
        // 8 column groups of a plane = 16 KB contiguous in the blocked arena: plain bulk copies (a tensor box with a
        // 16-byte inner dimension moved the same bytes 3x slower)
        const uint16_t* src = a.arena + ((((size_t)unit * NSLOT + p.a_slot) * 2) * 32 + kb * 8) * 1024;
        bulk_g2s_warp(st, src, G_PLANE, &full[s]);
        bulk_g2s_warp(st + G_PLANE, src + 32 * 1024, G_PLANE, &full[s]);
    };
    // every thread of the cluster passes the same sequence of cluster barriers
    auto csync = [&]() {
        tc_fence_before();
        cluster_sync_all();
        tc_fence_after();
    };

    uint32_t itg = 0;        // producer / MMA: ring uses so far (stage = itg % 3, phase parity = (itg / 3) & 1)
    int pre = 0;             // producer: first `pre` ring uses of the upcoming step already have their weight boxes in flight
    if (warp == 0) {         // weights do not depend on the previous kernel: start the ring before the grid dependency resolves
        const GStep& S = steps[0];
        pre = G_NSTG;
#pragma unroll 1
        for (int it = 0; it < pre; ++it) issue_w(S.p[it >> 2], it & 3, it);
    }
    pdl_wait();                  // the pooling partials / previous stage's kernels are visible from here on
    pdl_launch_dependents();
    stamp(1);

    const int ew = warp - 2, q = warp & 3, chunk = ew >> 2;
    const int row = q * 32 + lane;
    const bool rok = row < N;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int c32 = 32 * rank;                   // this CTA's first feature column
    const size_t grow = (size_t)b * N + row;     // row of [B][N][..] tensors
    float cnt = 0.f;

    // =================================================================== prep (workers)
    if (warp >= 2) {
        // pooled' = sum of the split-K pooling partials (fixed order), kernel operand (+ mask kernel for the depth
        // branch, kernel_update_head.py:250), mask pixel count; thread = (row, 8 columns)
        const int c0 = c32 + 8 * chunk;
        float p8[8], k8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) p8[i] = 0.f, k8[i] = 0.f;
        if (rok) {
            const int S = a.S;
            // partials are column-group-major ([unit][S][64][N][4], pf_pool.cu): consecutive rows = consecutive float4
            const float* src = a.partial + ((size_t)unit * S) * N * 256 + ((size_t)(c0 >> 2) * N + row) * 4;
            const size_t stride = (size_t)N * 256;
#pragma unroll 1
            for (int s = 0; s < S; s += 8) {   // 16 loads in flight per thread; added in slab order (deterministic)
                float4 u[8], v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const bool ok = s + i < S;
                    u[i] = ok ? ld4(src + (size_t)(s + i) * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[i] = ok ? ld4(src + (size_t)(s + i) * stride + (size_t)N * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    p8[0] += u[i].x, p8[1] += u[i].y, p8[2] += u[i].z, p8[3] += u[i].w;
                    p8[4] += v[i].x, p8[5] += v[i].y, p8[6] += v[i].z, p8[7] += v[i].w;
                }
            }
#pragma unroll 1
            for (int s2 = 0; s2 < S; s2 += 8) {    // the mask unit's counts; 8 loads in flight (exact integers: any order)
                float c8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) c8[i] = s2 + i < S ? __ldg(a.cntp + ((size_t)b * S + s2 + i) * N + row) : 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) cnt += c8[i];
            }
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
        gwait(vecbar, 0);        // the vector slices have landed (needed from the first epilogue on)
    }
    csync();
    stamp(2);

    if (warp == 0) {
        // =============================================================== TMA producer
#pragma unroll 1
        for (int st = 0; st < NSTEP; ++st) {
            const GStep& S = steps[st];
            if (S.npass) {
                if (dbg3 && st == ST_FC) dbg3[5] = gtime();
                // (the peers fenced their generic-proxy stores to the arena towards the async proxy BEFORE the cluster barrier
                // that released them; no second fence on the reading side)
                const int total = S.npass * 4;
#pragma unroll 1
                for (int it = 0; it < total; ++it) {
                    const uint32_t n = itg + it;
                    const int s = n % G_NSTG;
                    if (it >= pre) {
                        if (dbg3 && st == ST_FC) dbg3[6] = gtime();
                        gwait(&empty[s], ((n / G_NSTG) & 1) ^ 1);
                        issue_w(S.p[it >> 2], it & 3, s);
                    }
                    issue_a(S.p[it >> 2], it & 3, s);
                }
                if (dbg3 && st == ST_FC) dbg3[7] = gtime();
                itg += total, pre = 0;
                if (!S.noprefetch) {         // weights of the next GEMM step into the slots the MMAs release
                    int nx = st + 1;
                    while (nx < NSTEP && steps[nx].npass == 0) ++nx;
                    if (nx < NSTEP) {
                        const GStep& X = steps[nx];
                        const int tot = X.npass * 4;
                        pre = tot < G_NSTG ? tot : G_NSTG;
#pragma unroll 1
                        for (int it = 0; it < pre; ++it) {
                            const uint32_t n = itg + it;
                            gwait(&empty[n % G_NSTG], ((n / G_NSTG) & 1) ^ 1);
                            issue_w(X.p[it >> 2], it & 3, n % G_NSTG);
                        }
                    }
                }
            }
#pragma unroll 1
            for (int i = 0; i < S.nsync; ++i) csync();
        }
    } else if (warp == 1) {
        // =============================================================== MMA issuer: D = Al*Wh + Ah*Wl + Ah*Wh
#pragma unroll 1
        for (int st = 0; st < NSTEP; ++st) {
            const GStep& S = steps[st];
#pragma unroll 1
            for (int ps = 0; ps < S.npass; ++ps) {
                const GPass& p = S.p[ps];
                const uint32_t idesc = make_idesc_bf16(128, 32 * p.nbox, 0, 0);
                const uint32_t d = tmem_base + (uint32_t)p.tcol;
#pragma unroll 1
                for (int kb = 0; kb < 4; ++kb) {
                    const uint32_t n = itg++;
                    const int s = n % G_NSTG;
                    gwait(&full[s], (n / G_NSTG) & 1);
                    if (dbg3 && st == ST_FC) dbg3[kb] = gtime();
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * G_STAGE);
                    // A: [8 column groups][128 rows][16 bytes], no swizzle: K = 16 step = two groups = 4096 bytes;
                    // W: rows of 128 bytes, 128-byte swizzle: K = 16 step = 32 bytes
                    const uint64_t dah = make_smem_desc_nosw(sa, 2048, 128), dal = make_smem_desc_nosw(sa + G_PLANE, 2048, 128);
                    const uint64_t dwh = make_smem_desc_sw128(sa + 2 * G_PLANE, 16, 1024);
                    const uint64_t dwl = make_smem_desc_sw128(sa + 3 * G_PLANE, 16, 1024);
#pragma unroll
                    for (int k16 = 0; k16 < 4; ++k16) {
                        const uint64_t oa = (uint64_t)(k16 * 256), o = (uint64_t)(k16 * 2);
                        umma_bf16_ss_warp(d, dal + oa, dwh + o, idesc, (kb | k16) != 0);
                        umma_bf16_ss_warp(d, dah + oa, dwl + o, idesc, 1);
                        umma_bf16_ss_warp(d, dah + oa, dwh + o, idesc, 1);
                    }
                    umma_commit_warp(&empty[s]);
                }
                umma_commit_warp(&accfull[p.acc]);
                if (dbg3 && st == ST_FC) dbg3[4] = gtime();
            }
            if (st == ST_QKV) {
                // attention of head `rank` on the tensor cores, operands = the planes the workers stage in the ring
                const uint32_t sb = smem_u32(smem);
                att_sync();                  // Q, K, V planes are in shared memory
                tc_fence_after();
                {   // S[q][key] = sum_d Q[q][d] K[key][d]: M = 128, N = 128 keys, K = 32 (two K = 16 steps)
                    const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
                    const uint64_t qh = make_smem_desc_sw128(sb + AT_Q, 16, 1024), ql = make_smem_desc_sw128(sb + AT_Q + G_PLANE, 16, 1024);
                    const uint64_t kh = make_smem_desc_sw128(sb + AT_K, 16, 1024), kl = make_smem_desc_sw128(sb + AT_K + G_PLANE, 16, 1024);
#pragma unroll
                    for (int k16 = 0; k16 < 2; ++k16) {
                        const uint64_t o = (uint64_t)(k16 * 2);
                        umma_bf16_ss_warp(tmem_base + AT_S_TCOL, ql + o, kh + o, idesc, k16 != 0);
                        umma_bf16_ss_warp(tmem_base + AT_S_TCOL, qh + o, kl + o, idesc, 1);
                        umma_bf16_ss_warp(tmem_base + AT_S_TCOL, qh + o, kh + o, idesc, 1);
                    }
                    umma_commit_warp(&accfull[0]);
                }
                att_sync();                  // P planes are in shared memory
                tc_fence_after();
                {   // O[q][d] = sum_key P[q][key] V[key][d]: M = 128, N = 64 (32 in use), K = 128 keys;
                    // V is the MN-major operand (one 128-byte row of d per key: 16 keys = 2048 bytes per K = 16 step)
                    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);
                    const uint64_t vh = make_smem_desc_sw128(sb + AT_V, G_PLANE, 1024), vl = make_smem_desc_sw128(sb + AT_V + G_PLANE, G_PLANE, 1024);
#pragma unroll 1
                    for (int kc = 0; kc < 2; ++kc) {     // P: two 64-key chunks of [128][64], 16 KB apart
                        const uint64_t ph = make_smem_desc_sw128(sb + AT_P + kc * G_PLANE, 16, 1024);
                        const uint64_t pl = make_smem_desc_sw128(sb + AT_P + 2 * G_PLANE + kc * G_PLANE, 16, 1024);
#pragma unroll
                        for (int k16 = 0; k16 < 4; ++k16) {
                            const uint64_t oa = (uint64_t)(k16 * 2), ob = (uint64_t)((kc * 4 + k16) * 128);
                            umma_bf16_ss_warp(tmem_base + AT_O_TCOL, pl + oa, vh + ob, idesc, (kc | k16) != 0);
                            umma_bf16_ss_warp(tmem_base + AT_O_TCOL, ph + oa, vl + ob, idesc, 1);
                            umma_bf16_ss_warp(tmem_base + AT_O_TCOL, ph + oa, vh + ob, idesc, 1);
                        }
                    }
                    umma_commit_warp(&accfull[1]);
                }
            }
#pragma unroll 1
            for (int i = 0; i < S.nsync; ++i) csync();
        }
    } else {
        // =============================================================== workers: epilogue of every step, thread = (row, chunk)
        uint32_t par0 = 0, par1 = 0;     // parity of the next completion of accfull[0] / accfull[1]
        float y[32];
#pragma unroll 1
        for (int st = 0; st < NSTEP; ++st) {
            const GStep& S = steps[st];
            const int nsync = S.nsync;
            if (S.npass) {
                // which accumulator barriers this step's epilogue needs
                const bool w1 = st == ST_DUAL || st == ST_FFN1B || st == ST_FFN2B;
                const bool w0 = !(st == ST_FFN1B || st == ST_FFN2B);
                if (w0) gwait(&accfull[0], par0), par0 ^= 1u;
                if (w1) gwait(&accfull[1], par1), par1 ^= 1u;
                tc_fence_after();
            }
            stamp(3 + 5 * st);
            if (st == ST_QKV) {
                // ---- q, k, v of head `rank` -> operand planes; softmax(q k^T) v on the tensor cores
                // (mmcv MultiheadAttention -> nn.MultiheadAttention, seq-first; kernel_update_head.py:259-260)
                if (chunk < 3) {
                    tmem_ld32f(tlane + 32 * chunk, y);
                    add_vec32(y, vec + VS_QKVB + 32 * chunk);
                    const float scale = chunk == 0 ? 0.17677669529663687f : 1.f;   // 1/sqrt(32) on q, before q k^T as torch does
                    const bool keep = rok || chunk != 2;                            // V rows of padded keys must be exact zeros
#pragma unroll
                    for (int c = 0; c < 32; ++c) y[c] = keep ? y[c] * scale : 0.f;
                    smem_planes32(y, smem + 2 * G_PLANE * chunk, smem + 2 * G_PLANE * chunk + G_PLANE, row, 0);
                }
                fence_proxy_async_smem();
                tc_fence_before();
                att_arrive();
                gwait(&accfull[0], par0), par0 ^= 1u;        // S is in tensor memory
                tc_fence_after();
                tmem_ld32f(tlane + AT_S_TCOL + 32 * chunk, y);
                float mx = -INFINITY;
#pragma unroll
                for (int c = 0; c < 32; ++c) mx = (32 * chunk + c < N) ? fmaxf(mx, y[c]) : mx;
                st4[chunk][row].x = mx;
                worker_bar();
                mx = fmaxf(fmaxf(st4[0][row].x, st4[1][row].x), fmaxf(st4[2][row].x, st4[3][row].x));
                float sum = 0.f;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    y[c] = (32 * chunk + c < N) ? __expf(y[c] - mx) : 0.f;
                    sum += y[c];
                }
                st4[chunk][row].y = sum;
                // P[q][key]: keys 32 chunk .. + 31 = 16-byte chunks 4 (chunk & 1) .. + 3 of the row in 64-key plane (chunk >> 1)
                smem_planes32(y, smem + AT_P + (chunk >> 1) * G_PLANE, smem + AT_P + 2 * G_PLANE + (chunk >> 1) * G_PLANE, row, 4 * (chunk & 1));
                fence_proxy_async_smem();
                tc_fence_before();
                stamp(3 + 5 * st + 1);
                att_arrive();
                gwait(&accfull[1], par1), par1 ^= 1u;        // O is in tensor memory
                tc_fence_after();
                stamp(3 + 5 * st + 2);
                if (chunk == 0) {
                    tmem_ld32f(tlane + AT_O_TCOL, y);
                    const float inv = 1.f / ((st4[0][row].y + st4[1][row].y) + (st4[2][row].y + st4[3][row].y));
#pragma unroll
                    for (int c = 0; c < 32; ++c) y[c] *= inv;
                    planes32(y, rok, a.arena, unit, SLOT_ATT, row, c32);
                }
                fence_proxy_async_all();
            } else if (st == ST_FFN2A || st == ST_FFN2B) {
                // ---- split-K partial of FFN layer 2: this CTA's 256 hidden channels, all 256 output columns
                const int oc = 128 * (st - ST_FFN2A) + 32 * chunk;
                tmem_ld32f(tlane + oc, y);
                float* pb = a.part + ((size_t)unit * G_CL + rank) * 128 * 256 + (size_t)row * 4;   // [64 column groups][128 rows][4]
#pragma unroll
                for (int c = 0; c < 32; c += 4)
                    *reinterpret_cast<float4*>(pb + (size_t)((oc + c) >> 2) * 512) = make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]);
            } else if (st == ST_REDUCE) {
                // ---- ffn_norm(x + sum of partials + b2)   (:271-272); thread = (row, 8 columns)
                const int c0 = c32 + 8 * chunk;
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = 0.f;
                const float* pp = a.part + ((size_t)unit * G_CL) * 128 * 256 + (size_t)(c0 >> 2) * 512 + (size_t)row * 4;
                float4 u[G_CL], v[G_CL];
#pragma unroll
                for (int j = 0; j < G_CL; ++j) {   // written by the peers during THIS launch: coherent (L2) loads, not the read-only path
                    u[j] = __ldcg(reinterpret_cast<const float4*>(pp + (size_t)j * 128 * 256));
                    v[j] = __ldcg(reinterpret_cast<const float4*>(pp + (size_t)j * 128 * 256 + 512));
                }
#pragma unroll
                for (int j = 0; j < G_CL; ++j) {
                    x[0] += u[j].x, x[1] += u[j].y, x[2] += u[j].z, x[3] += u[j].w;
                    x[4] += v[j].x, x[5] += v[j].y, x[6] += v[j].z, x[7] += v[j].w;
                }
                float res[8];
                tmem_ld8f(tlane + TC_RES + 8 * chunk, res);
                const float* b2 = vec + VS_FFN2B + 8 * chunk;
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] += b2[i] + res[i];
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
                stamp(3 + 5 * st + 1);
                csync();
                stamp(3 + 5 * st + 2);
                float mean, rstd;
                mail.combine(0, row, mean, rstd);
                const float* ga = vec + VS_LN_FFN + 8 * chunk;
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = (x[i] - mean) * rstd * ga[i] + ga[32 + i];
                if (rok) {              // obj_feat / depth_feat_new: the stage's outputs
                    float* dst = (br == 0 ? a.obj_out : a.dep_out) + grow * 256 + c0;
                    *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
                    *reinterpret_cast<float4*>(dst + 4) = make_float4(x[4], x[5], x[6], x[7]);
                }
                planes8(x, rok, a.arena, unit, SLOT_OBJ2, row, c0);
                fence_proxy_async_all();
                stamp(3 + 5 * st + 3);
                csync();
            } else if (st == ST_KERN) {
                // ---- fc_mask / fc_depth (folded) -> dynamic kernels (fp32 + bf16 hi / lo planes), logit bias, fc_cls
                if (chunk == 0) {
                    tmem_ld32f(tlane + 0, y);
                    add_vec32(y, vec + VS_KERNB);
                    if (rok) {
                        uint16_t* hi = a.kern_split + (((size_t)unit * 2) * N + row) * 256 + c32;
                        uint16_t* lo = hi + (size_t)N * 256;
#pragma unroll
                        for (int c = 0; c < 32; c += 8) {
                            uint4 h, l;
                            split8(&y[c], h, l);
                            *reinterpret_cast<uint4*>(hi + c) = h;
                            *reinterpret_cast<uint4*>(lo + c) = l;
                        }
                        if (a.kern) store32(a.kern + ((size_t)unit * N + row) * 256 + c32, y);
                    }
                } else if (chunk == 1 && rank == 0) {
                    float t[8];
                    tmem_ld8f(tlane + 32, t);                                    // column 32 = the logit-bias row
                    if (rok) a.kbias[(size_t)unit * N + row] = t[0] + vec[VS_KBROWB];
                }
                if (S.npass == 2) {     // CTA 1 of the mask branch: fc_cls
                    gwait(&accfull[1], par1), par1 ^= 1u;
                    tc_fence_after();
                    if (chunk == 2 && a.cls_out) {
                        tmem_ld32f(tlane + 64, y);
                        add_vec32(y, vec + VS_CLSB);
                        const int ncls = a.w.num_classes;
                        if (rok) {
                            float* dst = a.cls_out + grow * ncls;
#pragma unroll
                            for (int c = 0; c < PF_MAX_CLASSES; ++c)
                                if (c < ncls) dst[c] = a.cls_sigmoid ? sigmoid_fast(y[c]) : y[c];
                        }
                    }
                }
            } else {
                // ---- generic step: accumulator (+ bias ...) -> [LayerNorm over the cluster] -> activation -> planes / parked columns
                const Epi e = get_epi(st, chunk, rank, br, a.w.br[br].head_relu);
                if (e.tcol >= 0) {
                    tmem_ld32f(tlane + e.tcol, y);
                    if (e.vb >= 0) add_vec32(y, vec + e.vb);
                    if (e.vcb >= 0) fma_vec32(y, cnt, vec + e.vcb);
                    if (e.res_tcol >= 0 || e.z_tcol >= 0) {
                        float t[32];
                        tmem_ld32f(tlane + (e.res_tcol >= 0 ? e.res_tcol : e.z_tcol), t);
                        if (e.res_tcol >= 0) {
#pragma unroll
                            for (int c = 0; c < 32; ++c) y[c] += t[c];
                        } else {
                            add_vec32(t, vec + e.vbz);
#pragma unroll
                            for (int c = 0; c < 32; ++c) y[c] *= t[c];
                        }
                    }
                    if (e.pub >= 0) {
                        float mu, m2;
                        stats32(y, mu, m2);
                        mail.publish(e.pub, rank, row, mu, m2, G_CL);
                    }
                }
                stamp(3 + 5 * st + 1);
                if (nsync == 2) csync();                     // LayerNorm pieces exchanged
                stamp(3 + 5 * st + 2);
                if (e.tcol >= 0) {
                    if (e.pub >= 0) {
                        float mean, rstd;
                        mail.combine(e.pub, row, mean, rstd);
                        const float* ga = vec + e.vln;
#pragma unroll
                        for (int c = 0; c < 32; ++c) y[c] = (y[c] - mean) * rstd * ga[c] + ga[32 + c];
                    }
                    if (e.act == 1) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) y[c] = fmaxf(y[c], 0.f);
                    } else if (e.act == 2) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) y[c] = sigmoid_fast(y[c]);
                    }
                    if (e.mul_tcol >= 0) {
                        float t[32];
                        tmem_ld32f(tlane + e.mul_tcol, t);
#pragma unroll
                        for (int c = 0; c < 32; ++c) y[c] *= t[c];
                    }
                    if (e.hand == 1) tmem_st32f(tlane + TC_PON, y);      // update_gate * param_out -> the chunk-0 warp of this quarter
                }
                if (st == ST_GATE) {
                    tc_fence_before();
                    worker_bar();
                    tc_fence_after();
                }
                if (e.tcol >= 0) {
                    if (e.hand == 2) {   // features = update_gate * param_out + input_gate * input_out
                        float t[32];
                        tmem_ld32f(tlane + TC_PON, t);
#pragma unroll
                        for (int c = 0; c < 32; ++c) y[c] = t[c] + y[c];
                    }
                    if (e.out_slot >= 0) planes32(y, rok, a.arena, unit, e.out_slot, row, e.out_col);
                    if (e.out_tcol >= 0) tmem_st32f(tlane + e.out_tcol, y);
                }
                fence_proxy_async_all();
            }
            if (st != ST_REDUCE) stamp(3 + 5 * st + 3);
            if (st != ST_REDUCE && nsync >= 1) csync();      // end of the step (FFN1A / FFN2A / KERN: none)
            stamp(3 + 5 * st + 4);
        }
    }
    tc_fence_before();
    __syncthreads();
    stamp(63);
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// [2 branches][8 ranks][VS_TOTAL]: the per-CTA slices of a stage's vectors
__global__ void __launch_bounds__(256) vec_slices_kernel(const __grid_constant__ pf_stage_weights w, float* __restrict__ out) {
    const int r = blockIdx.x, br = blockIdx.y;
    const pf_branch_weights& bw = w.br[br];
    float* o = out + ((size_t)br * G_CL + r) * VS_TOTAL;
    const int gpos = 128 * (r >> 1) + 32 * (r & 1);     // position of this CTA's features in the interleaved gate vectors
    for (int i = threadIdx.x; i < VS_TOTAL; i += blockDim.x) {
        const float* src = nullptr;
        int off = 0;
        auto seg = [&](int base, int len, const float* p, int o0) {
            if (i >= base && i < base + len) src = p, off = o0 + (i - base);
        };
        seg(VS_DYNB_IN, 32, bw.dyn_b, 32 * r), seg(VS_DYNB_OUT, 32, bw.dyn_b, 256 + 32 * r);
        seg(VS_DYNCB_IN, 32, bw.dyn_cb, 32 * r), seg(VS_DYNCB_OUT, 32, bw.dyn_cb, 256 + 32 * r);
        seg(VS_INPB_IN, 32, bw.inp_b, 32 * r), seg(VS_INPB_OUT, 32, bw.inp_b, 256 + 32 * r);
        seg(VS_GATEB_IG, 32, bw.gate_b, gpos), seg(VS_GATEB_UG, 32, bw.gate_b, gpos + 64);
        seg(VS_FCB, 32, bw.fc_b, 32 * r);
        seg(VS_QKVB, 32, bw.qkv_b, 32 * r), seg(VS_QKVB + 32, 32, bw.qkv_b, 256 + 32 * r), seg(VS_QKVB + 64, 32, bw.qkv_b, 512 + 32 * r);
        seg(VS_OUTB, 32, bw.out_b, 32 * r), seg(VS_FFN1B, 256, bw.ffn1_b, 256 * r), seg(VS_FFN2B, 32, bw.ffn2_b, 32 * r);
        seg(VS_KERNB, 32, bw.kern_b, 32 * r), seg(VS_CLSB, 32, br == 0 ? bw.cls_b : nullptr, 0), seg(VS_KBROWB, 1, bw.kbrow_b, 0);
        const float* lns[9] = {bw.ln_norm_out, bw.ln_input_norm_out, bw.ln_input_norm_in, bw.ln_norm_in, bw.ln_fc_norm,
                               bw.ln_attn, bw.ln_ffn, bw.ln_head_a, br == 0 ? bw.ln_head_b : nullptr};
        for (int k = 0; k < 9; ++k) {
            seg(VS_LN_NORM_OUT + 64 * k, 32, lns[k], 32 * r);              // gamma
            seg(VS_LN_NORM_OUT + 64 * k + 32, 32, lns[k], 256 + 32 * r);   // beta
        }
        o[i] = src ? __ldg(src + off) : 0.f;
    }
}

static int g_fused_update = 0;   // default: whichever is faster on the headline shape (see profiles/README.md)

}  // namespace pf
PF_DEFINE_DBG_SETTER(set_dbg_stage)

extern "C" int pf_set_fused_update(int on) {
    const int old = pf::g_fused_update;
    pf::g_fused_update = on ? 1 : 0;
    return old;
}

extern "C" size_t pf_vec_slices_bytes(void) { return (size_t)2 * pf::G_CL * pf::VS_TOTAL * sizeof(float); }

extern "C" int pf_pack_vec_slices(const pf_stage_weights* w, float* out, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(w && out, PF_ERR_ARG, "pf_pack_vec_slices: null pointer");
    PF_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, PF_ERR_ALIGN, "pf_pack_vec_slices: out not 16-byte aligned");
    vec_slices_kernel<<<dim3(G_CL, 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(*w, out);
    PF_CHECK_LAUNCH("vec_slices_kernel");
    return PF_OK;
}

namespace pf {

bool fused_update_enabled() { return g_fused_update != 0; }

// ffn2 partials + room for the vector slices when the caller did not pre-pack them
size_t stage_fused_ws_bytes(int B) {
    return (size_t)2 * B * G_CL * 128 * 256 * sizeof(float) + ((pf_vec_slices_bytes() + 255) / 256) * 256;
}

// one launch: grid (8, 2B), clusters of 8 along x
int launch_stage_fused(const pf_stage_weights* w, const float* partial, const float* cntp, int S, const float* obj_in,
                       const float* dep_in, float* obj_out, float* dep_out, float* cls_out, float* kern, uint16_t* kern_split,
                       float* kbias, uint16_t* arena, float* part, int B, int N, int cls_sigmoid, cudaStream_t st) {
    CUtensorMap mw, mf, ma;
    if (int e = cached_tmap_2d(&mw, w->wstack256, (uint64_t)w->wstack256_rows, 256, 32)) return e;
    if (int e = cached_tmap_2d(&mf, w->wstack_ffn, (uint64_t)w->wstack_ffn_rows, (uint64_t)w->ffn_channels, 32)) return e;
    {   // the blocked arena as a 3-D tensor [blocks = units * slots * 2 planes * 32 column groups][128 rows][8]; box = 8 groups
        static std::mutex mu;
        static const void* c_base = nullptr;
        static int c_B = 0;
        static CUtensorMap c_map;
        std::lock_guard<std::mutex> lock(mu);
        if (c_base != arena || c_B != B) {
            if (int e = make_tmap_bf16_blocked(&c_map, arena, (uint64_t)2 * B * NSLOT * 2 * 32)) return e;
            c_base = arena, c_B = B;
        }
        ma = c_map;
    }
    StageArgs a;
    memset(&a, 0, sizeof(a));
    a.w = *w;
    a.vec_slices = w->vec_slices;
    if (!a.vec_slices) {     // not pre-packed by the host (pf_pack_vec_slices): gather them now, behind the ffn2 partials
        float* vs = part + (size_t)2 * B * G_CL * 128 * 256;
        if (int e = pf_pack_vec_slices(w, vs, st)) return e;
        a.vec_slices = vs;
    }
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
