// Shared between the two implementations of the small-N block (pf_update.cu: one launch per layer; pf_stage.cu: the
// fused cluster kernel): arena layout, slot numbers, epilogue helpers and the LayerNorm mailbox.
#pragma once
#include "pf_internal.h"
#include "pf_sm100.cuh"

namespace pf {

constexpr int T_THREADS = 576;
constexpr int T_TN = 128;                       // output columns per CTA (per pass)
constexpr int T_KC = 64;                        // K per ring stage
constexpr int T_K = 256;                        // K per pass
constexpr int T_NSTG = 3;
constexpr int T_PLANE = 128 * T_KC * 2;         // 16384: one [128][64] bf16 box
constexpr int T_STAGE = 4 * T_PLANE;            // A hi | A lo | W hi | W lo
constexpr int T_BAR_OFF = T_NSTG * T_STAGE;     // 196608
constexpr int T_MAIL_OFF = T_BAR_OFF + 256;
constexpr int T_MAIL_BYTES = 2 * 128 * 9 * 8;   // [array][row][piece, pitch 9] x (mean, M2); also [128][17] for 64-column tiles
constexpr int T_VEC_OFF = T_MAIL_OFF + T_MAIL_BYTES;   // per-column parameter vectors of this tile: 8 x [128] floats
enum { V_BIAS0 = 0, V_CBIAS0, V_BIAS1, V_GA0, V_BE0, V_GA1, V_BE1, V_COUNT };
constexpr int T_SMEM_USED = T_VEC_OFF + V_COUNT * 128 * 4;
constexpr int T_SMEM = T_SMEM_USED + 1024;      // slack for the 1024-byte alignment of the ring
constexpr int T_SLD = 132;                      // floats per row of the epilogue staging tiles (conflict-free float4 rows)
constexpr float U_LN_EPS = 1e-5f;               // nn.LayerNorm default (mmcv build_norm_layer(dict(type='LN')))

enum { SLOT_POOLED = 0, SLOT_INP, SLOT_GATEIN, SLOT_MIX, SLOT_OBJ0, SLOT_ATT, SLOT_OBJ1, SLOT_HID0,
       SLOT_OBJ2 = SLOT_HID0 + 8, SLOT_HEAD0, SLOT_HEAD1, NSLOT };
constexpr size_t SLOT_ELEMS = 2 * 128 * 256;    // bf16 elements per slot (hi + lo)

enum { MODE_GENERIC = 0, MODE_DUAL = 1, MODE_GATE = 2, MODE_GENERIC64 = 3 };   // GENERIC64: 64-column tiles
constexpr int T_SLD64 = 68;                     // floats per row of the 64-column staging tile
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2 };

struct TcPass {
    int a_slot;             // arena slot of the A operand (+ split when ksplit > 1)
    int w_ffn;              // 0: weight stack with 256 columns, 1: the FFN-wide stack
    int w_row, w_lo, w_k0;  // row of output column 0 (hi plane), rows to the lo plane, first K column (+256*split)
};
struct TcJob {
    int unit0;              // arena unit of image 0 (= branch * B)
    int ntiles;             // 128-column tiles of this job; CTAs with blockIdx.x >= ntiles idle
    int npass;
    TcPass pass[2];
    // ---- epilogue
    const float *bias0, *cbias0, *count, *bias1;
    const float *ln0, *ln1;     // {gamma[256], beta[256]}: GENERIC -> 256-column group 0 / >= 1; DUAL -> acc0 / acc1
                                // (out halves); GATE -> input_norm_in / norm_in
    int act0, act1;
    const float* res;           // residual [R][ldr]
    int ldr;
    const float *mul0, *mul1;   // GATE: LN'ed input_out / param_out [R][256]
    float *Y, *Y1;              // fp32 outputs (Y1: DUAL acc1); split-K partials when ksplit > 1
    int ldy, nstore;
    int p_slot;                 // arena slot of 256-column group 0 of the bf16 output planes, or -1
    uint16_t* xplanes;          // external planes [unit][2][xrows][256] (kern_split), or null
    int xrows, xunit0;
};
constexpr int T_MAXJOBS = 5;
struct TcArgs {
    TcJob job[T_MAXJOBS];
    uint16_t* arena;
    int B, N, R, ksplit;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void st_cluster_f32x2(const void* local_smem_ptr, uint32_t rank, float a, float b) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_smem_ptr)), "r"(rank));
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }   // the 16 epilogue warps

__device__ __forceinline__ void tmem_ld32f(uint32_t taddr, float (&y)[32]) {
    uint32_t v[32];
    tmem_ld32(taddr, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(v[i]);
}
__device__ __forceinline__ void add_vec32(float (&y)[32], const float* v) {   // v: shared memory, same address in all lanes
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
        const float4 t = *reinterpret_cast<const float4*>(v + c);
        y[c] += t.x, y[c + 1] += t.y, y[c + 2] += t.z, y[c + 3] += t.w;
    }
}
__device__ __forceinline__ void fma_vec32(float (&y)[32], float s, const float* v) {
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
        const float4 t = *reinterpret_cast<const float4*>(v + c);
        y[c] += s * t.x, y[c + 1] += s * t.y, y[c + 2] += s * t.z, y[c + 3] += s * t.w;
    }
}
// (mean, M2) of 32 register values: 8 independent partial sums, two passes
__device__ __forceinline__ void stats32(const float (&y)[32], float& mean, float& m2) {
    float p[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) p[k] = y[k];
#pragma unroll
    for (int c = 8; c < 32; ++c) p[c & 7] += y[c];
    mean = (((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]))) * (1.f / 32.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) p[k] = (y[k] - mean) * (y[k] - mean);
#pragma unroll
    for (int c = 8; c < 32; ++c) p[c & 7] += (y[c] - mean) * (y[c] - mean);
    m2 = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
}
__device__ __forceinline__ void store32(float* d, const float (&y)[32]) {
#pragma unroll
    for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(d + c) = make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]);
}
__device__ __forceinline__ void load32(const float* d, float (&y)[32]) {
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
        const float4 t = *reinterpret_cast<const float4*>(d + c);
        y[c] = t.x, y[c + 1] = t.y, y[c + 2] = t.z, y[c + 3] = t.w;
    }
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
// (mean, M2) of the 4 * LANES values held by a group of LANES lanes (two-pass, in registers)
template <int LANES>
__device__ __forceinline__ void stats4(float4 v, float inv_n, float& mean, float& m2) {
    float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    mean = s * inv_n;
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    float q = (a * a + b * b) + (c * c + d * d);
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    m2 = q;
}
__device__ __forceinline__ float4 ln4(float4 v, float mean, float rstd, float4 ga, float4 be) {
    return make_float4((v.x - mean) * rstd * ga.x + be.x, (v.y - mean) * rstd * ga.y + be.y,
                       (v.z - mean) * rstd * ga.z + be.z, (v.w - mean) * rstd * ga.w + be.w);
}
__device__ __forceinline__ float4 act4(float4 v, int act) {
    if (act == ACT_RELU) return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    // ex2.approx / rcp.approx: ~1e-7 relative on the sigmoid (the exponent's range reduction error is scaled by
    // (1 - sigmoid) |x| <= 0.3); the IEEE division + expf version made the gate epilogue instruction-bound
    if (act == ACT_SIGMOID)
        return make_float4(__fdividef(1.f, 1.f + __expf(-v.x)), __fdividef(1.f, 1.f + __expf(-v.y)),
                           __fdividef(1.f, 1.f + __expf(-v.z)), __fdividef(1.f, 1.f + __expf(-v.w)));
    return v;
}
// 4 values -> bf16 hi / lo, 8 contiguous bytes in each plane
__device__ __forceinline__ void store_planes4(float4 v, uint16_t* hi_ptr, uint16_t* lo_ptr) {
    // two packed conversions give the four hi halves; widening a bf16 back to fp32 is a shift / a mask
    const uint32_t p01 = pack_bf16x2(v.x, v.y), p23 = pack_bf16x2(v.z, v.w);
    const float h0 = __uint_as_float(p01 << 16), h1 = __uint_as_float(p01 & 0xFFFF0000u);
    const float h2 = __uint_as_float(p23 << 16), h3 = __uint_as_float(p23 & 0xFFFF0000u);
    *reinterpret_cast<uint2*>(hi_ptr) = make_uint2(p01, p23);
    *reinterpret_cast<uint2*>(lo_ptr) = make_uint2(pack_bf16x2(v.x - h0, v.y - h1), pack_bf16x2(v.z - h2, v.w - h3));
}
__device__ __forceinline__ uint16_t* arena_row(uint16_t* arena, int unit, int slot, int plane, int r) {
    return arena + ((size_t)unit * NSLOT + slot) * SLOT_ELEMS + ((size_t)plane * 128 + r) * 256;
}

// LayerNorm statistics of a 256-wide group held as 8 pieces of 32 columns by (4 quarters x 2 CTAs) or (2 quarters x
// 4 CTAs): every holder publishes (mean, M2) of its piece into mailbox[arr][src][row] of every CTA of the cluster
// (DSMEM stores); after one cluster barrier each CTA merges the pieces (Chan et al.), which is as stable as a
// two-pass LayerNorm.
struct Mail {
    float2 (*box)[128][9];   // row pitch 9: the rows a warp publishes / merges fall into different banks
    __device__ __forceinline__ void publish(int arr, int src, int r, float mean, float m2, int csize) const {
        for (int k = 0; k < csize; ++k) st_cluster_f32x2(&box[arr][r][src], (uint32_t)k, mean, m2);
    }
    __device__ __forceinline__ void combine(int arr, int r, float& mean, float& rstd) const {
        float2 p[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) p[k] = box[arr][r][k];
        mean = (((p[0].x + p[1].x) + (p[2].x + p[3].x)) + ((p[4].x + p[5].x) + (p[6].x + p[7].x))) * 0.125f;
        float m2 = 0.f, dv = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) m2 += p[k].y, dv += (p[k].x - mean) * (p[k].x - mean);
        rstd = rsqrtf((m2 + 32.f * dv) * (1.f / 256.f) + U_LN_EPS);   // MUFU.RSQ, <= 2 ulp; no slow-path call
    }
};

// tensor map over a long-lived bf16 [rows][cols] buffer, box = [box_rows][64], 128-byte swizzle (cached by pf_update.cu)
int cached_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows = 128);

}  // namespace pf
