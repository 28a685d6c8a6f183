// K2 -- the small-N block of one decoder stage, both branches (mask / depth) per launch:
//   KernelUpdator (polyphonic/funcs/kernel_updator.py:55-93), inter-kernel multi-head attention + LN, FFN + LN and
//   the cls / mask / depth FC heads (polyphonic/kernel_update_head.py:245-288), with feat_transform folded into the
//   first and last linear layers (see include/pf_decoder.h).
//
// Building block: tcgemm_kernel -- Y[128 rows][128 cols] = epilogue(prologue(X...) @ W^T) per CTA on tcgen05.
//   * M = 128 = the N (=111) kernels of ONE image (rows >= N are zero): weights are streamed once per image;
//   * fp32-level accuracy from bf16 tensor cores: X = Xh + Xl, W = Wh + Wl (bf16 each) and
//     D = Xl*Wh + Xh*Wl + Xh*Wh (3 MMAs per K step, fp32 accumulate in TMEM), error ~2^-16 per product --
//     the reference computes these layers in fp32 and they feed LayerNorms;
//   * A operand: built by the prologue (all 6 warps) from fp32 activations -- a+b, a*b, a*b + c*d for the updator
//     gates, or LN(sum of split-K partials + bias + residual) after the FFN -- split into hi/lo and written K-major /
//     128B-swizzled into shared memory; for the FFN's second layer it is TMA-loaded from the bf16 hi/lo planes the
//     first layer emitted;
//   * B operand: weights pre-split into bf16 hi/lo planes at pack time, TMA-loaded [128 n][64 k] boxes, 3-stage ring,
//     issued BEFORE griddepcontrol.wait so the fetch overlaps the previous kernel (programmatic dependent launch);
//   * epilogue: thread = row (TMEM lane): + bias (+ count * folded bias) (+ residual) -> LayerNorm over the 256-wide
//     group (row statistics exchanged with the peer CTA of a 2-CTA cluster through DSMEM) -> ReLU / sigmoid ->
//     fp32 and/or bf16 hi/lo planes.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
// 13 launches per stage.  This block is latency / weight-streaming bound, not roofline bound; see DESIGN.md.
#include <string.h>

#include <mutex>

#include "pf_internal.h"
#include "pf_sm100.cuh"

namespace pf {

constexpr int T_TN = 128;                 // output columns per CTA
constexpr int T_KC = 64;                  // K per weight stage
constexpr int T_WST = 3;                  // weight ring depth
constexpr int T_K = 256;                  // K per CTA (FFN2 is split-K in slabs of 256)
constexpr int T_A_BYTES = 128 * T_K * 2;  // 65536 per plane: 4 K-blocks of [128 rows][64 k]
constexpr int T_W_PLANE = T_TN * T_KC * 2;      // 16384
constexpr int T_W_STAGE = 2 * T_W_PLANE;        // hi + lo
constexpr int T_THREADS = 192;
constexpr int T_SMEM_USED = 2 * T_A_BYTES + T_WST * T_W_STAGE + 128 /*barriers*/ + 2048 /*LN mailbox*/;   // 231552
constexpr int T_SMEM = 232448;            // the 227 KB opt-in maximum; the slack (896 B) absorbs the 1024-byte alignment
constexpr float U_LN_EPS = 1e-5f;         // nn.LayerNorm default (mmcv build_norm_layer(dict(type='LN')))

enum { PRO_PLAIN = 0, PRO_ADD = 1, PRO_MUL = 2, PRO_MIX = 3, PRO_SUMLN = 4, PRO_PLANES = 5 };
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2 };

struct TcBranch {
    // ---- A operand
    const float *X, *X2, *X3, *X4;
    int ldx, ldx2, ldx3, ldx4;
    const float* part;    // PRO_SUMLN: [nsplit][R][256] split-K partials
    int nsplit;
    const float* pbias;   // [256]
    const float* pres;    // residual [R][256]
    const float* pln;     // {gamma[256], beta[256]}
    float* xout;          // LN result written back [R][256] (by the column-tile-0 CTAs)
    int a_unit0;          // PRO_PLANES: first unit of this branch in the A planes
    const float* rowdot_w;  // optional: rowdot_out[row] = X'[row,:] . rowdot_w + rowdot_b
    float rowdot_b;
    float* rowdot_out;
    // ---- B operand: row of column 0 in the weight stack (hi plane); the lo plane starts w_lo_off rows later
    int w_row, w_lo_off;
    // ---- epilogue
    const float *bias, *cbias, *count, *res;
    int ldr;
    const float* ln[2];   // LayerNorm {gamma[256], beta[256]} of 256-column group min(group,1), or null
    int act[2];
    float* Y;             // fp32 output or null; split-K partials when ksplit > 1
    int ldy, Nout, nstore;
    uint16_t* planes;     // optional bf16 hi/lo output [unit][2][plane_rows][plane_ld]
    int plane_rows, plane_ld, plane_unit0;
};
struct TcArgs {
    TcBranch br[2];
    int B, N, R, pro, ksplit, cluster;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// programmatic dependent launch: everything before pdl_wait() may overlap the previous kernel of the stream
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f32(float* local_smem_ptr, uint32_t rank, float v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_smem_ptr)), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__global__ void __launch_bounds__(T_THREADS, 1)
tcgemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a, const TcArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA_hi = smem;
    uint8_t* sA_lo = smem + T_A_BYTES;
    uint8_t* sW = smem + 2 * T_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + T_WST * T_W_STAGE);
    uint64_t* full = bars;
    uint64_t* empty = bars + T_WST;
    uint64_t* accfull = bars + 2 * T_WST;
    uint64_t* abar = bars + 2 * T_WST + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * T_WST + 2);
    float(*s_mail)[2][128] = reinterpret_cast<float(*)[2][128]>(reinterpret_cast<uint8_t*>(bars) + 128);   // [LN pass][source CTA rank][row]
    if (smem + T_SMEM_USED > smem_raw + T_SMEM) __trap();   // dynamic smem base less aligned than assumed

    const TcBranch& g = args.br[blockIdx.z];
    const int tile = blockIdx.x, nb = tile * T_TN;
    const int b = blockIdx.y / args.ksplit, split = blockIdx.y % args.ksplit;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool active = nb < g.Nout;   // uniform per cluster (Nout is a multiple of 256 whenever clusters are used)
    const int N = args.N;
    constexpr int NK = T_K / T_KC;     // 4 weight stages per CTA
    const int kbase = split * T_K;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_w);
        for (int i = 0; i < T_WST; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(accfull, 1);
        mbar_init(abar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<T_TN>(tmem_slot);
    __syncthreads();
    // weights do not depend on the previous kernel: start the ring before the grid dependency is resolved
    if (active && threadIdx.x == 0) {
        for (int kc = 0; kc < T_WST; ++kc) {
            mbar_arrive_expect_tx(&full[kc], T_W_STAGE);
            tma_load_2d(sW + kc * T_W_STAGE, &tmap_w, &full[kc], kbase + kc * T_KC, g.w_row + nb, kEvictLast);
            tma_load_2d(sW + kc * T_W_STAGE + T_W_PLANE, &tmap_w, &full[kc], kbase + kc * T_KC, g.w_row + g.w_lo_off + nb,
                        kEvictLast);
        }
    }
    pdl_wait();                 // activations written by the previous kernel are visible from here on
    pdl_launch_dependents();    // let the next kernel start prefetching its weights

    if (active) {
        if (args.pro == PRO_PLANES) {
            if (threadIdx.x == 0) {   // A = bf16 hi/lo planes [unit][2][128][K_total], K slab of this split
                mbar_arrive_expect_tx(abar, 2 * T_A_BYTES);
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
                        tma_load_3d(sA_hi + h * T_A_BYTES + kb * (128 * 128), &tmap_a, abar, kbase + kb * 64, 0,
                                    (g.a_unit0 + b) * 2 + h, kEvictFirst);
            }
        } else {
            // ---------------- prologue: warp per row, lane owns k = 8*lane .. 8*lane+7
            const int k0 = lane * 8;
            const uint32_t aoff_k = (uint32_t)(lane >> 3) * (128 * 128);
            const bool do_side = (tile == 0 && split == 0);
#pragma unroll 2
            for (int r = warp; r < 128; r += T_THREADS / 32) {
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = 0.f;
                if (r < N) {
                    const size_t m = (size_t)b * N + r;
                    if (args.pro == PRO_SUMLN) {
                        for (int s2 = 0; s2 < g.nsplit; ++s2) {
                            const float* pp = g.part + ((size_t)s2 * args.R + m) * 256 + k0;
                            const float4 u = ld4(pp), v = ld4(pp + 4);
                            x[0] += u.x, x[1] += u.y, x[2] += u.z, x[3] += u.w;
                            x[4] += v.x, x[5] += v.y, x[6] += v.z, x[7] += v.w;
                        }
                        const float4 bu = ld4(g.pbias + k0), bv = ld4(g.pbias + k0 + 4);
                        const float4 ru = ld4(g.pres + m * 256 + k0), rv = ld4(g.pres + m * 256 + k0 + 4);
                        x[0] += bu.x + ru.x, x[1] += bu.y + ru.y, x[2] += bu.z + ru.z, x[3] += bu.w + ru.w;
                        x[4] += bv.x + rv.x, x[5] += bv.y + rv.y, x[6] += bv.z + rv.z, x[7] += bv.w + rv.w;
                        float sum = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) sum += x[i];
                        const float mean = warp_sum(sum) * (1.f / 256.f);
                        float sq = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) sq += (x[i] - mean) * (x[i] - mean);
                        const float rstd = 1.f / sqrtf(warp_sum(sq) * (1.f / 256.f) + U_LN_EPS);
                        const float4 gu = ld4(g.pln + k0), gv = ld4(g.pln + k0 + 4);
                        const float4 eu = ld4(g.pln + 256 + k0), ev = ld4(g.pln + 256 + k0 + 4);
                        const float gam[8] = {gu.x, gu.y, gu.z, gu.w, gv.x, gv.y, gv.z, gv.w};
                        const float bet[8] = {eu.x, eu.y, eu.z, eu.w, ev.x, ev.y, ev.z, ev.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) x[i] = (x[i] - mean) * rstd * gam[i] + bet[i];
                        if (do_side && g.xout) {
                            *reinterpret_cast<float4*>(g.xout + m * 256 + k0) = make_float4(x[0], x[1], x[2], x[3]);
                            *reinterpret_cast<float4*>(g.xout + m * 256 + k0 + 4) = make_float4(x[4], x[5], x[6], x[7]);
                        }
                    } else {
                        const float4 u = ld4(g.X + m * g.ldx + k0), v = ld4(g.X + m * g.ldx + k0 + 4);
                        x[0] = u.x, x[1] = u.y, x[2] = u.z, x[3] = u.w, x[4] = v.x, x[5] = v.y, x[6] = v.z, x[7] = v.w;
                        if (args.pro == PRO_ADD) {
                            if (g.X2) {
                                const float4 a = ld4(g.X2 + m * g.ldx2 + k0), c = ld4(g.X2 + m * g.ldx2 + k0 + 4);
                                x[0] += a.x, x[1] += a.y, x[2] += a.z, x[3] += a.w;
                                x[4] += c.x, x[5] += c.y, x[6] += c.z, x[7] += c.w;
                            }
                        } else if (args.pro == PRO_MUL || args.pro == PRO_MIX) {
                            const float4 a = ld4(g.X2 + m * g.ldx2 + k0), c = ld4(g.X2 + m * g.ldx2 + k0 + 4);
                            x[0] *= a.x, x[1] *= a.y, x[2] *= a.z, x[3] *= a.w;
                            x[4] *= c.x, x[5] *= c.y, x[6] *= c.z, x[7] *= c.w;
                            if (args.pro == PRO_MIX) {
                                const float4 p3 = ld4(g.X3 + m * g.ldx3 + k0), q3 = ld4(g.X3 + m * g.ldx3 + k0 + 4);
                                const float4 p4 = ld4(g.X4 + m * g.ldx4 + k0), q4 = ld4(g.X4 + m * g.ldx4 + k0 + 4);
                                x[0] += p3.x * p4.x, x[1] += p3.y * p4.y, x[2] += p3.z * p4.z, x[3] += p3.w * p4.w;
                                x[4] += q3.x * q4.x, x[5] += q3.y * q4.y, x[6] += q3.z * q4.z, x[7] += q3.w * q4.w;
                            }
                        }
                    }
                    if (do_side && g.rowdot_out) {
                        const float4 wu = ld4(g.rowdot_w + k0), wv = ld4(g.rowdot_w + k0 + 4);
                        float d = x[0] * wu.x + x[1] * wu.y + x[2] * wu.z + x[3] * wu.w + x[4] * wv.x + x[5] * wv.y +
                                  x[6] * wv.z + x[7] * wv.w;
                        d = warp_sum(d);
                        if (lane == 0) g.rowdot_out[m] = d + g.rowdot_b;
                    }
                }
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float h0 = bf16_round(x[2 * i]), h1 = bf16_round(x[2 * i + 1]);
                    hi[i] = pack_bf16x2(h0, h1);
                    lo[i] = pack_bf16x2(x[2 * i] - h0, x[2 * i + 1] - h1);
                }
                const uint32_t off = aoff_k + sw128_offset(r, lane & 7);
                *reinterpret_cast<uint4*>(sA_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(sA_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            fence_proxy_async_smem();
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (active && warp == 0 && lane == 0) {
        // ================= weight producer: remaining stages =================
        for (int kc = T_WST; kc < NK; ++kc) {
            const int s = kc % T_WST;
            mbar_wait(&empty[s], ((kc / T_WST) & 1) ^ 1);
            mbar_arrive_expect_tx(&full[s], T_W_STAGE);
            tma_load_2d(sW + s * T_W_STAGE, &tmap_w, &full[s], kbase + kc * T_KC, g.w_row + nb, kEvictLast);
            tma_load_2d(sW + s * T_W_STAGE + T_W_PLANE, &tmap_w, &full[s], kbase + kc * T_KC, g.w_row + g.w_lo_off + nb,
                        kEvictLast);
        }
    } else if (active && warp == 1 && lane == 0) {
        // ================= MMA issuer: D = Al*Bh + Ah*Bl + Ah*Bh =================
        constexpr uint32_t idesc = make_idesc_bf16(128, T_TN, 0, 0);
        const uint32_t a_hi = smem_u32(sA_hi), a_lo = smem_u32(sA_lo);
        if (args.pro == PRO_PLANES) mbar_wait(abar, 0);
        for (int kc = 0; kc < NK; ++kc) {
            const int s = kc % T_WST;
            mbar_wait(&full[s], (kc / T_WST) & 1);
            tc_fence_after();
            const uint32_t w_hi = smem_u32(sW + s * T_W_STAGE), w_lo = w_hi + T_W_PLANE;
#pragma unroll
            for (int k16 = 0; k16 < T_KC / 16; ++k16) {
                const uint32_t aoff = kc * (128 * 128) + k16 * 32;
                const uint64_t dah = make_smem_desc_sw128(a_hi + aoff, 16, 1024);
                const uint64_t dal = make_smem_desc_sw128(a_lo + aoff, 16, 1024);
                const uint64_t dbh = make_smem_desc_sw128(w_hi + k16 * 32, 16, 1024);
                const uint64_t dbl = make_smem_desc_sw128(w_lo + k16 * 32, 16, 1024);
                umma_bf16_ss(tmem_base, dal, dbh, idesc, (kc | k16) != 0);
                umma_bf16_ss(tmem_base, dah, dbl, idesc, 1);
                umma_bf16_ss(tmem_base, dah, dbh, idesc, 1);
            }
            umma_commit(&empty[s]);
        }
        umma_commit(accfull);
    }
    __syncwarp();

    // ================= epilogue (warps 2..5: thread = TMEM lane = kernel row) =================
    const int grp = tile >> 1;
    const int gi = grp < 1 ? grp : 1;
    const float* ln = active ? g.ln[gi] : nullptr;
    const uint32_t crank = tile & 1;
    const bool epi = active && warp >= 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const bool rok = r < N;
    const size_t m = (size_t)b * N + (rok ? r : 0);
    float y[T_TN];
    float mean = 0.f, rstd = 1.f;
    if (epi) {
        mbar_wait(accfull, 0);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < T_TN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) y[c0 + i] = __uint_as_float(v[i]);
        }
        if (args.ksplit == 1) {
            const float cnt = (g.count && rok) ? __ldg(g.count + m) : 0.f;
#pragma unroll
            for (int c = 0; c < T_TN; c += 4) {
                const int col = nb + c;
                if (col < g.Nout) {   // Nout is a multiple of 4
                    if (g.bias) {
                        const float4 t = ld4(g.bias + col);
                        y[c] += t.x, y[c + 1] += t.y, y[c + 2] += t.z, y[c + 3] += t.w;
                    }
                    if (g.cbias) {
                        const float4 t = ld4(g.cbias + col);
                        y[c] += cnt * t.x, y[c + 1] += cnt * t.y, y[c + 2] += cnt * t.z, y[c + 3] += cnt * t.w;
                    }
                    if (g.res && rok) {
                        const float4 t = ld4(g.res + m * g.ldr + col);
                        y[c] += t.x, y[c + 1] += t.y, y[c + 2] += t.z, y[c + 3] += t.w;
                    }
                }
            }
        }
        if (ln) {
            float s1 = 0.f;
#pragma unroll
            for (int c = 0; c < T_TN; ++c) s1 += y[c];
            s_mail[0][crank][r] = s1;
            st_cluster_f32(&s_mail[0][crank][r], crank ^ 1u, s1);
        }
    }
    if (ln) cluster_sync_all();   // every thread of both CTAs
    if (epi && ln) {
        mean = (s_mail[0][0][r] + s_mail[0][1][r]) * (1.f / 256.f);
        float s2 = 0.f;
#pragma unroll
        for (int c = 0; c < T_TN; ++c) s2 += (y[c] - mean) * (y[c] - mean);
        s_mail[1][crank][r] = s2;
        st_cluster_f32(&s_mail[1][crank][r], crank ^ 1u, s2);
    }
    if (ln) cluster_sync_all();
    if (epi) {
        if (ln) {
            rstd = 1.f / sqrtf((s_mail[1][0][r] + s_mail[1][1][r]) * (1.f / 256.f) + U_LN_EPS);
            const int cg = (tile & 1) * T_TN;   // column inside the 256-wide group
#pragma unroll
            for (int c = 0; c < T_TN; c += 4) {
                const float4 ga = ld4(ln + cg + c), be = ld4(ln + 256 + cg + c);
                y[c] = (y[c] - mean) * rstd * ga.x + be.x, y[c + 1] = (y[c + 1] - mean) * rstd * ga.y + be.y;
                y[c + 2] = (y[c + 2] - mean) * rstd * ga.z + be.z, y[c + 3] = (y[c + 3] - mean) * rstd * ga.w + be.w;
            }
        }
        const int act = g.act[gi];
        if (act == ACT_RELU) {
#pragma unroll
            for (int c = 0; c < T_TN; ++c) y[c] = fmaxf(y[c], 0.f);
        } else if (act == ACT_SIGMOID) {
#pragma unroll
            for (int c = 0; c < T_TN; ++c) y[c] = 1.f / (1.f + expf(-y[c]));
        }
        if (g.Y && rok) {
            float* dst = g.Y + ((size_t)split * args.R + m) * g.ldy + nb;
            if ((g.ldy & 3) == 0 && nb + T_TN <= g.nstore) {
#pragma unroll
                for (int c = 0; c < T_TN; c += 4)
                    *reinterpret_cast<float4*>(dst + c) = make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]);
            } else {
#pragma unroll
                for (int c = 0; c < T_TN; ++c)
                    if (nb + c < g.nstore) dst[c] = y[c];
            }
        }
        if (g.planes && r < g.plane_rows) {
            const size_t unit = (size_t)g.plane_unit0 + b;
            uint16_t* ph = g.planes + ((unit * 2) * g.plane_rows + r) * g.plane_ld + nb;
            uint16_t* pl = ph + (size_t)g.plane_rows * g.plane_ld;
#pragma unroll
            for (int c = 0; c < T_TN; c += 8) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float v0 = rok ? y[c + 2 * i] : 0.f, v1 = rok ? y[c + 2 * i + 1] : 0.f;
                    const float h0 = bf16_round(v0), h1 = bf16_round(v1);
                    hi[i] = pack_bf16x2(h0, h1);
                    lo[i] = pack_bf16x2(v0 - h0, v1 - h1);
                }
                *reinterpret_cast<uint4*>(ph + c) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(pl + c) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<T_TN>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// Inter-kernel self-attention of one (branch, image, head): softmax(q k^T / sqrt(32)) v over the N kernels of the
// image (mmcv MultiheadAttention -> nn.MultiheadAttention, seq-first; kernel_update_head.py:259-260).
// qkv [R][768] = [q | k | v]; out [R][256] (heads concatenated), before out_proj.
// 4 threads per query; keys are dealt to the 4 threads in blocks of 4 consecutive keys so that K^T rows and V rows
// are read as float4 (4 FMAs per shared-memory load).  Two-pass softmax with the scores kept in registers, quad
// reduction by shuffles.  Two CTAs per (branch, image, head) split the queries.
constexpr int ATT_NB = PF_MAX_N / 16;   // key blocks of 4 per thread (8 -> 32 keys per thread, 128 per quad)
constexpr int ATT_QPC = PF_MAX_N / 2;   // queries per CTA
__global__ void __launch_bounds__(ATT_QPC * 4) attention_kernel(const float* __restrict__ qkv0,
                                                                const float* __restrict__ qkv1,
                                                                float* __restrict__ out0, float* __restrict__ out1,
                                                                int N) {
    __shared__ __align__(16) float s_kt[32][PF_MAX_N + 4];   // K transposed: [d][key]
    __shared__ __align__(16) float s_v[PF_MAX_N][36];
    const int h = blockIdx.x >> 1, half = blockIdx.x & 1, b = blockIdx.y;
    pdl_wait();
    pdl_launch_dependents();
    const float* qkv = (blockIdx.z == 0 ? qkv0 : qkv1) + (size_t)b * N * 768;
    float* out = (blockIdx.z == 0 ? out0 : out1) + (size_t)b * N * 256;
    for (int i = threadIdx.x; i < PF_MAX_N * 32; i += ATT_QPC * 4) {
        const int n = i >> 5, d = i & 31;
        const bool ok = n < N;
        s_kt[d][n] = ok ? qkv[(size_t)n * 768 + 256 + h * 32 + d] : 0.f;
        s_v[n][d] = ok ? qkv[(size_t)n * 768 + 512 + h * 32 + d] : 0.f;
    }
    __syncthreads();
    const int n = half * ATT_QPC + (threadIdx.x >> 2), part = threadIdx.x & 3;
    const int nq = n < N ? n : N - 1;   // keep whole quads alive for the shuffles
    float q[32];
    const float scale = 0.17677669529663687f;  // 1/sqrt(32), applied to q before q k^T as torch does
#pragma unroll
    for (int d4 = 0; d4 < 8; ++d4) {
        const float4 t = *reinterpret_cast<const float4*>(qkv + (size_t)nq * 768 + h * 32 + d4 * 4);
        q[d4 * 4] = t.x * scale, q[d4 * 4 + 1] = t.y * scale, q[d4 * 4 + 2] = t.z * scale, q[d4 * 4 + 3] = t.w * scale;
    }
    float sc[ATT_NB][4];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < ATT_NB; ++i) {
        const int j0 = 4 * (part + 4 * i);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int d = 0; d < 32; ++d) {
            const float4 kk = *reinterpret_cast<const float4*>(&s_kt[d][j0]);
            a.x += q[d] * kk.x, a.y += q[d] * kk.y, a.z += q[d] * kk.z, a.w += q[d] * kk.w;
        }
        sc[i][0] = j0 < N ? a.x : -INFINITY, sc[i][1] = j0 + 1 < N ? a.y : -INFINITY;
        sc[i][2] = j0 + 2 < N ? a.z : -INFINITY, sc[i][3] = j0 + 3 < N ? a.w : -INFINITY;
        mx = fmaxf(fmaxf(mx, fmaxf(sc[i][0], sc[i][1])), fmaxf(sc[i][2], sc[i][3]));
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float o[32], den = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] = 0.f;
#pragma unroll
    for (int i = 0; i < ATT_NB; ++i) {
        const int j0 = 4 * (part + 4 * i);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float pj = expf(sc[i][e] - mx);   // exp(-inf) = 0 for padded keys (their V rows are zero)
            den += pj;
#pragma unroll
            for (int d4 = 0; d4 < 8; ++d4) {
                const float4 vv = *reinterpret_cast<const float4*>(&s_v[j0 + e][d4 * 4]);
                o[d4 * 4] += pj * vv.x, o[d4 * 4 + 1] += pj * vv.y, o[d4 * 4 + 2] += pj * vv.z, o[d4 * 4 + 3] += pj * vv.w;
            }
        }
    }
    den += __shfl_xor_sync(0xffffffffu, den, 1);
    den += __shfl_xor_sync(0xffffffffu, den, 2);
    const float inv = 1.f / den;
#pragma unroll
    for (int d = 0; d < 32; ++d) {
        o[d] += __shfl_xor_sync(0xffffffffu, o[d], 1);
        o[d] += __shfl_xor_sync(0xffffffffu, o[d], 2);
    }
    if (n < N) {
        float* orow = out + (size_t)n * 256 + h * 32;
#pragma unroll
        for (int pp = 0; pp < 4; ++pp)   // lane `part` writes dims [8 part, 8 part + 8)
            if (part == pp) {
                *reinterpret_cast<float4*>(orow + pp * 8) =
                    make_float4(o[pp * 8] * inv, o[pp * 8 + 1] * inv, o[pp * 8 + 2] * inv, o[pp * 8 + 3] * inv);
                *reinterpret_cast<float4*>(orow + pp * 8 + 4) =
                    make_float4(o[pp * 8 + 4] * inv, o[pp * 8 + 5] * inv, o[pp * 8 + 6] * inv, o[pp * 8 + 7] * inv);
            }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
struct UpdateScratch {
    float *params, *inp, *gate, *obj0, *qkv, *att, *obj1, *head, *part;
    uint16_t* hid;   // bf16 hi/lo planes [B][2][128][ffn]
};
constexpr int kFfnSplit = 8;

// tensor maps over the (long-lived) weight stacks are cached: encoding one costs ~1 us of host time per call otherwise
static int cached_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols) {
    struct Entry { const void* base; uint64_t rows, cols; CUtensorMap map; };
    static Entry cache[64];
    static int n = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < n; ++i)
        if (cache[i].base == base && cache[i].rows == rows && cache[i].cols == cols) {
            *out = cache[i].map;
            return PF_OK;
        }
    CUtensorMap m;
    if (int e = make_tmap_bf16_2d(&m, base, rows, cols, cols, T_TN, T_KC)) return e;
    if (n < 64) cache[n++] = Entry{base, rows, cols, m};
    *out = m;
    return PF_OK;
}

static int launch_tc(const CUtensorMap& tw, const CUtensorMap& ta, TcArgs a, int max_nout, int nbranch, cudaStream_t st) {
    const int tiles = (max_nout + T_TN - 1) / T_TN;
    a.cluster = (tiles % 2 == 0) ? 2 : 1;
    cudaError_t e = cudaFuncSetAttribute(tcgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM);
    if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "tcgemm smem attribute: %s", cudaGetErrorString(e));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(tiles, a.B * a.ksplit, nbranch);
    cfg.blockDim = dim3(T_THREADS);
    cfg.dynamicSmemBytes = T_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    attrs[1].id = cudaLaunchAttributeClusterDimension;
    attrs[1].val.clusterDim.x = a.cluster, attrs[1].val.clusterDim.y = 1, attrs[1].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 2;
    e = cudaLaunchKernelEx(&cfg, tcgemm_kernel, tw, ta, a);
    if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "tcgemm_kernel launch: %s", cudaGetErrorString(e));
    count_launch();
    return PF_OK;
}

static int launch_attention(const float* q0, const float* q1, float* o0, float* o1, int B, int N, int nbranch,
                            cudaStream_t st) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(PF_HEADS * 2, B, nbranch);
    cfg.blockDim = dim3(ATT_QPC * 4);
    cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, attention_kernel, q0, q1, o0, o1, N);
    if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "attention_kernel launch: %s", cudaGetErrorString(e));
    count_launch();
    return PF_OK;
}

static TcBranch blank() {
    TcBranch g;
    memset(&g, 0, sizeof(g));
    return g;
}
static size_t align256(size_t v) { return (v + 255) / 256 * 256; }
static size_t update_ws_bytes(int B, int N, int ffn) {
    const size_t R = (size_t)B * N;
    size_t per_branch = R * (512 * 3 + 256 * 3 + 768 + 512) * 4 + (size_t)kFfnSplit * R * 256 * 4;
    per_branch = align256(per_branch) + align256((size_t)B * 2 * 128 * ffn * 2);
    return 2 * per_branch + align256(2 * R * 256 * 4) + align256(R * 4) + 256;
}

}  // namespace pf

extern "C" size_t pf_update_workspace_bytes(int B, int N, int ffn_channels) {
    if (B <= 0 || N <= 0 || ffn_channels <= 0) return 0;
    return pf::update_ws_bytes(B, N, ffn_channels);
}

extern "C" int pf_kernel_update(const pf_stage_weights* w, const float* partial, const float* cntp, int S,
                                const float* obj_in, const float* dep_in, float* obj_out, float* dep_out,
                                float* cls_out, float* kern, uint16_t* kern_split, float* kbias, void* workspace,
                                size_t workspace_bytes, int B, int N, int cls_sigmoid, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(w && partial && cntp && obj_in && dep_in && obj_out && dep_out && kern_split && kbias && workspace, PF_ERR_ARG,
               "pf_kernel_update: null pointer");
    PF_REQUIRE(B > 0 && N > 0 && N <= PF_MAX_N && S > 0, PF_ERR_ARG, "pf_kernel_update: bad shape B=%d N=%d S=%d", B, N, S);
    const int ffn = w->ffn_channels;
    PF_REQUIRE(ffn == kFfnSplit * T_K, PF_ERR_ARG, "pf_kernel_update: ffn_channels=%d (this build supports %d)", ffn, kFfnSplit * T_K);
    PF_REQUIRE(w->num_classes > 0 && w->num_classes <= PF_MAX_CLASSES, PF_ERR_ARG, "pf_kernel_update: num_classes=%d", w->num_classes);
    PF_REQUIRE(w->wstack256 && w->wstack_ffn, PF_ERR_ARG, "pf_kernel_update: weight stacks missing");
    PF_REQUIRE(workspace_bytes >= update_ws_bytes(B, N, ffn), PF_ERR_WORKSPACE, "pf_kernel_update: workspace %zu < %zu",
               workspace_bytes, update_ws_bytes(B, N, ffn));
    PF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PF_ERR_ALIGN, "pf_kernel_update: workspace not 256-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int R = B * N;

    UpdateScratch sc[2];
    char* wp = static_cast<char*>(workspace);
    for (int b = 0; b < 2; ++b) {
        float* p = reinterpret_cast<float*>(wp);
        sc[b].params = p, p += (size_t)R * 512;
        sc[b].inp = p, p += (size_t)R * 512;
        sc[b].gate = p, p += (size_t)R * 512;
        sc[b].obj0 = p, p += (size_t)R * 256;
        sc[b].qkv = p, p += (size_t)R * 768;
        sc[b].att = p, p += (size_t)R * 256;
        sc[b].obj1 = p, p += (size_t)R * 256;
        sc[b].head = p, p += (size_t)R * 512;
        sc[b].part = p, p += (size_t)kFfnSplit * R * 256;
        wp += align256(reinterpret_cast<char*>(p) - wp);
        sc[b].hid = reinterpret_cast<uint16_t*>(wp);
        wp += align256((size_t)B * 2 * 128 * ffn * 2);
    }
    float* pooled = reinterpret_cast<float*>(wp);   // [2][R][256]
    wp += align256((size_t)2 * R * 256 * 4);
    float* count = reinterpret_cast<float*>(wp);    // [R]
    const float* in_[2] = {obj_in, dep_in};
    float* out_[2] = {obj_out, dep_out};

    CUtensorMap tw, tf, th;
    if (int e = cached_tmap_2d(&tw, w->wstack256, (uint64_t)w->wstack256_rows, 256)) return e;
    if (int e = cached_tmap_2d(&tf, w->wstack_ffn, (uint64_t)w->wstack_ffn_rows, (uint64_t)ffn)) return e;
    // hid planes of both branches are adjacent in the workspace only up to alignment: one map per branch
    CUtensorMap thid[2];
    for (int b = 0; b < 2; ++b)
        if (int e = make_tmap_bf16_3d(&thid[b], sc[b].hid, (uint64_t)B * 2, 128, (uint64_t)ffn, 128, 64)) return e;
    (void)th;

    // 0. deterministic sum of the split-K pooling partials (fixed order)
    if (int e = pf_pool_reduce(partial, cntp, pooled, count, B, N, 2, S, stream)) return e;

    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B, a.N = N, a.R = R, a.ksplit = 1;

    // 1. parameters = dynamic_layer(pooled W_t^T + count b_t); param_out -> norm_out     (kernel_updator.py:58-62,78)
    a.pro = PRO_PLAIN;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcBranch g = blank();
        g.X = pooled + (size_t)b * R * 256, g.ldx = 256, g.count = count;
        g.w_row = bw.dyn_w, g.w_lo_off = 512, g.bias = bw.dyn_b, g.cbias = bw.dyn_cb;
        g.ln[1] = bw.ln_norm_out;
        g.Y = sc[b].params, g.ldy = 512, g.Nout = 512, g.nstore = 512;
        a.br[b] = g;
    }
    if (int e = launch_tc(tw, tw, a, 512, 2, st)) return e;

    // 2. input_feats = input_layer(kernel); depth kernel += mask kernel (kernel_update_head.py:250); input_out -> input_norm_out
    a.pro = PRO_ADD;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcBranch g = blank();
        g.X = in_[b], g.ldx = 256;
        if (b == 1) g.X2 = obj_in, g.ldx2 = 256;
        g.w_row = bw.inp_w, g.w_lo_off = 512, g.bias = bw.inp_b;
        g.ln[1] = bw.ln_input_norm_out;
        g.Y = sc[b].inp, g.ldy = 512, g.Nout = 512, g.nstore = 512;
        a.br[b] = g;
    }
    if (int e = launch_tc(tw, tw, a, 512, 2, st)) return e;

    // 3. gate_feats = input_in * param_in; [input_gate | update_gate] = sigmoid(LN(W g + b))   (:69,73-77)
    a.pro = PRO_MUL;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcBranch g = blank();
        g.X = sc[b].inp, g.ldx = 512, g.X2 = sc[b].params, g.ldx2 = 512;
        g.w_row = bw.gate_w, g.w_lo_off = 512, g.bias = bw.gate_b;
        g.ln[0] = bw.ln_input_norm_in, g.ln[1] = bw.ln_norm_in;
        g.act[0] = g.act[1] = ACT_SIGMOID;
        g.Y = sc[b].gate, g.ldy = 512, g.Nout = 512, g.nstore = 512;
        a.br[b] = g;
    }
    if (int e = launch_tc(tw, tw, a, 512, 2, st)) return e;

    // 4. features = update_gate * param_out + input_gate * input_out; relu(fc_norm(fc_layer(.)))   (:86-91)
    a.pro = PRO_MIX;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcBranch g = blank();
        g.X = sc[b].gate + 256, g.ldx = 512, g.X2 = sc[b].params + 256, g.ldx2 = 512;
        g.X3 = sc[b].gate, g.ldx3 = 512, g.X4 = sc[b].inp + 256, g.ldx4 = 512;
        g.w_row = bw.fc_w, g.w_lo_off = 256, g.bias = bw.fc_b, g.ln[0] = bw.ln_fc_norm, g.act[0] = ACT_RELU;
        g.Y = sc[b].obj0, g.ldy = 256, g.Nout = 256, g.nstore = 256;
        a.br[b] = g;
    }
    if (int e = launch_tc(tw, tw, a, 256, 2, st)) return e;

    // 5. attention in-projection
    a.pro = PRO_PLAIN;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcBranch g = blank();
        g.X = sc[b].obj0, g.ldx = 256, g.w_row = bw.qkv_w, g.w_lo_off = 768, g.bias = bw.qkv_b;
        g.Y = sc[b].qkv, g.ldy = 768, g.Nout = 768, g.nstore = 768;
        a.br[b] = g;
    }
    if (int e = launch_tc(tw, tw, a, 768, 2, st)) return e;

    // 6. softmax(q k^T) v per (branch, image, head)
    if (int e = launch_attention(sc[0].qkv, sc[1].qkv, sc[0].att, sc[1].att, B, N, 2, st)) return e;

    // 7. attention_norm(x + out_proj(attn))
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcBranch g = blank();
        g.X = sc[b].att, g.ldx = 256, g.w_row = bw.out_w, g.w_lo_off = 256, g.bias = bw.out_b;
        g.res = sc[b].obj0, g.ldr = 256, g.ln[0] = bw.ln_attn;
        g.Y = sc[b].obj1, g.ldy = 256, g.Nout = 256, g.nstore = 256;
        a.br[b] = g;
    }
    if (int e = launch_tc(tw, tw, a, 256, 2, st)) return e;

    // 8. FFN layer 1 + ReLU -> bf16 hi/lo planes (the A operand of layer 2)
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcBranch g = blank();
        g.X = sc[b].obj1, g.ldx = 256, g.w_row = bw.ffn1_w, g.w_lo_off = ffn, g.bias = bw.ffn1_b;
        g.act[0] = g.act[1] = ACT_RELU;
        g.Nout = ffn, g.nstore = ffn;
        g.planes = sc[b].hid, g.plane_rows = 128, g.plane_ld = ffn, g.plane_unit0 = 0;
        a.br[b] = g;
    }
    if (int e = launch_tc(tw, tw, a, ffn, 2, st)) return e;

    // 9. FFN layer 2, split-K over kFfnSplit slabs of 256 -> fp32 partials.  One launch per branch (its own A map).
    a.pro = PRO_PLANES, a.ksplit = kFfnSplit;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcBranch g = blank();
        g.a_unit0 = 0, g.w_row = bw.ffn2_w, g.w_lo_off = 256;
        g.Y = sc[b].part, g.ldy = 256, g.Nout = 256, g.nstore = 256;
        a.br[0] = g;
        if (int e = launch_tc(tf, thid[b], a, 256, 1, st)) return e;
    }
    a.ksplit = 1;

    // 10. ffn_norm(x + sum of partials + b2) -> obj_feat / depth_feat_new (returned to the caller) in the prologue,
    //     then cls_fcs / mask_fcs / depth_regs: Linear(no bias) + LN (+ ReLU except depth_regs)
    a.pro = PRO_SUMLN;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcBranch g = blank();
        g.part = sc[b].part, g.nsplit = kFfnSplit, g.pbias = bw.ffn2_b, g.pres = sc[b].obj1, g.pln = bw.ln_ffn;
        g.xout = out_[b];
        g.w_row = bw.head_w, g.w_lo_off = (b == 0) ? 512 : 256;
        g.ln[0] = bw.ln_head_a, g.ln[1] = bw.ln_head_b;
        g.act[0] = g.act[1] = bw.head_relu ? ACT_RELU : ACT_NONE;
        g.Nout = (b == 0) ? 512 : 256;
        g.Y = sc[b].head, g.ldy = 512, g.nstore = g.Nout;
        a.br[b] = g;
    }
    if (int e = launch_tc(tw, tw, a, 512, 2, st)) return e;

    // 11. fc_mask / fc_depth with feat_transform folded in -> dynamic kernels (bf16 hi/lo planes) + their logit bias
    a.pro = PRO_PLAIN;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcBranch g = blank();
        g.X = sc[b].head + (b == 0 ? 256 : 0), g.ldx = 512, g.w_row = bw.kern_w, g.w_lo_off = 256, g.bias = bw.kern_b;
        g.Y = kern ? kern + (size_t)b * R * 256 : nullptr, g.ldy = 256, g.Nout = 256, g.nstore = 256;
        g.planes = kern_split, g.plane_rows = N, g.plane_ld = 256, g.plane_unit0 = b * B;
        g.rowdot_w = bw.kb_w, g.rowdot_b = bw.kb_b, g.rowdot_out = kbias + (size_t)b * R;
        a.br[b] = g;
    }
    if (int e = launch_tc(tw, tw, a, 256, 2, st)) return e;

    // 12. fc_cls (mask branch only)
    if (cls_out) {
        const pf_branch_weights& bw = w->br[0];
        TcBranch g = blank();
        g.X = sc[0].head, g.ldx = 512, g.w_row = bw.cls_w, g.w_lo_off = 128, g.bias = bw.cls_b;
        g.act[0] = g.act[1] = cls_sigmoid ? ACT_SIGMOID : ACT_NONE;
        g.Y = cls_out, g.ldy = w->num_classes, g.Nout = PF_MAX_CLASSES, g.nstore = w->num_classes;
        a.br[0] = g;
        if (int e = launch_tc(tw, tw, a, PF_MAX_CLASSES, 1, st)) return e;
    }
    return PF_OK;
}

extern "C" size_t pf_updator_workspace_bytes(int R) { return R > 0 ? (size_t)R * 1536 * sizeof(float) + 256 : 0; }

// KernelUpdator.forward on its own (polyphonic/funcs/kernel_updator.py:55-93): no feat_transform fold, one branch.
// Rows are processed in groups of <= 128 ("images" of N = 128 rows).
extern "C" int pf_kernel_updator(const pf_stage_weights* w, const float* update_feature, const float* input_feature,
                                 float* out, void* workspace, size_t workspace_bytes, int R, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(w && update_feature && input_feature && out && workspace, PF_ERR_ARG, "pf_kernel_updator: null pointer");
    PF_REQUIRE(R > 0 && w->wstack256, PF_ERR_ARG, "pf_kernel_updator: R=%d", R);
    PF_REQUIRE(workspace_bytes >= pf_updator_workspace_bytes(R), PF_ERR_WORKSPACE, "pf_kernel_updator: workspace %zu < %zu",
               workspace_bytes, pf_updator_workspace_bytes(R));
    PF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && (reinterpret_cast<uintptr_t>(update_feature) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(input_feature) & 15) == 0,
               PF_ERR_ALIGN, "pf_kernel_updator: pointers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const pf_branch_weights* bw = &w->br[0];
    CUtensorMap tw;
    if (int e = cached_tmap_2d(&tw, w->wstack256, (uint64_t)w->wstack256_rows, 256)) return e;
    float* params = static_cast<float*>(workspace);
    float* inp = params + (size_t)R * 512;
    float* gate = inp + (size_t)R * 512;
    // groups of 128 rows; the last group may be ragged: run it as a second launch set with its own N
    for (int r0 = 0; r0 < R;) {
        const int full_groups = (R - r0) / 128;
        const int Bg = full_groups > 0 ? full_groups : 1;
        const int Ng = full_groups > 0 ? 128 : (R - r0);
        const int Rg = Bg * Ng;
        TcArgs a;
        memset(&a, 0, sizeof(a));
        a.B = Bg, a.N = Ng, a.R = Rg, a.ksplit = 1;
        TcBranch g;

        a.pro = PRO_PLAIN;
        g = blank();
        g.X = update_feature + (size_t)r0 * 256, g.ldx = 256, g.w_row = bw->dyn_w, g.w_lo_off = 512, g.bias = bw->dyn_b;
        g.ln[1] = bw->ln_norm_out;
        g.Y = params + (size_t)r0 * 512, g.ldy = 512, g.Nout = 512, g.nstore = 512;
        a.br[0] = g;
        if (int e = launch_tc(tw, tw, a, 512, 1, st)) return e;

        g = blank();
        g.X = input_feature + (size_t)r0 * 256, g.ldx = 256, g.w_row = bw->inp_w, g.w_lo_off = 512, g.bias = bw->inp_b;
        g.ln[1] = bw->ln_input_norm_out;
        g.Y = inp + (size_t)r0 * 512, g.ldy = 512, g.Nout = 512, g.nstore = 512;
        a.br[0] = g;
        if (int e = launch_tc(tw, tw, a, 512, 1, st)) return e;

        a.pro = PRO_MUL;
        g = blank();
        g.X = inp + (size_t)r0 * 512, g.ldx = 512, g.X2 = params + (size_t)r0 * 512, g.ldx2 = 512;
        g.w_row = bw->gate_w, g.w_lo_off = 512, g.bias = bw->gate_b;
        g.ln[0] = bw->ln_input_norm_in, g.ln[1] = bw->ln_norm_in, g.act[0] = g.act[1] = ACT_SIGMOID;
        g.Y = gate + (size_t)r0 * 512, g.ldy = 512, g.Nout = 512, g.nstore = 512;
        a.br[0] = g;
        if (int e = launch_tc(tw, tw, a, 512, 1, st)) return e;

        a.pro = PRO_MIX;
        g = blank();
        g.X = gate + (size_t)r0 * 512 + 256, g.ldx = 512, g.X2 = params + (size_t)r0 * 512 + 256, g.ldx2 = 512;
        g.X3 = gate + (size_t)r0 * 512, g.ldx3 = 512, g.X4 = inp + (size_t)r0 * 512 + 256, g.ldx4 = 512;
        g.w_row = bw->fc_w, g.w_lo_off = 256, g.bias = bw->fc_b, g.ln[0] = bw->ln_fc_norm, g.act[0] = ACT_RELU;
        g.Y = out + (size_t)r0 * 256, g.ldy = 256, g.Nout = 256, g.nstore = 256;
        a.br[0] = g;
        if (int e = launch_tc(tw, tw, a, 256, 1, st)) return e;
        r0 += Rg;
    }
    return PF_OK;
}
