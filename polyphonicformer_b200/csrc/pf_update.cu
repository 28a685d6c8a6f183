// K2 -- the small-N block of one decoder stage, both branches (mask / depth) per launch:
//   KernelUpdator (polyphonic/funcs/kernel_updator.py:55-93), inter-kernel multi-head attention + LN, FFN + LN and
//   the cls / mask / depth FC heads (polyphonic/kernel_update_head.py:245-288), with feat_transform folded into the
//   first and last linear layers (see include/pf_decoder.h).
//
// Building block: rowgemm_kernel -- Y[16 rows][256 cols] = epilogue(prologue(X...) @ W^T) per CTA.
//   * prologue builds the [16][256] input tile in shared memory (sum of split-K pooling partials, a+b, a*b,
//     a*b + c*d for the updator gates) and splits it into tf32 hi + lo;
//   * the product runs on mma.sync.m16n8k8 tf32 with the 3-term split (hi*hi + lo*hi + hi*lo), fp32 accumulate,
//     i.e. fp32-level accuracy: the reference computes these layers in fp32 and they feed LayerNorms;
//   * epilogue: + bias (+ count * folded bias) (+ residual) -> LayerNorm over the 256-wide tile -> ReLU / sigmoid.
// 12 launches per stage; rows = B*N (111 per image).  This block is latency / weight-streaming bound, not
// roofline bound; see DESIGN.md.
#include <string.h>

#include "pf_internal.h"
#include "pf_sm100.cuh"

namespace pf {

constexpr int V_TM = 32;             // rows per CTA
constexpr int V_TN = 64;             // columns per CTA; 4 CTAs (one cluster) cover a 256-wide LayerNorm group
constexpr int V_KS = 64;             // K per pipeline stage
constexpr int V_NST = 3;             // cp.async ring depth for the weight tiles
constexpr int V_LDW = V_KS + 16;     // 80 floats: stride = 16 (mod 32) -> conflict-free float4 fragment loads
constexpr int V_LDX = 256 + 16;      // 272 floats, same property; 4 K-chunks of 64 side by side
constexpr int V_THREADS = 128;
constexpr int V_CL = 4;              // cluster size along the column tiles
constexpr int V_SMEM_X = V_TM * V_LDX * 4;             // 34816
constexpr int V_SMEM_W = V_NST * V_TN * V_LDW * 4;     // 61440
constexpr int V_SMEM = V_SMEM_X + V_SMEM_W;            // 96256 -> 2 CTAs / SM
constexpr float U_LN_EPS = 1e-5f;    // nn.LayerNorm default (mmcv build_norm_layer(dict(type='LN')))

enum { PRO_PLAIN = 0, PRO_ADD = 1, PRO_MUL = 2, PRO_MIX = 3 };
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2 };

struct GemmBranch {
    const float *X, *X2, *X3, *X4;
    int ldx, ldx2, ldx3, ldx4;
    const float* W;      // [Nout][K]
    const float* bias;   // [Nout] or null
    const float* cbias;  // [Nout] or null: + count[row] * cbias (feat_transform bias folded through the pooling)
    const float* count;  // [R] or null
    const float* res;    // residual [R][ldr] or null
    int ldr;
    const float* ln[2];  // LayerNorm {gamma[256], beta[256]} of 256-column group min(group,1), or null
    int act[2];
    float* Y;
    int ldy, Nout, nstore;
    uint16_t* split_out;    // optional: bf16 hi/lo copy of Y as [unit][2][N][256] for the tcgen05 einsum (Nout == 256)
    int split_unit0, split_N;
    const float* rowdot_w;  // optional: rowdot_out[row] = X'[row,:] . rowdot_w + rowdot_b   (K == 256 only)
    float rowdot_b;
    float* rowdot_out;
};
struct GemmArgs {
    GemmBranch br[2];
    int R, K, pro, cluster;
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// programmatic dependent launch: everything before pdl_wait() may overlap the previous kernel of the stream
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// cluster helpers (row statistics of a 256-wide LayerNorm live in 4 CTAs)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f32(float* local_smem_ptr, uint32_t rank, float v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_smem_ptr)), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

// sum of `v` over the 256 columns of a row that lives in (4 lanes of a quad) x (2 column-warps) x (V_CL CTAs).
// s_red: [2][V_TM] per-CTA scratch, s_cl: [V_CL][V_TM] per-CTA mailbox written by every CTA of the cluster.
__device__ __forceinline__ void row_allreduce(float (&v)[2], float* s_red, float* s_cl, int warp_m, int warp_n, int gq,
                                              int tq, uint32_t crank, float (&out)[2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        v[h] += __shfl_xor_sync(0xffffffffu, v[h], 1);
        v[h] += __shfl_xor_sync(0xffffffffu, v[h], 2);
    }
    if (tq == 0) {
        s_red[warp_n * V_TM + warp_m * 16 + gq] = v[0];
        s_red[warp_n * V_TM + warp_m * 16 + gq + 8] = v[1];
    }
    __syncthreads();
    if (threadIdx.x < V_TM) {
        const float t = s_red[threadIdx.x] + s_red[V_TM + threadIdx.x];
#pragma unroll
        for (uint32_t r = 0; r < V_CL; ++r) st_cluster_f32(&s_cl[crank * V_TM + threadIdx.x], r, t);
    }
    cluster_sync_all();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = warp_m * 16 + gq + 8 * h;
        out[h] = (s_cl[r] + s_cl[V_TM + r]) + (s_cl[2 * V_TM + r] + s_cl[3 * V_TM + r]);
    }
    cluster_sync_all();   // mailbox may be rewritten by the next reduction only after everyone has read it
}

__global__ void __launch_bounds__(V_THREADS, 2) rowgemm_kernel(const GemmArgs args) {
    extern __shared__ __align__(16) uint8_t dsm[];
    float* s_x = reinterpret_cast<float*>(dsm);              // [V_TM][V_LDX]: 4 K-chunks of 64 per row
    float* s_w = reinterpret_cast<float*>(dsm + V_SMEM_X);   // [V_NST][V_TN][V_LDW]
    __shared__ float s_cnt[V_TM];
    __shared__ float s_red[2 * V_TM];
    __shared__ float s_cl[V_CL * V_TM];

    const GemmBranch& g = args.br[blockIdx.z];
    const int tile = blockIdx.y;
    const int nb = tile * V_TN;
    const bool active = nb < g.Nout;     // uniform per cluster: Nout is a multiple of 256 whenever clusters are used
    const int m0 = blockIdx.x * V_TM;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int warp_m = warp >> 1, warp_n = warp & 1;
    const int gq = lane >> 2, tq = lane & 3;
    const int nk = args.K / V_KS;
    const bool streaming = args.K > 256;   // X is then a plain [R][K] matrix streamed with cp.async like W

    auto issue_w = [&](int kc) {   // weight tile [64 n][64 k] of K-chunk kc -> ring slot kc % V_NST
        float* dst = s_w + (kc % V_NST) * (V_TN * V_LDW);
#pragma unroll
        for (int it = 0; it < (V_TN * V_KS / 4) / V_THREADS; ++it) {
            const int idx = tid + it * V_THREADS;
            const int n = idx >> 4, k4 = (idx & 15) * 4;
            const int ng = min(nb + n, g.Nout - 1);
            cp_async16(dst + n * V_LDW + k4, g.W + (size_t)ng * args.K + kc * V_KS + k4);
        }
    };
    auto issue_x = [&](int kc) {   // streaming mode only: X tile [32 m][64 k] -> column block kc % 4
#pragma unroll
        for (int it = 0; it < (V_TM * V_KS / 4) / V_THREADS; ++it) {
            const int idx = tid + it * V_THREADS;
            const int r = idx >> 4, k4 = (idx & 15) * 4;
            const int mg = min(m0 + r, args.R - 1);
            cp_async16(s_x + r * V_LDX + (kc & 3) * V_KS + k4, g.X + (size_t)mg * g.ldx + kc * V_KS + k4);
        }
    };

    if (active) {
        issue_w(0);
        cp_async_commit();
        if (nk > 1) issue_w(1);
        cp_async_commit();
    }
    pdl_wait();                 // activations written by the previous kernel are visible from here on
    pdl_launch_dependents();    // let the next kernel start prefetching its weights
    if (!active) return;

    if (streaming) {
        issue_x(0);
        if (nk > 1) issue_x(1);
        cp_async_commit();
    } else {
        // ---------------- prologue: build the fp32 X' tile [32][256]
#pragma unroll 4
        for (int it = 0; it < (V_TM * 64) / V_THREADS; ++it) {
            const int idx = tid + it * V_THREADS;
            const int r = idx >> 6, k4 = (idx & 63) * 4;
            const int m = m0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < args.R) {
                v = __ldg(reinterpret_cast<const float4*>(g.X + (size_t)m * g.ldx + k4));
                if (args.pro == PRO_ADD) {
                    if (g.X2) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(g.X2 + (size_t)m * g.ldx2 + k4));
                        v.x += t.x, v.y += t.y, v.z += t.z, v.w += t.w;
                    }
                } else if (args.pro == PRO_MUL || args.pro == PRO_MIX) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(g.X2 + (size_t)m * g.ldx2 + k4));
                    v.x *= t.x, v.y *= t.y, v.z *= t.z, v.w *= t.w;
                    if (args.pro == PRO_MIX) {
                        const float4 c = __ldg(reinterpret_cast<const float4*>(g.X3 + (size_t)m * g.ldx3 + k4));
                        const float4 d = __ldg(reinterpret_cast<const float4*>(g.X4 + (size_t)m * g.ldx4 + k4));
                        v.x += c.x * d.x, v.y += c.y * d.y, v.z += c.z * d.z, v.w += c.w * d.w;
                    }
                }
            }
            *reinterpret_cast<float4*>(&s_x[r * V_LDX + k4]) = v;
        }
    }
    if (tid < V_TM) s_cnt[tid] = (g.count && m0 + tid < args.R) ? __ldg(g.count + m0 + tid) : 0.f;

    // three independent accumulator sets (hi*hi, lo*hi, hi*lo): 12 independent mma chains per warp hide the
    // mma.sync latency; they are summed once at the end (small terms first).
    float acc[4][4], acc_lh[4][4], acc_hl[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = acc_lh[i][jj] = acc_hl[i][jj] = 0.f;

    // ---------------- main loop: 3xTF32 (hi*hi + lo*hi + hi*lo), operands split at fragment-load time.
    // physical k = 16*k16 + 4*tq + {0,1 | 2,3} feeds the logical mma slots (tq, tq+4) of two k8 steps.
    for (int kc = 0; kc < nk; ++kc) {
        if (kc == 0) cp_async_wait<0>(); else cp_async_wait<1>();
        __syncthreads();
        if (kc + 2 < nk) {
            issue_w(kc + 2);
            if (streaming) issue_x(kc + 2);
        }
        cp_async_commit();
        const float* xw = s_x + (warp_m * 16 + gq) * V_LDX + (kc & 3) * V_KS + 4 * tq;
        const float* ww = s_w + (kc % V_NST) * (V_TN * V_LDW) + (warp_n * 32 + gq) * V_LDW + 4 * tq;
#pragma unroll
        for (int k16 = 0; k16 < V_KS / 16; ++k16) {
            const float4 a0 = *reinterpret_cast<const float4*>(xw + k16 * 16);
            const float4 a1 = *reinterpret_cast<const float4*>(xw + 8 * V_LDX + k16 * 16);
            uint32_t Ah_a[4], Al_a[4], Ah_b[4], Al_b[4];
            split_tf32(a0.x, Ah_a[0], Al_a[0]), split_tf32(a1.x, Ah_a[1], Al_a[1]);
            split_tf32(a0.y, Ah_a[2], Al_a[2]), split_tf32(a1.y, Ah_a[3], Al_a[3]);
            split_tf32(a0.z, Ah_b[0], Al_b[0]), split_tf32(a1.z, Ah_b[1], Al_b[1]);
            split_tf32(a0.w, Ah_b[2], Al_b[2]), split_tf32(a1.w, Ah_b[3], Al_b[3]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float4 w = *reinterpret_cast<const float4*>(ww + nt * 8 * V_LDW + k16 * 16);
                uint32_t bh[4], bl[4];
                split_tf32(w.x, bh[0], bl[0]), split_tf32(w.y, bh[1], bl[1]);
                split_tf32(w.z, bh[2], bl[2]), split_tf32(w.w, bh[3], bl[3]);
                mma_tf32(acc_lh[nt], Al_a, bh[0], bh[1]);
                mma_tf32(acc_hl[nt], Ah_a, bl[0], bl[1]);
                mma_tf32(acc[nt], Ah_a, bh[0], bh[1]);
                mma_tf32(acc_lh[nt], Al_b, bh[2], bh[3]);
                mma_tf32(acc_hl[nt], Ah_b, bl[2], bl[3]);
                mma_tf32(acc[nt], Ah_b, bh[2], bh[3]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] += acc_lh[i][jj] + acc_hl[i][jj];

    // ---------------- optional per-row dot of the input tile (the folded logit bias of the dynamic kernels)
    if (g.rowdot_out && tile == 0) {
        for (int r = warp; r < V_TM; r += V_THREADS / 32) {
            float sacc = 0.f;
            for (int k = lane; k < 256; k += 32) sacc += s_x[r * V_LDX + k] * __ldg(g.rowdot_w + k);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            if (lane == 0 && m0 + r < args.R) g.rowdot_out[m0 + r] = sacc + g.rowdot_b;
        }
    }

    // ---------------- epilogue in registers: thread owns rows (rA, rA+8), columns cb + nt*8 + 2*tq + {0,1}
    const int grp = tile / V_CL;                 // 256-column group
    const int gi = grp < 1 ? grp : 1;
    const float* ln = g.ln[gi];
    const int act = g.act[gi];
    const int rl[2] = {warp_m * 16 + gq, warp_m * 16 + gq + 8};
    const int cb = nb + warp_n * 32 + 2 * tq;
    float y[2][8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int m = m0 + rl[h];
        const bool mok = m < args.R;
        const float cnt = s_cnt[rl[h]];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = cb + nt * 8 + e;
                float v = acc[nt][2 * h + e];
                if (c < g.Nout) {
                    if (g.bias) v += __ldg(g.bias + c);
                    if (g.cbias) v += cnt * __ldg(g.cbias + c);
                    if (g.res && mok) v += __ldg(g.res + (size_t)m * g.ldr + c);
                }
                y[h][nt * 2 + e] = v;
            }
    }
    if (ln) {   // launched as a 4-CTA cluster covering the 256-wide row
        const uint32_t crank = cluster_ctarank();
        float part[2], tot[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            part[h] = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) part[h] += y[h][i];
        }
        row_allreduce(part, s_red, s_cl, warp_m, warp_n, gq, tq, crank, tot);
        const float mean[2] = {tot[0] * (1.f / 256.f), tot[1] * (1.f / 256.f)};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            part[h] = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) part[h] += (y[h][i] - mean[h]) * (y[h][i] - mean[h]);
        }
        row_allreduce(part, s_red, s_cl, warp_m, warp_n, gq, tq, crank, tot);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float rstd = 1.f / sqrtf(tot[h] * (1.f / 256.f) + U_LN_EPS);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = (cb + nt * 8 + e) & 255;   // column inside the 256-wide group
                    y[h][nt * 2 + e] = (y[h][nt * 2 + e] - mean[h]) * rstd * __ldg(ln + c) + __ldg(ln + 256 + c);
                }
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int m = m0 + rl[h];
        if (m >= args.R) continue;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            float v0 = y[h][nt * 2], v1 = y[h][nt * 2 + 1];
            if (act == ACT_RELU) v0 = fmaxf(v0, 0.f), v1 = fmaxf(v1, 0.f);
            else if (act == ACT_SIGMOID) v0 = 1.f / (1.f + expf(-v0)), v1 = 1.f / (1.f + expf(-v1));
            const int c = cb + nt * 8;
            if (g.split_out) {
                const int unit = g.split_unit0 + m / g.split_N, n = m % g.split_N;
                const float h0 = bf16_round(v0), h1 = bf16_round(v1);
                uint32_t* hi = reinterpret_cast<uint32_t*>(g.split_out + ((size_t)(unit * 2) * g.split_N + n) * 256 + c);
                hi[0] = pack_bf16x2(h0, h1);
                hi[(size_t)g.split_N * 128] = pack_bf16x2(v0 - h0, v1 - h1);   // lo plane: + N*256 bf16 = N*128 u32
            }
            if (!g.Y) continue;
            float* dst = g.Y + (size_t)m * g.ldy + c;
            if (c + 1 < g.nstore && (g.ldy & 1) == 0) {
                *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
            } else {
                if (c < g.nstore) dst[0] = v0;
                if (c + 1 < g.nstore) dst[1] = v1;
            }
        }
    }
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
struct UpdateScratch {
    float *params, *inp, *gate, *obj0, *qkv, *att, *obj1, *hid, *head;
};
static size_t scratch_floats_per_row(int ffn) { return 512 * 3 + 256 * 3 + 768 + (size_t)ffn + 512; }

// every K2 kernel is launched with programmatic stream serialisation (PDL): its weight prefetch overlaps the tail of
// the previous kernel; 256-wide LayerNorm groups are 4-CTA clusters along the column tiles.
static int launch_gemm(GemmArgs a, int max_nout, int nbranch, cudaStream_t st) {
    const int ytiles = (max_nout + V_TN - 1) / V_TN;
    a.cluster = (ytiles % V_CL == 0) ? V_CL : 1;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(rowgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V_SMEM);
        if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "rowgemm smem attribute: %s", cudaGetErrorString(e));
        attr_done = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((a.R + V_TM - 1) / V_TM, ytiles, nbranch);
    cfg.blockDim = dim3(V_THREADS);
    cfg.dynamicSmemBytes = V_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    attrs[1].id = cudaLaunchAttributeClusterDimension;
    attrs[1].val.clusterDim.x = 1, attrs[1].val.clusterDim.y = a.cluster, attrs[1].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, rowgemm_kernel, a);
    if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "rowgemm_kernel launch: %s", cudaGetErrorString(e));
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

static GemmBranch blank() {
    GemmBranch g;
    memset(&g, 0, sizeof(g));
    return g;
}

}  // namespace pf

extern "C" size_t pf_update_workspace_bytes(int B, int N, int ffn_channels) {
    if (B <= 0 || N <= 0 || ffn_channels <= 0) return 0;
    return (2 * (size_t)B * N * (pf::scratch_floats_per_row(ffn_channels) + 256) + (size_t)B * N) * sizeof(float) + 256;
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
    PF_REQUIRE(ffn > 0 && ffn % 256 == 0, PF_ERR_ARG, "pf_kernel_update: ffn_channels=%d must be a multiple of 256", ffn);
    PF_REQUIRE(w->num_classes > 0 && w->num_classes <= PF_MAX_CLASSES, PF_ERR_ARG, "pf_kernel_update: num_classes=%d", w->num_classes);
    PF_REQUIRE(workspace_bytes >= pf_update_workspace_bytes(B, N, ffn), PF_ERR_WORKSPACE,
               "pf_kernel_update: workspace %zu < %zu", workspace_bytes, pf_update_workspace_bytes(B, N, ffn));
    PF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, PF_ERR_ALIGN, "pf_kernel_update: workspace not 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int R = B * N;

    UpdateScratch sc[2];
    {
        float* p = static_cast<float*>(workspace);
        for (int b = 0; b < 2; ++b) {
            sc[b].params = p, p += (size_t)R * 512;
            sc[b].inp = p, p += (size_t)R * 512;
            sc[b].gate = p, p += (size_t)R * 512;
            sc[b].obj0 = p, p += (size_t)R * 256;
            sc[b].qkv = p, p += (size_t)R * 768;
            sc[b].att = p, p += (size_t)R * 256;
            sc[b].obj1 = p, p += (size_t)R * 256;
            sc[b].hid = p, p += (size_t)R * ffn;
            sc[b].head = p, p += (size_t)R * 512;
        }
    }
    const float* in_[2] = {obj_in, dep_in};
    float* out_[2] = {obj_out, dep_out};

    float* pooled = static_cast<float*>(workspace) + 2 * (size_t)R * scratch_floats_per_row(ffn);   // [2][R][256]
    float* count = pooled + 2 * (size_t)R * 256;                                                      // [R]
    // 0. deterministic sum of the split-K pooling partials (fixed order)
    if (int e = pf_pool_reduce(partial, cntp, pooled, count, B, N, 2, S, stream)) return e;

    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.R = R;

    // 1. parameters = dynamic_layer(pooled W_t^T + count b_t); param_out -> norm_out     (kernel_updator.py:58-62,78)
    a.K = 256, a.pro = PRO_PLAIN;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        GemmBranch g = blank();
        g.X = pooled + (size_t)b * R * 256, g.ldx = 256, g.count = count;
        g.W = bw.dyn_w, g.bias = bw.dyn_b, g.cbias = bw.dyn_cb;
        g.ln[1] = bw.ln_norm_out;
        g.Y = sc[b].params, g.ldy = 512, g.Nout = 512, g.nstore = 512;
        a.br[b] = g;
    }
    if (int e = launch_gemm(a, 512, 2, st)) return e;

    // 2. input_feats = input_layer(kernel); depth kernel += mask kernel (kernel_update_head.py:250); input_out -> input_norm_out
    a.pro = PRO_ADD;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        GemmBranch g = blank();
        g.X = in_[b], g.ldx = 256;
        if (b == 1) g.X2 = obj_in, g.ldx2 = 256;
        g.W = bw.inp_w, g.bias = bw.inp_b;
        g.ln[1] = bw.ln_input_norm_out;
        g.Y = sc[b].inp, g.ldy = 512, g.Nout = 512, g.nstore = 512;
        a.br[b] = g;
    }
    if (int e = launch_gemm(a, 512, 2, st)) return e;

    // 3. gate_feats = input_in * param_in; [input_gate | update_gate] = sigmoid(LN(W g + b))   (:69,73-77)
    a.pro = PRO_MUL;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        GemmBranch g = blank();
        g.X = sc[b].inp, g.ldx = 512, g.X2 = sc[b].params, g.ldx2 = 512;
        g.W = bw.gate_w, g.bias = bw.gate_b;
        g.ln[0] = bw.ln_input_norm_in, g.ln[1] = bw.ln_norm_in;
        g.act[0] = g.act[1] = ACT_SIGMOID;
        g.Y = sc[b].gate, g.ldy = 512, g.Nout = 512, g.nstore = 512;
        a.br[b] = g;
    }
    if (int e = launch_gemm(a, 512, 2, st)) return e;

    // 4. features = update_gate * param_out + input_gate * input_out; relu(fc_norm(fc_layer(.)))   (:86-91)
    a.pro = PRO_MIX;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        GemmBranch g = blank();
        g.X = sc[b].gate + 256, g.ldx = 512, g.X2 = sc[b].params + 256, g.ldx2 = 512;
        g.X3 = sc[b].gate, g.ldx3 = 512, g.X4 = sc[b].inp + 256, g.ldx4 = 512;
        g.W = bw.fc_w, g.bias = bw.fc_b, g.ln[0] = bw.ln_fc_norm, g.act[0] = ACT_RELU;
        g.Y = sc[b].obj0, g.ldy = 256, g.Nout = 256, g.nstore = 256;
        a.br[b] = g;
    }
    if (int e = launch_gemm(a, 256, 2, st)) return e;

    // 5. attention in-projection
    a.pro = PRO_PLAIN;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        GemmBranch g = blank();
        g.X = sc[b].obj0, g.ldx = 256, g.W = bw.qkv_w, g.bias = bw.qkv_b;
        g.Y = sc[b].qkv, g.ldy = 768, g.Nout = 768, g.nstore = 768;
        a.br[b] = g;
    }
    if (int e = launch_gemm(a, 768, 2, st)) return e;

    // 6. softmax(q k^T) v per (branch, image, head)
    if (int e = launch_attention(sc[0].qkv, sc[1].qkv, sc[0].att, sc[1].att, B, N, 2, st)) return e;

    // 7. attention_norm(x + out_proj(attn))
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        GemmBranch g = blank();
        g.X = sc[b].att, g.ldx = 256, g.W = bw.out_w, g.bias = bw.out_b;
        g.res = sc[b].obj0, g.ldr = 256, g.ln[0] = bw.ln_attn;
        g.Y = sc[b].obj1, g.ldy = 256, g.Nout = 256, g.nstore = 256;
        a.br[b] = g;
    }
    if (int e = launch_gemm(a, 256, 2, st)) return e;

    // 8. FFN layer 1 + ReLU
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        GemmBranch g = blank();
        g.X = sc[b].obj1, g.ldx = 256, g.W = bw.ffn1_w, g.bias = bw.ffn1_b;
        g.act[0] = g.act[1] = ACT_RELU;
        g.Y = sc[b].hid, g.ldy = ffn, g.Nout = ffn, g.nstore = ffn;
        a.br[b] = g;
    }
    if (int e = launch_gemm(a, ffn, 2, st)) return e;

    // 9. ffn_norm(x + FFN layer 2) -> obj_feat / depth_feat_new (returned to the caller)
    a.K = ffn;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        GemmBranch g = blank();
        g.X = sc[b].hid, g.ldx = ffn, g.W = bw.ffn2_w, g.bias = bw.ffn2_b;
        g.res = sc[b].obj1, g.ldr = 256, g.ln[0] = bw.ln_ffn;
        g.Y = out_[b], g.ldy = 256, g.Nout = 256, g.nstore = 256;
        a.br[b] = g;
    }
    if (int e = launch_gemm(a, 256, 2, st)) return e;
    a.K = 256;

    // 10. cls_fcs / mask_fcs / depth_regs: Linear(no bias) + LN (+ ReLU except depth_regs)
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        GemmBranch g = blank();
        g.X = out_[b], g.ldx = 256, g.W = bw.head_w;
        g.ln[0] = bw.ln_head_a, g.ln[1] = bw.ln_head_b;
        g.act[0] = g.act[1] = bw.head_relu ? ACT_RELU : ACT_NONE;
        g.Nout = (b == 0) ? 512 : 256;
        g.Y = sc[b].head, g.ldy = 512, g.nstore = g.Nout;
        a.br[b] = g;
    }
    if (int e = launch_gemm(a, 512, 2, st)) return e;

    // 11. fc_mask / fc_depth with feat_transform folded in -> dynamic kernels + their logit bias
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        GemmBranch g = blank();
        g.X = sc[b].head + (b == 0 ? 256 : 0), g.ldx = 512, g.W = bw.kern_w, g.bias = bw.kern_b;
        g.Y = kern ? kern + (size_t)b * R * 256 : nullptr, g.ldy = 256, g.Nout = 256, g.nstore = 256;
        g.split_out = kern_split, g.split_unit0 = b * B, g.split_N = N;
        g.rowdot_w = bw.kb_w, g.rowdot_b = bw.kb_b, g.rowdot_out = kbias + (size_t)b * R;
        a.br[b] = g;
    }
    if (int e = launch_gemm(a, 256, 2, st)) return e;

    // 12. fc_cls (mask branch only)
    if (cls_out) {
        const pf_branch_weights& bw = w->br[0];
        GemmBranch g = blank();
        g.X = sc[0].head, g.ldx = 512, g.W = bw.cls_w, g.bias = bw.cls_b;
        g.act[0] = g.act[1] = cls_sigmoid ? ACT_SIGMOID : ACT_NONE;
        g.Y = cls_out, g.ldy = w->num_classes, g.Nout = PF_MAX_CLASSES, g.nstore = w->num_classes;
        a.br[0] = g;
        if (int e = launch_gemm(a, PF_MAX_CLASSES, 1, st)) return e;
    }
    return PF_OK;
}

extern "C" size_t pf_updator_workspace_bytes(int R) { return R > 0 ? (size_t)R * 1536 * sizeof(float) : 0; }

// KernelUpdator.forward on its own (polyphonic/funcs/kernel_updator.py:55-93): no feat_transform fold, one branch.
extern "C" int pf_kernel_updator(const pf_branch_weights* bw, const float* update_feature, const float* input_feature,
                                 float* out, void* workspace, size_t workspace_bytes, int R, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(bw && update_feature && input_feature && out && workspace, PF_ERR_ARG, "pf_kernel_updator: null pointer");
    PF_REQUIRE(R > 0, PF_ERR_ARG, "pf_kernel_updator: R=%d", R);
    PF_REQUIRE(workspace_bytes >= pf_updator_workspace_bytes(R), PF_ERR_WORKSPACE, "pf_kernel_updator: workspace %zu < %zu",
               workspace_bytes, pf_updator_workspace_bytes(R));
    PF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && (reinterpret_cast<uintptr_t>(update_feature) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(input_feature) & 15) == 0,
               PF_ERR_ALIGN, "pf_kernel_updator: pointers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* params = static_cast<float*>(workspace);
    float* inp = params + (size_t)R * 512;
    float* gate = inp + (size_t)R * 512;
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.R = R, a.K = 256;
    GemmBranch g;

    a.pro = PRO_PLAIN;
    g = blank();
    g.X = update_feature, g.ldx = 256, g.W = bw->dyn_w, g.bias = bw->dyn_b, g.ln[1] = bw->ln_norm_out;
    g.Y = params, g.ldy = 512, g.Nout = 512, g.nstore = 512;
    a.br[0] = g;
    if (int e = launch_gemm(a, 512, 1, st)) return e;

    g = blank();
    g.X = input_feature, g.ldx = 256, g.W = bw->inp_w, g.bias = bw->inp_b, g.ln[1] = bw->ln_input_norm_out;
    g.Y = inp, g.ldy = 512, g.Nout = 512, g.nstore = 512;
    a.br[0] = g;
    if (int e = launch_gemm(a, 512, 1, st)) return e;

    a.pro = PRO_MUL;
    g = blank();
    g.X = inp, g.ldx = 512, g.X2 = params, g.ldx2 = 512, g.W = bw->gate_w, g.bias = bw->gate_b;
    g.ln[0] = bw->ln_input_norm_in, g.ln[1] = bw->ln_norm_in, g.act[0] = g.act[1] = ACT_SIGMOID;
    g.Y = gate, g.ldy = 512, g.Nout = 512, g.nstore = 512;
    a.br[0] = g;
    if (int e = launch_gemm(a, 512, 1, st)) return e;

    a.pro = PRO_MIX;
    g = blank();
    g.X = gate + 256, g.ldx = 512, g.X2 = params + 256, g.ldx2 = 512, g.X3 = gate, g.ldx3 = 512, g.X4 = inp + 256, g.ldx4 = 512;
    g.W = bw->fc_w, g.bias = bw->fc_b, g.ln[0] = bw->ln_fc_norm, g.act[0] = ACT_RELU;
    g.Y = out, g.ldy = 256, g.Nout = 256, g.nstore = 256;
    a.br[0] = g;
    return launch_gemm(a, 256, 1, st);
}
