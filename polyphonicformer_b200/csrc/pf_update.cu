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

constexpr int U_TM = 16;            // rows per CTA
constexpr int U_TN = 256;           // columns per CTA
constexpr int U_KC = 256;           // K chunk staged in shared memory
constexpr int U_LDS = U_KC + 16;    // smem row stride (floats): conflict-free float4 fragment loads
constexpr int U_LDY = U_TN + 8;
constexpr float U_LN_EPS = 1e-5f;   // nn.LayerNorm default (mmcv build_norm_layer(dict(type='LN')))

enum { PRO_PLAIN = 0, PRO_ADD = 1, PRO_MUL = 2, PRO_MIX = 3, PRO_POOLSUM = 4 };
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2 };

struct GemmBranch {
    const float *X, *X2, *X3, *X4;
    int ldx, ldx2, ldx3, ldx4;
    const float* W;      // [Nout][K]
    const float* bias;   // [Nout] or null
    const float* cbias;  // [Nout] or null (PRO_POOLSUM: + count[row] * cbias)
    const float* res;    // residual [R][ldr] or null
    int ldr;
    const float* ln[2];  // LayerNorm {gamma[256], beta[256]} of column tile min(tile,1), or null
    int act[2];
    float* Y;
    int ldy, Nout, nstore;
    const float* rowdot_w;  // optional: rowdot_out[row] = X'[row,:] . rowdot_w + rowdot_b   (K == 256 only)
    float rowdot_b;
    float* rowdot_out;
    const float* partial;  // PRO_POOLSUM
    const float* cntp;
    int unit0;
};
struct GemmArgs {
    GemmBranch br[2];
    int R, K, pro, B, N, S;
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256) rowgemm_kernel(const GemmArgs args) {
    __shared__ __align__(16) float s_hi[U_TM * U_LDS];
    __shared__ __align__(16) float s_lo[U_TM * U_LDS];
    __shared__ float s_cnt[U_TM];
    float* s_y = s_hi;  // [16][U_LDY] aliased after the K loop

    const GemmBranch& g = args.br[blockIdx.z];
    const int tile = blockIdx.y;
    const int nb = tile * U_TN;
    if (nb >= g.Nout) return;
    const int m0 = blockIdx.x * U_TM;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.f;

    const int nchunks = args.K / U_KC;
    for (int kc = 0; kc < nchunks; ++kc) {
        if (kc > 0) __syncthreads();
        // ---------------- prologue: build X' tile [16][256] -> tf32 hi / lo
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int idx = tid + it * 256;
            const int r = idx >> 6, k4 = (idx & 63) * 4;
            const int m = m0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < args.R) {
                const int kk = kc * U_KC + k4;
                if (args.pro == PRO_POOLSUM) {
                    const int b = m / args.N, n = m % args.N;
                    const float* pp = g.partial + (((size_t)(g.unit0 + b) * args.S) * args.N + n) * PF_C + kk;
                    for (int s = 0; s < args.S; ++s) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(pp + (size_t)s * args.N * PF_C));
                        v.x += t.x, v.y += t.y, v.z += t.z, v.w += t.w;
                    }
                } else {
                    v = __ldg(reinterpret_cast<const float4*>(g.X + (size_t)m * g.ldx + kk));
                    if (args.pro == PRO_ADD) {
                        if (g.X2) {
                            const float4 t = __ldg(reinterpret_cast<const float4*>(g.X2 + (size_t)m * g.ldx2 + kk));
                            v.x += t.x, v.y += t.y, v.z += t.z, v.w += t.w;
                        }
                    } else if (args.pro == PRO_MUL || args.pro == PRO_MIX) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(g.X2 + (size_t)m * g.ldx2 + kk));
                        v.x *= t.x, v.y *= t.y, v.z *= t.z, v.w *= t.w;
                        if (args.pro == PRO_MIX) {
                            const float4 c = __ldg(reinterpret_cast<const float4*>(g.X3 + (size_t)m * g.ldx3 + kk));
                            const float4 d = __ldg(reinterpret_cast<const float4*>(g.X4 + (size_t)m * g.ldx4 + kk));
                            v.x += c.x * d.x, v.y += c.y * d.y, v.z += c.z * d.z, v.w += c.w * d.w;
                        }
                    }
                }
            }
            float4 h, l;
            h.x = __uint_as_float(to_tf32(v.x)), h.y = __uint_as_float(to_tf32(v.y));
            h.z = __uint_as_float(to_tf32(v.z)), h.w = __uint_as_float(to_tf32(v.w));
            l.x = __uint_as_float(to_tf32(v.x - h.x)), l.y = __uint_as_float(to_tf32(v.y - h.y));
            l.z = __uint_as_float(to_tf32(v.z - h.z)), l.w = __uint_as_float(to_tf32(v.w - h.w));
            *reinterpret_cast<float4*>(&s_hi[r * U_LDS + k4]) = h;
            *reinterpret_cast<float4*>(&s_lo[r * U_LDS + k4]) = l;
        }
        if (kc == 0 && args.pro == PRO_POOLSUM && tid < U_TM) {
            const int m = m0 + tid;
            float c = 0.f;
            if (m < args.R) {
                const int b = m / args.N, n = m % args.N;
                for (int s = 0; s < args.S; ++s) c += g.cntp[((size_t)(g.unit0 + b) * args.S + s) * args.N + n];
            }
            s_cnt[tid] = c;
        }
        __syncthreads();

        // ---------------- 3xTF32 product; physical k = 16*k16 + 4*tq + {0,1 | 2,3} feeds logical slots (tq, tq+4)
        const float* wbase = g.W + (size_t)kc * U_KC + 4 * tq;
        float4 wv[4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int n = nb + warp * 32 + nt * 8 + gq;
            wv[nt] = (n < g.Nout) ? __ldg(reinterpret_cast<const float4*>(wbase + (size_t)n * args.K))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll 2
        for (int k16 = 0; k16 < U_KC / 16; ++k16) {
            float4 wn[4];
            if (k16 + 1 < U_KC / 16) {
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int n = nb + warp * 32 + nt * 8 + gq;
                    wn[nt] = (n < g.Nout)
                                 ? __ldg(reinterpret_cast<const float4*>(wbase + (size_t)n * args.K + (k16 + 1) * 16))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            const float4 ah0 = *reinterpret_cast<const float4*>(&s_hi[gq * U_LDS + k16 * 16 + 4 * tq]);
            const float4 ah1 = *reinterpret_cast<const float4*>(&s_hi[(gq + 8) * U_LDS + k16 * 16 + 4 * tq]);
            const float4 al0 = *reinterpret_cast<const float4*>(&s_lo[gq * U_LDS + k16 * 16 + 4 * tq]);
            const float4 al1 = *reinterpret_cast<const float4*>(&s_lo[(gq + 8) * U_LDS + k16 * 16 + 4 * tq]);
            const uint32_t Ah_a[4] = {__float_as_uint(ah0.x), __float_as_uint(ah1.x), __float_as_uint(ah0.y),
                                      __float_as_uint(ah1.y)};
            const uint32_t Ah_b[4] = {__float_as_uint(ah0.z), __float_as_uint(ah1.z), __float_as_uint(ah0.w),
                                      __float_as_uint(ah1.w)};
            const uint32_t Al_a[4] = {__float_as_uint(al0.x), __float_as_uint(al1.x), __float_as_uint(al0.y),
                                      __float_as_uint(al1.y)};
            const uint32_t Al_b[4] = {__float_as_uint(al0.z), __float_as_uint(al1.z), __float_as_uint(al0.w),
                                      __float_as_uint(al1.w)};
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float w4[4] = {wv[nt].x, wv[nt].y, wv[nt].z, wv[nt].w};
                uint32_t bh[4], bl[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    bh[e] = to_tf32(w4[e]);
                    bl[e] = to_tf32(w4[e] - __uint_as_float(bh[e]));
                }
                mma_tf32(acc[nt], Al_a, bh[0], bh[1]);
                mma_tf32(acc[nt], Ah_a, bl[0], bl[1]);
                mma_tf32(acc[nt], Ah_a, bh[0], bh[1]);
                mma_tf32(acc[nt], Al_b, bh[2], bh[3]);
                mma_tf32(acc[nt], Ah_b, bl[2], bl[3]);
                mma_tf32(acc[nt], Ah_b, bh[2], bh[3]);
            }
            if (k16 + 1 < U_KC / 16) {
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) wv[nt] = wn[nt];
            }
        }
    }
    __syncthreads();

    // ---------------- optional per-row dot of the input tile (the folded logit bias of the dynamic kernels)
    if (g.rowdot_out && tile == 0) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int r = warp * 2 + rr;
            float s = 0.f;
            for (int k = lane; k < U_KC; k += 32) s += (s_hi[r * U_LDS + k] + s_lo[r * U_LDS + k]) * __ldg(g.rowdot_w + k);
            s = warp_sum(s);
            if (lane == 0 && m0 + r < args.R) g.rowdot_out[m0 + r] = s + g.rowdot_b;
        }
        __syncthreads();
    }

    // ---------------- accumulators -> smem tile
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int c = warp * 32 + nt * 8 + 2 * tq;
        s_y[gq * U_LDY + c] = acc[nt][0];
        s_y[gq * U_LDY + c + 1] = acc[nt][1];
        s_y[(gq + 8) * U_LDY + c] = acc[nt][2];
        s_y[(gq + 8) * U_LDY + c + 1] = acc[nt][3];
    }
    __syncthreads();

    // ---------------- epilogue: warp w owns rows 2w, 2w+1; lane owns columns lane + 32 j
    const int ti = tile < 1 ? tile : 1;
    const float* ln = g.ln[ti];
    const int act = g.act[ti];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const int r = warp * 2 + rr;
        const int m = m0 + r;
        if (m >= args.R) continue;  // warp-uniform
        float y[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int c = lane + 32 * jj;
            const int n = nb + c;
            float v = 0.f;
            if (n < g.Nout) {
                v = s_y[r * U_LDY + c];
                if (g.bias) v += __ldg(g.bias + n);
                if (g.cbias) v += s_cnt[r] * __ldg(g.cbias + n);
                if (g.res) v += __ldg(g.res + (size_t)m * g.ldr + n);
            }
            y[jj] = v;
        }
        if (ln) {  // tile is a full 256-wide row by construction
            float s = 0.f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) s += y[jj];
            const float mean = warp_sum(s) * (1.f / U_TN);
            float q = 0.f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) q += (y[jj] - mean) * (y[jj] - mean);
            const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / U_TN) + U_LN_EPS);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int c = lane + 32 * jj;
                y[jj] = (y[jj] - mean) * rstd * __ldg(ln + c) + __ldg(ln + U_TN + c);
            }
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int n = nb + lane + 32 * jj;
            float v = y[jj];
            if (act == ACT_RELU) v = fmaxf(v, 0.f);
            else if (act == ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
            if (n < g.nstore) g.Y[(size_t)m * g.ldy + n] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Inter-kernel self-attention of one (branch, image, head): softmax(q k^T / sqrt(32)) v over the N kernels of the
// image (mmcv MultiheadAttention -> nn.MultiheadAttention, seq-first; kernel_update_head.py:259-260).
// qkv [R][768] = [q | k | v]; out [R][256] (heads concatenated), before out_proj.
__global__ void __launch_bounds__(128) attention_kernel(const float* __restrict__ qkv0, const float* __restrict__ qkv1,
                                                        float* __restrict__ out0, float* __restrict__ out1, int N) {
    __shared__ float s_k[PF_MAX_N][33];
    __shared__ float s_v[PF_MAX_N][33];
    const int h = blockIdx.x, b = blockIdx.y;
    const float* qkv = (blockIdx.z == 0 ? qkv0 : qkv1) + (size_t)b * N * 768;
    float* out = (blockIdx.z == 0 ? out0 : out1) + (size_t)b * N * 256;
    for (int i = threadIdx.x; i < N * 32; i += 128) {
        const int n = i >> 5, d = i & 31;
        s_k[n][d] = qkv[(size_t)n * 768 + 256 + h * 32 + d];
        s_v[n][d] = qkv[(size_t)n * 768 + 512 + h * 32 + d];
    }
    __syncthreads();
    const int n = threadIdx.x;
    if (n >= N) return;
    float q[32], o[32];
    const float scale = 0.17677669529663687f;  // 1/sqrt(32), applied to q before q k^T as torch does
#pragma unroll
    for (int d = 0; d < 32; ++d) {
        q[d] = qkv[(size_t)n * 768 + h * 32 + d] * scale;
        o[d] = 0.f;
    }
    float mx = -INFINITY, den = 0.f;
    for (int j = 0; j < N; ++j) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < 32; ++d) s += q[d] * s_k[j][d];
        const float mn = fmaxf(mx, s);
        const float corr = expf(mx - mn);
        const float pj = expf(s - mn);
        den = den * corr + pj;
#pragma unroll
        for (int d = 0; d < 32; ++d) o[d] = o[d] * corr + pj * s_v[j][d];
        mx = mn;
    }
    const float inv = 1.f / den;
#pragma unroll
    for (int d = 0; d < 32; ++d) out[(size_t)n * 256 + h * 32 + d] = o[d] * inv;
}

// ---------------------------------------------------------------------------------------------------------------
struct UpdateScratch {
    float *params, *inp, *gate, *obj0, *qkv, *att, *obj1, *hid, *head;
};
static size_t scratch_floats_per_row(int ffn) { return 512 * 3 + 256 * 3 + 768 + (size_t)ffn + 512; }

static int launch_gemm(const GemmArgs& a, int max_nout, int nbranch, cudaStream_t st) {
    dim3 grid((a.R + U_TM - 1) / U_TM, (max_nout + U_TN - 1) / U_TN, nbranch);
    rowgemm_kernel<<<grid, 256, 0, st>>>(a);
    PF_CHECK_LAUNCH("rowgemm_kernel");
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
    return 2 * (size_t)B * N * pf::scratch_floats_per_row(ffn_channels) * sizeof(float) + 256;
}

extern "C" int pf_kernel_update(const pf_stage_weights* w, const float* partial, const float* cntp, int S,
                                const float* obj_in, const float* dep_in, float* obj_out, float* dep_out,
                                float* cls_out, float* kern, float* kbias, void* workspace, size_t workspace_bytes,
                                int B, int N, int cls_sigmoid, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(w && partial && cntp && obj_in && dep_in && obj_out && dep_out && kern && kbias && workspace, PF_ERR_ARG,
               "pf_kernel_update: null pointer");
    PF_REQUIRE(B > 0 && N > 0 && N <= PF_MAX_N && S > 0, PF_ERR_ARG, "pf_kernel_update: bad shape B=%d N=%d S=%d", B, N, S);
    const int ffn = w->ffn_channels;
    PF_REQUIRE(ffn > 0 && ffn % U_KC == 0, PF_ERR_ARG, "pf_kernel_update: ffn_channels=%d must be a multiple of 256", ffn);
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

    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.R = R, a.B = B, a.N = N, a.S = S;

    // 1. parameters = dynamic_layer(pooled W_t^T + count b_t); param_out -> norm_out     (kernel_updator.py:58-62,78)
    a.K = 256, a.pro = PRO_POOLSUM;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        GemmBranch g = blank();
        g.partial = partial, g.cntp = cntp, g.unit0 = b * B;
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
    attention_kernel<<<dim3(PF_HEADS, B, 2), 128, 0, st>>>(sc[0].qkv, sc[1].qkv, sc[0].att, sc[1].att, N);
    PF_CHECK_LAUNCH("attention_kernel");

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
        g.Y = kern + (size_t)b * R * 256, g.ldy = 256, g.Nout = 256, g.nstore = 256;
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
    a.R = R, a.K = 256, a.B = 1, a.N = R, a.S = 1;
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
