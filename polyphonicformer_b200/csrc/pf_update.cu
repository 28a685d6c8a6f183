// K2 -- the small-N block of one decoder stage, both branches (mask / depth) per launch:
//   KernelUpdator (polyphonic/funcs/kernel_updator.py:55-93), inter-kernel multi-head attention + LN, FFN + LN and
//   the cls / mask / depth FC heads (polyphonic/kernel_update_head.py:245-288), with feat_transform folded into the
//   first and last linear layers (see include/pf_decoder.h).
//
// Data layout.  Every activation that feeds a GEMM lives in the "arena" as a pair of bf16 planes
//   arena[unit][slot][hi|lo][128 rows][256 cols]      (x = hi + lo to ~2^-17; rows >= N are padding)
// so that BOTH GEMM operands are TMA boxes and no kernel has a register-path prologue; values that are only used
// element-wise (LayerNorm'ed halves, residuals, q/k/v, split-K partials) stay fp32 [B*N][...].
//
// tcgemm_kernel<MODE>: Y[128 rows][128 cols] per CTA on tcgen05.
//   * M = 128 = the N (=111) kernels of ONE image;
//   * fp32-level accuracy from bf16 tensor cores: D = Al*Wh + Ah*Wl + Ah*Wh (3 MMAs per K=16 step, fp32 accumulate
//     in TMEM, error ~2^-16 per product) -- the reference runs these layers in fp32 and they feed LayerNorms;
//   * one ring stage = [A hi | A lo | W hi | W lo] boxes of [128][64 k] (64 KB), 3 stages; the weight boxes of the
//     first stages are issued BEFORE griddepcontrol.wait (programmatic dependent launch), the activation boxes after;
//   * a CTA runs 1 or 2 "passes" of K = 256 into separate TMEM accumulators (MODE_DUAL: dynamic_layer and
//     input_layer of the same output columns, so their product is formed in the epilogue);
//   * epilogue: 16 warps.  Phase 1, thread = (row, 32-column quarter): bias / count-bias, LayerNorm piece statistics
//     ((mean, M2) over 32 columns, merged across the quarters and the CTAs of the cluster through DSMEM with one
//     cluster barrier), values staged in shared memory.  Phase 2, warp = 8 rows x 128 columns: normalise, ReLU /
//     sigmoid, gate mixing, then fp32 rows and/or bf16 hi/lo planes with fully coalesced global accesses.
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..17 = epilogue.
// 12 launches per stage: prep, dyn+inp, gates, fc, qkv, attention, out-proj, ffn1, ffn2 (split-K), sum+LN, heads,
// kernels+cls.  This block is latency / weight-streaming bound, not roofline bound; see DESIGN.md.
#include <string.h>

#include <mutex>

#include "pf_internal.h"
#include "pf_sm100.cuh"
#include "pf_debug.cuh"
#include "pf_update.cuh"

namespace pf {

template <int MODE>
__global__ void __launch_bounds__(T_THREADS, 1)
tcgemm_kernel(const __grid_constant__ CUtensorMap tmap_w256, const __grid_constant__ CUtensorMap tmap_wffn,
              const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ TcArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T_BAR_OFF);
    uint64_t* full = bars;
    uint64_t* empty = bars + T_NSTG;
    uint64_t* accfull = bars + 2 * T_NSTG;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * T_NSTG + 1);
    Mail mail{reinterpret_cast<float2(*)[128][9]>(smem + T_MAIL_OFF)};
    if (smem + T_SMEM_USED > smem_raw + T_SMEM) __trap();   // dynamic smem base less aligned than assumed

    long long* dbg = nullptr;
    __shared__ long long* s_dbg;
    if (g_dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
        if (threadIdx.x == 0) {
            const unsigned long long slot = atomicAdd(reinterpret_cast<unsigned long long*>(g_dbg), 1ull);
            s_dbg = g_dbg + 16 + slot * 16;
            s_dbg[0] = gtime();
            s_dbg[15] = MODE * 1000 + gridDim.x * 100 + gridDim.z;
        }
        __syncthreads();
        dbg = s_dbg;
    }
    const TcJob& g = args.job[blockIdx.z];
    constexpr int TN = MODE == MODE_GENERIC64 ? 64 : T_TN;   // output columns of this CTA
    const int tile = blockIdx.x, nb = tile * TN;
    const int b = blockIdx.y / args.ksplit, split = blockIdx.y % args.ksplit;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool active = tile < g.ntiles;   // uniform per cluster (ntiles is a multiple of the cluster size)
    const int N = args.N;
    const int unit = g.unit0 + b;
    const int total_it = g.npass * (T_K / T_KC);
    // LayerNorm anywhere in this CTA?  (decides the cluster barriers; uniform per cluster)
    bool has_ln;
    if (MODE == MODE_GATE) has_ln = true;
    else if (MODE == MODE_DUAL) has_ln = nb >= 256;
    else has_ln = (nb < 256 ? g.ln0 : g.ln1) != nullptr;
    has_ln = has_ln && active;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_w256);
        tma_prefetch_desc(&tmap_wffn);
        tma_prefetch_desc(&tmap_a);
        for (int i = 0; i < T_NSTG; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(accfull, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);   // warp-uniform (REDUX -> uniform register): tcgen05 operands need no per-instruction R2UR
    if (threadIdx.x == 0) DBG(1);
    if (has_ln) cluster_arrive();   // matched by the cluster_wait before the first DSMEM store: peers have started
    // per-column parameters of this tile (weights: independent of the previous kernel) -> shared memory, now
    float* vec = reinterpret_cast<float*>(smem + T_VEC_OFF);
    if (active && warp >= 2 && warp < 2 + V_COUNT) {
        const int which = warp - 2;
        const float* src = nullptr;
        int off = nb;   // column offset inside the source vector
        if (MODE == MODE_GENERIC || MODE == MODE_GENERIC64) {
            const float* ln = nb < 256 ? g.ln0 : g.ln1;
            if (which == V_BIAS0) src = g.bias0;
            else if (which == V_GA0 && ln) src = ln, off = nb & 255;
            else if (which == V_BE0 && ln) src = ln + 256, off = nb & 255;
        } else if (MODE == MODE_DUAL) {
            if (which == V_BIAS0) src = g.bias0;
            else if (which == V_CBIAS0) src = g.cbias0;
            else if (which == V_BIAS1) src = g.bias1;
            else if (nb >= 256) {
                off = nb - 256;
                if (which == V_GA0) src = g.ln0;
                else if (which == V_BE0) src = g.ln0 + 256;
                else if (which == V_GA1) src = g.ln1;
                else if (which == V_BE1) src = g.ln1 + 256;
            }
        } else {   // MODE_GATE: tile columns [0,64) -> input_norm_in, [64,128) -> norm_in, both of features 64*tile..
            if (which == V_BIAS0) src = g.bias0;
            else if (which == V_GA0) src = (lane < 16 ? g.ln0 : g.ln1), off = tile * 64 - (lane < 16 ? 0 : 64);
            else if (which == V_BE0) src = (lane < 16 ? g.ln0 : g.ln1) + 256, off = tile * 64 - (lane < 16 ? 0 : 64);
        }
        const float4 v = src ? ld4(src + off + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(vec + which * 128 + lane * 4) = v;
    }

    // (both run by the WHOLE producer warp; one lane elected inside each PTX block issues)
    auto issue_w = [&](int it) {
        const TcPass& p = g.pass[it >> 2];
        const int kb = it & 3, s = it % T_NSTG;
        uint8_t* st = smem + s * T_STAGE;
        const CUtensorMap* wm = p.w_ffn ? &tmap_wffn : &tmap_w256;
        const int k = p.w_k0 + (p.w_ffn ? split * T_K : 0) + kb * T_KC;
        mbar_arrive_expect_tx_warp(&full[s], T_STAGE);
        tma_load_2d_warp(st + 2 * T_PLANE, wm, &full[s], k, p.w_row + nb, kEvictNormal);
        tma_load_2d_warp(st + 3 * T_PLANE, wm, &full[s], k, p.w_row + p.w_lo + nb, kEvictNormal);
    };
    auto issue_a = [&](int it) {
        const TcPass& p = g.pass[it >> 2];
        const int kb = it & 3, s = it % T_NSTG;
        uint8_t* st = smem + s * T_STAGE;
        const int row = ((unit * NSLOT + p.a_slot + (args.ksplit > 1 ? split : 0)) * 2) * 128;
        tma_load_2d_warp(st, &tmap_a, &full[s], kb * T_KC, row, kEvictFirst);
        tma_load_2d_warp(st + T_PLANE, &tmap_a, &full[s], kb * T_KC, row + 128, kEvictFirst);
    };

    // weights do not depend on the previous kernel: start the ring before the grid dependency is resolved
    const int pre = total_it < T_NSTG ? total_it : T_NSTG;
    if (active && warp == 0)
        for (int it = 0; it < pre; ++it) issue_w(it);
    pdl_wait();                 // activations written by the previous kernels are visible from here on
    pdl_launch_dependents();    // let the next kernel start prefetching its weights
    if (threadIdx.x == 0) DBG(2);
    // operands of the epilogue written by previous kernels: issue their loads now, they land while the MMAs run
    float4 pref[8];                     // GENERIC: residual, GATE: LN'ed input_out / param_out (phase-2 mapping)
    float cnt = 0.f;                    // DUAL: mask pixel count of this thread's row (phase-1 mapping)
    if (active && warp >= 2) {
        const int ew_ = warp - 2;
        if (MODE == MODE_GENERIC64) {
            if (g.res) {   // half-warp per row: lane & 15 -> 4 columns, lane >> 4 -> row parity
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int rr = ew_ * 8 + 2 * t + (lane >> 4);
                    pref[t] = rr < N ? ld4(g.res + ((size_t)b * N + rr) * g.ldr + nb + (lane & 15) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        } else if (MODE == MODE_GENERIC) {
            if (g.res && args.ksplit == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = ew_ * 8 + i;
                    pref[i] = rr < N ? ld4(g.res + ((size_t)b * N + rr) * g.ldr + nb + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        } else if (MODE == MODE_GATE) {
            const float* mul = ((lane >> 4) ? g.mul1 : g.mul0) + tile * 64 + (lane & 15) * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = ew_ * 8 + i;
                pref[i] = rr < N ? ld4(mul + ((size_t)b * N + rr) * 256) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            const int r_ = (warp & 3) * 32 + lane;
            if (g.count && r_ < N) cnt = __ldg(g.count + (size_t)b * N + r_);
        }
    }

    if (active && warp == 0) {
        // ================= TMA producer =================
        for (int it = 0; it < pre; ++it) issue_a(it);
        for (int it = pre; it < total_it; ++it) {
            mbar_wait(&empty[it % T_NSTG], ((it / T_NSTG) & 1) ^ 1);
            issue_w(it);
            issue_a(it);
        }
    } else if (active && warp == 1) {
        // ================= MMA issuer: D = Al*Wh + Ah*Wl + Ah*Wh =================
        // the whole warp runs the loop with warp-uniform operands, one lane elected inside the PTX block issues
        constexpr uint32_t idesc = make_idesc_bf16(128, TN, 0, 0);   // TN = 64: the first 64 rows of the weight box
        for (int it = 0; it < total_it; ++it) {
            const int s = it % T_NSTG, kb = it & 3;
            mbar_wait(&full[s], (it / T_NSTG) & 1);
            tc_fence_after();
            if (it < 4 && lane == 0) DBG(3 + it);
            const uint32_t st = smem_u32(smem + s * T_STAGE);
            // K-major operands: the start-address field advances by 32 bytes >> 4 per K = 16 step
            const uint64_t dah = make_smem_desc_sw128(st, 16, 1024), dal = make_smem_desc_sw128(st + T_PLANE, 16, 1024);
            const uint64_t dwh = make_smem_desc_sw128(st + 2 * T_PLANE, 16, 1024);
            const uint64_t dwl = make_smem_desc_sw128(st + 3 * T_PLANE, 16, 1024);
            const uint32_t d = tmem_base + (uint32_t)(it >> 2) * T_TN;
#pragma unroll
            for (int k16 = 0; k16 < T_KC / 16; ++k16) {
                const uint64_t o = (uint64_t)(k16 * 2);
                umma_bf16_ss_warp(d, dal + o, dwh + o, idesc, (kb | k16) != 0);
                umma_bf16_ss_warp(d, dah + o, dwl + o, idesc, 1);
                umma_bf16_ss_warp(d, dah + o, dwh + o, idesc, 1);
            }
            umma_commit_warp(&empty[s]);
        }
        umma_commit_warp(accfull);
    }
    __syncwarp();

    // ================= epilogue: warps 2..17 =================
    // phase 1: thread = (TMEM lane = kernel row, 32-column quarter): accumulator + bias in registers, LayerNorm piece
    //          statistics straight from the registers (no shuffles), values -> fp32 tile(s) in the (now idle) ring;
    // phase 2: warp = 8 rows, lane = 4 consecutive columns, so every global access is a contiguous row segment.
    const bool epi = active && warp >= 2;
    const int ew = warp - 2;
    const uint32_t crank = has_ln ? cluster_ctarank() : 0u;
    if (has_ln) cluster_wait();   // matches the arrive after setup: every CTA of the cluster is running (DSMEM is valid)
    float* S0 = reinterpret_cast<float*>(smem);
    float* S1 = S0 + 128 * T_SLD;
    const int cl = lane * 4;            // phase 2: column inside the tile
    const int col = nb + cl;            //          output column
    const int cg = col & 255;           //          column inside the 256-wide group
    if (MODE == MODE_GENERIC64) {
        // ---- 64-column tile: thread = (row, 16-column piece) in phase 1; half-warp per row in phase 2; LayerNorm
        //      groups of 256 columns span the 4 CTAs of the cluster (16 pieces of 16 columns).
        // mailbox [128 rows][16 pieces] with a row pitch of 17 float2: the 8 consecutive rows a warp merges (and the 32
        // consecutive rows a warp publishes) fall into different banks
        float2 (*box16)[17] = reinterpret_cast<float2(*)[17]>(smem + T_MAIL_OFF);
        const int hl = lane >> 4, c4 = (lane & 15) * 4;
        if (epi) {
            mbar_wait(accfull, 0);
            tc_fence_after();
            epi_bar();
            if (threadIdx.x == 64) DBG(8);
            const int q = warp & 3, pc = ew >> 2;
            const int r = q * 32 + lane;
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)pc * 16, v);
            tmem_ld_wait();
            float y[16];
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
                const float4 t = *reinterpret_cast<const float4*>(vec + V_BIAS0 * 128 + pc * 16 + c);
                y[c] = __uint_as_float(v[c]) + t.x, y[c + 1] = __uint_as_float(v[c + 1]) + t.y;
                y[c + 2] = __uint_as_float(v[c + 2]) + t.z, y[c + 3] = __uint_as_float(v[c + 3]) + t.w;
            }
            auto stats_publish = [&]() {
                float sm = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) sm += y[c];
                const float mu = sm * (1.f / 16.f);
                float m2 = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) m2 += (y[c] - mu) * (y[c] - mu);
                for (int k = 0; k < 4; ++k) st_cluster_f32x2(&box16[r][(int)crank * 4 + pc], (uint32_t)k, mu, m2);
            };
            if (has_ln && !g.res) stats_publish();
            float* d0 = S0 + r * T_SLD64 + pc * 16;
#pragma unroll
            for (int c = 0; c < 16; c += 4) *reinterpret_cast<float4*>(d0 + c) = make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]);
            epi_bar();
            if (g.res) {   // residual: coalesced add into the tile, then the statistics of the finished rows
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float4* p4 = reinterpret_cast<float4*>(S0 + (ew * 8 + 2 * t + hl) * T_SLD64 + c4);
                    *p4 = add4(*p4, pref[t]);
                }
                epi_bar();
                if (has_ln) {
#pragma unroll
                    for (int c = 0; c < 16; c += 4) {
                        const float4 t = *reinterpret_cast<const float4*>(d0 + c);
                        y[c] = t.x, y[c + 1] = t.y, y[c + 2] = t.z, y[c + 3] = t.w;
                    }
                    stats_publish();
                }
            }
            if (threadIdx.x == 64) DBG(9);
        }
        if (has_ln) {
            cluster_arrive();
            cluster_wait();
        }
        if (threadIdx.x == 64) DBG(11);
        float pm = 0.f, pr = 1.f;
        if (epi && has_ln && lane < 8) {   // (mean, rstd) of row 8 ew + lane: merge of the 16 pieces (Chan et al.)
            const int rr = ew * 8 + lane;
            float2 pcs[16];
            float sm = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) pcs[k] = box16[rr][k], sm += pcs[k].x;
            pm = sm * (1.f / 16.f);
            float m2 = 0.f, dv = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) m2 += pcs[k].y, dv += (pcs[k].x - pm) * (pcs[k].x - pm);
            pr = rsqrtf((m2 + 16.f * dv) * (1.f / 256.f) + U_LN_EPS);
        }
        if (threadIdx.x == 64) DBG(10);
        if (epi) {
            const int col64 = nb + c4, cg64 = col64 & 255;
            const int act = nb < 256 ? g.act0 : g.act1;
            const float4 ga = *reinterpret_cast<const float4*>(vec + V_GA0 * 128 + c4);
            const float4 be = *reinterpret_cast<const float4*>(vec + V_BE0 * 128 + c4);
            const int ldy = g.ldy, nstore = g.nstore, xrows = g.xrows;
            const bool vecst = (ldy & 3) == 0 && col64 + 4 <= nstore;
            float* Yb = g.Y ? g.Y + ((size_t)b * N) * ldy + col64 : nullptr;
            uint16_t* Pb = g.p_slot >= 0 ? arena_row(args.arena, unit, g.p_slot + (nb >> 8), 0, 0) + cg64 : nullptr;
            uint16_t* Xb = g.xplanes ? g.xplanes + (((size_t)(g.xunit0 + b) * 2) * xrows) * 256 + col64 : nullptr;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int i = 2 * t + hl;
                const int rr = ew * 8 + i;
                const bool rok = rr < N;
                float4 v = *reinterpret_cast<const float4*>(S0 + rr * T_SLD64 + c4);
                const float mean = __shfl_sync(0xffffffffu, pm, i), rstd = __shfl_sync(0xffffffffu, pr, i);
                if (has_ln) v = ln4(v, mean, rstd, ga, be);
                v = act4(v, act);
                if (Yb && rok) {
                    float* dst = Yb + (size_t)rr * ldy;
                    if (vecst) {
                        *reinterpret_cast<float4*>(dst) = v;
                    } else {
                        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (col64 + j < nstore) dst[j] = e[j];
                    }
                }
                if (Pb) store_planes4(rok ? v : make_float4(0.f, 0.f, 0.f, 0.f), Pb + rr * 256, Pb + (128 + rr) * 256);
                if (Xb && rr < xrows) store_planes4(v, Xb + (size_t)rr * 256, Xb + ((size_t)xrows + rr) * 256);
            }
        }
    } else {
    if (epi) {
        mbar_wait(accfull, 0);
        tc_fence_after();
        epi_bar();                      // the staged parameter vectors are visible to all epilogue warps
        if (threadIdx.x == 64) DBG(8);
        const int q = warp & 3, qt = ew >> 2;
        const int r = q * 32 + lane;
        const int c0 = nb + qt * 32;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)qt * 32;
        float* d0 = S0 + r * T_SLD + qt * 32;
        float y[32];
        float mu, m2;
        tmem_ld32f(taddr, y);
        if (MODE == MODE_GENERIC) {
            const bool res = g.res && args.ksplit == 1;
            add_vec32(y, vec + V_BIAS0 * 128 + qt * 32);
            if (has_ln && !res) {
                stats32(y, mu, m2);
                mail.publish(0, (int)crank * 4 + qt, r, mu, m2, 2);
            }
            store32(d0, y);
            epi_bar();
            if (res) {   // residual: coalesced add into the tile, then the statistics of the finished rows
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4* p4 = reinterpret_cast<float4*>(S0 + (ew * 8 + i) * T_SLD + cl);
                    *p4 = add4(*p4, pref[i]);
                }
                epi_bar();
                if (has_ln) {
                    load32(d0, y);
                    stats32(y, mu, m2);
                    mail.publish(0, (int)crank * 4 + qt, r, mu, m2, 2);
                }
            }
        } else if (MODE == MODE_DUAL) {
            // acc0 = dynamic_layer(pooled) (+ count * folded bias), acc1 = input_layer(kernel); kernel_updator.py:58-69
            float z[32];
            tmem_ld32f(taddr + T_TN, z);
            add_vec32(y, vec + V_BIAS0 * 128 + qt * 32);
            fma_vec32(y, cnt, vec + V_CBIAS0 * 128 + qt * 32);
            add_vec32(z, vec + V_BIAS1 * 128 + qt * 32);
            if (!has_ln) {   // gate_feats = input_in * param_in -> A operand of the gates
#pragma unroll
                for (int c = 0; c < 32; ++c) y[c] *= z[c];
                store32(d0, y);
            } else {
                stats32(y, mu, m2);
                mail.publish(0, (int)crank * 4 + qt, r, mu, m2, 2);
                stats32(z, mu, m2);
                mail.publish(1, (int)crank * 4 + qt, r, mu, m2, 2);
                store32(d0, y);
                store32(S1 + r * T_SLD + qt * 32, z);
            }
            epi_bar();
        } else {   // MODE_GATE: columns [0,64) = input_gate, [64,128) = update_gate of features 64*tile .. +63
            add_vec32(y, vec + V_BIAS0 * 128 + qt * 32);
            stats32(y, mu, m2);
            mail.publish(qt >> 1, (int)crank * 2 + (qt & 1), r, mu, m2, 4);
            store32(d0, y);
            epi_bar();
        }
        if (threadIdx.x == 64) DBG(9);
    }
    if (has_ln) {
        cluster_arrive();
        cluster_wait();
    }
    if (threadIdx.x == 64) DBG(11);
    // LayerNorm (mean, rstd) of this warp's 8 rows, one row per lane (lanes 8..15: the second array)
    float pm = 0.f, pr = 1.f;
    if (epi && has_ln && lane < (MODE == MODE_GENERIC ? 8 : 16)) mail.combine((lane >> 3) & 1, ew * 8 + (lane & 7), pm, pr);
    if (threadIdx.x == 64) DBG(10);

    if (MODE == MODE_GENERIC) {
        if (epi) {
            const int act = nb < 256 ? g.act0 : g.act1;
            const float4 ga = *reinterpret_cast<const float4*>(vec + V_GA0 * 128 + cl);
            const float4 be = *reinterpret_cast<const float4*>(vec + V_BE0 * 128 + cl);
            const int ldy = g.ldy, nstore = g.nstore, xrows = g.xrows;
            const bool vec = (ldy & 3) == 0 && col + 4 <= nstore;
            float* yp = g.Y ? g.Y + ((size_t)split * args.R + (size_t)b * N + ew * 8) * ldy + col : nullptr;
            uint16_t* pp = g.p_slot >= 0 ? arena_row(args.arena, unit, g.p_slot + (nb >> 8), 0, ew * 8) + cg : nullptr;
            uint16_t* xp = g.xplanes ? g.xplanes + (((size_t)(g.xunit0 + b) * 2) * xrows + ew * 8) * 256 + col : nullptr;
            const float* sp = S0 + ew * 8 * T_SLD + cl;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = ew * 8 + i;
                const bool rok = rr < N;
                float4 v = *reinterpret_cast<const float4*>(sp + i * T_SLD);
                const float mean = __shfl_sync(0xffffffffu, pm, i), rstd = __shfl_sync(0xffffffffu, pr, i);
                if (has_ln) v = ln4(v, mean, rstd, ga, be);
                v = act4(v, act);
                if (yp && rok) {
                    float* dst = yp + (size_t)i * ldy;
                    if (vec) {
                        *reinterpret_cast<float4*>(dst) = v;
                    } else {
                        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (col + j < nstore) dst[j] = e[j];
                    }
                }
                if (pp) store_planes4(rok ? v : make_float4(0.f, 0.f, 0.f, 0.f), pp + i * 256, pp + (128 + i) * 256);
                if (xp && rr < xrows) store_planes4(v, xp + i * 256, xp + ((size_t)xrows + i) * 256);
            }
        }
    } else if (MODE == MODE_DUAL) {
        if (epi && !has_ln) {
            uint16_t* pp = arena_row(args.arena, unit, g.p_slot, 0, ew * 8) + col;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = ew * 8 + i;
                const float4 v = rr < N ? *reinterpret_cast<float4*>(S0 + rr * T_SLD + cl) : make_float4(0.f, 0.f, 0.f, 0.f);
                store_planes4(v, pp + i * 256, pp + (128 + i) * 256);
            }
        } else if (epi) {   // param_out = norm_out(.), input_out = input_norm_out(.)   (:78-79)
            const int c2 = col - 256;
            const float4 ga0 = *reinterpret_cast<const float4*>(vec + V_GA0 * 128 + cl);
            const float4 be0 = *reinterpret_cast<const float4*>(vec + V_BE0 * 128 + cl);
            const float4 ga1 = *reinterpret_cast<const float4*>(vec + V_GA1 * 128 + cl);
            const float4 be1 = *reinterpret_cast<const float4*>(vec + V_BE1 * 128 + cl);
            float* y0 = g.Y + ((size_t)b * N + ew * 8) * 256 + c2;
            float* y1 = g.Y1 + ((size_t)b * N + ew * 8) * 256 + c2;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = ew * 8 + i;
                const float mean0 = __shfl_sync(0xffffffffu, pm, i), rstd0 = __shfl_sync(0xffffffffu, pr, i);
                const float mean1 = __shfl_sync(0xffffffffu, pm, i + 8), rstd1 = __shfl_sync(0xffffffffu, pr, i + 8);
                if (rr < N) {
                    *reinterpret_cast<float4*>(y0 + i * 256) =
                        ln4(*reinterpret_cast<float4*>(S0 + rr * T_SLD + cl), mean0, rstd0, ga0, be0);
                    *reinterpret_cast<float4*>(y1 + i * 256) =
                        ln4(*reinterpret_cast<float4*>(S1 + rr * T_SLD + cl), mean1, rstd1, ga1, be1);
                }
            }
        }
    } else {   // MODE_GATE   (:73-88)
        if (epi) {
            const int half = lane >> 4;
            const int f = tile * 64 + (lane & 15) * 4;   // feature column of this lane
            const float4 ga = *reinterpret_cast<const float4*>(vec + V_GA0 * 128 + cl);
            const float4 be = *reinterpret_cast<const float4*>(vec + V_BE0 * 128 + cl);
            uint16_t* pp = arena_row(args.arena, unit, g.p_slot, 0, ew * 8) + f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = ew * 8 + i;
                const float mean = __shfl_sync(0xffffffffu, pm, i + 8 * half);
                const float rstd = __shfl_sync(0xffffffffu, pr, i + 8 * half);
                float4 v = ln4(*reinterpret_cast<float4*>(S0 + rr * T_SLD + cl), mean, rstd, ga, be);
                v = act4(v, ACT_SIGMOID);
                const float4 t = pref[i];
                v = make_float4(v.x * t.x, v.y * t.y, v.z * t.z, v.w * t.w);
                // features = update_gate * param_out + input_gate * input_out: lanes l and l + 16 hold the same feature
                v.x += __shfl_xor_sync(0xffffffffu, v.x, 16), v.y += __shfl_xor_sync(0xffffffffu, v.y, 16);
                v.z += __shfl_xor_sync(0xffffffffu, v.z, 16), v.w += __shfl_xor_sync(0xffffffffu, v.w, 16);
                if (half == 0) store_planes4(v, pp + i * 256, pp + (128 + i) * 256);
            }
        }
    }
    }   // MODE != MODE_GENERIC64
    if (threadIdx.x == 64) DBG(12);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<256>(tmem_base);
    if (threadIdx.x == 0) DBG(13);
}

// ---------------------------------------------------------------------------------------------------------------
// prep: fp32 rows -> arena planes.
//   job 0/1: pooled[branch] = sum over the S pooling partials (fixed order) -> SLOT_POOLED (+ count, job 0)
//   job 2  : proposal_feat -> SLOT_INP of the mask branch
//   job 3  : depth_proposal + proposal_feat (kernel_update_head.py:250) -> SLOT_INP of the depth branch
// Also serves pf_kernel_updator (S = 1, two plain copies).  Thread = one float4 of one row.
struct PrepArgs {
    const float* src[4];     // job sources: partial base (jobs 0,1) or rows (jobs 2,3)
    const float* add[4];     // optional second addend
    int S[4];                // > 0: sum over S partials laid out [unit][S][N][256]
    int unit0[4], slot[4];
    const float* cntp;       // [B][S][N] (job 0 only) or null
    float* count;
    uint16_t* arena;
    int B, N, njobs;
};
__global__ void __launch_bounds__(256) prep_kernel(const __grid_constant__ PrepArgs a) {
    long long* dbg = dbg_claim(1);
    pdl_wait();
    pdl_launch_dependents();
    DBG(2);
    const int job = blockIdx.y;
    const int idx = blockIdx.x * 256 + threadIdx.x;   // (b, c4, r) with r < 128 fastest: the pooling partials are
    const int r = idx & 127, c4 = (idx >> 7) & 63, b = idx >> 13;   // column-group-major ([unit][S][64][N][4], pf_pool.cu)
    if (b >= a.B) return;
    const int N = a.N, S = a.S[job];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < N) {
        if (S > 0) {
            const float4* src = reinterpret_cast<const float4*>(a.src[job] + ((size_t)b * S) * N * 256) + (size_t)c4 * N + r;
            const size_t stride4 = (size_t)N * 64;
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
            if (c4 == 0 && a.cntp && a.count && job == 0) {
                float k = 0.f;
                for (int s2 = 0; s2 < S; ++s2) k += __ldg(a.cntp + ((size_t)b * S + s2) * N + r);
                a.count[b * N + r] = k;
            }
        } else {
            acc = __ldg(reinterpret_cast<const float4*>(a.src[job] + ((size_t)b * N + r) * 256) + c4);
            if (a.add[job]) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(a.add[job] + ((size_t)b * N + r) * 256) + c4);
                acc.x += t.x, acc.y += t.y, acc.z += t.z, acc.w += t.w;
            }
        }
    }
    const float h0 = bf16_round(acc.x), h1 = bf16_round(acc.y), h2 = bf16_round(acc.z), h3 = bf16_round(acc.w);
    const int unit = a.unit0[job] + b;
    uint16_t* ph = arena_row(a.arena, unit, a.slot[job], 0, r) + c4 * 4;
    uint16_t* pl = arena_row(a.arena, unit, a.slot[job], 1, r) + c4 * 4;
    *reinterpret_cast<uint2*>(ph) = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
    *reinterpret_cast<uint2*>(pl) = make_uint2(pack_bf16x2(acc.x - h0, acc.y - h1), pack_bf16x2(acc.z - h2, acc.w - h3));
    DBG(13);
}

// ---------------------------------------------------------------------------------------------------------------
// ffn_norm(x + ffn2 split-K partials + b2) (kernel_update_head.py:271-272): warp = row, lane owns 8 columns.
// Writes obj_feat / depth_feat_new (the stage's outputs, fp32) and SLOT_OBJ2 (A operand of the FC heads).
struct SumLnArgs {
    const float* part[2];    // [nsplit][R][256]
    const float* res[2];     // [R][256]
    const float* bias[2];
    const float* ln[2];
    float* out[2];
    uint16_t* arena;
    int B, N, R, nsplit;
};
__global__ void __launch_bounds__(256) sumln_kernel(const __grid_constant__ SumLnArgs a) {
    long long* dbg = dbg_claim(2);
    pdl_wait();
    pdl_launch_dependents();
    DBG(2);
    const int br = blockIdx.y;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);   // b * 128 + r
    const int lane = threadIdx.x & 31;
    const int b = row >> 7, r = row & 127;
    if (b >= a.B) return;
    const int k0 = lane * 8;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 0.f;
    if (r < a.N) {
        const size_t m = (size_t)b * a.N + r;
        float4 u[8], v[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            if (s < a.nsplit) {
                const float* pp = a.part[br] + ((size_t)s * a.R + m) * 256 + k0;
                u[s] = ld4(pp), v[s] = ld4(pp + 4);
            } else {
                u[s] = v[s] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        const float4 bu = ld4(a.bias[br] + k0), bv = ld4(a.bias[br] + k0 + 4);
        const float4 ru = ld4(a.res[br] + m * 256 + k0), rv = ld4(a.res[br] + m * 256 + k0 + 4);
        const float4 gu = ld4(a.ln[br] + k0), gv = ld4(a.ln[br] + k0 + 4);
        const float4 eu = ld4(a.ln[br] + 256 + k0), ev = ld4(a.ln[br] + 256 + k0 + 4);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            x[0] += u[s].x, x[1] += u[s].y, x[2] += u[s].z, x[3] += u[s].w;
            x[4] += v[s].x, x[5] += v[s].y, x[6] += v[s].z, x[7] += v[s].w;
        }
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
        const float gam[8] = {gu.x, gu.y, gu.z, gu.w, gv.x, gv.y, gv.z, gv.w};
        const float bet[8] = {eu.x, eu.y, eu.z, eu.w, ev.x, ev.y, ev.z, ev.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = (x[i] - mean) * rstd * gam[i] + bet[i];
        *reinterpret_cast<float4*>(a.out[br] + m * 256 + k0) = make_float4(x[0], x[1], x[2], x[3]);
        *reinterpret_cast<float4*>(a.out[br] + m * 256 + k0 + 4) = make_float4(x[4], x[5], x[6], x[7]);
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float h0 = bf16_round(x[2 * i]), h1 = bf16_round(x[2 * i + 1]);
        hi[i] = pack_bf16x2(h0, h1);
        lo[i] = pack_bf16x2(x[2 * i] - h0, x[2 * i + 1] - h1);
    }
    const int unit = br * a.B + b;
    *reinterpret_cast<uint4*>(arena_row(a.arena, unit, SLOT_OBJ2, 0, r) + k0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(arena_row(a.arena, unit, SLOT_OBJ2, 1, r) + k0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    DBG(13);
}

// ---------------------------------------------------------------------------------------------------------------
// Inter-kernel self-attention of one (branch, image, head): softmax(q k^T / sqrt(32)) v over the N kernels of the
// image (mmcv MultiheadAttention -> nn.MultiheadAttention, seq-first; kernel_update_head.py:259-260).
// qkv [R][768] = [q | k | v] fp32; output (heads concatenated, before out_proj) -> SLOT_ATT planes.
// 8 threads per query; keys are dealt to the 8 threads in blocks of 4 consecutive keys so that K^T rows and V rows
// are read as float4 (4 FMAs per shared-memory load; V chunks XOR-swizzled by key block: conflict-free).  Two-pass
// softmax with the scores kept in registers, octet reduction by shuffles.  Four CTAs per (branch, image, head)
// split the queries, so the 64 (image, head) pairs of a batch of 4 spread over 512 CTAs.
constexpr int ATT_TPQ = 8;                        // threads per query
constexpr int ATT_NB = PF_MAX_N / (4 * ATT_TPQ);  // key blocks of 4 per thread (4 -> 16 keys per thread)
constexpr int ATT_QPT = 2;                        // queries per thread: every K / V chunk read from shared memory feeds two
                                                  // queries (the kernel is bound by LDS.128 quarter-warp phases, not FMAs)
constexpr int ATT_QPC = 64;                       // queries per CTA
constexpr int ATT_THREADS = ATT_QPC / ATT_QPT * ATT_TPQ;   // 256
constexpr int ATT_CPH = PF_MAX_N / ATT_QPC;       // CTAs per (branch, image, head)
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const float* __restrict__ qkv0,
                                                                const float* __restrict__ qkv1,
                                                                uint16_t* __restrict__ arena, int B, int N) {
    __shared__ __align__(16) float s_kt[32][PF_MAX_N + 4];   // K transposed: [d][key]
    __shared__ __align__(16) float s_v[PF_MAX_N][32];        // V: 16-byte chunk c of key j stored at chunk c ^ ((j >> 2) & 7)
    const int h = blockIdx.x / ATT_CPH, qblk = blockIdx.x % ATT_CPH, b = blockIdx.y;
    long long* dbg = dbg_claim(3);
    pdl_wait();
    pdl_launch_dependents();
    DBG(2);
    const float* qkv = (blockIdx.z == 0 ? qkv0 : qkv1) + (size_t)b * N * 768;
    // this thread's two query rows first: their loads are in flight together with the K / V loads below
    const int part = threadIdx.x % ATT_TPQ;
    int nrow[ATT_QPT];
    float q[ATT_QPT][32];
    const float scale = 0.17677669529663687f;  // 1/sqrt(32), applied to q before q k^T as torch does
#pragma unroll
    for (int t = 0; t < ATT_QPT; ++t) {
        nrow[t] = qblk * ATT_QPC + t * (ATT_QPC / ATT_QPT) + threadIdx.x / ATT_TPQ;
        const int nq = nrow[t] < N ? nrow[t] : N - 1;   // keep whole octets alive for the shuffles
#pragma unroll
        for (int d4 = 0; d4 < 8; ++d4) {
            const float4 v = ld4(qkv + (size_t)nq * 768 + h * 32 + d4 * 4);
            q[t][d4 * 4] = v.x * scale, q[t][d4 * 4 + 1] = v.y * scale, q[t][d4 * 4 + 2] = v.z * scale, q[t][d4 * 4 + 3] = v.w * scale;
        }
    }
    {   // K and V of this head: 128 keys x 8 float4 each; all 8 loads of a thread are issued before the first store
        float4 kk[4], vv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = threadIdx.x + j * ATT_THREADS;   // 0 .. 1023
            const int n = i >> 3, d4 = i & 7;
            const bool ok = n < N;
            kk[j] = ok ? ld4(qkv + (size_t)n * 768 + 256 + h * 32 + d4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            vv[j] = ok ? ld4(qkv + (size_t)n * 768 + 512 + h * 32 + d4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = threadIdx.x + j * ATT_THREADS;
            const int n = i >> 3, d4 = i & 7;
            s_kt[d4 * 4][n] = kk[j].x, s_kt[d4 * 4 + 1][n] = kk[j].y, s_kt[d4 * 4 + 2][n] = kk[j].z, s_kt[d4 * 4 + 3][n] = kk[j].w;
            *reinterpret_cast<float4*>(&s_v[n][(d4 ^ ((n >> 2) & 7)) * 4]) = vv[j];
        }
    }
    __syncthreads();
    DBG(3);
    float sc[ATT_QPT][ATT_NB][4];
    float mx[ATT_QPT];
#pragma unroll
    for (int t = 0; t < ATT_QPT; ++t) mx[t] = -INFINITY;
#pragma unroll
    for (int i = 0; i < ATT_NB; ++i) {
        const int j0 = 4 * (part + ATT_TPQ * i);
        float4 a[ATT_QPT];
#pragma unroll
        for (int t = 0; t < ATT_QPT; ++t) a[t] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int d = 0; d < 32; ++d) {
            const float4 kk = *reinterpret_cast<const float4*>(&s_kt[d][j0]);
#pragma unroll
            for (int t = 0; t < ATT_QPT; ++t)
                a[t].x += q[t][d] * kk.x, a[t].y += q[t][d] * kk.y, a[t].z += q[t][d] * kk.z, a[t].w += q[t][d] * kk.w;
        }
#pragma unroll
        for (int t = 0; t < ATT_QPT; ++t) {
            sc[t][i][0] = j0 < N ? a[t].x : -INFINITY, sc[t][i][1] = j0 + 1 < N ? a[t].y : -INFINITY;
            sc[t][i][2] = j0 + 2 < N ? a[t].z : -INFINITY, sc[t][i][3] = j0 + 3 < N ? a[t].w : -INFINITY;
            mx[t] = fmaxf(fmaxf(mx[t], fmaxf(sc[t][i][0], sc[t][i][1])), fmaxf(sc[t][i][2], sc[t][i][3]));
        }
    }
#pragma unroll
    for (int t = 0; t < ATT_QPT; ++t)
#pragma unroll
        for (int o = 1; o < ATT_TPQ; o <<= 1) mx[t] = fmaxf(mx[t], __shfl_xor_sync(0xffffffffu, mx[t], o));
    DBG(4);
    float o[ATT_QPT][32], den[ATT_QPT];
#pragma unroll
    for (int t = 0; t < ATT_QPT; ++t) {
        den[t] = 0.f;
#pragma unroll
        for (int d = 0; d < 32; ++d) o[t][d] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < ATT_NB; ++i) {
        const int j0 = 4 * (part + ATT_TPQ * i);   // (j0 >> 2) & 7 == part
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float pj[ATT_QPT];
#pragma unroll
            for (int t = 0; t < ATT_QPT; ++t) {
                pj[t] = __expf(sc[t][i][e] - mx[t]);   // exp(-inf) = 0 for padded keys (their V rows are zero)
                den[t] += pj[t];
            }
#pragma unroll
            for (int d4 = 0; d4 < 8; ++d4) {
                const float4 vv = *reinterpret_cast<const float4*>(&s_v[j0 + e][(d4 ^ part) * 4]);
#pragma unroll
                for (int t = 0; t < ATT_QPT; ++t)
                    o[t][d4 * 4] += pj[t] * vv.x, o[t][d4 * 4 + 1] += pj[t] * vv.y, o[t][d4 * 4 + 2] += pj[t] * vv.z,
                        o[t][d4 * 4 + 3] += pj[t] * vv.w;
            }
        }
    }
    const int unit = blockIdx.z * B + b;
#pragma unroll
    for (int t = 0; t < ATT_QPT; ++t) {
        float dn = den[t];
#pragma unroll
        for (int s2 = 1; s2 < ATT_TPQ; s2 <<= 1) dn += __shfl_xor_sync(0xffffffffu, dn, s2);
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
        // lane `part` writes dims [4 part, 4 part + 4) of query row n as bf16 hi / lo (rows >= N: zeros)
        const int n = nrow[t];
        const bool ok = n < N;
        store_planes4(ok ? make_float4(r4[0] * inv, r4[1] * inv, r4[2] * inv, r4[3] * inv) : make_float4(0.f, 0.f, 0.f, 0.f),
                      arena_row(arena, unit, SLOT_ATT, 0, n) + h * 32 + part * 4,
                      arena_row(arena, unit, SLOT_ATT, 1, n) + h * 32 + part * 4);
    }
    DBG(13);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
constexpr int kFfnSplit = 8;

// tensor maps over long-lived buffers (weight stacks, workspaces) are cached: encoding one costs ~1 us per call
int cached_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    struct Entry { const void* base; uint64_t rows, cols; uint32_t box_rows; CUtensorMap map; };
    static Entry cache[64];
    static int n = 0, next = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < n; ++i)
        if (cache[i].base == base && cache[i].rows == rows && cache[i].cols == cols && cache[i].box_rows == box_rows) {
            *out = cache[i].map;
            return PF_OK;
        }
    CUtensorMap m;
    if (int e = make_tmap_bf16_2d(&m, base, rows, cols, cols, box_rows, T_KC)) return e;
    cache[next] = Entry{base, rows, cols, box_rows, m};
    next = (next + 1) % 64;
    if (n < 64) ++n;
    *out = m;
    return PF_OK;
}

struct Maps {
    CUtensorMap w256, wffn, a;
};

template <int MODE>
static int launch_tc(const Maps& mp, const TcArgs& a, int tiles, int njobs, int cluster, cudaStream_t st) {
    // the attribute is per (function, device): remember which devices have it (a process may drive several GPUs)
    static std::mutex attr_mu;
    static bool attr_done[64] = {};
    {
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lock(attr_mu);
        if (dev < 0 || dev >= 64 || !attr_done[dev]) {
            cudaError_t e = cudaFuncSetAttribute(tcgemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM);
            if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "tcgemm smem attribute: %s", cudaGetErrorString(e));
            if (dev >= 0 && dev < 64) attr_done[dev] = true;
        }
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(tiles, a.B * a.ksplit, njobs);
    cfg.blockDim = dim3(T_THREADS);
    cfg.dynamicSmemBytes = T_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    attrs[1].id = cudaLaunchAttributeClusterDimension;
    attrs[1].val.clusterDim.x = cluster, attrs[1].val.clusterDim.y = 1, attrs[1].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tcgemm_kernel<MODE>, mp.w256, mp.wffn, mp.a, a);
    if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "tcgemm_kernel<%d> launch: %s", MODE, cudaGetErrorString(e));
    count_launch();
    return PF_OK;
}

template <typename Args>
static int launch_pdl(void (*kern)(Args), dim3 grid, dim3 block, const Args& a, const char* name, cudaStream_t st) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid, cfg.blockDim = block, cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
    if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "%s launch: %s", name, cudaGetErrorString(e));
    count_launch();
    return PF_OK;
}

static int launch_attention(const float* q0, const float* q1, uint16_t* arena, int B, int N, int nbranch, cudaStream_t st) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(PF_HEADS * ATT_CPH, B, nbranch);
    cfg.blockDim = dim3(ATT_THREADS);
    cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, attention_kernel, q0, q1, arena, B, N);
    if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "attention_kernel launch: %s", cudaGetErrorString(e));
    count_launch();
    return PF_OK;
}

static TcJob blank_job() {
    TcJob g;
    memset(&g, 0, sizeof(g));
    g.npass = 1;
    g.p_slot = -1;
    return g;
}
static size_t align256(size_t v) { return (v + 255) / 256 * 256; }
static size_t arena_bytes(int units) { return align256((size_t)units * NSLOT * SLOT_ELEMS * 2); }
// fused cluster kernel (pf_stage.cu)
bool fused_update_enabled();
size_t stage_fused_ws_bytes(int B);
int launch_stage_fused(const pf_stage_weights* w, const float* partial, const float* cntp, int S, const float* obj_in,
                       const float* dep_in, float* obj_out, float* dep_out, float* cls_out, float* kern, uint16_t* kern_split,
                       float* kbias, uint16_t* arena, float* part, int B, int N, int cls_sigmoid, cudaStream_t st);

static size_t update_ws_bytes(int B, int N) {
    const size_t R = (size_t)B * N;
    // per-layer kernels, per branch: pon, ion, obj0, obj1 [R][256]; qkv [R][768]; part [8][R][256]
    const size_t per_branch = align256(R * (4 * 256 + 768 + kFfnSplit * 256) * 4);
    const size_t legacy = 2 * per_branch + align256(R * 4);
    // fused kernel: ffn2 partials [2B][8][128][256] + two [R][256] temporaries for in-place callers
    const size_t fused = align256(stage_fused_ws_bytes(B)) + 2 * align256(R * 256 * 4);
    return arena_bytes(2 * B) + (legacy > fused ? legacy : fused) + 256;
}

struct BranchScratch {
    float *pon, *ion, *obj0, *obj1, *qkv, *part;
};

static int make_maps(Maps* mp, const pf_stage_weights* w, const void* arena, int units) {
    if (int e = cached_tmap_2d(&mp->w256, w->wstack256, (uint64_t)w->wstack256_rows, 256)) return e;
    if (w->wstack_ffn) {
        if (int e = cached_tmap_2d(&mp->wffn, w->wstack_ffn, (uint64_t)w->wstack_ffn_rows, (uint64_t)w->ffn_channels)) return e;
    } else {
        mp->wffn = mp->w256;
    }
    return cached_tmap_2d(&mp->a, arena, (uint64_t)units * NSLOT * 256, 256);
}

// KernelUpdator.forward for `nb` branches on arena units [unit0_of_branch(b) ...): launches dyn+inp, gates, fc.
// fc output: fp32 rows into out[b] (may be null) and/or planes into slot out_slot (or -1).
static int run_updator(const Maps& mp, const pf_stage_weights* w, uint16_t* arena, BranchScratch* sc, const float* count,
                       float* const* out, int out_slot, int nbr, int B, int N, cudaStream_t st) {
    const int R = B * N;
    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.arena = arena, a.B = B, a.N = N, a.R = R, a.ksplit = 1;
    // 1. parameters = dynamic_layer(pooled') | input_feats = input_layer(kernel)   (kernel_updator.py:58-69, 78-79)
    for (int b = 0; b < nbr; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcJob g = blank_job();
        g.unit0 = b * B, g.ntiles = 4, g.npass = 2;
        g.pass[0] = TcPass{SLOT_POOLED, 0, bw.dyn_w, 512, 0};
        g.pass[1] = TcPass{SLOT_INP, 0, bw.inp_w, 512, 0};
        g.bias0 = bw.dyn_b, g.cbias0 = bw.dyn_cb, g.count = bw.dyn_cb ? count : nullptr, g.bias1 = bw.inp_b;
        g.ln0 = bw.ln_norm_out, g.ln1 = bw.ln_input_norm_out;
        g.Y = sc[b].pon, g.Y1 = sc[b].ion, g.p_slot = SLOT_GATEIN;
        a.job[b] = g;
    }
    if (int e = launch_tc<MODE_DUAL>(mp, a, 4, nbr, 2, st)) return e;
    // 2. gates + mixing   (:69-88)
    for (int b = 0; b < nbr; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcJob g = blank_job();
        g.unit0 = b * B, g.ntiles = 4;
        g.pass[0] = TcPass{SLOT_GATEIN, 0, bw.gate_w, 512, 0};
        g.bias0 = bw.gate_b, g.ln0 = bw.ln_input_norm_in, g.ln1 = bw.ln_norm_in;
        g.mul0 = sc[b].ion, g.mul1 = sc[b].pon, g.p_slot = SLOT_MIX;
        a.job[b] = g;
    }
    if (int e = launch_tc<MODE_GATE>(mp, a, 4, nbr, 4, st)) return e;
    // 3. relu(fc_norm(fc_layer(features)))   (:89-92)
    for (int b = 0; b < nbr; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcJob g = blank_job();
        g.unit0 = b * B, g.ntiles = 4;
        g.pass[0] = TcPass{SLOT_MIX, 0, bw.fc_w, 256, 0};
        g.bias0 = bw.fc_b, g.ln0 = bw.ln_fc_norm, g.act0 = ACT_RELU;
        g.Y = out[b], g.ldy = 256, g.nstore = 256, g.p_slot = out_slot;
        a.job[b] = g;
    }
    return launch_tc<MODE_GENERIC64>(mp, a, 4, nbr, 4, st);
}

}  // namespace pf

// Debug: device buffer of int64 [16 + 16 * capacity] (zeroed by the caller) receiving the in-kernel timeline of CTA
// (0,0,0) of every tcgemm launch; pass NULL to switch it off.  Not part of the product path.
PF_DEFINE_DBG_SETTER(set_dbg_update)
namespace pf {
int set_dbg_pool(long long* p);
int set_dbg_einsum(long long* p);
int set_dbg_stage(long long* p);
}
extern "C" int pf_debug_timeline(long long* device_buffer) {
    const int e = pf::set_dbg_update(device_buffer) | pf::set_dbg_pool(device_buffer) | pf::set_dbg_einsum(device_buffer) |
                  pf::set_dbg_stage(device_buffer);
    return e == 0 ? PF_OK : pf::set_error(PF_ERR_CUDA, "pf_debug_timeline: cudaMemcpyToSymbol failed (%d)", e);
}

extern "C" size_t pf_update_workspace_bytes(int B, int N, int ffn_channels) {
    if (B <= 0 || N <= 0 || ffn_channels <= 0) return 0;
    return pf::update_ws_bytes(B, N);
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
    PF_REQUIRE(workspace_bytes >= update_ws_bytes(B, N), PF_ERR_WORKSPACE, "pf_kernel_update: workspace %zu < %zu",
               workspace_bytes, update_ws_bytes(B, N));
    PF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PF_ERR_ALIGN, "pf_kernel_update: workspace not 256-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int R = B * N;

    char* wp = static_cast<char*>(workspace);
    uint16_t* arena = reinterpret_cast<uint16_t*>(wp);
    wp += arena_bytes(2 * B);
    if (fused_update_enabled()) {
        // ONE launch (pf_stage.cu).  The kernel reads obj_in / dep_in in its first phase and writes obj_out / dep_out in
        // a late one, cluster by cluster: an in-place caller goes through two temporaries.
        float* part = reinterpret_cast<float*>(wp);
        wp += align256(stage_fused_ws_bytes(B));
        float* tmp_obj = reinterpret_cast<float*>(wp);
        wp += align256((size_t)R * 256 * 4);
        float* tmp_dep = reinterpret_cast<float*>(wp);
        const bool alias = obj_out == obj_in || dep_out == dep_in || obj_out == dep_in || dep_out == obj_in;
        if (int e = launch_stage_fused(w, partial, cntp, S, obj_in, dep_in, alias ? tmp_obj : obj_out, alias ? tmp_dep : dep_out,
                                       cls_out, kern, kern_split, kbias, arena, part, B, N, cls_sigmoid, st))
            return e;
        if (alias) {
            cudaError_t e1 = cudaMemcpyAsync(obj_out, tmp_obj, (size_t)R * 256 * 4, cudaMemcpyDeviceToDevice, st);
            cudaError_t e2 = cudaMemcpyAsync(dep_out, tmp_dep, (size_t)R * 256 * 4, cudaMemcpyDeviceToDevice, st);
            if (e1 != cudaSuccess || e2 != cudaSuccess)
                return set_error(PF_ERR_CUDA, "pf_kernel_update: output copy: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
            count_launch(2);
        }
        return PF_OK;
    }
    BranchScratch sc[2];
    for (int b = 0; b < 2; ++b) {
        float* p = reinterpret_cast<float*>(wp);
        sc[b].pon = p, p += (size_t)R * 256;
        sc[b].ion = p, p += (size_t)R * 256;
        sc[b].obj0 = p, p += (size_t)R * 256;
        sc[b].obj1 = p, p += (size_t)R * 256;
        sc[b].qkv = p, p += (size_t)R * 768;
        sc[b].part = p, p += (size_t)kFfnSplit * R * 256;
        wp += align256(reinterpret_cast<char*>(p) - wp);
    }
    float* count = reinterpret_cast<float*>(wp);    // [R]
    float* out_[2] = {obj_out, dep_out};

    Maps mp;
    if (int e = make_maps(&mp, w, arena, 2 * B)) return e;

    // 0. pooled' = sum of the split-K pooling partials (fixed order); kernel operands; depth kernel += mask kernel
    {
        PrepArgs p;
        memset(&p, 0, sizeof(p));
        p.arena = arena, p.B = B, p.N = N, p.njobs = 4, p.cntp = cntp, p.count = count;
        p.src[0] = partial, p.S[0] = S, p.unit0[0] = 0, p.slot[0] = SLOT_POOLED;
        p.src[1] = partial + (size_t)B * S * N * 256, p.S[1] = S, p.unit0[1] = B, p.slot[1] = SLOT_POOLED;
        p.src[2] = obj_in, p.unit0[2] = 0, p.slot[2] = SLOT_INP;
        p.src[3] = dep_in, p.add[3] = obj_in, p.unit0[3] = B, p.slot[3] = SLOT_INP;
        if (int e = launch_pdl(prep_kernel, dim3((B * 128 * 64 + 255) / 256, 4), dim3(256), p, "prep_kernel", st)) return e;
    }
    // 1-3. KernelUpdator x2 -> obj0 (fp32 residual + SLOT_OBJ0 planes)
    float* obj0_[2] = {sc[0].obj0, sc[1].obj0};
    if (int e = run_updator(mp, w, arena, sc, count, obj0_, SLOT_OBJ0, 2, B, N, st)) return e;

    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.arena = arena, a.B = B, a.N = N, a.R = R, a.ksplit = 1;

    // 4. attention in-projection
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcJob g = blank_job();
        g.unit0 = b * B, g.ntiles = 12;
        g.pass[0] = TcPass{SLOT_OBJ0, 0, bw.qkv_w, 768, 0};
        g.bias0 = bw.qkv_b, g.Y = sc[b].qkv, g.ldy = 768, g.nstore = 768;
        a.job[b] = g;
    }
    if (int e = launch_tc<MODE_GENERIC64>(mp, a, 12, 2, 1, st)) return e;

    // 5. softmax(q k^T) v per (branch, image, head) -> SLOT_ATT
    if (int e = launch_attention(sc[0].qkv, sc[1].qkv, arena, B, N, 2, st)) return e;

    // 6. attention_norm(x + out_proj(attn))
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcJob g = blank_job();
        g.unit0 = b * B, g.ntiles = 4;
        g.pass[0] = TcPass{SLOT_ATT, 0, bw.out_w, 256, 0};
        g.bias0 = bw.out_b, g.res = sc[b].obj0, g.ldr = 256, g.ln0 = bw.ln_attn;
        g.Y = sc[b].obj1, g.ldy = 256, g.nstore = 256, g.p_slot = SLOT_OBJ1;
        a.job[b] = g;
    }
    if (int e = launch_tc<MODE_GENERIC64>(mp, a, 4, 2, 4, st)) return e;

    // 7. FFN layer 1 + ReLU -> SLOT_HID0..7 (one slot per 256 hidden channels)
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcJob g = blank_job();
        g.unit0 = b * B, g.ntiles = ffn / T_TN;
        g.pass[0] = TcPass{SLOT_OBJ1, 0, bw.ffn1_w, ffn, 0};
        g.bias0 = bw.ffn1_b, g.act0 = g.act1 = ACT_RELU, g.p_slot = SLOT_HID0;
        a.job[b] = g;
    }
    if (int e = launch_tc<MODE_GENERIC>(mp, a, ffn / T_TN, 2, 1, st)) return e;

    // 8. FFN layer 2, split-K over the 8 hidden slots -> fp32 partials
    a.ksplit = kFfnSplit;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcJob g = blank_job();
        g.unit0 = b * B, g.ntiles = 2;
        g.pass[0] = TcPass{SLOT_HID0, 1, bw.ffn2_w, 256, 0};
        g.Y = sc[b].part, g.ldy = 256, g.nstore = 256;
        a.job[b] = g;
    }
    if (int e = launch_tc<MODE_GENERIC>(mp, a, 2, 2, 1, st)) return e;
    a.ksplit = 1;

    // 9. ffn_norm(x + sum of partials + b2) -> obj_feat / depth_feat_new (returned to the caller) + SLOT_OBJ2
    {
        SumLnArgs s;
        memset(&s, 0, sizeof(s));
        s.arena = arena, s.B = B, s.N = N, s.R = R, s.nsplit = kFfnSplit;
        for (int b = 0; b < 2; ++b) {
            s.part[b] = sc[b].part, s.res[b] = sc[b].obj1, s.bias[b] = w->br[b].ffn2_b, s.ln[b] = w->br[b].ln_ffn;
            s.out[b] = out_[b];
        }
        if (int e = launch_pdl(sumln_kernel, dim3(B * 128 / 8, 2), dim3(256), s, "sumln_kernel", st)) return e;
    }

    // 10. cls_fcs / mask_fcs / depth_regs: Linear(no bias) + LN (+ ReLU except depth_regs) -> SLOT_HEAD0 / SLOT_HEAD1
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcJob g = blank_job();
        g.unit0 = b * B, g.ntiles = (b == 0) ? 8 : 4;
        g.pass[0] = TcPass{SLOT_OBJ2, 0, bw.head_w, (b == 0) ? 512 : 256, 0};
        g.ln0 = bw.ln_head_a, g.ln1 = bw.ln_head_b;
        g.act0 = g.act1 = bw.head_relu ? ACT_RELU : ACT_NONE;
        g.p_slot = (b == 0) ? SLOT_HEAD0 : SLOT_HEAD1;
        a.job[b] = g;
    }
    if (int e = launch_tc<MODE_GENERIC64>(mp, a, 8, 2, 4, st)) return e;

    // 11. fc_mask / fc_depth with feat_transform folded in -> dynamic kernels (bf16 hi/lo planes) + their logit bias
    //     (one extra weight row), fc_cls
    int nj = 0;
    for (int b = 0; b < 2; ++b) {
        const pf_branch_weights& bw = w->br[b];
        TcJob g = blank_job();
        g.unit0 = b * B, g.ntiles = 4;
        g.pass[0] = TcPass{SLOT_HEAD1, 0, bw.kern_w, 256, 0};
        g.bias0 = bw.kern_b;
        g.Y = kern ? kern + (size_t)b * R * 256 : nullptr, g.ldy = 256, g.nstore = 256;
        g.xplanes = kern_split, g.xrows = N, g.xunit0 = b * B;
        a.job[nj++] = g;
        TcJob k = blank_job();
        k.unit0 = b * B, k.ntiles = 1;
        k.pass[0] = TcPass{SLOT_HEAD1, 0, bw.kbrow_w, 128, 0};
        k.bias0 = bw.kbrow_b;
        k.Y = kbias + (size_t)b * R, k.ldy = 1, k.nstore = 1;
        a.job[nj++] = k;
    }
    if (cls_out) {
        const pf_branch_weights& bw = w->br[0];
        TcJob g = blank_job();
        g.unit0 = 0, g.ntiles = 1;
        g.pass[0] = TcPass{SLOT_HEAD0, 0, bw.cls_w, 128, 0};
        g.bias0 = bw.cls_b, g.act0 = cls_sigmoid ? ACT_SIGMOID : ACT_NONE;
        g.Y = cls_out, g.ldy = w->num_classes, g.nstore = w->num_classes;
        a.job[nj++] = g;
    }
    return launch_tc<MODE_GENERIC64>(mp, a, 4, nj, 1, st);
}

extern "C" size_t pf_updator_workspace_bytes(int R) {
    if (R <= 0) return 0;
    const int units = (R + 127) / 128;
    return pf::arena_bytes(units) + pf::align256((size_t)R * 512 * sizeof(float)) + 256;
}

// KernelUpdator.forward on its own (polyphonic/funcs/kernel_updator.py:55-93): no feat_transform fold, one branch.
// Rows are processed in groups of <= 128 ("images" of N = 128 rows); a ragged tail is a second launch set.
extern "C" int pf_kernel_updator(const pf_stage_weights* w, const float* update_feature, const float* input_feature,
                                 float* out, void* workspace, size_t workspace_bytes, int R, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(w && update_feature && input_feature && out && workspace, PF_ERR_ARG, "pf_kernel_updator: null pointer");
    PF_REQUIRE(R > 0 && w->wstack256, PF_ERR_ARG, "pf_kernel_updator: R=%d", R);
    PF_REQUIRE(workspace_bytes >= pf_updator_workspace_bytes(R), PF_ERR_WORKSPACE, "pf_kernel_updator: workspace %zu < %zu",
               workspace_bytes, pf_updator_workspace_bytes(R));
    PF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(update_feature) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(input_feature) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               PF_ERR_ALIGN, "pf_kernel_updator: workspace must be 256-byte, tensors 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int units = (R + 127) / 128;
    uint16_t* arena = static_cast<uint16_t*>(workspace);
    float* pon = reinterpret_cast<float*>(static_cast<char*>(workspace) + arena_bytes(units));
    float* ion = pon + (size_t)R * 256;
    Maps mp;
    if (int e = make_maps(&mp, w, arena, units)) return e;
    for (int r0 = 0; r0 < R;) {
        const int full_groups = (R - r0) / 128;
        const int Bg = full_groups > 0 ? full_groups : 1;
        const int Ng = full_groups > 0 ? 128 : (R - r0);
        const int u0 = r0 / 128;
        uint16_t* ar = arena + (size_t)u0 * NSLOT * SLOT_ELEMS;   // this group's units start at arena unit 0 of `ar`
        Maps mg = mp;
        if (int e = cached_tmap_2d(&mg.a, ar, (uint64_t)Bg * NSLOT * 256, 256)) return e;
        PrepArgs p;
        memset(&p, 0, sizeof(p));
        p.arena = ar, p.B = Bg, p.N = Ng, p.njobs = 2;
        p.src[0] = update_feature + (size_t)r0 * 256, p.slot[0] = SLOT_POOLED;
        p.src[1] = input_feature + (size_t)r0 * 256, p.slot[1] = SLOT_INP;
        if (int e = launch_pdl(prep_kernel, dim3((Bg * 128 * 64 + 255) / 256, 2), dim3(256), p, "prep_kernel", st)) return e;
        BranchScratch sc;
        memset(&sc, 0, sizeof(sc));
        sc.pon = pon + (size_t)r0 * 256, sc.ion = ion + (size_t)r0 * 256;
        float* o = out + (size_t)r0 * 256;
        if (int e = run_updator(mg, w, ar, &sc, nullptr, &o, -1, 1, Bg, Ng, st)) return e;
        r0 += Bg * Ng;
    }
    return PF_OK;
}
