// Panoptic + depth post-processing of one image, replacing
//   KernelUpdateIterHead.get_panoptic / merge_stuff_thing_stuff_joint   polyphonic/kernel_update.py:421-535
//   KernelUpdateHead.rescale_masks / rescale_depth                      polyphonic/kernel_update_head.py:593-626
//   depth_act                                                           polyphonic/funcs/depth_utils.py:1-19
//
// The reference materialises 111 sigmoid masks and 111 depth maps at full resolution twice (1.9 GB of fp32 per
// 1024x2048 frame and pass) and then walks the segments in a Python loop with ~300 host synchronisations.  Here the
// x4 bilinear up-sampling is fused into the consumers and nothing full-resolution is ever materialised except the
// three result maps:
//   pp_select  (1 CTA)    top-k of the 100x8 thing scores (bitonic sort), stuff scores sorted, one "champion" entry
//                         per proposal: of all (mask, label) entries that share a mask only the best-scoring one can
//                         win the per-pixel argmax of score * mask, the others end with area 0 and are skipped by the
//                         reference's loop as well (kernel_update.py:505-507);
//   pp_argmax  (tiles)    per output pixel: up-sampled sigmoid of every proposal's mask from a shared-memory patch of
//                         the low-resolution logits, argmax of score * mask -> winner id (u8), per-proposal winner
//                         area and per-proposal count of mask >= 0.5 (integer atomics: deterministic);
//   pp_merge   (1 thread) the reference's sequential loop over entries by descending score (score / overlap
//                         thresholds, running segment id) -> segment id per proposal + the segments_info records;
//   pp_paint   (pixels)   panoptic id, depth_basic (up-sampled initial depth) and depth_final (the winner's depth map
//                         where a segment was painted), depth_act applied before the interpolation as the reference.
// Supported geometry: predictions at 1/4 of the padded input (Hb = 4h, Wb = 4w -- fixed by the model), output = the
// top-left H0 x W0 crop of the padded input, ori_shape == img_shape (the second F.interpolate is then the identity).
#include <math.h>

#include "pf_internal.h"
#include "pf_sm100.cuh"
#include "pf_up2.cuh"

namespace pf {

constexpr int PP_MAXN = PF_MAX_N;       // proposals (masks)
constexpr int PP_MAXE = 128;            // entries: max_per_img things + stuff
constexpr int PP_TY = 16, PP_TX = 64;   // output tile
constexpr int PP_PR = PP_TY / 4 + 2, PP_PC = PP_TX / 4 + 2;   // low-resolution patch incl. the bilinear halo
constexpr int PP_THREADS = 256;

struct PpTables {                       // lives in the workspace
    int entry_mask[PP_MAXE];            // proposal index of entry e
    int entry_label[PP_MAXE];
    float entry_score[PP_MAXE];
    int n_entries, n_thing_entries;
    int champ_entry[PP_MAXN];           // best entry of proposal n, or -1
    float champ_score[PP_MAXN];
    int area[PP_MAXN];                  // pixels won by proposal n
    int orig[PP_MAXN];                  // pixels with mask_n >= 0.5
    int segid[PP_MAXN];                 // segment id painted for proposal n (0 = none)
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float depth_act_(float x, int mode) {
    const float min_depth = 0.01f, max_depth = 80.f;
    const float s = sigmoidf_(x);
    if (mode == 0) {   // monodepth: 1 / (min_disp + (max_disp - min_disp) * sigmoid)
        const float min_disp = 1.f / max_depth, max_disp = 1.f / min_depth;
        return 1.f / (min_disp + (max_disp - min_disp) * s);
    }
    return s * (max_depth - min_depth) + min_depth;   // 'sigmoid'
}

// ATen upsample_bilinear2d, align_corners = False, scale = in / out = 0.25
struct Tap {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ Tap make_tap(int dst, int in_size) {
    const float s = fmaxf(0.25f * ((float)dst + 0.5f) - 0.5f, 0.f);
    Tap t;
    t.i0 = (int)s;
    t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
    t.l1 = s - (float)t.i0;
    t.l0 = 1.f - t.l1;
    return t;
}

// ---------------------------------------------------------------------------------------------------------------
// one CTA per frame (blockIdx.x); tables of frame f at tb_base + f * ws_stride bytes
__global__ void __launch_bounds__(1024) pp_select_kernel(const float* __restrict__ cls_all, uint8_t* __restrict__ tb_base,
                                                         size_t ws_stride, int N, int P, int T, int ncls, int max_per_img) {
    const float* cls = cls_all + (size_t)blockIdx.x * N * ncls;
    PpTables* tb = reinterpret_cast<PpTables*>(tb_base + (size_t)blockIdx.x * ws_stride);
    __shared__ float s_key[1024];
    __shared__ int s_idx[1024];
    const int t = threadIdx.x;
    const int cand = P * T;
    s_key[t] = t < cand ? cls[(t / T) * ncls + (t % T)] : -INFINITY;
    s_idx[t] = t;
    __syncthreads();
    // bitonic sort, descending by score, ascending by index among equal scores
    for (int k = 2; k <= 1024; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int o = t ^ j;
            if (o > t) {
                const float a = s_key[t], b = s_key[o];
                const int ia = s_idx[t], ib = s_idx[o];
                const bool a_first = a > b || (a == b && ia < ib);   // a should come before b
                const bool up = (t & k) == 0;
                if (up ? !a_first : a_first) {
                    s_key[t] = b, s_key[o] = a;
                    s_idx[t] = ib, s_idx[o] = ia;
                }
            }
            __syncthreads();
        }
    const int nt = max_per_img < cand ? max_per_img : cand;
    if (t < nt) {
        tb->entry_mask[t] = s_idx[t] / T;
        tb->entry_label[t] = s_idx[t] % T;
        tb->entry_score[t] = s_key[t];
    }
    if (t < PP_MAXN) tb->area[t] = 0, tb->orig[t] = 0, tb->segid[t] = 0;
    __shared__ float s_stuff[PP_MAXN];
    __shared__ int s_emask[PP_MAXE];
    __shared__ float s_escore[PP_MAXE];
    const int S = N - P;   // stuff kernels: score = cls[P + i][T + i], sorted descending (kernel_update.py:449-451)
    if (t < S) s_stuff[t] = cls[(P + t) * ncls + T + t];
    if (t < nt) s_emask[t] = s_idx[t] / T, s_escore[t] = s_key[t];
    __syncthreads();
    if (t == 0) {
        int order[PP_MAXN];
        for (int i = 0; i < S; ++i) order[i] = i;
        for (int i = 1; i < S; ++i) {
            const int v = order[i];
            const float sv = s_stuff[v];
            int j = i - 1;
            while (j >= 0 && s_stuff[order[j]] < sv) order[j + 1] = order[j], --j;
            order[j + 1] = v;
        }
        for (int i = 0; i < S; ++i) {
            tb->entry_mask[nt + i] = P + order[i];
            tb->entry_label[nt + i] = T + order[i];
            tb->entry_score[nt + i] = s_stuff[order[i]];
            s_emask[nt + i] = P + order[i], s_escore[nt + i] = s_stuff[order[i]];
        }
        tb->n_entries = nt + S, tb->n_thing_entries = nt;
    }
    __syncthreads();
    if (t < N) {   // champion of proposal t: its lowest-numbered entry (thing entries are sorted by descending score)
        int ce = -1;
        for (int e = nt + S - 1; e >= 0; --e)
            if (s_emask[e] == t) ce = e;
        tb->champ_entry[t] = ce;
        tb->champ_score[t] = ce >= 0 ? s_escore[ce] : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// in_stride2: mask_logits are [N][h/2][w/2] (the decoder's own resolution) and their x2 bilinear up-sampling
// (kernel_update.py:133-143) is evaluated on the fly with pf_upsample2x's arithmetic instead of being read from memory
__global__ void __launch_bounds__(PP_THREADS) pp_argmax_kernel(const float* __restrict__ mask_all, uint8_t* __restrict__ ws_base,
                                                               size_t ws_stride, size_t tb_bytes, int N, int h, int w, int H0,
                                                               int W0, int in_stride2) {
    const int hs = in_stride2 ? h / 2 : h, wsrc = in_stride2 ? w / 2 : w;
    const float* mask_logits = mask_all + (size_t)blockIdx.z * N * hs * wsrc;
    PpTables* tb = reinterpret_cast<PpTables*>(ws_base + (size_t)blockIdx.z * ws_stride);
    uint8_t* ids = ws_base + (size_t)blockIdx.z * ws_stride + tb_bytes;
    extern __shared__ float s_sig[];   // [N][PP_PR][PP_PC] sigmoid of the low-resolution logits under this tile
    __shared__ float s_score[PP_MAXN];
    __shared__ int s_entry[PP_MAXN];
    __shared__ int s_area[PP_MAXN], s_orig[PP_MAXN];
    __shared__ float s_lo[PP_MAXN], s_hi[PP_MAXN];   // min / max of the sigmoid patch of proposal n
    __shared__ int s_list[PP_MAXN], s_nlist;
    const int tid = threadIdx.x, lane = tid & 31;
    const int Y0 = blockIdx.y * PP_TY, X0 = blockIdx.x * PP_TX;
    const int py0 = make_tap(Y0, h).i0, px0 = make_tap(X0, w).i0;
    for (int i = tid; i < PP_MAXN; i += PP_THREADS) {
        s_area[i] = 0, s_orig[i] = 0;
        s_score[i] = i < N ? tb->champ_score[i] : 0.f;
        s_entry[i] = i < N ? tb->champ_entry[i] : -1;
    }
    constexpr int PATCH = PP_PR * PP_PC;
    for (int i = tid; i < N * PATCH; i += PP_THREADS) {
        const int n = i / PATCH, r = (i % PATCH) / PP_PC, c = i % PP_PC;
        const int yy = min(py0 + r, h - 1), xx = min(px0 + c, w - 1);
        s_sig[i] = sigmoidf_(in_stride2 ? up2_val(mask_logits + (size_t)n * hs * wsrc, hs, wsrc, yy, xx)
                                        : __ldg(mask_logits + ((size_t)n * h + yy) * w + xx));
    }
    __syncthreads();
    // Tile-level pruning (exact).  A bilinear sample is a convex combination of patch values, so inside this tile
    // score_n * min(patch_n) <= score_n * mask_n(p) <= score_n * max(patch_n).  With LB = max_n score_n * min(patch_n)
    // every pixel's winner has a product >= LB; proposal n cannot win anywhere in the tile if score_n * max(patch_n)
    // < LB, and it adds nothing to the mask >= 0.5 counts if max(patch_n) < 0.5: such proposals are skipped.
    for (int n = tid >> 5; n < N; n += PP_THREADS / 32) {
        float lo = INFINITY, hi = -INFINITY;
        for (int i = lane; i < PATCH; i += 32) {
            const float v = s_sig[n * PATCH + i];
            lo = fminf(lo, v), hi = fmaxf(hi, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) s_lo[n] = lo, s_hi[n] = hi;
    }
    __syncthreads();
    if (tid < 32) {
        float lb = -1.f;
        for (int n = lane; n < N; n += 32)
            if (s_entry[n] >= 0) lb = fmaxf(lb, s_score[n] * s_lo[n]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lb = fmaxf(lb, __shfl_xor_sync(0xffffffffu, lb, o));
        int cnt = 0;
        for (int n0 = 0; n0 < N; n0 += 32) {   // ordered compaction: the list keeps ascending n
            const int n = n0 + lane;
            const bool keep = n < N && (s_hi[n] >= 0.5f || (s_entry[n] >= 0 && s_score[n] * s_hi[n] >= lb));
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) s_list[cnt + __popc(m & ((1u << lane) - 1u))] = n;
            cnt += __popc(m);
        }
        if (lane == 0) s_nlist = cnt;
    }
    __syncthreads();
    const int nlist = s_nlist;
    // 4 vertically ADJACENT pixels per thread (rows 4 ty0 .. 4 ty0 + 3 of the tile): with a x4 up-sampling they sample at
    // most THREE source rows (Y0 is a multiple of 4), so a mask costs 6 shared-memory loads and 3 horizontal + 4 vertical
    // interpolations per thread instead of 16 loads and 4 x 3.  The arithmetic per pixel is unchanged:
    // l0y * (l0x a + l1x b) + l1y * (l0x c + l1x d), ATen's order.
    const int tx = tid & (PP_TX - 1), ty0 = tid / PP_TX;
    const int X = X0 + tx;
    const Tap cx = make_tap(X, w);
    const int ox0 = cx.i0 - px0, ox1 = cx.i1 - px0;
    int r0[4], r1[4];          // source row of each pixel's two taps, relative to the first one (0..2)
    float ly0[4], ly1[4];
    bool inside[4];
    int rbase = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int Y = Y0 + 4 * ty0 + j;
        const Tap cy = make_tap(Y, h);
        if (j == 0) rbase = cy.i0;
        r0[j] = cy.i0 - rbase, r1[j] = cy.i1 - rbase;
        ly0[j] = cy.l0, ly1[j] = cy.l1;
        inside[j] = Y < H0 && X < W0;
    }
    // rows beyond the map (bottom border) are never selected (r <= i1 <= h - 1); clamp the LOAD to the patch all the same
    const int orow0 = (rbase - py0) * PP_PC;
    const int orow1 = min(rbase + 1 - py0, PP_PR - 1) * PP_PC, orow2 = min(rbase + 2 - py0, PP_PR - 1) * PP_PC;
    float best[4] = {-1.f, -1.f, -1.f, -1.f};
    int best_e[4] = {1 << 30, 1 << 30, 1 << 30, 1 << 30}, best_n[4] = {0, 0, 0, 0};
    for (int li = 0; li < nlist; ++li) {
        const int n = s_list[li];
        const float* sp = s_sig + n * PATCH;
        const float sc = s_score[n];
        const int e = s_entry[n];
        float hr[3];
        hr[0] = cx.l0 * sp[orow0 + ox0] + cx.l1 * sp[orow0 + ox1];
        hr[1] = cx.l0 * sp[orow1 + ox0] + cx.l1 * sp[orow1 + ox1];
        hr[2] = cx.l0 * sp[orow2 + ox0] + cx.l1 * sp[orow2 + ox1];
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float ha = r0[j] == 0 ? hr[0] : (r0[j] == 1 ? hr[1] : hr[2]);
            const float hb = r1[j] == 0 ? hr[0] : (r1[j] == 1 ? hr[1] : hr[2]);
            const float m = ly0[j] * ha + ly1[j] * hb;
            cnt += __popc(__ballot_sync(0xffffffffu, inside[j] && m >= 0.5f));
            const float prob = sc * m;
            if (e >= 0 && (prob > best[j] || (prob == best[j] && e < best_e[j]))) best[j] = prob, best_e[j] = e, best_n[j] = n;
        }
        if (lane == 0 && cnt) atomicAdd(&s_orig[n], cnt);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int Y = Y0 + 4 * ty0 + j;
        if (inside[j]) {
            ids[(size_t)Y * W0 + X] = (uint8_t)best_n[j];
            atomicAdd(&s_area[best_n[j]], 1);
        }
    }
    __syncthreads();
    for (int i = tid; i < N; i += PP_THREADS) {
        if (s_area[i]) atomicAdd(&tb->area[i], s_area[i]);
        if (s_orig[i]) atomicAdd(&tb->orig[i], s_orig[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pp_merge_kernel(uint8_t* __restrict__ tb_base, size_t ws_stride,
                                                       pf_segment* __restrict__ segs_all, int seg_stride,
                                                       int* __restrict__ n_segs_all, int T, float instance_score_thr,
                                                       float overlap_thr) {
    PpTables* tb = reinterpret_cast<PpTables*>(tb_base + (size_t)blockIdx.x * ws_stride);
    pf_segment* segs = segs_all + (size_t)blockIdx.x * seg_stride;
    int* n_segs = n_segs_all + blockIdx.x;
    __shared__ float s_score[PP_MAXE];
    __shared__ int s_mask[PP_MAXE], s_label[PP_MAXE], s_champ[PP_MAXN], s_area[PP_MAXN], s_orig[PP_MAXN], s_segid[PP_MAXN];
    __shared__ int s_order[PP_MAXE];
    const int t = threadIdx.x;
    const int E = tb->n_entries, nth = tb->n_thing_entries;
    if (t < E) s_score[t] = tb->entry_score[t], s_mask[t] = tb->entry_mask[t], s_label[t] = tb->entry_label[t];
    if (t < PP_MAXN) s_champ[t] = tb->champ_entry[t], s_area[t] = tb->area[t], s_orig[t] = tb->orig[t], s_segid[t] = 0;
    __syncthreads();
    if (t == 0) {
        // argsort(-total_scores), stable: the thing entries [0, nth) and the stuff entries [nth, E) are each sorted
        // by descending score already (pp_select), so this is a two-way merge
        int i0 = 0, i1 = nth, no = 0;
        while (i0 < nth || i1 < E) {
            const bool take0 = i1 >= E || (i0 < nth && s_score[i0] >= s_score[i1]);
            s_order[no++] = take0 ? i0++ : i1++;
        }
        int seg = 0;
        for (int i = 0; i < E; ++i) {
            const int k = s_order[i];
            const int label = s_label[k];
            const bool isthing = label < T;
            const float score = s_score[k];
            if (isthing && score < instance_score_thr) continue;
            const int n = s_mask[k];
            if (s_champ[n] != k) continue;          // a better entry owns this mask: this one won no pixel
            const int area = s_area[n], orig = s_orig[n];
            if (area > 0 && orig > 0) {
                if ((double)area / (double)orig < (double)overlap_thr) continue;
                ++seg;
                s_segid[n] = seg;
                pf_segment sg;
                sg.id = seg, sg.isthing = isthing ? 1 : 0, sg.category_id = label, sg.instance_id = k, sg.area = area, sg.score = score;
                segs[seg - 1] = sg;
            }
        }
        *n_segs = seg;
    }
    __syncthreads();
    if (t < PP_MAXN) tb->segid[t] = s_segid[t];
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pp_paint_kernel(const float* __restrict__ depth_all, const float* __restrict__ dinit_all,
                                                       const uint8_t* __restrict__ ws_base, size_t ws_stride, size_t tb_bytes,
                                                       int32_t* __restrict__ panoptic_all, float* __restrict__ dfinal_all,
                                                       float* __restrict__ dbasic_all, int N, int h, int w, int H0, int W0,
                                                       int depth_mode, int in_stride2) {
    const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y, f = blockIdx.z;
    if (X >= W0 || Y >= H0) return;
    const int hs = in_stride2 ? h / 2 : h, wsrc = in_stride2 ? w / 2 : w;
    const float* depth_logits = depth_all + (size_t)f * N * hs * wsrc;
    const float* depth_init = dinit_all + (size_t)f * hs * wsrc;
    const PpTables* tb = reinterpret_cast<const PpTables*>(ws_base + (size_t)f * ws_stride);
    const uint8_t* ids = ws_base + (size_t)f * ws_stride + tb_bytes;
    const Tap cy = make_tap(Y, h), cx = make_tap(X, w);
    auto tap = [&](const float* map, int yi, int xi) {   // value of the (h, w) map at (yi, xi)
        return depth_act_(in_stride2 ? up2_val(map, hs, wsrc, yi, xi) : __ldg(map + yi * w + xi), depth_mode);
    };
    auto sample = [&](const float* map) {
        const float a = tap(map, cy.i0, cx.i0), b = tap(map, cy.i0, cx.i1), c = tap(map, cy.i1, cx.i0), d = tap(map, cy.i1, cx.i1);
        return cy.l0 * (cx.l0 * a + cx.l1 * b) + cy.l1 * (cx.l0 * c + cx.l1 * d);
    };
    const size_t p = (size_t)Y * W0 + X;
    const size_t fo = (size_t)f * H0 * W0;
    const int n = ids[p];
    const int s = tb->segid[n];
    const float dinit = sample(depth_init);
    panoptic_all[fo + p] = s;
    dbasic_all[fo + p] = dinit;
    dfinal_all[fo + p] = s > 0 ? sample(depth_logits + (size_t)n * hs * wsrc) : dinit;
}

static size_t pp_align(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace pf

extern "C" size_t pf_panoptic_workspace_bytes(int H0, int W0) {
    if (H0 <= 0 || W0 <= 0) return 0;
    return pf::pp_align(sizeof(pf::PpTables)) + pf::pp_align((size_t)H0 * W0);
}

extern "C" int pf_panoptic_batch(const float* cls_scores, const float* mask_logits, const float* depth_logits,
                                 const float* depth_init, int B, int N, int num_proposals, int num_thing_classes,
                                 int num_classes, int h, int w, int H0, int W0, int max_per_img, float instance_score_thr,
                                 float overlap_thr, int depth_mode, int in_stride2, int32_t* panoptic, float* depth_final,
                                 float* depth_basic, pf_segment* segments, int segment_stride, int* n_segments,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(cls_scores && mask_logits && depth_logits && depth_init && panoptic && depth_final && depth_basic && segments &&
                   n_segments && workspace,
               PF_ERR_ARG, "pf_panoptic: null pointer");
    PF_REQUIRE(B > 0 && B <= 65535, PF_ERR_ARG, "pf_panoptic: batch %d", B);
    PF_REQUIRE(N > 0 && N <= PP_MAXN && num_proposals > 0 && num_proposals <= N && num_thing_classes > 0 &&
                   num_thing_classes + (N - num_proposals) <= num_classes && num_proposals * num_thing_classes <= 1024,
               PF_ERR_ARG, "pf_panoptic: bad class / proposal counts N=%d P=%d T=%d classes=%d", N, num_proposals,
               num_thing_classes, num_classes);
    PF_REQUIRE(h > 0 && w > 0 && H0 > 0 && W0 > 0 && H0 <= 4 * h && W0 <= 4 * w, PF_ERR_ARG,
               "pf_panoptic: output %dx%d must be a crop of the x4 up-sampled %dx%d predictions", H0, W0, h, w);
    PF_REQUIRE(!in_stride2 || (h % 2 == 0 && w % 2 == 0), PF_ERR_ARG, "pf_panoptic: in_stride2 needs even h, w (got %dx%d)", h, w);
    PF_REQUIRE(max_per_img > 0 && max_per_img + (N - num_proposals) <= PP_MAXE, PF_ERR_ARG, "pf_panoptic: max_per_img=%d", max_per_img);
    PF_REQUIRE(segment_stride >= max_per_img + (N - num_proposals) || B == 1, PF_ERR_ARG, "pf_panoptic: segment_stride=%d", segment_stride);
    PF_REQUIRE(depth_mode == 0 || depth_mode == 1, PF_ERR_ARG, "pf_panoptic: depth_mode must be 0 (monodepth) or 1 (sigmoid)");
    const size_t per = pf_panoptic_workspace_bytes(H0, W0);
    PF_REQUIRE(workspace_bytes >= (size_t)B * per, PF_ERR_WORKSPACE, "pf_panoptic: workspace %zu < %zu", workspace_bytes, (size_t)B * per);
    PF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PF_ERR_ALIGN, "pf_panoptic: workspace not 256-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* wsb = static_cast<uint8_t*>(workspace);
    const size_t tb_bytes = pp_align(sizeof(PpTables));

    pp_select_kernel<<<B, 1024, 0, st>>>(cls_scores, wsb, per, N, num_proposals, num_thing_classes, num_classes, max_per_img);
    PF_CHECK_LAUNCH("pp_select_kernel");
    const size_t smem = (size_t)N * PP_PR * PP_PC * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(pp_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "pp_argmax smem attribute: %s", cudaGetErrorString(e));
    dim3 grid((W0 + PP_TX - 1) / PP_TX, (H0 + PP_TY - 1) / PP_TY, B);
    pp_argmax_kernel<<<grid, PP_THREADS, smem, st>>>(mask_logits, wsb, per, tb_bytes, N, h, w, H0, W0, in_stride2);
    PF_CHECK_LAUNCH("pp_argmax_kernel");
    pp_merge_kernel<<<B, 128, 0, st>>>(wsb, per, segments, segment_stride, n_segments, num_thing_classes, instance_score_thr, overlap_thr);
    PF_CHECK_LAUNCH("pp_merge_kernel");
    pp_paint_kernel<<<dim3((W0 + 255) / 256, H0, B), 256, 0, st>>>(depth_logits, depth_init, wsb, per, tb_bytes, panoptic, depth_final,
                                                                    depth_basic, N, h, w, H0, W0, depth_mode, in_stride2);
    PF_CHECK_LAUNCH("pp_paint_kernel");
    return PF_OK;
}

extern "C" int pf_panoptic(const float* cls_scores, const float* mask_logits, const float* depth_logits,
                           const float* depth_init, int N, int num_proposals, int num_thing_classes, int num_classes, int h,
                           int w, int H0, int W0, int max_per_img, float instance_score_thr, float overlap_thr,
                           int depth_mode, int32_t* panoptic, float* depth_final, float* depth_basic,
                           pf_segment* segments, int* n_segments, void* workspace, size_t workspace_bytes, void* stream) {
    return pf_panoptic_batch(cls_scores, mask_logits, depth_logits, depth_init, 1, N, num_proposals, num_thing_classes,
                             num_classes, h, w, H0, W0, max_per_img, instance_score_thr, overlap_thr, depth_mode, 0, panoptic,
                             depth_final, depth_basic, segments, 128, n_segments, workspace, workspace_bytes, stream);
}
