// The video model's tracking path -- reference: polyphonic/polyphonic_former_video.py:364-451,
// polyphonic/video/utils.py:40-82, polyphonic/funcs/utils.py:4-22, polyphonic/video/track_heads.py:92-102,
// polyphonic/video/qdtrack/trackers/quasi_dense_embed_tracker.py:46-207, mmdet SingleRoIExtractor + mmcv RoIAlign.
//
//   mask -> box      per-item column / row pixel histograms (integer atomics, deterministic) -> centre, mean absolute
//                    deviation, tight box; items are masks [K][H][W] or segment ids of pf_panoptic's map
//   RoIAlign         7 x 7 bins, 2 x 2 samples, aligned, pyramid level from the box size; written straight into the bf16
//                    hi / lo activation planes of the first convolution: item = 64 rows (7 grid rows of pitch 8 + zero
//                    padding) x 256 channels, so that a 3 x 3 tap is a row shift of dy * 8 + dx (pf_sgemm.cuh)
//   embedding head   4 x [3x3 conv + GroupNorm32 + ReLU] = 4 launches of sgemm_kernel<GN64> (9 taps x 4 channel blocks,
//                    GroupNorm inside the epilogue: an item's 49 positions are 64 rows of ONE tile), FC 12544 -> 1024 =
//                    sgemm_kernel<PARTIAL> (split-K x 7 over the grid rows, 56 CTAs stream the 51 MB of weights once),
//                    then one small fp32 kernel: split-K sum + bias + ReLU + FC 1024 -> 256
//   association      ONE CTA: score sort, IoU duplicate removal, embeds x memo^T, bi-softmax, category mask, the sequential
//                    greedy assignment, new ids and the memo update; the memo (tracklets + backdrops) lives on the device
#include <math.h>

#include "pf_internal.h"
#include "pf_sgemm.cuh"
#include "pf_sm100.cuh"

namespace pf {

constexpr int TR_RP = 64;            // rows per item in the activation planes
constexpr int TR_PITCH = 8;          // row pitch of the 7 x 7 grid inside an item
constexpr int TR_GRID = 7;
constexpr int TR_FC1 = 1024;
constexpr int TR_SPLITS = 7;         // FC1 split-K: one grid row (7 positions x 256 channels) per split
constexpr int MAXK = PF_TRACK_MAX_K, MAXT = PF_TRACK_MAX_TRACKS, MAXBF = PF_TRACK_MAX_BACK_FRAMES;
constexpr int MAXM = MAXT + MAXBF * MAXK;
constexpr int EMB = PF_TRACK_EMBED;

// ------------------------------------------------------------------------------------------------ mask -> box
constexpr int BX_COLS = 256, BX_ROWS = 32;

// thread = one column of a strip of BX_ROWS rows: run-length compressed column counts, warp-aggregated row counts
template <bool FROM_IDS>
__global__ void __launch_bounds__(BX_COLS) box_hist_kernel(const float* __restrict__ masks, const int32_t* __restrict__ pan,
                                                           const int32_t* __restrict__ seg_ids, int K, int H, int W,
                                                           int* __restrict__ hx, int* __restrict__ hy) {
    __shared__ int s_slot[256];
    const int x = blockIdx.x * BX_COLS + threadIdx.x, y0 = blockIdx.y * BX_ROWS, k_mask = blockIdx.z;
    if (FROM_IDS) {
        s_slot[threadIdx.x] = -1;
        __syncthreads();
        if (threadIdx.x < K) {
            const int id = seg_ids[threadIdx.x];
            if (id > 0 && id < 256) s_slot[id] = threadIdx.x;
        }
        __syncthreads();
    }
    int cur = -1, run = 0;
    const int y1 = min(y0 + BX_ROWS, H);
    for (int y = y0; y < y1; ++y) {
        int slot = -1;
        if (x < W) {
            if (FROM_IDS) {
                const int id = __ldg(pan + (size_t)y * W + x);
                slot = (id > 0 && id < 256) ? s_slot[id] : -1;
            } else {
                slot = __ldg(masks + ((size_t)k_mask * H + y) * W + x) != 0.f ? k_mask : -1;
            }
        }
        if (slot != cur) {
            if (cur >= 0) atomicAdd(hx + (size_t)cur * W + x, run);
            cur = slot, run = 0;
        }
        ++run;
        const unsigned peers = __match_any_sync(0xffffffffu, slot);
        if (slot >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(hy + (size_t)slot * H + y, __popc(peers));
    }
    if (cur >= 0) atomicAdd(hx + (size_t)cur * W + x, run);
}

// one CTA per item: centre, mean absolute deviation and tight box from the two histograms
__global__ void __launch_bounds__(256) box_finalize_kernel(const int* __restrict__ hx, const int* __restrict__ hy, int H, int W,
                                                           float* __restrict__ rois, float* __restrict__ tight) {
    const int k = blockIdx.x, t = threadIdx.x;
    __shared__ double s_a[256], s_b[256];
    __shared__ int s_lo[256], s_hi[256];
    __shared__ float s_c[2], s_d[2];
    __shared__ int s_box[4];
    __shared__ double s_n;
    for (int axis = 0; axis < 2; ++axis) {
        const int* h = axis == 0 ? hx + (size_t)k * W : hy + (size_t)k * H;
        const int L = axis == 0 ? W : H;
        double n = 0, sx = 0;
        int lo = 1 << 30, hi = -1;
        for (int i = t; i < L; i += 256) {
            const int c = h[i];
            if (c) n += c, sx += (double)c * i, lo = min(lo, i), hi = max(hi, i);
        }
        s_a[t] = n, s_b[t] = sx, s_lo[t] = lo, s_hi[t] = hi;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (t < o) s_a[t] += s_a[t + o], s_b[t] += s_b[t + o], s_lo[t] = min(s_lo[t], s_lo[t + o]), s_hi[t] = max(s_hi[t], s_hi[t + o]);
            __syncthreads();
        }
        if (t == 0) {
            s_n = s_a[0];
            s_c[axis] = s_a[0] > 0 ? (float)(s_b[0] / s_a[0]) : 0.f;     // torch.mean of the fp32 coordinates
            s_box[axis] = s_lo[0], s_box[2 + axis] = s_hi[0];
        }
        __syncthreads();
        const float c = s_c[axis];
        double dev = 0;
        for (int i = t; i < L; i += 256) {
            const int cnt = h[i];
            if (cnt) dev += (double)cnt * (double)fabsf((float)i - c);   // |coord - centre| rounded to fp32 as the reference
        }
        s_a[t] = dev;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (t < o) s_a[t] += s_a[t + o];
            __syncthreads();
        }
        if (t == 0) s_d[axis] = s_n > 0 ? fmaxf((float)(s_a[0] / s_n), 1.f) : 0.f;
        __syncthreads();
    }
    if (t == 0) {
        const bool empty = s_n <= 0;
        float* r = rois + k * 5;
        r[0] = 0.f;
        if (empty) {
            r[1] = r[2] = r[3] = r[4] = 0.f;
            tight[k * 4 + 0] = -1.f, tight[k * 4 + 1] = -1.f, tight[k * 4 + 2] = 10.f, tight[k * 4 + 3] = 10.f;
        } else {
            const float cx = s_c[0], cy = s_c[1], dx = s_d[0], dy = s_d[1];
            r[1] = fmaxf(__fsub_rn(cx, __fmul_rn(dx, 2.f)), 0.f), r[2] = fmaxf(__fsub_rn(cy, __fmul_rn(dy, 2.f)), 0.f);
            r[3] = fmaxf(__fadd_rn(cx, __fmul_rn(dx, 2.f)), 0.f), r[4] = fmaxf(__fadd_rn(cy, __fmul_rn(dy, 2.f)), 0.f);
            tight[k * 4 + 0] = (float)s_box[0], tight[k * 4 + 1] = (float)s_box[1];
            tight[k * 4 + 2] = (float)s_box[2], tight[k * 4 + 3] = (float)s_box[3];
        }
    }
}

// ------------------------------------------------------------------------------------------------ RoIAlign
struct RoiArgs {
    const float* feat[PF_TRACK_LEVELS];
    int h[PF_TRACK_LEVELS], w[PF_TRACK_LEVELS];
    float scale[PF_TRACK_LEVELS];
    const float* rois;
    int K;
    uint16_t *hi, *lo;       // activation planes [Kpad * 64][256]
    float* roi_feats;        // optional fp32 [K][256][7][7]
};

// mmcv / torchvision bilinear_interpolate of RoIAlign (zero outside [-1, size], clamped inside)
__device__ __forceinline__ void roi_taps(float y, float x, int H, int W, int (&idx)[4], float (&wt)[4]) {
    if (y < -1.f || y > (float)H || x < -1.f || x > (float)W) {
        idx[0] = idx[1] = idx[2] = idx[3] = 0;
        wt[0] = wt[1] = wt[2] = wt[3] = 0.f;
        return;
    }
    y = fmaxf(y, 0.f), x = fmaxf(x, 0.f);
    int yl = (int)y, xl = (int)x, yh, xh;
    if (yl >= H - 1) yh = yl = H - 1, y = (float)yl;
    else yh = yl + 1;
    if (xl >= W - 1) xh = xl = W - 1, x = (float)xl;
    else xh = xl + 1;
    const float ly = y - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
    idx[0] = yl * W + xl, idx[1] = yl * W + xh, idx[2] = yh * W + xl, idx[3] = yh * W + xh;
    wt[0] = hy * hx, wt[1] = hy * lx, wt[2] = ly * hx, wt[3] = ly * lx;
}

// grid (8 grid rows, Kpad items), 256 threads = channels.  Row 7 of an item and items >= K are zero padding.
__global__ void __launch_bounds__(256) roi_align_kernel(const __grid_constant__ RoiArgs a) {
    const int k = blockIdx.y, ph = blockIdx.x, c = threadIdx.x;
    const size_t row0 = (size_t)k * TR_RP + ph * TR_PITCH;
    if (k >= a.K || ph >= TR_GRID) {
        for (int pw = 0; pw < TR_PITCH; ++pw) a.hi[(row0 + pw) * 256 + c] = 0, a.lo[(row0 + pw) * 256 + c] = 0;
        return;
    }
    const float* r = a.rois + k * 5;
    const float x1 = r[1], y1 = r[2], x2 = r[3], y2 = r[4];
    // SingleRoIExtractor.map_roi_levels: floor(log2(sqrt(w h) / 56 + 1e-6)) clamped to the pyramid
    const float sc = sqrtf((x2 - x1) * (y2 - y1));
    int lvl = (int)floorf(log2f(sc / 56.f + 1e-6f));
    lvl = min(max(lvl, 0), PF_TRACK_LEVELS - 1);
    const float* __restrict__ f = a.feat[0];
    int H = a.h[0], W = a.w[0];
    float s = a.scale[0];
#pragma unroll
    for (int l = 1; l < PF_TRACK_LEVELS; ++l)
        if (lvl == l) f = a.feat[l], H = a.h[l], W = a.w[l], s = a.scale[l];
    const float sw = x1 * s - 0.5f, sh = y1 * s - 0.5f;              // aligned = True
    const float bw = (x2 * s - 0.5f - sw) / TR_GRID, bh = (y2 * s - 0.5f - sh) / TR_GRID;
    const float* fc = f + (size_t)c * H * W;
    for (int pw = 0; pw < TR_GRID; ++pw) {
        float acc = 0.f;
#pragma unroll
        for (int iy = 0; iy < 2; ++iy)
#pragma unroll
            for (int ix = 0; ix < 2; ++ix) {
                const float y = sh + ph * bh + (iy + 0.5f) * bh / 2.f, x = sw + pw * bw + (ix + 0.5f) * bw / 2.f;
                int idx[4];
                float wt[4];
                roi_taps(y, x, H, W, idx, wt);
                acc += wt[0] * __ldg(fc + idx[0]) + wt[1] * __ldg(fc + idx[1]) + wt[2] * __ldg(fc + idx[2]) + wt[3] * __ldg(fc + idx[3]);
            }
        acc *= 0.25f;
        const float h = bf16_round(acc);
        a.hi[(row0 + pw) * 256 + c] = (uint16_t)(__float_as_uint(h) >> 16);
        a.lo[(row0 + pw) * 256 + c] = (uint16_t)(__float_as_uint(bf16_round(acc - h)) >> 16);
        if (a.roi_feats) a.roi_feats[(((size_t)k * 256 + c) * TR_GRID + ph) * TR_GRID + pw] = acc;
    }
    a.hi[(row0 + TR_GRID) * 256 + c] = 0, a.lo[(row0 + TR_GRID) * 256 + c] = 0;   // the zero column of the grid
}

// fp32 [K][256][7][7] (the output of an external RoI extractor) -> the same activation planes; grid as roi_align_kernel
__global__ void __launch_bounds__(256) roi_feats_to_planes_kernel(const float* __restrict__ x, int K, uint16_t* __restrict__ hi,
                                                                  uint16_t* __restrict__ lo) {
    const int k = blockIdx.y, ph = blockIdx.x, c = threadIdx.x;
    const size_t row0 = (size_t)k * TR_RP + ph * TR_PITCH;
    for (int pw = 0; pw < TR_PITCH; ++pw) {
        float v = 0.f;
        if (k < K && ph < TR_GRID && pw < TR_GRID) v = __ldg(x + (((size_t)k * 256 + c) * TR_GRID + ph) * TR_GRID + pw);
        const float h = bf16_round(v);
        hi[(row0 + pw) * 256 + c] = (uint16_t)(__float_as_uint(h) >> 16);
        lo[(row0 + pw) * 256 + c] = (uint16_t)(__float_as_uint(bf16_round(v - h)) >> 16);
    }
}

// ------------------------------------------------------------------------------------------------ FC tail
// one CTA per item: h = ReLU(b1 + sum over splits of the FC1 partials), embed = b2 + W2 h   (track_heads.py:97-101)
__global__ void __launch_bounds__(256) fc_tail_kernel(const float* __restrict__ part, int part_rows, const float* __restrict__ b1,
                                                      const float* __restrict__ w2t, const float* __restrict__ b2,
                                                      float* __restrict__ embeds) {
    pdl_wait();
    __shared__ float s_h[TR_FC1];
    const int k = blockIdx.x, t = threadIdx.x;
    for (int j = t; j < TR_FC1; j += 256) {
        float pv[TR_SPLITS];
#pragma unroll
        for (int s = 0; s < TR_SPLITS; ++s) pv[s] = part[((size_t)s * part_rows + k) * TR_FC1 + j];
        float v = __ldg(b1 + j);
#pragma unroll
        for (int s = 0; s < TR_SPLITS; ++s) v += pv[s];
        s_h[j] = fmaxf(v, 0.f);
    }
    __syncthreads();
    // 8 independent partial sums (fixed order: deterministic) keep 8 loads of the L2-resident W2^T in flight per thread
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int j = 0; j < TR_FC1; j += 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(s_h[j + i], __ldg(w2t + (size_t)(j + i) * EMB + t), acc[i]);
    }
    embeds[(size_t)k * EMB + t] = __ldg(b2 + t) + (((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7])));
}

// ------------------------------------------------------------------------------------------------ association
struct TrackerState {
    int next_id, n_tracks, cur, n_back_frames, back_head, overflow, pad[2];
    int back_n[MAXBF];
    int t_id[2][MAXT], t_label[2][MAXT], t_last[2][MAXT];
    int b_label[MAXBF][MAXK];
    float t_embed[2][MAXT][EMB];
    float b_embed[MAXBF][MAXK][EMB];
};

// mmdet bbox_overlaps(mode='iou', eps=1e-6)
__device__ __forceinline__ float box_iou(const float* a, const float* b) {
    const float aa = (a[2] - a[0]) * (a[3] - a[1]), ab = (b[2] - b[0]) * (b[3] - b[1]);
    const float w = fmaxf(fminf(a[2], b[2]) - fmaxf(a[0], b[0]), 0.f), h = fmaxf(fminf(a[3], b[3]) - fmaxf(a[1], b[1]), 0.f);
    const float ov = w * h;
    return ov / fmaxf(aa + ab - ov, 1e-6f);
}

constexpr int TM_THREADS = 512;
constexpr int TM_SC_CAP = 32768;      // floats of dynamic shared memory for the score matrix (128 KB)

__global__ void __launch_bounds__(TM_THREADS) tracker_match_kernel(const pf_tracker_config cfg, TrackerState* __restrict__ st,
                                                                   const float* __restrict__ bboxes,
                                                                   const int32_t* __restrict__ labels,
                                                                   const float* __restrict__ embeds, int K, int frame_id,
                                                                   int32_t* __restrict__ order_out, int32_t* __restrict__ ids_out,
                                                                   int32_t* __restrict__ n_kept_out, int32_t* __restrict__ status_out,
                                                                   float* __restrict__ scores) {
    __shared__ float s_score[MAXK], s_kscore[MAXK], s_kbox[MAXK][4], s_rmax[MAXK], s_rsum[MAXK];
    __shared__ int s_order[MAXK], s_valid[MAXK], s_kidx[MAXK], s_klabel[MAXK], s_ids[MAXK], s_kfound[MAXK], s_kdst[MAXK], s_bkeep[MAXK];
    __shared__ int s_mid[MAXM], s_mlabel[MAXM];
    __shared__ const float* s_mptr[MAXM];
    __shared__ float s_cmax[MAXM], s_csum[MAXM];
    __shared__ int s_tmatch[MAXT], s_tdst[MAXT];
    __shared__ int s_nk, s_M, s_ntracks_new, s_nback;
    __shared__ uint8_t s_taken[MAXM];
    extern __shared__ float s_sc[];      // TM_SC_CAP floats: the score matrix when it fits (it does unless K * memo > 32 k)
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int cur = st->cur, n_tracks = st->n_tracks;

    // ---- :139-142 sort by score, descending (stable)
    if (t < K) s_score[t] = bboxes[t * 5 + 4];
    __syncthreads();
    if (t < K) {
        int r = 0;
        const float v = s_score[t];
        for (int j = 0; j < K; ++j) r += (s_score[j] > v) || (s_score[j] == v && j < t);
        s_order[r] = t;
    }
    __syncthreads();
    // ---- :146-153 duplicate removal against every earlier box
    if (t < K) {
        const int o = s_order[t];
        const float thr = s_score[o] < cfg.obj_score_thr ? cfg.nms_backdrop_iou_thr : cfg.nms_class_iou_thr;
        int valid = 1;
        for (int j = 0; j < t; ++j)
            if (box_iou(bboxes + o * 5, bboxes + s_order[j] * 5) > thr) valid = 0;
        s_valid[t] = valid;
    }
    __syncthreads();
    if (t == 0) {
        int n = 0;
        for (int i = 0; i < K; ++i)
            if (s_valid[i]) s_kidx[n++] = s_order[i];
        s_nk = n;
        // memo: tracklets in insertion order, then the backdrops, most recent frame first (:103-135)
        int M = n_tracks;
        for (int f = 0; f < st->n_back_frames; ++f) M += st->back_n[(st->back_head + f) % MAXBF];
        s_M = M;
    }
    __syncthreads();
    const int nk = s_nk, M = s_M;
    float* sc = (nk * M <= TM_SC_CAP) ? s_sc : scores;      // row pitch M in shared memory, MAXM in the global workspace
    const int ld = (nk * M <= TM_SC_CAP) ? M : MAXM;
    for (int m = t; m < M; m += TM_THREADS) s_taken[m] = 0;
    if (t < nk) {
        const int o = s_kidx[t];
        s_kscore[t] = s_score[o], s_klabel[t] = labels[o], s_ids[t] = -1, s_kfound[t] = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) s_kbox[t][c] = bboxes[o * 5 + c];
    }
    for (int m = t; m < M; m += TM_THREADS) {
        if (m < n_tracks) {
            s_mid[m] = st->t_id[cur][m], s_mlabel[m] = st->t_label[cur][m], s_mptr[m] = st->t_embed[cur][m];
        } else {
            int r = m - n_tracks, f = 0, slot = st->back_head;
            while (r >= st->back_n[slot]) r -= st->back_n[slot], ++f, slot = (st->back_head + f) % MAXBF;
            s_mid[m] = -1, s_mlabel[m] = st->b_label[slot][r], s_mptr[m] = st->b_embed[slot][r];
        }
    }
    __syncthreads();

    if (nk > 0 && n_tracks > 0) {
        // ---- :166-170 feats = embeds memo^T, bi-softmax
        for (int idx = t; idx < nk * M; idx += TM_THREADS) {
            const int i = idx / M, m = idx - i * M;
            const float4* e = reinterpret_cast<const float4*>(embeds + (size_t)s_kidx[i] * EMB);
            const float4* q = reinterpret_cast<const float4*>(s_mptr[m]);
            float acc = 0.f;
            for (int c = 0; c < EMB / 4; ++c) {
                const float4 u = e[c], v = q[c];
                acc = fmaf(u.x, v.x, acc), acc = fmaf(u.y, v.y, acc), acc = fmaf(u.z, v.z, acc), acc = fmaf(u.w, v.w, acc);
            }
            sc[(size_t)i * ld + m] = acc;
        }
        __syncthreads();
        for (int i = warp; i < nk; i += TM_THREADS / 32) {          // softmax over the memo (dim = 1)
            float mx = -INFINITY;
            for (int m = lane; m < M; m += 32) mx = fmaxf(mx, sc[(size_t)i * ld + m]);
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sum = 0.f;
            for (int m = lane; m < M; m += 32) sum += expf(sc[(size_t)i * ld + m] - mx);
            sum = sg_warp_sum(sum);
            if (lane == 0) s_rmax[i] = mx, s_rsum[i] = sum;
        }
        for (int m = t; m < M; m += TM_THREADS) {                   // softmax over the detections (dim = 0)
            float mx = -INFINITY;
            for (int i = 0; i < nk; ++i) mx = fmaxf(mx, sc[(size_t)i * ld + m]);
            float sum = 0.f;
            for (int i = 0; i < nk; ++i) sum += expf(sc[(size_t)i * ld + m] - mx);
            s_cmax[m] = mx, s_csum[m] = sum;
        }
        __syncthreads();
        for (int idx = t; idx < nk * M; idx += TM_THREADS) {
            const int i = idx / M, m = idx - i * M;
            const float v = sc[(size_t)i * ld + m];
            float bs = (expf(v - s_rmax[i]) / s_rsum[i] + expf(v - s_cmax[m]) / s_csum[m]) / 2.f;
            if (cfg.with_cats && s_klabel[i] != s_mlabel[m]) bs = 0.f;          // :185-187
            sc[(size_t)i * ld + m] = bs;
        }
        __syncthreads();
        // ---- :189-201 greedy assignment in score order: sequential over the detections, so ONE warp runs it without block
        // barriers.  Zeroing column j in every other row (:196-197) = marking the memo entry taken: rows before i are done,
        // rows after i see 0 there.
        if (warp == 0) {
            for (int i = 0; i < nk; ++i) {
                float bv = -INFINITY;
                int bi = 0x7fffffff;
                for (int m = lane; m < M; m += 32) {
                    const float v = s_taken[m] ? 0.f : sc[(size_t)i * ld + m];
                    if (v > bv) bv = v, bi = m;
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi;     // torch.max: the first maximum
                }
                if (lane == 0 && bv > cfg.match_score_thr && s_mid[bi] > -1) {
                    if (s_kscore[i] > cfg.obj_score_thr) s_ids[i] = s_mid[bi], s_taken[bi] = 1;
                    else if (bv > cfg.nms_conf_thr) s_ids[i] = -2;
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }
    // ---- :202-208 new tracklets
    if (t == 0) {
        int next = st->next_id;
        for (int i = 0; i < nk; ++i)
            if (s_ids[i] == -1 && s_kscore[i] > cfg.init_score_thr) s_ids[i] = next++;
        st->next_id = next;
    }
    __syncthreads();
    // ---- update_memo :46-101, rebuilt into the other track buffer: survivors in order, then the new tracklets
    const int nxt = cur ^ 1;
    for (int tr = t; tr < n_tracks; tr += TM_THREADS) {
        const int id = st->t_id[cur][tr];
        int k = -1;
        for (int i = 0; i < nk; ++i)
            if (s_ids[i] == id) k = i;
        if (k >= 0) s_kfound[k] = 1;
        const int last = k >= 0 ? frame_id : st->t_last[cur][tr];
        s_tmatch[tr] = (frame_id - last >= cfg.memo_tracklet_frames) ? -2 : k;     // -2: popped (:92-98)
    }
    __syncthreads();
    if (t == 0) {
        int n = 0;
        for (int tr = 0; tr < n_tracks; ++tr) s_tdst[tr] = s_tmatch[tr] == -2 ? -1 : n++;
        int overflow = 0;
        for (int i = 0; i < nk; ++i) {
            s_kdst[i] = -1;
            if (s_ids[i] > -1 && !s_kfound[i] && cfg.memo_tracklet_frames > 0) {
                if (n < MAXT) s_kdst[i] = n++;
                else overflow = 1;
            }
        }
        s_ntracks_new = n;
        if (overflow) st->overflow = 1;
    }
    // backdrops :74-86: unmatched detections that do not overlap an earlier kept box
    if (t < nk) {
        int keep = s_ids[t] == -1;
        if (keep)
            for (int j = 0; j < t; ++j)
                if (box_iou(s_kbox[t], s_kbox[j]) > cfg.nms_backdrop_iou_thr) keep = 0;
        s_bkeep[t] = keep;
    }
    __syncthreads();
    const float mom = cfg.memo_momentum, omm = (float)(1.0 - (double)cfg.memo_momentum);
    for (int idx = t; idx < n_tracks * (EMB / 4); idx += TM_THREADS) {
        const int tr = idx / (EMB / 4), c = idx - tr * (EMB / 4), d = s_tdst[tr];
        if (d < 0) continue;
        float4 o = reinterpret_cast<const float4*>(st->t_embed[cur][tr])[c];
        const int k = s_tmatch[tr];
        if (k >= 0) {
            const float4 e = reinterpret_cast<const float4*>(embeds + (size_t)s_kidx[k] * EMB)[c];
            o.x = __fadd_rn(__fmul_rn(omm, o.x), __fmul_rn(mom, e.x)), o.y = __fadd_rn(__fmul_rn(omm, o.y), __fmul_rn(mom, e.y));
            o.z = __fadd_rn(__fmul_rn(omm, o.z), __fmul_rn(mom, e.z)), o.w = __fadd_rn(__fmul_rn(omm, o.w), __fmul_rn(mom, e.w));
        }
        reinterpret_cast<float4*>(st->t_embed[nxt][d])[c] = o;
        if (c == 0) {
            st->t_id[nxt][d] = st->t_id[cur][tr];
            st->t_label[nxt][d] = k >= 0 ? s_klabel[k] : st->t_label[cur][tr];
            st->t_last[nxt][d] = k >= 0 ? frame_id : st->t_last[cur][tr];
        }
    }
    for (int idx = t; idx < nk * (EMB / 4); idx += TM_THREADS) {
        const int i = idx / (EMB / 4), c = idx - i * (EMB / 4), d = s_kdst[i];
        if (d < 0) continue;
        reinterpret_cast<float4*>(st->t_embed[nxt][d])[c] = reinterpret_cast<const float4*>(embeds + (size_t)s_kidx[i] * EMB)[c];
        if (c == 0) st->t_id[nxt][d] = s_ids[i], st->t_label[nxt][d] = s_klabel[i], st->t_last[nxt][d] = frame_id;
    }
    // the new backdrop frame goes in front (:81-86); the oldest one falls out (:100-101)
    const int slot = (st->back_head + MAXBF - 1) % MAXBF;
    if (t == 0) {
        int n = 0;
        for (int i = 0; i < nk; ++i) s_bkeep[i] = s_bkeep[i] ? n++ : -1;
        s_nback = n;
    }
    __syncthreads();
    for (int idx = t; idx < nk * (EMB / 4); idx += TM_THREADS) {
        const int i = idx / (EMB / 4), c = idx - i * (EMB / 4), d = s_bkeep[i];
        if (d < 0) continue;
        reinterpret_cast<float4*>(st->b_embed[slot][d])[c] = reinterpret_cast<const float4*>(embeds + (size_t)s_kidx[i] * EMB)[c];
        if (c == 0) st->b_label[slot][d] = s_klabel[i];
    }
    if (t < nk) order_out[t] = s_kidx[t], ids_out[t] = s_ids[t];
    __syncthreads();
    if (t == 0) {
        st->back_n[slot] = s_nback;
        st->back_head = slot;
        st->n_back_frames = min(st->n_back_frames + 1, cfg.memo_backdrop_frames);
        st->cur = nxt;
        st->n_tracks = s_ntracks_new;
        n_kept_out[0] = nk;
        if (status_out) status_out[0] = st->overflow;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) track_paint_kernel(const int32_t* __restrict__ pan, const uint8_t* __restrict__ sem_lut,
                                                          const int32_t* __restrict__ track_lut, int n, uint8_t* __restrict__ sem,
                                                          T* __restrict__ track) {
    __shared__ uint8_t s_sem[256];
    __shared__ int32_t s_trk[256];
    s_sem[threadIdx.x] = sem_lut[threadIdx.x], s_trk[threadIdx.x] = track_lut[threadIdx.x];
    __syncthreads();
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const int id = __ldg(pan + i) & 255;
        sem[i] = s_sem[id], track[i] = static_cast<T>(s_trk[id]);
    }
}

static size_t al256(size_t v) { return (v + 255) / 256 * 256; }

static int boxes_common(bool from_ids, const float* masks, const int32_t* pan, const int32_t* seg_ids, int K, int H, int W,
                        float* rois, float* tight, void* ws, size_t ws_bytes, void* stream) {
    if (int e = check_device()) return e;
    PF_REQUIRE(K > 0 && K <= MAXK && H > 0 && W > 0, PF_ERR_ARG, "pf_track_boxes: K=%d H=%d W=%d", K, H, W);
    PF_REQUIRE(rois && tight && ws && (from_ids ? (pan && seg_ids) : masks != nullptr), PF_ERR_ARG, "pf_track_boxes: null pointer");
    PF_REQUIRE(ws_bytes >= pf_track_boxes_workspace_bytes(K, H, W), PF_ERR_WORKSPACE, "pf_track_boxes: workspace %zu < %zu", ws_bytes,
               pf_track_boxes_workspace_bytes(K, H, W));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int* hx = static_cast<int*>(ws);
    int* hy = hx + (size_t)K * W;
    cudaError_t e = cudaMemsetAsync(ws, 0, (size_t)K * (H + W) * sizeof(int), st);
    if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "pf_track_boxes memset: %s", cudaGetErrorString(e));
    dim3 grid((W + BX_COLS - 1) / BX_COLS, (H + BX_ROWS - 1) / BX_ROWS, from_ids ? 1 : K);
    if (from_ids) box_hist_kernel<true><<<grid, BX_COLS, 0, st>>>(nullptr, pan, seg_ids, K, H, W, hx, hy);
    else box_hist_kernel<false><<<grid, BX_COLS, 0, st>>>(masks, nullptr, nullptr, K, H, W, hx, hy);
    PF_CHECK_LAUNCH("box_hist_kernel");
    box_finalize_kernel<<<K, 256, 0, st>>>(hx, hy, H, W, rois, tight);
    PF_CHECK_LAUNCH("box_finalize_kernel");
    return PF_OK;
}

struct EmbedScratch {
    uint16_t* act[2][2];   // [ping-pong][hi, lo] planes [Kpad * 64][256]
    float* part;           // [TR_SPLITS][Kpad][1024]
    size_t total;
};
static EmbedScratch carve_embed(void* base, int K) {
    EmbedScratch s;
    const int Kpad = (K + 1) / 2 * 2;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? static_cast<char*>(base) + off : nullptr;
        off = al256(off + bytes);
        return p;
    };
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) s.act[i][j] = static_cast<uint16_t*>(take((size_t)Kpad * TR_RP * 256 * 2));
    s.part = static_cast<float*>(take((size_t)TR_SPLITS * Kpad * TR_FC1 * 4));
    s.total = off;
    return s;
}

template <int EPI>
static int launch_sgemm(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& w, const SgArgs& a, dim3 grid,
                        cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(sgemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SG_SMEM);
    if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "sgemm smem attribute: %s", cudaGetErrorString(e));
    return launch_pdl("sgemm_kernel", sgemm_kernel<EPI>, grid, dim3(SG_THREADS), SG_SMEM, st, ah, al, w, a);
}

// QuasiDenseMaskEmbedHeadGTMask.forward (track_heads.py:92-102) on the activation planes sc.act[0]
static int run_track_head(const pf_track_weights* w, const EmbedScratch& sc, int K, float* embeds, cudaStream_t st) {
    const int Kpad = (K + 1) / 2 * 2;
    // ---- 4 x [3x3 conv + GN32 + ReLU]: tap (ky, kx) = row shift (ky - 1) * 8 + (kx - 1)
    const uint64_t rows = (uint64_t)Kpad * TR_RP;
    for (int l = 0; l < 4; ++l) {
        CUtensorMap ah, al, wm;
        const uint64_t adims[3] = {256, rows, 1}, astr[2] = {512, rows * 512};
        const uint32_t abox[3] = {SG_KC, 128, 1};
        if (int e = make_tmap_bf16_nd(&ah, sc.act[l & 1][0], 3, adims, astr, abox)) return e;
        if (int e = make_tmap_bf16_nd(&al, sc.act[l & 1][1], 3, adims, astr, abox)) return e;
        if (int e = make_tmap_bf16_2d(&wm, w->conv_w + (size_t)l * 2 * 9 * 256 * 256, 2 * 9 * 256, 256, 256, 128, SG_KC)) return e;
        SgArgs a = {};
        a.mode = SG_CONV, a.n_kb = 9 * 4, a.cin_blocks = 4, a.w_tap_rows = 256, a.w_lo = 9 * 256;
        a.rows_per_img = (int)rows, a.planes_per_img = 1;
        for (int tp = 0; tp < 9; ++tp) a.shift[tp] = (tp / 3 - 1) * TR_PITCH + (tp % 3 - 1);
        a.n_items = K, a.gamma = w->gn_gamma + l * 256, a.beta = w->gn_beta + l * 256, a.eps = w->gn_eps;
        a.out_hi = sc.act[(l + 1) & 1][0], a.out_lo = sc.act[(l + 1) & 1][1];
        if (int e = launch_sgemm<SG_EPI_GN64>(ah, al, wm, a, dim3(2, Kpad / 2, 1), st)) return e;
    }
    // ---- FC 12544 -> 1024 (split-K over the 7 grid rows), then the tail
    {
        CUtensorMap ah, al, wm;
        const uint64_t adims[3] = {256, TR_RP, (uint64_t)Kpad}, astr[2] = {512, (uint64_t)TR_RP * 512};
        const uint32_t abox[3] = {SG_KC, 1, 128};
        if (int e = make_tmap_bf16_nd(&ah, sc.act[0][0], 3, adims, astr, abox)) return e;
        if (int e = make_tmap_bf16_nd(&al, sc.act[0][1], 3, adims, astr, abox)) return e;
        if (int e = make_tmap_bf16_2d(&wm, w->fc1_w, 2 * TR_FC1, 49 * 256, 49 * 256, 128, SG_KC)) return e;
        SgArgs a = {};
        a.mode = SG_FC, a.n_kb = TR_GRID * 4, a.fc_pos_per_split = TR_GRID, a.fc_grid = TR_GRID, a.fc_pitch = TR_PITCH, a.w_lo = TR_FC1;
        a.partial = sc.part, a.part_rows = Kpad, a.part_ld = TR_FC1;
        if (int e = launch_sgemm<SG_EPI_PARTIAL>(ah, al, wm, a, dim3(TR_FC1 / SG_TN, 1, TR_SPLITS), st)) return e;
    }
    return launch_pdl("fc_tail_kernel", fc_tail_kernel, dim3(K), dim3(256), 0, st, (const float*)sc.part, Kpad, w->fc1_b, w->fc2_wt,
                      w->fc2_b, embeds);
}

}  // namespace pf

extern "C" size_t pf_track_boxes_workspace_bytes(int K, int H, int W) {
    if (K <= 0 || H <= 0 || W <= 0) return 0;
    return pf::al256((size_t)K * (H + W) * sizeof(int));
}

extern "C" int pf_track_boxes_from_masks(const float* masks, int K, int H, int W, float* rois, float* tight, void* workspace,
                                         size_t workspace_bytes, void* stream) {
    return pf::boxes_common(false, masks, nullptr, nullptr, K, H, W, rois, tight, workspace, workspace_bytes, stream);
}

extern "C" int pf_track_boxes_from_panoptic(const int32_t* panoptic, const int32_t* seg_ids, int K, int H, int W, float* rois,
                                            float* tight, void* workspace, size_t workspace_bytes, void* stream) {
    return pf::boxes_common(true, nullptr, panoptic, seg_ids, K, H, W, rois, tight, workspace, workspace_bytes, stream);
}

extern "C" size_t pf_track_embed_workspace_bytes(int K) {
    if (K <= 0 || K > PF_TRACK_MAX_K) return 0;
    return pf::carve_embed(nullptr, K).total;
}

extern "C" int pf_track_embed(const pf_track_weights* w, const float* const* feats_host, const int* feat_h_host,
                              const int* feat_w_host, const int* strides_host, const float* rois, int K, float* embeds,
                              float* roi_feats, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(w && feats_host && feat_h_host && feat_w_host && strides_host && rois && embeds && workspace, PF_ERR_ARG,
               "pf_track_embed: null pointer");
    PF_REQUIRE(K > 0 && K <= MAXK, PF_ERR_ARG, "pf_track_embed: K=%d (1..%d)", K, MAXK);
    PF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PF_ERR_ALIGN, "pf_track_embed: workspace not 256-byte aligned");
    const EmbedScratch sc = carve_embed(workspace, K);
    PF_REQUIRE(workspace_bytes >= sc.total, PF_ERR_WORKSPACE, "pf_track_embed: workspace %zu < %zu", workspace_bytes, sc.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int Kpad = (K + 1) / 2 * 2;

    RoiArgs ra;
    for (int l = 0; l < PF_TRACK_LEVELS; ++l) {
        PF_REQUIRE(feats_host[l] && feat_h_host[l] > 0 && feat_w_host[l] > 0 && strides_host[l] > 0, PF_ERR_ARG,
                   "pf_track_embed: level %d", l);
        ra.feat[l] = feats_host[l], ra.h[l] = feat_h_host[l], ra.w[l] = feat_w_host[l], ra.scale[l] = 1.f / (float)strides_host[l];
    }
    ra.rois = rois, ra.K = K, ra.hi = sc.act[0][0], ra.lo = sc.act[0][1], ra.roi_feats = roi_feats;
    roi_align_kernel<<<dim3(TR_RP / TR_PITCH, Kpad), 256, 0, st>>>(ra);
    PF_CHECK_LAUNCH("roi_align_kernel");

    return run_track_head(w, sc, K, embeds, st);
}

extern "C" int pf_track_head(const pf_track_weights* w, const float* roi_feats, int K, float* embeds, void* workspace,
                             size_t workspace_bytes, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(w && roi_feats && embeds && workspace, PF_ERR_ARG, "pf_track_head: null pointer");
    PF_REQUIRE(K > 0 && K <= MAXK, PF_ERR_ARG, "pf_track_head: K=%d (1..%d)", K, MAXK);
    PF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PF_ERR_ALIGN, "pf_track_head: workspace not 256-byte aligned");
    const EmbedScratch sc = carve_embed(workspace, K);
    PF_REQUIRE(workspace_bytes >= sc.total, PF_ERR_WORKSPACE, "pf_track_head: workspace %zu < %zu", workspace_bytes, sc.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int Kpad = (K + 1) / 2 * 2;
    roi_feats_to_planes_kernel<<<dim3(TR_RP / TR_PITCH, Kpad), 256, 0, st>>>(roi_feats, K, sc.act[0][0], sc.act[0][1]);
    PF_CHECK_LAUNCH("roi_feats_to_planes_kernel");
    return run_track_head(w, sc, K, embeds, st);
}

extern "C" size_t pf_tracker_state_bytes(void) { return pf::al256(sizeof(pf::TrackerState)); }
extern "C" size_t pf_tracker_workspace_bytes(void) { return (size_t)pf::MAXK * pf::MAXM * sizeof(float); }

extern "C" int pf_tracker_reset(void* state, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(state, PF_ERR_ARG, "pf_tracker_reset: null state");
    // the header (counters, back_n) is all that has to be zero
    cudaError_t e = cudaMemsetAsync(state, 0, offsetof(TrackerState, t_id), static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "pf_tracker_reset: %s", cudaGetErrorString(e));
    return PF_OK;
}

extern "C" int pf_tracker_match(const pf_tracker_config* cfg, void* state, const float* bboxes, const int32_t* labels,
                                const float* embeds, int K, int frame_id, int32_t* order, int32_t* ids, int32_t* n_kept,
                                int32_t* status_out, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(cfg && state && order && ids && n_kept && workspace, PF_ERR_ARG, "pf_tracker_match: null pointer");
    PF_REQUIRE(K >= 0 && K <= MAXK && (K == 0 || (bboxes && labels && embeds)), PF_ERR_ARG, "pf_tracker_match: K=%d", K);
    PF_REQUIRE(K == 0 || (reinterpret_cast<uintptr_t>(embeds) & 15) == 0, PF_ERR_ALIGN, "pf_tracker_match: embeds not 16-byte aligned");
    PF_REQUIRE(cfg->memo_backdrop_frames >= 0 && cfg->memo_backdrop_frames <= MAXBF && cfg->memo_tracklet_frames >= 0, PF_ERR_ARG,
               "pf_tracker_match: memo_backdrop_frames=%d (0..%d)", cfg->memo_backdrop_frames, MAXBF);
    PF_REQUIRE(workspace_bytes >= pf_tracker_workspace_bytes(), PF_ERR_WORKSPACE, "pf_tracker_match: workspace %zu < %zu",
               workspace_bytes, pf_tracker_workspace_bytes());
    cudaError_t ea = cudaFuncSetAttribute(tracker_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SC_CAP * 4);
    if (ea != cudaSuccess) return set_error(PF_ERR_CUDA, "tracker smem attribute: %s", cudaGetErrorString(ea));
    tracker_match_kernel<<<1, TM_THREADS, TM_SC_CAP * 4, static_cast<cudaStream_t>(stream)>>>(*cfg, static_cast<TrackerState*>(state), bboxes,
                                                                                  labels, embeds, K, frame_id, order, ids, n_kept,
                                                                                  status_out, static_cast<float*>(workspace));
    PF_CHECK_LAUNCH("tracker_match_kernel");
    return PF_OK;
}

extern "C" int pf_track_paint(const int32_t* panoptic, const uint8_t* sem_lut, const int32_t* track_lut, int n_pixels,
                              uint8_t* sem, void* track, int track_f64, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    PF_REQUIRE(panoptic && sem_lut && track_lut && sem && track && n_pixels > 0, PF_ERR_ARG, "pf_track_paint: bad argument");
    const int blocks = min((n_pixels + 255) / 256, num_sms() * 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (track_f64) track_paint_kernel<double><<<blocks, 256, 0, st>>>(panoptic, sem_lut, track_lut, n_pixels, sem, static_cast<double*>(track));
    else track_paint_kernel<int32_t><<<blocks, 256, 0, st>>>(panoptic, sem_lut, track_lut, n_pixels, sem, static_cast<int32_t*>(track));
    PF_CHECK_LAUNCH("track_paint_kernel");
    return PF_OK;
}
