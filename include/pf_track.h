/* pf_track.h -- C ABI of the video model's tracking path (SURVEY.md section 8f rank 3, section 8e video mode), part of
 * libpf_decoder.so.  Conventions as in pf_decoder.h: device pointers, caller-owned buffers, launches on `stream` only,
 * no allocation, no synchronisation, PF_OK or a negative pf_status, no CPU fallback.
 *
 * What each entry point replaces in the reference (paths relative to the reference tree):
 *
 *   pf_track_boxes_*     batch_mask2boxlist / coords2bboxTensor   polyphonic/video/utils.py:40-82   (RoI boxes)
 *                        tensor_mask2box / coords2bbox_all        polyphonic/funcs/utils.py:4-22    (tracker boxes)
 *                        PolyphonicVideo.get_things_id_for_tracking  polyphonic/polyphonic_former_video.py:421-434
 *   pf_track_embed       bboxlist2roi + clamp + SingleRoIExtractor (mmcv RoIAlign 7x7, sampling_ratio 2, aligned) +
 *                        QuasiDenseMaskEmbedHeadGTMask.forward    polyphonic_former_video.py:408-419,
 *                                                                 polyphonic/video/track_heads.py:92-102
 *   pf_tracker_match     QuasiDenseEmbedTracker.match + update_memo + memo
 *                                                polyphonic/video/qdtrack/trackers/quasi_dense_embed_tracker.py:46-207
 *   pf_track_paint       generate_track_id_maps + get_semantic_seg    polyphonic_former_video.py:436-451
 */
#ifndef PF_TRACK_H
#define PF_TRACK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define PF_TRACK_MAX_K 128        /* detections per frame (max_per_img = 100 in the shipped configs) */
#define PF_TRACK_MAX_TRACKS 512   /* live tracklets */
#define PF_TRACK_MAX_BACK_FRAMES 4
#define PF_TRACK_EMBED 256
#define PF_TRACK_LEVELS 4

/* ---- mask -> box ---------------------------------------------------------------------------------------------------
 * Both variants build per-item column / row pixel histograms with integer atomics (deterministic) and derive from them
 *   rois  [K][5]  (0, cx - 2 dx, cy - 2 dy, cx + 2 dx, cy + 2 dy) clamped at 0: centre of the mask pixels +- 2 x their
 *                 mean absolute deviation (at least 1 pixel) per axis; an empty mask gives zeros   (video/utils.py:40-82,
 *                 polyphonic_former_video.py:412-415)
 *   tight [K][4]  (xmin, ymin, xmax, ymax); an empty mask gives (-1, -1, 10, 10)                   (funcs/utils.py:4-22)
 * workspace: pf_track_boxes_workspace_bytes(K, H, W), zero-filled by the call itself. */
size_t pf_track_boxes_workspace_bytes(int K, int H, int W);
/* masks fp32 [K][H][W], a pixel belongs to the mask iff its value is non-zero */
int pf_track_boxes_from_masks(const float* masks, int K, int H, int W, float* rois, float* tight, void* workspace,
                              size_t workspace_bytes, void* stream);
/* panoptic int32 [H][W] (pf_panoptic's map); item k is the set of pixels equal to seg_ids[k] (1 <= id <= 255) */
int pf_track_boxes_from_panoptic(const int32_t* panoptic, const int32_t* seg_ids, int K, int H, int W, float* rois,
                                 float* tight, void* workspace, size_t workspace_bytes, void* stream);

/* ---- RoI features + embedding head -----------------------------------------------------------------------------------
 * Parameters of QuasiDenseMaskEmbedHeadGTMask after host-side packing (track.py: PackedTrackHead):
 *   conv_w   bf16 [4 layers][2 planes (hi, lo)][9 taps (ky*3+kx)][256 out][256 in]   convs.{l}.conv.weight
 *   gn_gamma / gn_beta  fp32 [4][256]                                                convs.{l}.gn.{weight,bias}
 *   fc1_w    bf16 [2 planes][1024][49*256]  fcs.0.weight with the input index permuted from (c, y, x) to (y, x, c)
 *   fc1_b    fp32 [1024];  fc2_wt fp32 [1024][256] = fc_embed.weight transposed;  fc2_b fp32 [256] */
typedef struct pf_track_weights {
    const uint16_t* conv_w;
    const float* gn_gamma;
    const float* gn_beta;
    const uint16_t* fc1_w;
    const float* fc1_b;
    const float* fc2_wt;
    const float* fc2_b;
    float gn_eps;
} pf_track_weights;

size_t pf_track_embed_workspace_bytes(int K);
/*   feats    4 device pointers (host array): the FPN levels fp32 [256][feat_h[l]][feat_w[l]] of ONE image, strides[l] =
 *            their down-sampling factors (4, 8, 16, 32)
 *   rois     [K][5] as written by pf_track_boxes_*;  embeds out fp32 [K][256];  K <= PF_TRACK_MAX_K
 *   roi_feats optional out fp32 [K][256][7][7] (the RoIAlign result, for tests; may be NULL) */
int pf_track_embed(const pf_track_weights* w, const float* const* feats_host, const int* feat_h_host,
                   const int* feat_w_host, const int* strides_host, const float* rois, int K, float* embeds,
                   float* roi_feats, void* workspace, size_t workspace_bytes, void* stream);

/* the embedding head alone (QuasiDenseMaskEmbedHeadGTMask.forward, track_heads.py:92-102) on RoI features an external
 * extractor produced: roi_feats fp32 [K][256][7][7] -> embeds fp32 [K][256]; workspace as pf_track_embed */
int pf_track_head(const pf_track_weights* w, const float* roi_feats, int K, float* embeds, void* workspace,
                  size_t workspace_bytes, void* stream);

/* ---- association ----------------------------------------------------------------------------------------------------- */
typedef struct pf_tracker_config {
    float init_score_thr, obj_score_thr, match_score_thr, memo_momentum, nms_conf_thr, nms_backdrop_iou_thr,
        nms_class_iou_thr;
    int memo_tracklet_frames, memo_backdrop_frames, with_cats;
} pf_tracker_config;

size_t pf_tracker_state_bytes(void);
size_t pf_tracker_workspace_bytes(void);
/* empties the memo (QuasiDenseEmbedTracker.__init__ / PolyphonicVideo.init_tracker) */
int pf_tracker_reset(void* state, void* stream);
/* One frame of QuasiDenseEmbedTracker.match (bisoftmax metric).  bboxes [K][5] (x1, y1, x2, y2, score), labels int32 [K],
 * embeds [K][256] (16-byte aligned).  Outputs: n_kept[0] = detections that survive the duplicate removal; for i < n_kept: order[i] = index
 * of kept detection i in the inputs (descending score), ids[i] = its track id (>= 0), -1 (backdrop) or -2 (duplicate of
 * a confident track).  status_out[0] != 0 if the memo overflowed PF_TRACK_MAX_TRACKS (new tracks were dropped). */
int pf_tracker_match(const pf_tracker_config* cfg, void* state, const float* bboxes, const int32_t* labels,
                     const float* embeds, int K, int frame_id, int32_t* order, int32_t* ids, int32_t* n_kept,
                     int32_t* status_out, void* workspace, size_t workspace_bytes, void* stream);

/* sem[p] = sem_lut[panoptic[p]] (uint8), track[p] = track_lut[panoptic[p]]; luts have 256 entries; `track` is int32
 * [n_pixels], or float64 [n_pixels] when track_f64 != 0 (the reference's track map is np.zeros(shape): float64) */
int pf_track_paint(const int32_t* panoptic, const uint8_t* sem_lut, const int32_t* track_lut, int n_pixels, uint8_t* sem,
                   void* track, int track_f64, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* PF_TRACK_H */
