/* pf_decoder.h -- C ABI of the B200-native PolyphonicFormer decoder (libpf_decoder.so).
 *
 * The reference (HarborYuan/PolyphonicFormer) is pure Python on top of mmcv/mmdet and has no FFI of its own; the
 * hot path is the body of these Python methods (paths relative to the reference tree):
 *
 *   KernelUpdateIterHead.simple_test / _mask_forward   polyphonic/kernel_update.py:282-354, :125-157
 *   KernelUpdateHead.forward                           polyphonic/kernel_update_head.py:212-353
 *   KernelUpdator.forward                              polyphonic/funcs/kernel_updator.py:55-93
 *   KernelHead initial pooling                         polyphonic/kernel_head.py:313-320
 *
 * Each entry point below replaces a contiguous slice of those bodies (cited per function).  The binding a reference
 * maintainer would add is a ctypes stub inside those methods -- see INTEGRATION.md; the in-tree host mirror is
 * polyphonicformer_b200/_cabi.py + decoder.py.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller (PyTorch host code) owns all
 *     buffers; the library never allocates, frees or synchronises;
 *   - `stream` is a cudaStream_t passed as void*; every launch goes to that stream only, so calls can be captured
 *     into a CUDA graph;
 *   - returns PF_OK (0) or a negative pf_status; pf_last_error_string() describes the last failure on the calling
 *     thread.  No exceptions cross the boundary.  There is no CPU fallback: on a machine without an sm_100 device
 *     every compute entry point returns PF_ERR_ARCH;
 *   - fixed by the model family (configs/_base_/models/polyphonic_former.py): C = 256 channels, 8 heads of 32,
 *     N <= 128 kernels (111 at inference), num_classes <= 32.
 *
 * Device data layouts
 *   feats   bf16  [2][B][256][HWp]   branch 0 = x_feats, branch 1 = depth_feats; HWp = row pitch in elements,
 *                                    HWp % 8 == 0, HWp >= HW; columns >= HW are never read as data
 *   bits    u32   [B][WORDS][128]    WORDS = ceil(HW/32); bit j of word w of row n = (mask logit[n][32w+j] > 0)
 *                                    == sigmoid(logit) > 0.5 (kernel_update_head.py:236-238); rows >= N are zero
 *   partial f32   [G][S][64][N][4]   per-split pooled sums, G = n_branch*B units (unit = branch*B + b); element (n, c) of a
 *                                    split lives at [c / 4][n][c % 4] (column-group-major: a warp of 32 kernel rows reads /
 *                                    writes 512 contiguous bytes).  Opaque to callers: produced by pf_mask_pool, consumed by
 *                                    pf_kernel_update / pf_pool_reduce / pf_init_proposals
 *   cntp    f32   [G][S][N]          per-split mask pixel counts
 *   kern    f32   [G][N][256]        dynamic 1x1-conv kernels with feat_transform already folded in
 *   kern_split bf16 [G][2][N][256]   the same kernels as bf16 hi (= bf16(x)) and lo (= bf16(x - hi)) planes: the
 *                                    tensor-core einsum runs hi and lo as two MMAs so the product keeps ~16 mantissa bits
 *   kbias   f32   [G][N]             per-kernel logit bias produced by that fold
 *   logits  f32   [G][N][HW]         unit-major: branch 0 = new mask logits, branch 1 = new depth logits
 */
#ifndef PF_DECODER_H
#define PF_DECODER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define PF_C 256
#define PF_MAX_N 128
#define PF_MAX_CLASSES 32
#define PF_HEADS 8

typedef enum pf_status {
    PF_OK = 0,
    PF_ERR_ARG = -1,       /* bad shape / null pointer */
    PF_ERR_ALIGN = -2,     /* pointer or pitch alignment */
    PF_ERR_CUDA = -3,      /* CUDA runtime / driver error (message has the code) */
    PF_ERR_ARCH = -4,      /* no sm_100 device */
    PF_ERR_WORKSPACE = -5  /* workspace too small */
} pf_status;

int pf_version(void);
const char* pf_last_error_string(void);

/* ---- parameters of one KernelUpdateHead after host-side packing (decoder.py: PackedStage).
 * Weight MATRICES live in two bf16 "stacks" (row-major, one matrix after the other):
 *   wstack256  [rows][256]   every matrix with 256 input features,
 *   wstack_ffn [rows][FFN]   ffn.layers.1 (256 x FFN),
 * each matrix [out][in] (nn.Linear layout) stored as a hi plane (bf16(w)) followed by a lo plane (bf16(w - hi)), out
 * padded to a multiple of 128 rows.  The tensor-core GEMMs run hi and lo as separate MMAs (fp32-level accuracy).  The
 * struct holds the ROW of the hi plane; the lo plane starts `padded out` rows later.  Vectors stay fp32 (each padded
 * to a multiple of 64 floats).
 * feat_transform W_t,b_t (kernel_update_head.py:224-226) never touches the feature map; it is folded (in fp64):
 *   dyn_w  = dynamic_layer.weight @ W_t,  dyn_cb = dynamic_layer.weight @ b_t   (pooled' = pooled W_t^T + count b_t)
 *   kern_w = W_t^T @ fc_{mask,depth}.weight, kern_b = W_t^T @ fc.bias,
 *   kbrow_w = (fc.weight^T @ b_t) as ONE weight row, kbrow_b[0] = fc.bias . b_t    (the per-kernel logit bias)
 * gate_w / gate_b are stored interleaved in blocks of 64 output rows: [input_gate 0..63 | update_gate 0..63 |
 * input_gate 64..127 | ...], so that one 128-column GEMM tile holds both gates of the same 64 features.
 */
typedef struct pf_branch_weights {
    int dyn_w;    /* (*) [512][256]  dynamic_layer, folded; kernel_updator.py:58 */
    int inp_w;    /*     [512][256]  input_layer; :64 */
    int gate_w;   /*     [512][256]  input_gate / update_gate interleaved by 64; :73-74 */
    int fc_w;     /*     [256][256]  fc_layer; :89 */
    int qkv_w;    /*     [768][256]  attn.in_proj; kernel_update_head.py:259-260 */
    int out_w;    /*     [256][256]  attn.out_proj */
    int ffn1_w;   /*     [FFN][256]  ffn.layers.0.0; :271-272 */
    int head_w;   /*     mask: [cls_fcs.0; mask_fcs.0] [512][256]; depth: depth_regs.0 [256][256]; :278-283 */
    int cls_w;    /*     mask branch only: fc_cls padded to [128][256]; :285 */
    int kern_w;   /* (*) [256][256]  fc_mask / fc_depth with the fold; :287-288 */
    int kbrow_w;  /* (*) [1][256] padded to [128][256] */
    int ffn2_w;   /*     [256][FFN]  ffn.layers.1 -- row in wstack_ffn */
    int head_relu;                                  /* 1 for the mask branch (mask_fcs has ReLU), 0 for depth_regs */
    const float *dyn_b, *dyn_cb;                    /* [512], (*) [512] (dyn_cb may be NULL: no fold) */
    const float *inp_b, *gate_b;                    /* [512], [512] (interleaved like gate_w) */
    const float *ln_input_norm_in, *ln_norm_in;     /* each [2][256] = gamma, beta; kernel_updator.py:75-77 */
    const float *ln_norm_out, *ln_input_norm_out;   /* :78-79 */
    const float *fc_b, *ln_fc_norm;                 /* :89-91 */
    const float *qkv_b, *out_b, *ln_attn;           /* attention_norm */
    const float *ffn1_b, *ffn2_b, *ln_ffn;          /* ffn_norm */
    const float *ln_head_a, *ln_head_b;             /* mask: cls_fcs.1, mask_fcs.1; depth: depth_regs.1, unused */
    const float *cls_b;                             /* [32] (padded) */
    const float *kern_b, *kbrow_b;                  /* (*) [256], (*) [1] */
} pf_branch_weights;

typedef struct pf_stage_weights {
    pf_branch_weights br[2];   /* 0 = mask branch, 1 = depth branch */
    const uint16_t* wstack256; /* bf16 [wstack256_rows][256] */
    const uint16_t* wstack_ffn;/* bf16 [wstack_ffn_rows][ffn_channels] */
    int wstack256_rows, wstack_ffn_rows;
    int ffn_channels;          /* 2048 */
    int num_classes;           /* 19 */
    const float* vec_slices;   /* optional: the vectors above re-gathered per (branch, 32-column slice) for the fused
                                * small-N kernel, pf_vec_slices_bytes() bytes written by pf_pack_vec_slices(); NULL: the
                                * kernel launch gathers them into its workspace first (one extra small launch per call) */
} pf_stage_weights;

/* gather the per-(branch, column slice) copies of a stage's fp32 vectors (biases, LayerNorm gamma | beta) that the fused
 * small-N kernel bulk-copies into shared memory: out = device buffer of pf_vec_slices_bytes() bytes, 16-byte aligned;
 * w->vec_slices itself is ignored.  Call once per set of weights and store `out` in pf_stage_weights.vec_slices. */
size_t pf_vec_slices_bytes(void);
int pf_pack_vec_slices(const pf_stage_weights* w_host, float* out, void* stream);

/* fp32 NCHW feature maps -> the bf16 [2][B][256][HWp] layout.  Replaces nothing in the reference (storage cast). */
int pf_cast_feats(const float* x_feats, const float* depth_feats, uint16_t* feats, int B, int HW, int HWp,
                  void* stream);

/* kernel_update_head.py:236-238 (sigmoid > hard_mask_thr=0.5, .float()) as a packed bit mask. */
int pf_binarise(const float* mask_logits, uint32_t* bits, int B, int N, int HW, void* stream);

/* number of HW splits pf_mask_pool uses for this shape (size of the S dimension of partial / cntp) */
int pf_pool_splits(int B, int n_branch, int HW);

/* kernel_update_head.py:241-242 (and kernel_head.py:313-320): pooled[g][n][c] = sum_hw bit[b][n][hw] * feats[g][c][hw]
 * on tcgen05 tensor cores, split over S slabs of HW per unit; sum over S is taken by pf_kernel_update / pf_pool_reduce. */
int pf_mask_pool(const uint16_t* feats, const uint32_t* bits, float* partial, float* cntp, int B, int N, int HW,
                 int HWp, int n_branch, int S, void* stream);

/* sum the S partials: pooled [G][N][256], count [B][N] (used by KernelHead's init pooling and by tests) */
int pf_pool_reduce(const float* partial, const float* cntp, float* pooled, float* count, int B, int N, int n_branch,
                   int S, void* stream);

/* kernel_head.py:313-336 (the producer of the decoder's kernels): proposal_feats [B][P + n_stuff][256] =
 * [init_kernels [P][256] + pooled features of the P proposal masks ; stuff_kernels [n_stuff][256] for every image].
 * partial / cntp: output of pf_mask_pool(n_branch = 1, N = P) over x_feats and the sign bits of the initial masks. */
int pf_init_proposals(const float* partial, const float* cntp, const float* init_kernels, const float* stuff_kernels,
                      float* proposal_feats, int B, int P, int n_stuff, int S, void* stream);

/* bytes of scratch pf_kernel_update needs */
size_t pf_update_workspace_bytes(int B, int N, int ffn_channels);

/* kernel_update_head.py:245-288 + kernel_updator.py:55-93: both branches of the small-N block.
 *   obj_in / dep_in    [B][N][256]   proposal_feat, depth_proposal (before the "+ proposal_feat" of :250)
 *   obj_out / dep_out  [B][N][256]   obj_feat, depth_feat_new (post FFN+LN; next stage's inputs)
 *   cls_out            [B][N][num_classes]  (sigmoid applied iff cls_sigmoid != 0, kernel_update.py:333-334)
 *   kern (optional, may be NULL) / kern_split / kbias    folded dynamic kernels for pf_mask_einsum */
int pf_kernel_update(const pf_stage_weights* w_host, const float* partial, const float* cntp, int S,
                     const float* obj_in, const float* dep_in, float* obj_out, float* dep_out, float* cls_out,
                     float* kern, uint16_t* kern_split, float* kbias, void* workspace, size_t workspace_bytes, int B,
                     int N, int cls_sigmoid, void* stream);

/* pf_kernel_update has two implementations with the same results (tests/test_stage_fused_gpu.py):
 *   fused (default): ONE launch, an 8-CTA thread-block cluster per (image, branch) keeps the 128 x 256 activation block
 *                    on chip across the ten layers (csrc/pf_stage.cu);
 *   per-layer:       12 dependent launches (csrc/pf_update.cu), also what pf_kernel_updator uses.
 * Returns the previous setting.  Process-wide; meant for A/B measurements and tests. */
int pf_set_fused_update(int on);

/* fp32 kernels [n_units][N][256] -> kern_split (for callers that produce the dynamic kernels themselves) */
int pf_split_kernels(const float* kern, uint16_t* kern_split, int n_units, int N, void* stream);

/* KernelUpdator.forward alone (kernel_updator.py:55-93), for callers that use the module outside the stage:
 * `w->br[0]` holds the module's own (un-folded) dyn_w / dyn_b, inp_*, gate_*, the four gate LayerNorms, fc_*,
 * ln_fc_norm (rows in w->wstack256); other fields are ignored.  update_feature, input_feature, out: [R][256]. */
size_t pf_updator_workspace_bytes(int R);
int pf_kernel_updator(const pf_stage_weights* w_host, const float* update_feature, const float* input_feature,
                      float* out, void* workspace, size_t workspace_bytes, int R, void* stream);

/* kernel_update_head.py:308-334: logits[g][n][hw] = sum_c kern[g][n][c] * feats[g][c][hw] + kbias[g][n] on tcgen05.
 * n_units = B (mask branch only) or 2B.  logits and/or bits_out may be NULL (bits are taken from units < B). */
int pf_mask_einsum(const uint16_t* feats, const uint16_t* kern_split, const float* kbias, float* logits,
                   uint32_t* bits_out, int B, int N, int HW, int HWp, int n_units, void* stream);

/* kernel_update.py:133-143: F.interpolate(scale_factor=2, bilinear, align_corners=False) on `maps` [H][W] planes */
int pf_upsample2x(const float* in, float* out, int maps, int H, int W, void* stream);

/* bytes of scratch pf_decoder_forward needs */
size_t pf_decoder_workspace_bytes(int B, int N, int HW, int ffn_channels);

/* kernel_update.py:316-336: the whole stage loop of KernelUpdateIterHead.simple_test (no post-processing).
 *   stages_host   array of n_stages pf_stage_weights (host memory, device pointers inside)
 *   mask_logits   [B][N][H*W] initial mask logits (only their sign is used)
 *   obj / dep     [B][N][256] in: proposal_feats, depth_proposal; out: final object_feats, depth_proposal
 *   cls_out       [B][N][num_classes] sigmoid scores of the last stage
 *   logits_out    [2][B][N][H*W] last-stage mask and depth logits
 *   scaled_out    [2][B][N][2H*2W] their x2 upsampling, or NULL when mask_upsample_stride == 1
 *   flags         PF_FWD_* below */
#define PF_FWD_ALL_STAGE_OUTPUTS 1 /* also run the (unobservable) depth einsum + fp32 logits of stages 0..S-2 */
int pf_decoder_forward(const pf_stage_weights* stages_host, int n_stages, const uint16_t* feats,
                       const float* mask_logits, float* obj, float* dep, float* cls_out, float* logits_out,
                       float* scaled_out, void* workspace, size_t workspace_bytes, int B, int N, int H, int W, int HWp,
                       int flags, void* stream);

/* The same loop over the batch window [b0, b0+B) of full-batch tensors: feats [2][B_total][256][HWp], mask_logits /
 * obj / dep / cls_out [B_total][...], logits_out / scaled_out [2][B_total][N][..]; `workspace` sized for B images
 * (pf_decoder_workspace_bytes(B, ...)) and private to the call.  Lets the host decode disjoint windows of one batch
 * concurrently on several streams: the latency-bound small-N block of one window overlaps the HBM-bound pooling /
 * einsum of another (DecoderEngine.decode_inplace(splits=...)). */
int pf_decoder_forward_slice(const pf_stage_weights* stages_host, int n_stages, const uint16_t* feats,
                             const float* mask_logits, float* obj, float* dep, float* cls_out, float* logits_out,
                             float* scaled_out, void* workspace, size_t workspace_bytes, int B_total, int b0, int B,
                             int N, int H, int W, int HWp, int flags, void* stream);

/* ---- post-processing of one image ------------------------------------------------
 * KernelUpdateIterHead.get_panoptic + merge_stuff_thing_stuff_joint (polyphonic/kernel_update.py:421-535) with
 * rescale_masks / rescale_depth (polyphonic/kernel_update_head.py:593-626) and depth_act
 * (polyphonic/funcs/depth_utils.py:1-19) fused in: nothing full-resolution is materialised except the three results.
 *   cls_scores   [N][num_classes]  sigmoid class scores of the last stage
 *   mask_logits  [N][h][w]         scaled_mask_preds (1/4 of the padded network input)
 *   depth_logits [N][h][w]         scaled_depth_preds
 *   depth_init   [h][w]            the initial depth prediction at the same resolution (kernel_update.py:302-307)
 *   H0, W0                         img_shape = ori_shape: the top-left crop of the x4 up-sampled maps that is returned
 *   panoptic [H0][W0] int32, depth_final / depth_basic [H0][W0] fp32, segments [max_per_img + stuff] records in
 *   painting order, *n_segments their count (all DEVICE pointers).  depth_mode: 0 = 'monodepth', 1 = 'sigmoid'. */
typedef struct pf_segment {
    int id, isthing, category_id, instance_id, area;
    float score;
} pf_segment;
size_t pf_panoptic_workspace_bytes(int H0, int W0);
int pf_panoptic(const float* cls_scores, const float* mask_logits, const float* depth_logits, const float* depth_init,
                int N, int num_proposals, int num_thing_classes, int num_classes, int h, int w, int H0, int W0,
                int max_per_img, float instance_score_thr, float overlap_thr, int depth_mode, int32_t* panoptic,
                float* depth_final, float* depth_basic, pf_segment* segments, int* n_segments, void* workspace,
                size_t workspace_bytes, void* stream);

/* The same for B frames in one set of launches (frame f at offset f * (frame size) in every tensor; `segments` holds
 * `segment_stride` records per frame; `workspace` = B * pf_panoptic_workspace_bytes(H0, W0) bytes).
 * in_stride2 != 0: mask_logits / depth_logits / depth_init are the decoder's OWN maps [..][h/2][w/2] (stride 8) and the x2
 * bilinear up-sampling of kernel_update.py:131-143 / :302-307 is evaluated on the fly with pf_upsample2x's arithmetic, so
 * the 4x larger scaled_mask_preds / scaled_depth_preds are never written or read; results are bit-identical to
 * pf_upsample2x followed by pf_panoptic. */
int pf_panoptic_batch(const float* cls_scores, const float* mask_logits, const float* depth_logits, const float* depth_init,
                      int B, int N, int num_proposals, int num_thing_classes, int num_classes, int h, int w, int H0, int W0,
                      int max_per_img, float instance_score_thr, float overlap_thr, int depth_mode, int in_stride2,
                      int32_t* panoptic, float* depth_final, float* depth_basic, pf_segment* segments, int segment_stride,
                      int* n_segments, void* workspace, size_t workspace_bytes, void* stream);

/* ---- the producer of the decoder's inputs -----------------------------------------
 * The tail of KernelHead._decode_init_proposals (polyphonic/kernel_head.py:250-336) after SemanticFPN:
 *   loc / sem / dep = ReLU(GroupNorm32(conv1x1(maps[0 / 1 / 2])))       kernel_head.py:250-251, 264-265, 277-278
 *   mask_preds = init_kernels(loc)                                      :256      (rows 0 .. P-1 of head_w)
 *   seg_preds = conv_seg(sem), depth_pred = conv_direct_depth(dep)      :295, :285 (rows 112 .., row 144 of head_w)
 *   x_feats = sem + loc                                                 :303
 *   mask_preds = cat(mask_preds, seg_preds[:, num_thing_classes:])      :329-331 (cat_stuff_mask, eval)
 * pf_mask_pool(bits) + pf_init_proposals finish :313-336.  Static weights, packed once by the host:
 *   conv_split  bf16 [6][2][128][256]: block (half * 3 + map) = rows [128 half, 128 half + 128) of that map's
 *               {loc,seg,depth}_convs.0.conv.weight as hi / lo planes (hi = bf16(w), lo = bf16(w - hi))
 *   gn_gamma / gn_beta  fp32 [3][256]: {loc,seg,depth}_convs.0.gn.{weight,bias};  gn_eps = 1e-5
 *   head_w      bf16 [2][160][256] hi / lo planes: rows 0..111 init_kernels.weight (zero-padded), 112..143
 *               conv_seg.weight, 144..159 conv_direct_depth.weight;  head_b fp32 [160] the matching biases */
typedef struct pf_head_weights {
    const uint16_t* conv_split;
    const float* gn_gamma;
    const float* gn_beta;
    const uint16_t* head_w;
    const float* head_b;
    int num_proposals;      /* P = 100 */
    int num_classes;        /* 19 */
    int num_thing_classes;  /* 8 */
    float gn_eps;
} pf_head_weights;

/* fp32 [rows][HW] -> bf16 [rows][HWp] (pad columns zero): the storage cast of the SemanticFPN maps */
int pf_cast_maps(const float* maps, uint16_t* out, int rows, int HW, int HWp, void* stream);

size_t pf_kernel_head_workspace_bytes(int B, int HW);
/*   maps        bf16 [3][B][256][HWp]   localization_feats (loc, semantic, depth) in the storage dtype
 *   feats       out bf16 [2][B][256][HWp]   x_feats, depth_feats: the layout pf_decoder_forward consumes
 *   x32 / d32   optional fp32 [B][256][HW] copies of the same (may be NULL)
 *   mask_preds  out [B][P + num_classes - num_thing_classes][HW];  seg_preds out [B][num_classes][HW];
 *   depth_pred  out [B][HW];  bits optional out u32 [B][ceil(HW/32)][128]: sigmoid(mask_preds[:, :P]) > 0.5 (:314-317) */
int pf_kernel_head(const pf_head_weights* w, const uint16_t* maps, uint16_t* feats, float* x32, float* d32,
                   float* mask_preds, float* seg_preds, float* depth_pred, uint32_t* bits, void* workspace,
                   size_t workspace_bytes, int B, int HW, int HWp, void* stream);

/* ---- the last step of SemanticFPN (partial) ------------------------------------------
 * SemanticFPNWrapper.forward, polyphonic/funcs/semantic_fpn.py:221-229: conv_pred and the two aux_convs (mmcv
 * ConvModule 1x1 conv without bias + GN32 + ReLU, built :159-178) applied to the SAME fused map -- the three
 * localization_feats pf_kernel_head consumes.
 *   conv_split / gn_gamma / gn_beta   as in struct pf_head_weights, for (conv_pred, aux_convs.0, aux_convs.1)
 *   fused    bf16 [B][256][HWp]       feature_add_all_level in the storage dtype (pf_cast_maps)
 *   maps     out bf16 [3][B][256][HWp];  maps32 optional fp32 [3][B][256][HW] (may be NULL)
 *   workspace: pf_kernel_head_workspace_bytes(B, HW) bytes */
int pf_fpn_pred(const uint16_t* conv_split, const float* gn_gamma, const float* gn_beta, float gn_eps,
                const uint16_t* fused, uint16_t* maps, float* maps32, void* workspace, size_t workspace_bytes, int B,
                int HW, int HWp, void* stream);

/* debug only: int64 device buffer [16 + 16*capacity], zero-filled by the caller; CTA (0,0,0) of every GEMM launch of
 * the small-N block appends 16 %globaltimer samples (see scripts/k2_timeline.py).  NULL switches it off. */
int pf_debug_timeline(long long* device_buffer);

/* number of kernels the last call on this thread launched (for bench.py's gpu_launches) */
int pf_last_launch_count(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* PF_DECODER_H */
