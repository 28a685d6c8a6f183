/* pf_fpn.h -- C ABI of the SemanticFPN pyramid, part of libpf_decoder.so.  Conventions as in
 * pf_decoder.h: device pointers, caller-owned buffers, launches on `stream` only, no allocation, no synchronisation, PF_OK
 * or a negative pf_status, no CPU fallback.
 *
 * pf_semantic_fpn replaces SemanticFPNWrapper.forward up to `feature_add_all_level` (polyphonic/funcs/semantic_fpn.py:198-219
 * of the reference) in the shipped configuration (configs/_base_/models/polyphonic_former.py:78-96: levels 0..3,
 * upsample_times = 2, sine positional encoding added to level 3, no coordinate channels, sum fusion): seven
 * [3x3 conv (no bias) + GroupNorm(32) + ReLU] modules, the bilinear x2 steps between them and the four-level sum.  Its output
 * is the input of pf_fpn_pred (conv_pred + aux_convs, :221-229), whose outputs feed pf_kernel_head.
 */
#ifndef PF_FPN_H
#define PF_FPN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

/* Parameters after host-side packing (kernel_head.py: PackedSemanticFpn), conv order
 *   0 convs_all_levels.0.conv0 (stride 2)   1 convs_all_levels.1.conv0   2, 3 convs_all_levels.2.conv{0,1}
 *   4, 5, 6 convs_all_levels.3.conv{0,1,2}
 *   conv_w   bf16 [7][2 planes (hi, lo)][9 taps (ky*3+kx)][256 out][256 in]     <name>.conv.weight
 *   gn_gamma / gn_beta  fp32 [7][256]                                           <name>.gn.{weight,bias} */
typedef struct pf_fpn_weights {
    const uint16_t* conv_w;
    const float* gn_gamma;
    const float* gn_beta;
    float gn_eps;
} pf_fpn_weights;

/* H, W: the decoder map (stride 8 of the frame), both multiples of 4 */
size_t pf_semantic_fpn_workspace_bytes(int B, int H, int W);
/*   p0..p3   the FPN levels fp32 NCHW: [B][256][2H][2W], [B][256][H][W], [B][256][H/2][W/2], [B][256][H/4][W/4]
 *   fused    out bf16 [B][256][HWp] = feature_add_all_level in the storage dtype (pad columns zero)
 *   fused32  optional out fp32 [B][256][H*W] (may be NULL) */
int pf_semantic_fpn(const pf_fpn_weights* w, const float* p0, const float* p1, const float* p2, const float* p3,
                    uint16_t* fused, float* fused32, void* workspace, size_t workspace_bytes, int B, int H, int W, int HWp,
                    void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* PF_FPN_H */
