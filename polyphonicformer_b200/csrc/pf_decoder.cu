// The stage loop of KernelUpdateIterHead.simple_test (polyphonic/kernel_update.py:316-336 of the reference) as one
// C call: binarise -> S x (pool -> update -> einsum) -> x2 upsample.  Launch-only (no allocation, no sync), so the
// host may capture it into a CUDA graph.
#include "pf_internal.h"

namespace pf {
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct DecoderScratch {
    uint32_t* bits;
    float *partial, *cntp, *kbias;
    uint16_t* kern;
    void* update_ws;
    float *pp_obj[2], *pp_dep[2];   // ping-pong kernels between stages (the fused small-N kernel must not run in place)
    size_t update_ws_bytes, total;
};

static DecoderScratch carve(void* base, int B, int N, int HW, int ffn) {
    DecoderScratch s;
    const int words = (HW + 31) / 32;
    const int S = pf_pool_splits(B, 2, HW);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? static_cast<char*>(base) + off : nullptr;
        off = align_up(off + bytes, 256);
        return p;
    };
    s.bits = static_cast<uint32_t*>(take((size_t)B * words * 128 * 4));
    s.partial = static_cast<float*>(take((size_t)2 * B * S * N * PF_C * 4));
    s.cntp = static_cast<float*>(take((size_t)2 * B * S * N * 4));
    s.kern = static_cast<uint16_t*>(take((size_t)2 * B * 2 * N * PF_C * 2));
    s.kbias = static_cast<float*>(take((size_t)2 * B * N * 4));
    s.update_ws_bytes = pf_update_workspace_bytes(B, N, ffn);
    s.update_ws = take(s.update_ws_bytes);
    for (int i = 0; i < 2; ++i) {
        s.pp_obj[i] = static_cast<float*>(take((size_t)B * N * PF_C * 4));
        s.pp_dep[i] = static_cast<float*>(take((size_t)B * N * PF_C * 4));
    }
    s.total = off;
    return s;
}
}  // namespace pf

extern "C" size_t pf_decoder_workspace_bytes(int B, int N, int HW, int ffn_channels) {
    if (B <= 0 || N <= 0 || HW <= 0 || ffn_channels <= 0) return 0;
    return pf::carve(nullptr, B, N, HW, ffn_channels).total;
}

extern "C" int pf_decoder_forward(const pf_stage_weights* stages, int n_stages, const uint16_t* feats,
                                  const float* mask_logits, float* obj, float* dep, float* cls_out, float* logits_out,
                                  float* scaled_out, void* workspace, size_t workspace_bytes, int B, int N, int H, int W,
                                  int HWp, int flags, void* stream) {
    return pf_decoder_forward_slice(stages, n_stages, feats, mask_logits, obj, dep, cls_out, logits_out, scaled_out,
                                    workspace, workspace_bytes, B, 0, B, N, H, W, HWp, flags, stream);
}

extern "C" int pf_decoder_forward_slice(const pf_stage_weights* stages, int n_stages, const uint16_t* feats,
                                        const float* mask_logits, float* obj, float* dep, float* cls_out,
                                        float* logits_out, float* scaled_out, void* workspace, size_t workspace_bytes,
                                        int B_total, int b0, int B, int N, int H, int W, int HWp, int flags, void* stream) {
    using namespace pf;
    if (int e = check_device()) return e;
    reset_launch_count();
    PF_REQUIRE(stages && n_stages > 0 && feats && mask_logits && obj && dep && cls_out && logits_out && workspace,
               PF_ERR_ARG, "pf_decoder_forward: null pointer");
    PF_REQUIRE(B > 0 && N > 0 && N <= PF_MAX_N && H > 0 && W > 0, PF_ERR_ARG, "pf_decoder_forward: bad shape");
    PF_REQUIRE(b0 >= 0 && b0 + B <= B_total, PF_ERR_ARG, "pf_decoder_forward: bad batch window %d+%d of %d", b0, B, B_total);
    PF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PF_ERR_ALIGN, "pf_decoder_forward: workspace not 256-byte aligned");
    const int HW = H * W;
    const int ffn = stages[0].ffn_channels;
    const int ncls = stages[0].num_classes;
    const DecoderScratch s = carve(workspace, B, N, HW, ffn);
    PF_REQUIRE(workspace_bytes >= s.total, PF_ERR_WORKSPACE, "pf_decoder_forward: workspace %zu < %zu", workspace_bytes, s.total);
    const int S = pf_pool_splits(B, 2, HW);
    // batch-major tensors: plain offsets; [2][B_total] tensors (feats, logits, scaled): addressed through the window
    mask_logits += (size_t)b0 * N * HW;
    obj += (size_t)b0 * N * PF_C, dep += (size_t)b0 * N * PF_C, cls_out += (size_t)b0 * N * ncls;

    // Every kernel of the loop is launched with programmatic stream serialization.  binarise waits for everything
    // before this call (the producers of feats / mask_logits / obj / dep) BEFORE it releases its dependents, so the
    // pooling and einsum kernels may stream the feature maps ahead of their own grid dependency (early_feats = 1).
    if (int e = pf_binarise(mask_logits, s.bits, B, N, HW, stream)) return e;
    for (int st = 0; st < n_stages; ++st) {
        const bool last = st == n_stages - 1;
        if (int e = mask_pool_window(feats, s.bits, s.partial, s.cntp, B_total, b0, B, N, HW, HWp, 2, S, 1, stream)) return e;
        // stage st reads the kernels stage st-1 wrote and writes the other buffer; the last stage writes obj / dep
        const float* obj_i = st == 0 ? obj : s.pp_obj[(st - 1) & 1];
        const float* dep_i = st == 0 ? dep : s.pp_dep[(st - 1) & 1];
        float* obj_o = last ? obj : s.pp_obj[st & 1];
        float* dep_o = last ? dep : s.pp_dep[st & 1];
        if (int e = pf_kernel_update(&stages[st], s.partial, s.cntp, S, obj_i, dep_i, obj_o, dep_o, cls_out, nullptr, s.kern,
                                     s.kbias, s.update_ws, s.update_ws_bytes, B, N, last ? 1 : 0, stream))
            return e;
        int e = PF_OK;
        if (last && scaled_out && !(flags & PF_FWD_ALL_STAGE_OUTPUTS)) {
            // last stage, branch by branch: einsum -> x2 up-sampling while the 58 MB of logits are still L2-resident
            for (int br = 0; br < 2 && !e; ++br) {
                e = mask_einsum_window(feats, s.kern, s.kbias, logits_out, nullptr, B_total, b0, B, N, HW, HWp, B, br, 1, stream);
                const size_t u0 = ((size_t)br * B_total + b0) * N;
                if (!e) e = pf_upsample2x(logits_out + u0 * HW, scaled_out + u0 * 4 * HW, B * N, H, W, stream);
            }
            if (e) return e;
            return PF_OK;
        }
        if (last)
            e = mask_einsum_window(feats, s.kern, s.kbias, logits_out, nullptr, B_total, b0, B, N, HW, HWp, 2 * B, 0, 1, stream);
        else if (flags & PF_FWD_ALL_STAGE_OUTPUTS)
            e = mask_einsum_window(feats, s.kern, s.kbias, logits_out, s.bits, B_total, b0, B, N, HW, HWp, 2 * B, 0, 1, stream);
        else  // only the sign of the next mask is observable (kernel_update_head.py:236-238)
            e = mask_einsum_window(feats, s.kern, s.kbias, nullptr, s.bits, B_total, b0, B, N, HW, HWp, B, 0, 1, stream);
        if (e) return e;
    }
    if (scaled_out)
        for (int br = 0; br < 2; ++br) {
            const size_t u0 = ((size_t)br * B_total + b0) * N;
            if (int e = pf_upsample2x(logits_out + u0 * HW, scaled_out + u0 * 4 * HW, B * N, H, W, stream)) return e;
        }
    return PF_OK;
}
