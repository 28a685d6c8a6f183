// Host-side plumbing shared by the translation units of libpf_decoder.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pf_decoder.h"
#include "../../include/pf_track.h"
#include "../../include/pf_fpn.h"

namespace pf {

int set_error(int code, const char* fmt, ...);
int check_device();             // PF_OK iff the current device is sm_100
int num_sms();                  // SM count of the current device
void count_launch(int n = 1);   // per-thread launch counter (pf_last_launch_count)
void reset_launch_count();

// cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (no link-time libcuda dependency).
// 2-D bf16 tensor [rows][cols] with row pitch `pitch_elems`; box = [box_rows][box_cols]; 128-byte swizzle.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                      uint32_t box_rows, uint32_t box_cols);

// 3-D bf16 tensor [d2][d1][d0] (dense); box = [1][box1][box0]; rows >= d1 read as zero.
int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t d2, uint64_t d1, uint64_t d0, uint32_t box1,
                      uint32_t box0);

// bf16 "blocked" activation arena: 3-D tensor [blocks][128 rows][8] (dense), box = [8 blocks][128 rows][8], no swizzle:
// in shared memory the box is the canonical no-swizzle K-major UMMA operand ([8 column groups][128 rows][16 bytes])
int make_tmap_bf16_blocked(CUtensorMap* out, const void* base, uint64_t blocks);

// 3-D fp32 tensor [d2][d1][d0] (dense); box = [1][box1][box0], 128-byte swizzle (box0 * 4 bytes must be 128)
int make_tmap_f32_3d(CUtensorMap* out, const void* base, uint64_t d2, uint64_t d1, uint64_t d0, uint32_t box1,
                     uint32_t box0);

// rank-N bf16 tensor (dims / box innermost first, strides_bytes[i] = byte stride of dimension i + 1), 128-byte swizzle,
// out-of-range elements (negative coordinates included) read as zero
int make_tmap_bf16_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box);

// batch-window variants of the streaming kernels (pf_decoder_forward_slice): feats / logits are full-batch tensors
int mask_pool_window(const uint16_t* feats, const uint32_t* bits, float* partial, float* cntp, int Btot, int b0, int B,
                     int N, int HW, int HWp, int n_branch, int S, int early_feats, void* stream);
int mask_einsum_window(const uint16_t* feats, const uint16_t* kern, const float* kbias, float* logits, uint32_t* bits_out,
                       int Btot, int b0, int B, int N, int HW, int HWp, int n_units, int branch0, int early_feats,
                       void* stream);

// the three 1x1 convolutions of pf_kernel_head as one einsum launch (pf_einsum.cu)
// stats: [6B][conv1x1_ctas_per_unit][128] per-row (sum, sum of squares) partials for the GroupNorm that follows
int conv1x1_ctas_per_unit(int B, int HW);
int conv1x1_maps(const uint16_t* maps, int n_inputs, const uint16_t* conv_split, float* Y, float2* stats, int B, int HW,
                 int HWp, void* stream);

// the same convolutions without the fp32 intermediate: affine == null -> statistics only; else conv + (scale, shift) + ReLU ->
// bf16 out [3][B][256][HWp] (+ optional fp32 out32 [3][B][256][HW])
int conv1x1_fused(const uint16_t* maps, int n_inputs, const uint16_t* conv_split, float2* stats, const float2* affine,
                  uint16_t* out, float* out32, int B, int HW, int HWp, void* stream);

// Launch with programmatic stream serialization (PDL): the kernel may begin before its predecessor in the stream has
// finished; it must execute griddepcontrol.wait (pdl_wait) before touching anything the predecessor wrote.
template <typename... KArgs, typename... Args>
int launch_pdl(const char* name, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
    if (e != cudaSuccess) return set_error(PF_ERR_CUDA, "%s launch: %s", name, cudaGetErrorString(e));
    count_launch();
    return PF_OK;
}

#define PF_CHECK_LAUNCH(name)                                                          \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess)                                                        \
            return pf::set_error(PF_ERR_CUDA, "%s launch: %s", name, cudaGetErrorString(e__)); \
        pf::count_launch();                                                            \
    } while (0)

#define PF_REQUIRE(cond, code, ...) \
    do {                            \
        if (!(cond)) return pf::set_error(code, __VA_ARGS__); \
    } while (0)

}  // namespace pf
