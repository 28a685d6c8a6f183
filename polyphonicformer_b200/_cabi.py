"""ctypes binding of include/pf_decoder.h and include/pf_track.h (libpf_decoder.so).

This is the only place the package touches the native library.  There is NO fallback: if the library is missing or
a call fails, a ``PFError`` is raised -- nothing in this package computes the decoder with PyTorch ops.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'lib', 'libpf_decoder.so')

PF_C = 256
PF_MAX_N = 128
PF_MAX_CLASSES = 32
PF_HEADS = 8
PF_FWD_ALL_STAGE_OUTPUTS = 1

STATUS = {0: 'PF_OK', -1: 'PF_ERR_ARG', -2: 'PF_ERR_ALIGN', -3: 'PF_ERR_CUDA', -4: 'PF_ERR_ARCH',
          -5: 'PF_ERR_WORKSPACE'}


class PFError(RuntimeError):
    def __init__(self, code, fn, msg):
        super().__init__('%s failed with %s (%d): %s' % (fn, STATUS.get(code, '?'), code, msg))
        self.code = code


_fp = POINTER(c_float)


class BranchWeights(Structure):
    """struct pf_branch_weights -- field order must match include/pf_decoder.h."""
    _ROWS = ['dyn_w', 'inp_w', 'gate_w', 'fc_w', 'qkv_w', 'out_w', 'ffn1_w', 'head_w', 'cls_w', 'kern_w', 'kbrow_w',
             'ffn2_w']
    _PTRS = ['dyn_b', 'dyn_cb', 'inp_b', 'gate_b', 'ln_input_norm_in', 'ln_norm_in', 'ln_norm_out',
             'ln_input_norm_out', 'fc_b', 'ln_fc_norm', 'qkv_b', 'out_b', 'ln_attn', 'ffn1_b', 'ffn2_b', 'ln_ffn',
             'ln_head_a', 'ln_head_b', 'cls_b', 'kern_b', 'kbrow_b']
    _fields_ = [(n, c_int) for n in _ROWS] + [('head_relu', c_int)] + [(n, c_void_p) for n in _PTRS]


class StageWeights(Structure):
    """struct pf_stage_weights."""
    _fields_ = [('br', BranchWeights * 2), ('wstack256', c_void_p), ('wstack_ffn', c_void_p),
                ('wstack256_rows', c_int), ('wstack_ffn_rows', c_int), ('ffn_channels', c_int), ('num_classes', c_int),
                ('vec_slices', c_void_p)]


class HeadWeights(Structure):
    """struct pf_head_weights."""
    _fields_ = [('conv_split', c_void_p), ('gn_gamma', c_void_p), ('gn_beta', c_void_p), ('head_w', c_void_p),
                ('head_b', c_void_p), ('num_proposals', c_int), ('num_classes', c_int), ('num_thing_classes', c_int),
                ('gn_eps', c_float)]


class TrackWeights(Structure):
    """struct pf_track_weights (include/pf_track.h)."""
    _fields_ = [('conv_w', c_void_p), ('gn_gamma', c_void_p), ('gn_beta', c_void_p), ('fc1_w', c_void_p),
                ('fc1_b', c_void_p), ('fc2_wt', c_void_p), ('fc2_b', c_void_p), ('gn_eps', c_float)]


class FpnWeights(Structure):
    """struct pf_fpn_weights (include/pf_fpn.h)."""
    _fields_ = [('conv_w', c_void_p), ('gn_gamma', c_void_p), ('gn_beta', c_void_p), ('gn_eps', c_float)]


class TrackerConfig(Structure):
    """struct pf_tracker_config (include/pf_track.h)."""
    _FLOATS = ['init_score_thr', 'obj_score_thr', 'match_score_thr', 'memo_momentum', 'nms_conf_thr',
               'nms_backdrop_iou_thr', 'nms_class_iou_thr']
    _INTS = ['memo_tracklet_frames', 'memo_backdrop_frames', 'with_cats']
    _fields_ = [(n, c_float) for n in _FLOATS] + [(n, c_int) for n in _INTS]


_SIGS = {
    # ---- include/pf_track.h
    'pf_track_boxes_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'pf_track_boxes_from_masks': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'pf_track_boxes_from_panoptic': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                             c_void_p]),
    'pf_track_embed_workspace_bytes': (c_size_t, [c_int]),
    'pf_track_embed': (c_int, [POINTER(TrackWeights), POINTER(c_void_p), POINTER(c_int), POINTER(c_int), POINTER(c_int),
                               c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'pf_track_head': (c_int, [POINTER(TrackWeights), c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'pf_tracker_state_bytes': (c_size_t, []),
    'pf_tracker_workspace_bytes': (c_size_t, []),
    'pf_tracker_reset': (c_int, [c_void_p, c_void_p]),
    'pf_tracker_match': (c_int, [POINTER(TrackerConfig), c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'pf_track_paint': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    # ---- include/pf_fpn.h
    'pf_semantic_fpn_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'pf_semantic_fpn': (c_int, [POINTER(FpnWeights), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_size_t, c_int, c_int, c_int, c_int, c_void_p]),
    # ---- include/pf_decoder.h
    'pf_cast_maps': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'pf_fpn_pred': (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int,
                            c_int, c_int, c_void_p]),
    'pf_kernel_head_workspace_bytes': (c_size_t, [c_int, c_int]),
    'pf_kernel_head': (c_int, [POINTER(HeadWeights), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_void_p]),
    'pf_version': (c_int, []),
    'pf_last_error_string': (c_char_p, []),
    'pf_last_launch_count': (c_int, []),
    'pf_panoptic_workspace_bytes': (c_size_t, [c_int, c_int]),
    'pf_panoptic': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                            c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_size_t, c_void_p]),
    'pf_panoptic_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'pf_debug_timeline': (c_int, [c_void_p]),
    'pf_cast_feats': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'pf_binarise': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'pf_pool_splits': (c_int, [c_int, c_int, c_int]),
    'pf_mask_pool': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                             c_void_p]),
    'pf_pool_reduce': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'pf_init_proposals': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'pf_update_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'pf_kernel_update': (c_int, [POINTER(StageWeights), c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int,
                                 c_int, c_void_p]),
    'pf_set_fused_update': (c_int, [c_int]),
    'pf_vec_slices_bytes': (c_size_t, []),
    'pf_pack_vec_slices': (c_int, [POINTER(StageWeights), c_void_p, c_void_p]),
    'pf_split_kernels': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'pf_updator_workspace_bytes': (c_size_t, [c_int]),
    'pf_kernel_updator': (c_int, [POINTER(StageWeights), c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int,
                                  c_void_p]),
    'pf_mask_einsum': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                               c_void_p]),
    'pf_upsample2x': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'pf_decoder_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'pf_decoder_forward_slice': (c_int, [POINTER(StageWeights), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_int,
                                         c_int, c_int, c_int, c_void_p]),
    'pf_decoder_forward': (c_int, [POINTER(StageWeights), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p]),
}
EXPORTS = tuple(_SIGS)

_lib = None


def load():
    """Load libpf_decoder.so (once) and declare the prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PFError(-3, 'load', 'native library %s not found -- run `python -m polyphonicformer_b200.build` '
                                   '(there is no CPU/PyTorch fallback for the decoder)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, fn):
    if code != 0:
        raise PFError(code, fn, load().pf_last_error_string().decode('utf-8', 'replace'))


def call(fn, *args):
    """Call an int-returning entry point and raise PFError on a non-zero status."""
    check(getattr(load(), fn)(*args), fn)
