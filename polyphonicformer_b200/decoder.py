"""Host side of the B200 decoder: weight packing (with the feat_transform fold), buffer management and the calls
into libpf_decoder.so.  PyTorch is used for device memory and streams only; every arithmetic step of the decoder
runs in the CUDA kernels behind include/pf_decoder.h.

Reference being replaced (paths relative to the reference tree):
  KernelUpdateHead.forward                         polyphonic/kernel_update_head.py:212-353
  KernelUpdator.forward                            polyphonic/funcs/kernel_updator.py:55-93
  KernelUpdateIterHead._mask_forward / simple_test polyphonic/kernel_update.py:125-157, 282-354
"""
import ctypes

import torch

from . import _cabi
from ._cabi import PF_C, PF_MAX_CLASSES, PF_MAX_N, BranchWeights, StageWeights

_BRANCHES = (
    # suffix, updator, feat transform, head fcs, head norm, final fc
    dict(sfx='', upd='kernel_update_conv', ft='feat_transform', fc='fc_mask'),
    dict(sfx='_depth', upd='kernel_update_conv_depth', ft='feat_depth_transform', fc='fc_depth'),
)


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def round_up(v, m):
    return (v + m - 1) // m * m


def gate_interleave_index():
    """Packed row order of the concatenated [input_gate; update_gate] (512 rows): blocks of 64 alternate between the
    two gates so that one 128-column GEMM tile holds both gates of the same 64 features (pf_decoder.h)."""
    idx = []
    for t in range(PF_C // 64):
        idx += list(range(64 * t, 64 * t + 64)) + list(range(PF_C + 64 * t, PF_C + 64 * t + 64))
    return torch.tensor(idx, dtype=torch.long)


def _split_planes(w64, pad_rows):
    """fp64 [out][in] -> bf16 [2*pad_rows][in]: hi plane (bf16(w)) then lo plane (bf16(w - hi)), rows zero-padded."""
    w32 = w64.to(torch.float32)
    hi = w32.to(torch.bfloat16)
    lo = (w32 - hi.to(torch.float32)).to(torch.bfloat16)
    out = torch.zeros((2 * pad_rows, w32.shape[1]), dtype=torch.bfloat16)
    out[:w32.shape[0]] = hi
    out[pad_rows:pad_rows + w32.shape[0]] = lo
    return out


class _Packer:
    """Builds the bf16 weight stacks + one flat fp32 vector buffer and fills ``struct pf_branch_weights``."""

    def __init__(self):
        self.mats256, self.mats_ffn, self.vecs = [], [], []
        self.rows256 = self.rows_ffn = self.nvec = 0

    def add_matrix(self, w64, ffn=False):
        pad = round_up(w64.shape[0], 128)
        planes = _split_planes(w64, pad)
        if ffn:
            row, self.rows_ffn = self.rows_ffn, self.rows_ffn + planes.shape[0]
            self.mats_ffn.append(planes)
        else:
            assert w64.shape[1] == PF_C
            row, self.rows256 = self.rows256, self.rows256 + planes.shape[0]
            self.mats256.append(planes)
        return row

    def add_vector(self, v64):
        off, self.nvec = self.nvec, self.nvec + round_up(v64.numel(), 64)
        self.vecs.append((off, v64.reshape(-1).to(torch.float32)))
        return off

    def finish(self, device, ffn_channels):
        flat = torch.zeros(self.nvec + 128, dtype=torch.float32)   # slack: kernels read whole 64-float groups
        for off, v in self.vecs:
            flat[off:off + v.numel()] = v
        self.flat = flat.to(device)
        self.stack256 = (torch.cat(self.mats256) if self.mats256 else torch.zeros((128, PF_C), dtype=torch.bfloat16)).to(device)
        self.stack_ffn = (torch.cat(self.mats_ffn) if self.mats_ffn
                          else torch.zeros((128, ffn_channels), dtype=torch.bfloat16)).to(device)


class PackedStage:
    """One KernelUpdateHead's parameters in the layout of ``struct pf_stage_weights``.

    ``sd`` maps the reference's state-dict keys of one stage (e.g.
    ``kernel_update_conv.dynamic_layer.weight``) to tensors.  Folding is done in fp64 and rounded once to fp32
    (``views``: the folded fp32 parameters, used by the CPU algebra test), then matrices are split into bf16 hi/lo.
    """

    MATS = ('dyn_w', 'inp_w', 'gate_w', 'fc_w', 'qkv_w', 'out_w', 'ffn1_w', 'head_w', 'cls_w', 'kern_w', 'kbrow_w')
    PERMUTED = ('gate_w', 'gate_b')     # stored in gate_interleave_index() order; ``views`` keep the logical order

    def __init__(self, sd, device, num_classes, ffn_channels):
        f64 = {k: v.detach().to('cpu', torch.float64) for k, v in sd.items()}
        parts = {}

        def ln(name):
            return torch.stack([f64[name + '.weight'], f64[name + '.bias']])

        for bi, br in enumerate(_BRANCHES):
            sfx, upd, ft, fc = br['sfx'], br['upd'], br['ft'], br['fc']
            if ft + '.conv.weight' in f64:   # 1x1 conv with bias, no norm, no activation (kernel_update_head.py:124-140)
                Wt = f64[ft + '.conv.weight'].reshape(PF_C, PF_C)
                bt = f64[ft + '.conv.bias']
            else:                            # feat_transform_cfg=None
                Wt = torch.eye(PF_C, dtype=torch.float64)
                bt = torch.zeros(PF_C, dtype=torch.float64)
            Wdyn = f64[upd + '.dynamic_layer.weight']
            assert Wdyn.shape == (2 * PF_C, PF_C), 'KernelUpdator with feat_channels != 256 is not supported'
            Wfc, bfc = f64[fc + '.weight'], f64[fc + '.bias']
            p = dict(
                dyn_w=Wdyn @ Wt, dyn_b=f64[upd + '.dynamic_layer.bias'], dyn_cb=Wdyn @ bt,
                inp_w=f64[upd + '.input_layer.weight'], inp_b=f64[upd + '.input_layer.bias'],
                gate_w=torch.cat([f64[upd + '.input_gate.weight'], f64[upd + '.update_gate.weight']]),
                gate_b=torch.cat([f64[upd + '.input_gate.bias'], f64[upd + '.update_gate.bias']]),
                ln_input_norm_in=ln(upd + '.input_norm_in'), ln_norm_in=ln(upd + '.norm_in'),
                ln_norm_out=ln(upd + '.norm_out'), ln_input_norm_out=ln(upd + '.input_norm_out'),
                fc_w=f64[upd + '.fc_layer.weight'], fc_b=f64[upd + '.fc_layer.bias'], ln_fc_norm=ln(upd + '.fc_norm'),
                qkv_w=f64['attention%s.attn.in_proj_weight' % sfx], qkv_b=f64['attention%s.attn.in_proj_bias' % sfx],
                out_w=f64['attention%s.attn.out_proj.weight' % sfx], out_b=f64['attention%s.attn.out_proj.bias' % sfx],
                ln_attn=ln('attention_norm%s' % sfx),
                ffn1_w=f64['ffn%s.layers.0.0.weight' % sfx], ffn1_b=f64['ffn%s.layers.0.0.bias' % sfx],
                ffn2_w=f64['ffn%s.layers.1.weight' % sfx], ffn2_b=f64['ffn%s.layers.1.bias' % sfx],
                ln_ffn=ln('ffn_norm%s' % sfx),
                kern_w=Wt.t() @ Wfc, kern_b=Wt.t() @ bfc, kb_w=Wfc.t() @ bt,
            )
            p['kbrow_w'] = p['kb_w'].reshape(1, PF_C)
            p['kbrow_b'] = (bfc @ bt).reshape(1)
            assert p['ffn1_w'].shape == (ffn_channels, PF_C)
            if bi == 0:
                p['head_w'] = torch.cat([f64['cls_fcs.0.weight'], f64['mask_fcs.0.weight']])
                p['ln_head_a'], p['ln_head_b'] = ln('cls_fcs.1'), ln('mask_fcs.1')
                cw = torch.zeros(PF_MAX_CLASSES, PF_C, dtype=torch.float64)
                cb = torch.zeros(PF_MAX_CLASSES, dtype=torch.float64)
                ncls = f64['fc_cls.weight'].shape[0]
                assert ncls == num_classes <= PF_MAX_CLASSES
                cw[:ncls], cb[:ncls] = f64['fc_cls.weight'], f64['fc_cls.bias']
                p['cls_w'], p['cls_b'] = cw, cb
            else:
                p['head_w'] = f64['depth_regs.0.weight']
                p['ln_head_a'] = ln('depth_regs.1')
            parts[bi] = (p, float(bfc @ bt))

        pk = _Packer()
        rows, vecs = {}, {}
        perm = gate_interleave_index()
        for bi, (p, _) in parts.items():
            for k, v in p.items():
                if k in self.PERMUTED:
                    v = v[perm]
                if k in self.MATS:
                    rows[(bi, k)] = pk.add_matrix(v)
                elif k == 'ffn2_w':
                    rows[(bi, k)] = pk.add_matrix(v, ffn=True)
                else:
                    vecs[(bi, k)] = pk.add_vector(v)
        pk.finish(device, ffn_channels)
        self.packer = pk                      # keeps the device buffers alive
        self.views = {k: v.to(torch.float32) for bi, (p, _) in parts.items() for k, v in
                      (((bi, n), t) for n, t in p.items())}
        self.kb_b = {bi: kb for bi, (_, kb) in parts.items()}
        self.struct = StageWeights()
        self.struct.ffn_channels = ffn_channels
        self.struct.num_classes = num_classes
        self.struct.wstack256 = pk.stack256.data_ptr()
        self.struct.wstack_ffn = pk.stack_ffn.data_ptr()
        self.struct.wstack256_rows = pk.stack256.shape[0]
        self.struct.wstack_ffn_rows = pk.stack_ffn.shape[0]
        base = pk.flat.data_ptr()
        for bi, (p, kb_b) in parts.items():
            bw = self.struct.br[bi]
            for name in BranchWeights._ROWS:
                setattr(bw, name, rows.get((bi, name), 0))
            for name in BranchWeights._PTRS:
                setattr(bw, name, base + 4 * vecs[(bi, name)] if (bi, name) in vecs else None)
            bw.head_relu = 1 if bi == 0 else 0
        # per-(branch, column slice) copies of the vectors for the fused small-N kernel (gathered on the device, once);
        # a host-only packing (the CPU algebra test) leaves the field NULL
        self.vec_slices = None
        if torch.device(device).type == 'cuda':
            lib = _cabi.load()
            self.vec_slices = torch.empty(lib.pf_vec_slices_bytes() // 4, dtype=torch.float32, device=device)
            with torch.cuda.device(device):
                _cabi.call('pf_pack_vec_slices', ctypes.byref(self.struct), _ptr(self.vec_slices), _stream_ptr())
            self.struct.vec_slices = self.vec_slices.data_ptr()


class PackedUpdator:
    """A standalone KernelUpdator's parameters (no feat_transform fold) as a ``struct pf_stage_weights`` (branch 0)."""

    def __init__(self, sd, device):
        f = {k: v.detach().to('cpu', torch.float64) for k, v in sd.items()}

        def ln(name):
            return torch.stack([f[name + '.weight'], f[name + '.bias']])

        pk = _Packer()
        perm = gate_interleave_index()
        rows = dict(dyn_w=pk.add_matrix(f['dynamic_layer.weight']), inp_w=pk.add_matrix(f['input_layer.weight']),
                    gate_w=pk.add_matrix(torch.cat([f['input_gate.weight'], f['update_gate.weight']])[perm]),
                    fc_w=pk.add_matrix(f['fc_layer.weight']))
        vecs = dict(dyn_b=pk.add_vector(f['dynamic_layer.bias']), inp_b=pk.add_vector(f['input_layer.bias']),
                    gate_b=pk.add_vector(torch.cat([f['input_gate.bias'], f['update_gate.bias']])[perm]),
                    ln_input_norm_in=pk.add_vector(ln('input_norm_in')), ln_norm_in=pk.add_vector(ln('norm_in')),
                    ln_norm_out=pk.add_vector(ln('norm_out')), ln_input_norm_out=pk.add_vector(ln('input_norm_out')),
                    fc_b=pk.add_vector(f['fc_layer.bias']), ln_fc_norm=pk.add_vector(ln('fc_norm')))
        pk.finish(device, 2048)
        self.packer = pk
        self.struct = StageWeights()
        self.struct.wstack256 = pk.stack256.data_ptr()
        self.struct.wstack256_rows = pk.stack256.shape[0]
        self.struct.wstack_ffn = pk.stack_ffn.data_ptr()
        self.struct.wstack_ffn_rows = pk.stack_ffn.shape[0]
        self.struct.ffn_channels, self.struct.num_classes = 2048, 1
        bw = self.struct.br[0]
        for k, r in rows.items():
            setattr(bw, k, r)
        for k, o in vecs.items():
            setattr(bw, k, pk.flat.data_ptr() + 4 * o)


def run_kernel_updator(packed, update_feature, input_feature):
    """[R,256], [R,256] fp32 CUDA tensors -> [R,256]."""
    lib = _cabi.load()
    R = update_feature.shape[0]
    out = torch.empty((R, PF_C), dtype=torch.float32, device=update_feature.device)
    nbytes = lib.pf_updator_workspace_bytes(R)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=update_feature.device)
    _cabi.call('pf_kernel_updator', ctypes.byref(packed.struct), _ptr(update_feature), _ptr(input_feature), _ptr(out),
               _ptr(ws), nbytes, R, _stream_ptr())
    return out


class DecoderEngine:
    """Runs decoder stages on the current CUDA device/stream through the C ABI.

    Parameters: ``stage_dicts`` -- list (one per stage) of stage-local state dicts in the reference's key names.
    """

    def __init__(self, stage_dicts, device, num_classes=19, ffn_channels=2048):
        _cabi.load()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _cabi.PFError(-4, 'DecoderEngine', 'the decoder runs on sm_100a CUDA devices only; there is no '
                                                      'CPU path (got device %s)' % device)
        self.num_classes = num_classes
        self.ffn_channels = ffn_channels
        self.stages = [PackedStage(sd, self.device, num_classes, ffn_channels) for sd in stage_dicts]
        self.stage_array = (StageWeights * len(self.stages))(*[s.struct for s in self.stages])
        self._ws = {}

    # ------------------------------------------------------------------ buffers
    def _scratch(self, key, nbytes):
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = buf
        return buf

    @staticmethod
    def pitch(HW):
        return round_up(HW, 8)

    def prepare_feats(self, x_feats, depth_feats):
        """[B,256,H,W] x2 (fp32 or bf16) -> bf16 [2,B,256,HWp] in the library's layout."""
        B, C, H, W = x_feats.shape
        assert C == PF_C and depth_feats.shape == x_feats.shape
        HW = H * W
        HWp = self.pitch(HW)
        feats = torch.empty((2, B, PF_C, HWp), dtype=torch.bfloat16, device=self.device)
        if x_feats.dtype == torch.float32 and depth_feats.dtype == torch.float32:
            x, d = x_feats.contiguous(), depth_feats.contiguous()
            _cabi.call('pf_cast_feats', _ptr(x), _ptr(d), _ptr(feats), B, HW, HWp, _stream_ptr())
        else:   # already in storage precision: plain strided copy (no arithmetic)
            feats[0, :, :, :HW].copy_(x_feats.reshape(B, C, HW))
            feats[1, :, :, :HW].copy_(depth_feats.reshape(B, C, HW))
            if HWp != HW:
                feats[:, :, :, HW:].zero_()
        return feats

    # ------------------------------------------------------------------ one stage (KernelUpdateHead.forward)
    def stage_forward(self, stage, feats, mask_logits, obj, dep, H, W, cls_sigmoid=False):
        """feats from prepare_feats; mask_logits [B,N,H,W] fp32; obj/dep [B,N,256] fp32.
        Returns cls_score [B,N,classes], logits [2,B,N,H,W], obj_out, dep_out."""
        B, N = obj.shape[:2]
        assert N <= PF_MAX_N
        HW, HWp = H * W, feats.shape[-1]
        st = _stream_ptr()
        lib = _cabi.load()
        words = (HW + 31) // 32
        S = lib.pf_pool_splits(B, 2, HW)
        mask_logits = mask_logits.contiguous()
        obj, dep = obj.contiguous(), dep.contiguous()
        bits = torch.empty((B, words, 128), dtype=torch.int32, device=self.device)
        partial = torch.empty((2 * B, S, N, PF_C), dtype=torch.float32, device=self.device)
        cntp = torch.empty((2 * B, S, N), dtype=torch.float32, device=self.device)
        ksplit = torch.empty((2 * B, 2, N, PF_C), dtype=torch.bfloat16, device=self.device)
        kbias = torch.empty((2, B, N), dtype=torch.float32, device=self.device)
        obj_out, dep_out = torch.empty_like(obj), torch.empty_like(dep)
        cls = torch.empty((B, N, self.num_classes), dtype=torch.float32, device=self.device)
        logits = torch.empty((2, B, N, H, W), dtype=torch.float32, device=self.device)
        ws_bytes = lib.pf_update_workspace_bytes(B, N, self.ffn_channels)
        ws = self._scratch('update', ws_bytes)
        _cabi.call('pf_binarise', _ptr(mask_logits), _ptr(bits), B, N, HW, st)
        _cabi.call('pf_mask_pool', _ptr(feats), _ptr(bits), _ptr(partial), _ptr(cntp), B, N, HW, HWp, 2, S, st)
        _cabi.call('pf_kernel_update', ctypes.byref(self.stages[stage].struct), _ptr(partial), _ptr(cntp), S,
                   _ptr(obj), _ptr(dep), _ptr(obj_out), _ptr(dep_out), _ptr(cls), None, _ptr(ksplit), _ptr(kbias),
                   _ptr(ws), ws_bytes, B, N, 1 if cls_sigmoid else 0, st)
        _cabi.call('pf_mask_einsum', _ptr(feats), _ptr(ksplit), _ptr(kbias), _ptr(logits), None, B, N, HW, HWp,
                   2 * B, st)
        return cls, logits, obj_out, dep_out

    def init_proposals(self, feats, mask_preds, init_kernels_weight, seg_preds, conv_seg_weight, depth_kernel_weight,
                       num_thing_classes):
        """The tail of KernelHead._decode_init_proposals (polyphonic/kernel_head.py:313-336) on the decoder's kernels:
        binarise the P initial masks, pool x_feats under them (pf_mask_pool, one branch), add init_kernels.weight
        (pf_init_proposals), append the stuff masks / stuff kernels, expand the depth kernel.
        feats: prepare_feats(...) layout; mask_preds [B,P,H,W]; returns (proposal_feats [B,N,256,1,1],
        mask_preds [B,N,H,W], depth_proposal [B,N,256,1,1]) as the reference hands them to the decoder."""
        lib = _cabi.load()
        B, P, H, W = mask_preds.shape
        HW, HWp = H * W, feats.shape[-1]
        T = num_thing_classes
        n_stuff = conv_seg_weight.shape[0] - T
        N = P + n_stuff
        st = _stream_ptr()
        mask_preds = mask_preds.float().contiguous()
        bits = torch.empty((B, (HW + 31) // 32, 128), dtype=torch.int32, device=self.device)
        S = lib.pf_pool_splits(B, 1, HW)
        partial = torch.empty((B, S, P, PF_C), dtype=torch.float32, device=self.device)
        cntp = torch.empty((B, S, P), dtype=torch.float32, device=self.device)
        wk = init_kernels_weight.reshape(P, PF_C).float().contiguous()
        sk = conv_seg_weight.reshape(-1, PF_C)[T:].float().contiguous()
        prop = torch.empty((B, N, PF_C), dtype=torch.float32, device=self.device)
        _cabi.call('pf_binarise', _ptr(mask_preds), _ptr(bits), B, P, HW, st)
        _cabi.call('pf_mask_pool', _ptr(feats), _ptr(bits), _ptr(partial), _ptr(cntp), B, P, HW, HWp, 1, S, st)
        _cabi.call('pf_init_proposals', _ptr(partial), _ptr(cntp), _ptr(wk), _ptr(sk), _ptr(prop), B, P, n_stuff, S, st)
        masks = torch.cat([mask_preds, seg_preds[:, T:T + n_stuff].float()], dim=1)             # data movement only
        dprop = depth_kernel_weight.reshape(1, 1, PF_C, 1, 1).float().expand(B, N, PF_C, 1, 1)
        return prop.reshape(B, N, PF_C, 1, 1), masks, dprop

    def upsample2x(self, maps):
        """[..., H, W] fp32 -> [..., 2H, 2W] (bilinear, align_corners=False)."""
        maps = maps.contiguous()
        H, W = maps.shape[-2:]
        out = torch.empty(maps.shape[:-2] + (2 * H, 2 * W), dtype=torch.float32, device=self.device)
        _cabi.call('pf_upsample2x', _ptr(maps), _ptr(out), maps.numel() // (H * W), H, W, _stream_ptr())
        return out

    # ------------------------------------------------------------------ the whole stage loop
    @staticmethod
    def default_splits(B):
        """Batch windows decoded concurrently (see decode_inplace).  Measured at B=4, 1024x2048: 1 window 0.663 ms,
        2 windows 0.660 ms, 4 windows 0.741 ms per step -- the small-N block is latency-bound, so a half batch costs
        as much as a full one and concurrency buys nothing; one window is the default."""
        return 1

    @staticmethod
    def windows(B, splits):
        """[(b0, Bsub)] -- contiguous, near-equal batch windows."""
        splits = max(1, min(splits, B))
        edges = [B * i // splits for i in range(splits + 1)]
        return [(edges[i], edges[i + 1] - edges[i]) for i in range(splits)]

    def alloc_decode_buffers(self, B, N, H, W, upsample=True, splits=None):
        HW = H * W
        lib = _cabi.load()
        splits = self.default_splits(B) if splits is None else splits
        wins = self.windows(B, splits)
        ws = []
        for _, bs in wins:
            nbytes = lib.pf_decoder_workspace_bytes(bs, N, HW, self.ffn_channels)
            ws.append((torch.empty(nbytes, dtype=torch.uint8, device=self.device), nbytes))
        return dict(
            ws=ws, windows=wins,
            obj=torch.empty((B, N, PF_C), dtype=torch.float32, device=self.device),
            dep=torch.empty((B, N, PF_C), dtype=torch.float32, device=self.device),
            cls=torch.empty((B, N, self.num_classes), dtype=torch.float32, device=self.device),
            logits=torch.empty((2, B, N, H, W), dtype=torch.float32, device=self.device),
            scaled=(torch.empty((2, B, N, 2 * H, 2 * W), dtype=torch.float32, device=self.device)
                    if upsample else None))

    def decode(self, feats, mask_logits, obj, dep, H, W, upsample=True, all_stage_outputs=False, buffers=None):
        """KernelUpdateIterHead.simple_test's stage loop (kernel_update.py:316-336).  Returns a dict with
        cls_score (sigmoid), mask_preds, depth_preds, scaled_mask_preds, scaled_depth_preds, object_feats,
        depth_proposal; tensors alias ``buffers`` when given."""
        B, N = obj.shape[:2]
        buf = buffers or self.alloc_decode_buffers(B, N, H, W, upsample)
        buf['obj'].copy_(obj.reshape(B, N, PF_C))
        buf['dep'].copy_(dep.reshape(B, N, PF_C))
        self.decode_inplace(feats, mask_logits.contiguous(), buf, H, W, all_stage_outputs)
        scaled = buf['scaled'] if upsample else buf['logits']
        return dict(cls_score=buf['cls'], mask_preds=buf['logits'][0], depth_preds=buf['logits'][1],
                    scaled_mask_preds=scaled[0], scaled_depth_preds=scaled[1],
                    object_feats=buf['obj'], depth_proposal=buf['dep'])

    def decode_inplace(self, feats, mask_logits, buf, H, W, all_stage_outputs=False):
        """Launch-only variant (graph-capturable): obj/dep in ``buf`` are updated in place.

        The batch is decoded as ``len(buf['windows'])`` disjoint windows, each through ``pf_decoder_forward_slice`` on
        its own stream (window 0 on the current stream, the others on side streams forked from / joined to it with
        events).  Frames are independent, so this changes no result; it lets the latency-bound small-N block of one
        window run under the HBM-bound pooling / einsum of another."""
        B, N = buf['obj'].shape[:2]
        flags = _cabi.PF_FWD_ALL_STAGE_OUTPUTS if all_stage_outputs else 0
        cur = torch.cuda.current_stream()
        wins = buf['windows']
        side = self._side_streams(len(wins) - 1)
        fork = None
        if side:
            fork = torch.cuda.Event()
            fork.record(cur)
        launches = 0
        for i, ((b0, bs), (ws, ws_bytes)) in enumerate(zip(wins, buf['ws'])):
            st = cur if i == 0 else side[i - 1]
            if i > 0:
                st.wait_event(fork)
            _cabi.call('pf_decoder_forward_slice', self.stage_array, len(self.stages), _ptr(feats), _ptr(mask_logits),
                       _ptr(buf['obj']), _ptr(buf['dep']), _ptr(buf['cls']), _ptr(buf['logits']), _ptr(buf['scaled']),
                       _ptr(ws), ws_bytes, B, b0, bs, N, H, W, feats.shape[-1], flags,
                       ctypes.c_void_p(st.cuda_stream))
            launches += _cabi.load().pf_last_launch_count()
            if i > 0:
                join = torch.cuda.Event()
                join.record(st)
                cur.wait_event(join)
        self.last_launches = launches

    def _side_streams(self, n):
        pool = self.__dict__.setdefault('_streams', [])
        while len(pool) < n:
            pool.append(torch.cuda.Stream(self.device))
        return pool[:n]

class HostPipeline:
    """Decoder over HOST buffers: the call a user of the C ABI makes when feature maps and proposals live in pinned
    host memory (e.g. handed over by another process).  ``submit`` enqueues, on three streams, the host->device copy of
    one batch, the 3-stage decode and the device->host copy of its results; consecutive submissions overlap (copy of
    batch i+1 | decode of batch i | read-back of batch i-1) using ``depth`` device slots.  Nothing here computes: it
    is stream/event plumbing around ``DecoderEngine.decode_inplace``.

    host_in : dict of pinned tensors  x, d [B,256,H,W] bf16; mask [B,N,H,W] f32; prop, dprop [B,N,256] f32
    host_out: dict of pinned tensors  cls [B,N,classes] f32; scaled [2,B,N,2H,2W] f32 (or logits [2,B,N,H,W])
    """

    def __init__(self, engine, B, N, H, W, upsample=True, depth=2):
        self.eng, self.B, self.N, self.H, self.W, self.upsample = engine, B, N, H, W, upsample
        dev = engine.device
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.slots = []
        HWp = engine.pitch(H * W)
        self.direct = HWp == H * W      # the host maps can be copied straight into the library's [2,B,256,HWp] layout
        for _ in range(depth):
            feats = torch.empty((2, B, PF_C, HWp), dtype=torch.bfloat16, device=dev)
            if self.direct:
                x, d = feats[0].view(B, PF_C, H, W), feats[1].view(B, PF_C, H, W)
            else:
                feats.zero_()
                x = torch.empty((B, PF_C, H, W), dtype=torch.bfloat16, device=dev)
                d = torch.empty_like(x)
            self.slots.append(dict(
                x=x, d=d, feats=feats, mask=torch.empty((B, N, H, W), dtype=torch.float32, device=dev),
                buf=engine.alloc_decode_buffers(B, N, H, W, upsample),
                ev_in=torch.cuda.Event(), ev_run=torch.cuda.Event(), ev_out=torch.cuda.Event()))
        self.n = 0

    def h2d_bytes(self):
        s = self.slots[0]
        return sum(t.numel() * t.element_size() for t in (s['x'], s['d'], s['mask'], s['buf']['obj'], s['buf']['dep']))

    def d2h_bytes(self):
        b = self.slots[0]['buf']
        out = b['scaled'] if self.upsample else b['logits']
        return b['cls'].numel() * 4 + out.numel() * 4

    def submit(self, host_in, host_out):
        """Enqueue one batch; returns the event that fires when ``host_out`` holds its results."""
        s = self.slots[self.n % len(self.slots)]
        self.n += 1
        B, H, W = self.B, self.H, self.W
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(s['ev_run'])        # the slot's previous decode has consumed its inputs
            s['x'].copy_(host_in['x'], non_blocking=True), s['d'].copy_(host_in['d'], non_blocking=True)
            s['mask'].copy_(host_in['mask'], non_blocking=True)
            self.s_in.wait_event(s['ev_out'])        # ... and its previous results have left the output buffers
            s['buf']['obj'].copy_(host_in['prop'].reshape(B, self.N, PF_C), non_blocking=True)
            s['buf']['dep'].copy_(host_in['dprop'].reshape(B, self.N, PF_C), non_blocking=True)
            s['ev_in'].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(s['ev_in'])
            self.s_run.wait_event(s['ev_out'])
            if not self.direct:                      # ragged row pitch: one strided device copy into the layout
                HW = H * W
                s['feats'][0, :, :, :HW].copy_(s['x'].reshape(B, PF_C, HW))
                s['feats'][1, :, :, :HW].copy_(s['d'].reshape(B, PF_C, HW))
            self.eng.decode_inplace(s['feats'], s['mask'], s['buf'], H, W)
            s['ev_run'].record(self.s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(s['ev_run'])
            host_out['cls'].copy_(s['buf']['cls'], non_blocking=True)
            src = s['buf']['scaled'] if self.upsample else s['buf']['logits']
            host_out['scaled' if self.upsample else 'logits'].copy_(src, non_blocking=True)
            s['ev_out'].record(self.s_out)
        return s['ev_out']

    def drain(self):
        for st in (self.s_in, self.s_run, self.s_out):
            st.synchronize()


class PanopticPipeline(HostPipeline):
    """``KernelUpdateIterHead.simple_test`` over HOST buffers (kernel_update.py:282-354): like HostPipeline, but the
    decoder's logits never leave the device -- ``pf_panoptic`` turns them into the panoptic map and the two depth maps
    per frame, and only those (12 bytes per pixel) plus the segment records are copied back.

    host_in : HostPipeline's inputs + depth_pred [B,1,H,W] f32 (KernelHead's initial depth prediction)
    host_out: panoptic [B,H0,W0] i32; depth_final, depth_basic [B,H0,W0] f32; segments [B,128,24] u8 (struct
              pf_segment records); nseg [B] i32 -- all pinned;  H0 = 8H, W0 = 8W (no padding crop)
    """

    def __init__(self, engine, B, N, H, W, num_proposals=100, num_thing_classes=8, max_per_img=100,
                 instance_score_thr=0.3, overlap_thr=0.6, depth_act_mode='sigmoid', depth=2):
        # upsample=False: the x2 up-sampled logits are never materialised -- pf_panoptic_batch samples the decoder's own
        # stride-8 maps with the composed taps (bit-identical to pf_upsample2x followed by pf_panoptic)
        super().__init__(engine, B, N, H, W, upsample=False, depth=depth)
        dev = engine.device
        lib = _cabi.load()
        self.cfg = (num_proposals, num_thing_classes, max_per_img, float(instance_score_thr), float(overlap_thr),
                    {'monodepth': 0, 'sigmoid': 1}[depth_act_mode])
        H0, W0 = 8 * H, 8 * W
        self.pws_bytes = B * lib.pf_panoptic_workspace_bytes(H0, W0)
        for s in self.slots:
            s.update(dpred=torch.empty((B, 1, H, W), dtype=torch.float32, device=dev),
                     pan=torch.empty((B, H0, W0), dtype=torch.int32, device=dev),
                     dfinal=torch.empty((B, H0, W0), dtype=torch.float32, device=dev),
                     dbasic=torch.empty((B, H0, W0), dtype=torch.float32, device=dev),
                     segs=torch.zeros((B, 128, 24), dtype=torch.uint8, device=dev),
                     nseg=torch.zeros(B, dtype=torch.int32, device=dev),
                     pws=torch.empty(self.pws_bytes, dtype=torch.uint8, device=dev))

    def h2d_bytes(self):
        return super().h2d_bytes() + self.slots[0]['dpred'].numel() * 4

    def d2h_bytes(self):
        s = self.slots[0]
        return sum(s[k].numel() * s[k].element_size() for k in ('pan', 'dfinal', 'dbasic', 'segs', 'nseg'))

    def submit(self, host_in, host_out):
        s = self.slots[self.n % len(self.slots)]
        self.n += 1
        B, N, H, W = self.B, self.N, self.H, self.W
        P, T, max_per_img, thr_i, thr_o, dmode = self.cfg
        ncls = self.eng.num_classes
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(s['ev_run'])
            s['x'].copy_(host_in['x'], non_blocking=True), s['d'].copy_(host_in['d'], non_blocking=True)
            s['mask'].copy_(host_in['mask'], non_blocking=True)
            s['dpred'].copy_(host_in['depth_pred'], non_blocking=True)
            s['buf']['obj'].copy_(host_in['prop'].reshape(B, N, PF_C), non_blocking=True)
            s['buf']['dep'].copy_(host_in['dprop'].reshape(B, N, PF_C), non_blocking=True)
            s['ev_in'].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(s['ev_in'])
            self.s_run.wait_event(s['ev_out'])           # the previous results of this slot have been read back
            if not self.direct:
                HW = H * W
                s['feats'][0, :, :, :HW].copy_(s['x'].reshape(B, PF_C, HW))
                s['feats'][1, :, :, :HW].copy_(s['d'].reshape(B, PF_C, HW))
            self.eng.decode_inplace(s['feats'], s['mask'], s['buf'], H, W)
            logits, cls = s['buf']['logits'], s['buf']['cls']
            _cabi.call('pf_panoptic_batch', _ptr(cls), _ptr(logits[0]), _ptr(logits[1]), _ptr(s['dpred']), B, N, P, T, ncls,
                       2 * H, 2 * W, 8 * H, 8 * W, max_per_img, thr_i, thr_o, dmode, 1, _ptr(s['pan']), _ptr(s['dfinal']),
                       _ptr(s['dbasic']), _ptr(s['segs']), 128, _ptr(s['nseg']), _ptr(s['pws']), self.pws_bytes,
                       _stream_ptr())
            s['ev_run'].record(self.s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(s['ev_run'])
            for k, hk in (('pan', 'panoptic'), ('dfinal', 'depth_final'), ('dbasic', 'depth_basic'), ('segs', 'segments'),
                          ('nseg', 'nseg')):
                host_out[hk].copy_(s[k], non_blocking=True)
            s['ev_out'].record(self.s_out)
        return s['ev_out']
