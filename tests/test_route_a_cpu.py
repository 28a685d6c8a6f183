"""INTEGRATION.md route A, end to end on CPU: the reference's own, unmodified `Polyphonic` / `PolyphonicVideo` detectors
are built from its unmodified configs through its own registries (mmcv restated by oracle/mmcv_shim.py) AFTER
`polyphonicformer_b200.register_all(force=True)`, so their `rpn_head` / `roi_head` are this package's modules; a state
dict built by the reference's own classes then loads with strict=True.  Needs /root/reference (build container only)."""
import os
import subprocess
import sys
import textwrap

import pytest

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'polyphonic')),
                                reason='the reference checkout is only present in the build container')

SCRIPT = textwrap.dedent('''
    import sys, copy
    sys.path.insert(0, %(root)r)
    import torch
    from oracle import mmcv_shim as shim
    shim.install()
    import polyphonic                                   # the reference registers ITS modules
    from mmdet.models.builder import build_detector, HEADS
    ref_cls = {k: HEADS.get(k) for k in ('KernelHead', 'KernelUpdateIterHead', 'KernelUpdateHead')}
    cfg_path = %(cfg)r
    cfg = shim.load_config(cfg_path)
    cfg.model.train_cfg = None                          # tools/test.py:200
    torch.manual_seed(0)
    ref_model = build_detector(copy.deepcopy(cfg.model))          # all-reference model
    ref_sd = ref_model.state_dict()

    import polyphonicformer_b200 as pf
    assert pf.register_all(force=True) is True          # INTEGRATION.md route A
    for k, c in ref_cls.items():
        assert HEADS.get(k) is not c and HEADS.get(k).__module__.startswith('polyphonicformer_b200'), k
    model = build_detector(copy.deepcopy(cfg.model))
    assert type(model).__module__.startswith('polyphonic.'), type(model)          # the reference's detector class
    assert isinstance(model.rpn_head, pf.KernelHead) and isinstance(model.roi_head, pf.KernelUpdateIterHead)
    assert all(isinstance(h, pf.KernelUpdateHead) for h in model.roi_head.mask_head)
    assert all(isinstance(h.kernel_update_conv, pf.KernelUpdator) for h in model.roi_head.mask_head)
    assert type(model.rpn_head.localization_fpn).__module__.startswith('polyphonic.')   # the neck stays the reference's
    from polyphonicformer_b200.modules import _pyramid_supported
    assert _pyramid_supported(model.rpn_head.localization_fpn)      # ... and KernelHead runs it on pf_semantic_fpn
    sd = model.state_dict()
    assert list(sd.keys()) == list(ref_sd.keys()), set(sd) ^ set(ref_sd)
    assert all(sd[k].shape == ref_sd[k].shape and sd[k].dtype == ref_sd[k].dtype for k in sd)
    res = model.load_state_dict(ref_sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    print('ROUTE_A_OK', type(model).__name__, len(sd))
''')


@pytest.mark.parametrize('cfg,name', [('configs/polyphonic_image/poly_r50_cityscapes_2x.py', 'Polyphonic'),
                                      ('configs/polyphonic_video/poly_r50_cityscapes_1x.py', 'PolyphonicVideo')])
def test_route_a_builds_reference_detector_with_b200_heads(cfg, name):
    # a subprocess: shim.install() rewires sys.modules / sys.meta_path for the whole interpreter
    code = SCRIPT % dict(root=ROOT, cfg=os.path.join(REF, cfg))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert 'ROUTE_A_OK ' + name in r.stdout, r.stdout[-500:]
