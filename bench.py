#!/usr/bin/env python
"""bench.py -- decoder throughput on synthetic Cityscapes-shape frames (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--height H] [--width W]

One "step" = the 3-stage kernel-update decoder (KernelUpdateIterHead.simple_test's stage loop incl. the x2 upsampling
and cls sigmoid) over one batch of B frames.  Workload at N=1: BASELINE.json configs[1] -- 1024x2048 frames
(128x256 decoder map, N=111 kernels, C=256), batch 4, bf16 feature maps.  N>1: one process per GPU (torchrun), each
rank decodes its own batch (frames shard batch-wise, no data-path collective) -> "scaling": "weak".

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM, timed with CUDA events, max over
ranks.  `e2e` = frames/s through the call the reference's users make (KernelUpdateIterHead.simple_test: decoder +
panoptic merge, PanopticPipeline) with pinned HOST buffers, H2D of the step's inputs and D2H of the step's results
(panoptic map + 2 depth maps + segments) inside the timed region, at every N; `e2e_logits` = the decoder-only variant
that reads the fp32 up-sampled logits back (HostPipeline).  `roofline` = dominant kernel against MEASURED_PEAKS.json; `kernels` lists
every kernel of the step.  `cpu_baseline` / `--impl reference` = the CPU oracle (oracle/decoder_ref.py, a pinned
restatement of the reference's forward; the reference itself needs mmcv, absent on the box) on the host cores.
Non-headline extras on rank 0 at N=1: `library_baseline` (the same oracle through PyTorch eager on the same GPU),
`postprocess` (pf_panoptic alone), `kernel_head` (the KernelHead tail that produces the decoder's inputs).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

C, N_KERNELS, STAGES, NUM_CLASSES = 256, 111, 3, 19
METRIC = 'decoder frames/sec (1024x2048, 100 queries -> 111 kernels, 3 stages)'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--height', type=int, default=1024)
    ap.add_argument('--width', type=int, default=2048)
    ap.add_argument('--all-stage-outputs', action='store_true',
                    help='also materialise the (unobservable) fp32 logits + depth einsum of stages 0..S-2')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-video', action='store_true')
    ap.add_argument('--no-model', action='store_true')
    ap.add_argument('--no-check', action='store_true', help='skip the oracle check of one frame of the timed batch')
    ap.add_argument('--no-postprocess', action='store_true', help='skip the (non-headline) post-processing timing')
    ap.add_argument('--no-kernel-head', action='store_true', help='skip the (non-headline) KernelHead-tail timing')
    ap.add_argument('--no-graph', action='store_true', help='launch every step from the host instead of replaying a CUDA graph')
    ap.add_argument('--update-mode', default='default', choices=['default', 'fused', 'per-layer'],
                    help='A/B: small-N block as ONE fused cluster kernel (csrc/pf_stage.cu) or as 12 per-layer launches '
                         '(csrc/pf_update.cu); default = the library default')
    ap.add_argument('--splits', type=int, default=None, help='batch windows decoded concurrently (default: 1)')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d['hbm_gbs']), tf=float(d.get('bf16_tflops_sustained', d.get('bf16_tflops'))),
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, tf=1400.0, source='fallback (B200_PROFILING.md)')


def synth_state(seed=0):
    from oracle import synth
    sd = synth.synth_decoder_state(STAGES, seed)
    stage_dicts = [{k[len('mask_head.%d.' % s):]: v for k, v in sd.items() if k.startswith('mask_head.%d.' % s)}
                   for s in range(STAGES)]
    return sd, stage_dicts


def host_inputs(B, H, W, seed):
    """Synthetic decoder inputs in HOST memory (same generator as the parity tests, cheap variant for big shapes)."""
    g = torch.Generator().manual_seed(1000 + seed)
    x = (torch.relu(torch.randn(B, C, H, W, generator=g)) + torch.relu(torch.randn(B, C, H, W, generator=g)))
    d = (torch.relu(torch.randn(B, C, H, W, generator=g)) + 0.5 * torch.relu(torch.randn(B, C, H, W, generator=g)))
    coarse = torch.randn(B, N_KERNELS, max(H // 4, 1), max(W // 4, 1), generator=g)
    mask = torch.nn.functional.interpolate(coarse, size=(H, W), mode='bilinear', align_corners=False) * 3 \
        + torch.randn(B, N_KERNELS, H, W, generator=g) - 0.4
    prop = torch.randn(B, N_KERNELS, C, generator=g) * 0.5
    dprop = (torch.randn(1, 1, C, generator=g) * 0.1).expand(B, N_KERNELS, C).contiguous()
    return dict(x=x.to(torch.bfloat16), d=d.to(torch.bfloat16), mask=mask.contiguous(), prop=prop, dprop=dprop)


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own algorithm on the host cores: oracle/decoder_ref.py, a restatement pinned to the REAL
    reference's outputs by tests/test_oracle_golden.py (the reference itself needs mmcv, absent on the box).  Same
    workload as our arm: every step decodes the same batch of `--batch` frames, `--steps` steps after `--warmup`."""
    if rank != 0:
        return
    from oracle import decoder_ref as ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd, _ = synth_state()
    B, H, W = args.batch, args.height // 8, args.width // 8
    inp = host_inputs(B, H, W, 0)
    x, d = inp['x'].float(), inp['d'].float()
    prop, dprop = inp['prop'].reshape(B, N_KERNELS, C, 1, 1), inp['dprop'].reshape(B, N_KERNELS, C, 1, 1)
    steps, warm = max(1, args.steps), max(0, args.warmup)

    def one_step():   # frame by frame: the reference's own test loop runs samples_per_gpu=1 (configs/_base_/datasets)
        for b in range(B):
            ref.decoder_forward(sd, x[b:b + 1], prop[b:b + 1], inp['mask'][b:b + 1], d[b:b + 1], dprop[b:b + 1])

    with torch.no_grad():
        for _ in range(warm):
            one_step()
        t0 = time.perf_counter()
        for _ in range(steps):
            one_step()
        dt = time.perf_counter() - t0
    fps = steps * B / dt
    sample = '%d steps x %d frames of %dx%d (decoder map %dx%d), fp32, torch %d threads' % (
        steps, B, args.height, args.width, H, W, cores)
    line = dict(impl='reference', metric=METRIC, value=fps, unit='frames/s', n_gpus=args.gpus, steps=steps, warmup=warm,
                ms_per_step=1e3 * dt / steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic', config=workload_config(args, B),
                cpu_baseline=dict(value=fps, unit='frames/s', cores=cores, kind='port', sample=sample,
                                  pinned='oracle/decoder_ref.py is pinned to the unmodified reference by '
                                         'tests/test_oracle_golden.py (fixtures regenerated bit-identically from '
                                         '/root/reference by oracle/make_golden.py)'),
                e2e=dict(value=fps, unit='frames/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


def workload_config(args, B):
    """The workload only -- identical for both arms (the driver compares the two lines)."""
    return dict(workload='poly_r50 image decoder, %dx%d synthetic Cityscapes-shape frames, batch=%d per GPU, '
                         'N=111 kernels (100 queries + 11 stuff), C=256, 3 stages, bf16 feature maps' %
                         (args.height, args.width, B),
                global_batch=B * args.gpus, decoder_map='%dx%d' % (args.height // 8, args.width // 8),
                parallelism='dp%d (batch-sharded frames, no collective)' % args.gpus,
                l2='per-step working set %.0f MB > 126 MB L2 (no flush needed)' % working_set_mb(args, B)
                if working_set_mb(args, B) > 126 else 'L2 flushed between timed iterations')


def launch_info(args, B):
    """How OUR arm runs the workload (not part of `config`)."""
    return dict(stage_outputs='all' if args.all_stage_outputs else 'observable-only',
                batch_windows=args.splits if args.splits else 1,
                small_n_block='fused 8-CTA-cluster kernel (1 launch per stage)' if getattr(args, 'fused', False) else '12 per-layer launches',
                launch='eager' if args.no_graph else 'cuda-graph replay')


def working_set_mb(args, B):
    HW = (args.height // 8) * (args.width // 8)
    return (2 * B * C * HW * 2 + 2 * B * N_KERNELS * HW * 4 * 5) / 1e6


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = 'timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
        'clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '20'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self, window=None):
        """window = (t0, t1) host time.time() bounds of the busy region: only samples inside it are used."""
        import datetime
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.05)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(', ') for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                if window is not None:
                    ts = datetime.datetime.strptime(r[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                    if not (window[0] <= ts <= window[1]):
                        continue
                sm.append(float(r[1])), mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                if v.strip().lower() == 'active':
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no samples'])
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


class NvmlClockSampler:
    """In-process sampler (a thread polling NVML every ~2 ms between start and stop): every sample lies inside the
    busy region by construction, however short it is.  `nvidia-smi -lms` (ClockSampler) delivers a sample only every
    ~100 ms on some boxes, which can miss a 110 ms timed region entirely; it stays as the fallback."""
    BITS = ((0x8, 'hw_slowdown'), (0x40, 'hw_thermal_slowdown'), (0x20, 'sw_thermal_slowdown'), (0x4, 'sw_power_cap'))

    def __init__(self, gpu):
        import threading
        import pynvml
        pynvml.nvmlInit()
        self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(gpu)
        self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        self.sm, self.bits, self.run = [], 0, True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def _loop(self):
        nv = self.nv
        reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while self.run:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.bits |= int(reasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self, window=None):
        self.run = False
        self.t.join()
        if not self.sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no samples'])
        sm = sorted(self.sm)
        return dict(sm_mhz=float(sm[len(sm) // 2]), sm_max_mhz=float(self.max_mhz),
                    reasons=sorted(n for b, n in self.BITS if self.bits & b), samples=len(sm), source='nvml thread, 2 ms period')


def make_clock_sampler(gpu):
    try:
        return NvmlClockSampler(gpu)
    except Exception:
        return ClockSampler(gpu)


# ----------------------------------------------------------------------------------------------- our arm
def flush_l2(buf):
    buf.zero_()


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs NVML reports as local to its GPU, BEFORE any pinned host buffer is allocated
    (first touch puts the pages on that NUMA node).  Matters for the host-buffer e2e path when 8 ranks share one host:
    without it every rank's 659 MB per step crosses the socket interconnect.  Best effort: silently skipped."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from polyphonicformer_b200 import _cabi
    from polyphonicformer_b200.decoder import DecoderEngine, _ptr, _stream_ptr
    lib = _cabi.load()
    if args.update_mode != 'default':
        lib.pf_set_fused_update(1 if args.update_mode == 'fused' else 0)
    fused = bool(lib.pf_set_fused_update(0))     # the setter returns the previous value: read it, then restore it
    lib.pf_set_fused_update(1 if fused else 0)
    args.fused = fused
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    B, H, W = args.batch, args.height // 8, args.width // 8
    HW = H * W
    N = N_KERNELS
    sd, stage_dicts = synth_state()
    eng = DecoderEngine(stage_dicts, dev, NUM_CLASSES, 2048)
    hin = host_inputs(B, H, W, rank)
    # resident inputs
    feats = eng.prepare_feats(hin['x'].to(dev), hin['d'].to(dev))
    mask = hin['mask'].to(dev)
    prop, dprop = hin['prop'].to(dev), hin['dprop'].to(dev)
    buf = eng.alloc_decode_buffers(B, N, H, W, upsample=True, splits=args.splits)
    need_flush = working_set_mb(args, B) <= 126
    l2buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if need_flush else None

    def step_eager():
        buf['obj'].copy_(prop), buf['dep'].copy_(dprop)
        eng.decode_inplace(feats, mask, buf, H, W, args.all_stage_outputs)

    step = step_eager
    if not args.no_graph:
        # the step is launch-only (no allocation, no sync): capture it once, replay it per step
        graph = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(dev)
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            step_eager()
            cap.synchronize()
            with torch.cuda.graph(graph, stream=cap):
                step_eager()
        torch.cuda.current_stream().wait_stream(cap)
        step = graph.replay

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = make_clock_sampler(local_rank) if rank == 0 else None   # samples the timed steps only
    if isinstance(sampler, ClockSampler):
        time.sleep(0.25)                                       # let nvidia-smi start before the GPU gets busy
    t_busy0 = time.time()
    launches_per_step = eng.last_launches + 2               # + the two 57 KB proposal copies (torch memcpy kernels)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        if need_flush:
            flush_l2(l2buf)
        ev[i][0].record()
        step()
        ev[i][1].record()
    barrier()
    clocks = sampler.stop((t_busy0, time.time())) if sampler else None
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    value = world * B * args.steps / (ms_total / 1e3)

    # ---------------- parity of the step that was just timed (outside the timed region): frame 0 against the oracle
    check = None
    if rank == 0 and not args.no_check:
        step()
        torch.cuda.synchronize()
        check = parity_check(sd, hin, buf, args.all_stage_outputs)

    # ---------------- per-kernel timing (same arguments as inside the step), CUDA events on the launching stream
    kernels = kernel_breakdown(args, eng, lib, feats, mask, prop, dprop, buf, B, N, H, W, dev) if rank == 0 else None

    # ---------------- e2e through the public API with pinned host buffers
    e2e = e2e_logits = None
    if not args.no_e2e:
        e2e = run_e2e(args, eng, hin, B, N, H, W, dev, world, barrier)
        e2e_logits = run_e2e_logits(args, eng, hin, B, N, H, W, dev, world, barrier)
    video = None
    if not args.no_video:
        try:
            video = run_video(args, eng, dev, world, rank, barrier)
        except Exception as e:       # an extra section, never a reason to lose the headline line
            if world > 1:
                raise                # ... but ranks must not diverge around a collective
            video = dict(unavailable=repr(e)[:300])
    mfps = None
    if not args.no_model:
        try:
            mfps = model_fps(args, dev, world, barrier)
        except Exception as e:
            if world > 1:
                raise
            mfps = dict(unavailable=repr(e)[:300])

    if rank != 0:
        return
    pk = peaks()
    for k in kernels:
        if k['bound'] == 'hbm':
            k['achieved'] = k['alg_bytes'] / (k['ms'] * 1e-3) / 1e9
            k['peak'], k['unit'] = pk['hbm'], 'GB/s'
            k['frac'] = k['achieved'] / k['peak']
    # the dominant ROOFLINE-bound kernel of the step (the small-N block is latency-bound and listed in ms only)
    dom = max((k for k in kernels if k['bound'] == 'hbm'), key=lambda k: k['ms'] * k['calls_per_step'])
    roofline = dict(kernel=dom['name'], bound=dom['bound'], achieved=dom['achieved'], peak=dom['peak'],
                    unit=dom['unit'], frac=dom['frac'], traffic=dom['traffic'], peak_source=pk['source'],
                    alg_bytes_per_launch=dom['alg_bytes'], ms_per_launch=dom['ms'],
                    timing=dom['timing'] + '; CUDA events on the launching stream',
                    ms_single_launch_l2_flushed=dom['ms_single_launch_l2_flushed'],
                    traffic_source='profiles/r2_traffic.json (ncu --set full, dram__bytes_read.sum + '
                                   'dram__bytes_write.sum per launch)' if dom['traffic'] else None)
    line = dict(metric=METRIC, value=value,
                unit='frames/s', n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms_total / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='bf16', data='synthetic', config=workload_config(args, B), launch=launch_info(args, B), clocks=clocks,
                gpu_launches=launches_per_step * args.steps, launches_per_step=launches_per_step,
                roofline=roofline, kernels=kernels, ms_per_stage=ms_total / args.steps / STAGES)
    if check:
        line['check'] = check
    if e2e:
        if numa_cpus:
            e2e['host_cpus_bound_per_rank'] = numa_cpus
        line['e2e'] = e2e
        line['e2e_logits'] = e2e_logits
    if video:
        line['video'] = video
    if mfps:
        line['model_fps'] = mfps
    if not args.no_cpu_baseline and world == 1:
        line['cpu_baseline'] = cpu_baseline(args)
    if not args.no_postprocess and world == 1:
        line['postprocess'] = postprocess_timing(args, dev, cpu=not args.no_cpu_baseline)
        if e2e and 'cpu_baseline' in line and 'cpu_port_ms_per_frame' in line['postprocess']:
            per_frame = 1e3 / line['cpu_baseline']['value'] + line['postprocess']['cpu_port_ms_per_frame']
            line['e2e']['cpu_port_frames_per_s_same_work'] = 1e3 / per_frame
    if not args.no_cpu_baseline and world == 1:
        torch.cuda.empty_cache()
        try:
            line['library_baseline'] = library_baseline(args, B, dev)
        except Exception as e:   # a baseline, never a reason to lose the line
            line['library_baseline'] = dict(unavailable=repr(e)[:200])
        torch.cuda.empty_cache()
    if not args.no_kernel_head and world == 1:
        torch.cuda.empty_cache()
        line['kernel_head'] = kernel_head_timing(args, B, dev, pk['hbm'], cpu=not args.no_cpu_baseline)
    emit(line)


def parity_check(sd, hin, buf, all_stage_outputs):
    """Frame 0 of the batch the timed steps decoded, against oracle/decoder_ref.py (fp32, host) on the same bf16-rounded
    feature maps: norm-wise relative error of every output of the step and the number of final mask bits
    (logit > 0) that differ.  The north-star gate is 1e-3 on the mask and depth logits."""
    from oracle import decoder_ref as ref
    x, d = hin['x'][:1].float(), hin['d'][:1].float()
    prop, dprop = hin['prop'][:1].reshape(1, N_KERNELS, C, 1, 1), hin['dprop'][:1].reshape(1, N_KERNELS, C, 1, 1)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        want = ref.decoder_forward(sd, x, prop, hin['mask'][:1], d, dprop, return_all_stages=True)
    got = dict(cls_score=buf['cls'][:1], mask_preds=buf['logits'][0, :1], depth_preds=buf['logits'][1, :1],
               scaled_mask_preds=buf['scaled'][0, :1], scaled_depth_preds=buf['scaled'][1, :1],
               object_feats=buf['obj'][:1].reshape(1, N_KERNELS, C, 1, 1),
               depth_proposal=buf['dep'][:1].reshape(1, N_KERNELS, C, 1, 1))
    rel = {}
    for k, v in got.items():
        a, b = v.detach().cpu().double().flatten(), want[k].double().flatten()
        rel[k] = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
    m_got, m_ref = got['mask_preds'].cpu() > 0, want['mask_preds'] > 0
    near = [float((st['mask_preds'].abs() < 1e-4).sum()) for st in want['stages']]
    return dict(frame=0, oracle='oracle/decoder_ref.py (fp32, host) on the same bf16-rounded feature maps',
                rel_l2=rel, worst_rel_l2=max(rel.values()), gate=1e-3, passed=max(rel.values()) < 1e-3,
                final_mask_bits=int(m_ref.numel()), final_mask_bits_flipped=int((m_got != m_ref).sum()),
                oracle_logits_within_1e4_of_0_per_stage=near)


def run_video(args, eng, dev, world, rank, barrier):
    """Video mode (BASELINE.json config 5: SemKITTI-DVPS-shape 376x1241 frames padded to 384x1248, 5-frame clips), frame
    sharded: every rank owns 5 frames of each wave of 5 * N consecutive frames (N clips), runs decoder + batched panoptic
    merge + tracking head (mask -> box, RoIAlign, embedding head) on them, ONE NCCL all_gather of the per-frame records,
    the association replayed in frame order on every rank (pf_tracker_match, memo reset per clip), track-id / semantic
    maps painted and read back.  Decoder inputs and the FPN pyramid are synthetic and resident; backbone / FPN / KernelHead
    are not part of it.  Device time from the first launch to the last read-back, max over ranks (the step has host
    synchronisations of its own: segment lists after the merge, the record headers, the ids)."""
    import json as _json
    from types import SimpleNamespace
    import torch.distributed as dist
    from oracle import synth
    from polyphonicformer_b200.registry import to_config
    from polyphonicformer_b200.track import TrackHeadEngine
    from polyphonicformer_b200.video import VideoShardRunner
    gold = os.path.join(ROOT, 'tests', 'golden')
    test_cfg = to_config(_json.load(open(os.path.join(gold, 'roi_head_cfg.json')))['test_cfg'])
    tracker_cfg = _json.load(open(os.path.join(gold, 'video_cfg.json')))['tracker']
    F, H, W, clip = 5, 48, 156, 5
    roi = SimpleNamespace(num_proposals=synth.N_PROPOSALS, num_thing_classes=synth.NUM_THING, merge_joint=True)
    last = SimpleNamespace(depth_act_mode='sigmoid', num_classes=synth.NUM_CLASSES)
    trk = TrackHeadEngine(synth.synth_track_head_state(0), dev)
    runner = VideoShardRunner(eng, trk, roi, last, test_cfg, tracker_cfg, synth.NUM_THING, synth.NUM_STUFF, clip_len=clip)
    hin = host_inputs(F, H, W, 50 + rank)
    g = torch.Generator().manual_seed(60 + rank)
    meta = dict(img_shape=(8 * H, 8 * W, 3), ori_shape=(8 * H, 8 * W, 3), pad_shape=(8 * H, 8 * W, 3), scale_factor=1.0,
                flip=False, batch_input_shape=(8 * H, 8 * W))
    batch = dict(feats=eng.prepare_feats(hin['x'].to(dev), hin['d'].to(dev)), mask=hin['mask'].to(dev), prop=hin['prop'].to(dev),
                 dprop=hin['dprop'].to(dev), depth_pred=torch.randn(F, 1, H, W, generator=g).to(dev), img_metas=[meta] * F,
                 fpn=[torch.randn(F, C, 8 * H // s, 8 * W // s, generator=g).to(dev) for s in (4, 8, 16, 32)])
    steps = max(3, min(args.steps, 10))
    things = 0
    for wave in range(2):
        out = runner.step(batch, H, W, wave)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for wave in range(2, 2 + steps):
        out = runner.step(batch, H, W, wave)
    b.record()
    barrier()
    things = sum(int(len(np.unique(o['track'])) - 1) for o in out) * steps
    runner.timing = {}
    for wave in range(2 + steps, 5 + steps):           # three more waves with a synchronisation after every section
        runner.step(batch, H, W, wave)
    sections = {k: 1e3 * v / 3 for k, v in runner.timing.items()}
    runner.timing = None
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # the tracking head alone on a fixed 30 RoIs per frame (random-weight decoders merge into few segments, so the step
    # above exercises it lightly): mask -> box from a 30-segment panoptic map, RoIAlign + embedding head, association
    K = 30
    pan = torch.zeros((8 * H, 8 * W), dtype=torch.int32)
    for k in range(K):
        pan[(k // 6) * 70 + 5:(k // 6) * 70 + 60, (k % 6) * 200 + 10:(k % 6) * 200 + 150] = k + 1
    pan = pan.to(dev)
    fpn1 = [lv[0] for lv in batch['fpn']]
    runner.tracker.reset()
    labels = torch.arange(K, device=dev) % 8
    scores = torch.linspace(0.95, 0.4, K, device=dev).view(-1, 1)

    def track_once(fid):
        rois, tight = trk.boxes_from_panoptic(pan, list(range(1, K + 1)))
        emb = trk.embed(fpn1, rois)
        return runner.tracker.match_async(torch.cat([tight, scores], 1), labels, emb, fid)

    for i in range(3):
        track_once(i + 1)
    torch.cuda.synchronize()
    c, d = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c.record()
    for i in range(20):
        track_once(4 + i)
    d.record()
    torch.cuda.synchronize()
    runner.tracker.reset()
    ms = t.item() / steps
    return dict(metric='video frames/sec (384x1248, 5-frame clips: decoder + panoptic merge + tracking head + association)',
                value=world * F * steps / (t.item() / 1e3), unit='frames/s', ms_per_wave=ms, frames_per_rank_per_wave=F,
                waves=steps, clip_len=clip, decoder_map='%dx%d' % (H, W), collective='one all_gather_into_tensor per wave '
                '(%d B per rank) on a side stream' % (F * (2 + 100 * 262) * 4) if world > 1 else 'none (1 rank)',
                tracked_things_per_frame=things / (F * steps), section_ms_per_wave_synchronised=sections,
                tracking_head_ms_per_frame_30_rois=c.elapsed_time(d) / 20,
                tracking_head_what='pf_track_boxes_from_panoptic + pf_track_embed (RoIAlign, 4 conv+GN, FC, FC) + '
                                   'pf_tracker_match on 30 RoIs, 9 launches')


class StandInBackbone(torch.nn.Module):
    """MEASUREMENT SCAFFOLDING, not product code: the backbone + FPN of poly_r50 stay the reference's PyTorch modules
    (north-star) and mmdet's cannot be imported on the bench box, so `model_fps` runs torchvision's ResNet-50 (the same
    architecture, random init) and a plain PyTorch FPN (mmdet FPN semantics: 1x1 laterals, nearest top-down, 3x3 outputs,
    256 channels, strides 4 / 8 / 16 / 32) in front of this package's heads.  PyTorch eager fp32 (cuDNN), as the reference
    would run them."""

    def __init__(self):
        super().__init__()
        from torchvision.models import resnet50
        r = resnet50(weights=None)
        self.stem = torch.nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool)
        self.layers = torch.nn.ModuleList([r.layer1, r.layer2, r.layer3, r.layer4])
        self.lateral = torch.nn.ModuleList([torch.nn.Conv2d(c, 256, 1) for c in (256, 512, 1024, 2048)])
        self.output = torch.nn.ModuleList([torch.nn.Conv2d(256, 256, 3, padding=1) for _ in range(4)])

    def forward(self, img):
        x, feats = self.stem(img), []
        for layer in self.layers:
            x = layer(x)
            feats.append(x)
        lat = [l(f) for l, f in zip(self.lateral, feats)]
        for i in range(3, 0, -1):
            lat[i - 1] = lat[i - 1] + torch.nn.functional.interpolate(lat[i], size=lat[i - 1].shape[-2:], mode='nearest')
        return [o(l) for o, l in zip(self.output, lat)]


def model_fps(args, dev, world, barrier):
    """SURVEY.md section 8(d) "model fps": the whole image model through the detector API the reference's users call,
    `Polyphonic.simple_test(img, img_metas)` (polyphonic_former.py:130-161) -- images in pinned HOST memory -> backbone +
    FPN (PyTorch, see StandInBackbone) -> SemanticFPN neck + KernelHead + 3-stage decoder + batched panoptic merge (this
    package's kernels) -> panoptic map + segments + two depth maps as numpy on the host.  Random-init weights of the
    poly_r50 architecture; per-GPU batch as the headline.  Wall-clock between barriers (the call synchronises itself when
    it reads the results back), max over ranks."""
    import json as _json
    import torch.distributed as dist
    from polyphonicformer_b200 import registry
    gold = os.path.join(ROOT, 'tests', 'golden')
    pd = _json.load(open(os.path.join(gold, 'rpn_head_cfg.json')))
    rd = _json.load(open(os.path.join(gold, 'roi_head_cfg.json')))
    torch.manual_seed(0)
    model = registry.build_detector(dict(type='Polyphonic', backbone=StandInBackbone(), neck=None, rpn_head=pd['rpn_head'],
                                         roi_head=rd['roi_head'], test_cfg=dict(rpn=pd['test_cfg'], rcnn=rd['test_cfg']),
                                         num_thing_classes=8, num_stuff_classes=11))
    model.rpn_head.init_weights()
    model.roi_head.init_weights()
    model = model.to(dev).eval()
    B, Hi, Wi = args.batch, args.height, args.width
    img_host = torch.randn(B, 3, Hi, Wi).pin_memory()
    metas = [dict(img_shape=(Hi, Wi, 3), ori_shape=(Hi, Wi, 3), pad_shape=(Hi, Wi, 3), scale_factor=1.0, flip=False,
                  batch_input_shape=(Hi, Wi)) for _ in range(B)]

    def step(parts=None):
        t0 = time.perf_counter()
        img = img_host.to(dev, non_blocking=True)
        with torch.no_grad():
            if parts is None:
                return model.simple_test(img, metas)
            x = model.extract_feat(img)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            rpn = model.rpn_head.simple_test_rpn(x, metas)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            out = model.roi_head.simple_test(rpn[1], rpn[0], rpn[2], rpn[3], metas, depth_preds=rpn[7], depth_feats=rpn[5],
                                             depth_proposal=rpn[6])
            t3 = time.perf_counter()
            parts['backbone_fpn_pytorch'] += t1 - t0
            parts['neck_kernel_head'] += t2 - t1
            parts['decoder_panoptic_readback'] += t3 - t2
            return out

    for _ in range(2):
        out = step()
    n = 4
    barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        out = step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    parts = dict(backbone_fpn_pytorch=0.0, neck_kernel_head=0.0, decoder_panoptic_readback=0.0)
    for _ in range(2):
        step(parts)
    segs = [len(o[2][1]) for o in out]
    del model
    torch.cuda.empty_cache()
    return dict(value=world * B * n / t.item(), unit='frames/s', ms_per_step=1e3 * t.item() / n, steps=n, batch_per_gpu=B,
                api='Polyphonic.simple_test(img, img_metas) -> [(None, None, (panoptic, segments_info), depth_basic, depth_final)]',
                h2d_bytes_per_step=img_host.numel() * 4, d2h_bytes_per_step=B * (Hi * Wi * 12 + 128 * 24 + 4),
                section_ms_synchronised={k: 1e3 * v / 2 for k, v in parts.items()}, segments_per_frame=segs,
                backbone='torchvision resnet50 + PyTorch FPN, random init, eager fp32 (stand-in for the reference\'s mmdet ResNet-50 '
                         '+ FPN, which stay PyTorch by north-star and cannot be imported on this box)')


def kernel_head_timing(args, B, dev, peak_gbs, cpu=True):
    """NOT part of the headline metric: the producer of the decoder's inputs, i.e. the
    tail of KernelHead._decode_init_proposals (kernel_head.py:250-336) = pf_kernel_head + pf_mask_pool +
    pf_init_proposals on B frames, device time with CUDA events over rotating input sets (3 x 201 MB of bf16 maps at
    B=4 > L2), next to oracle/kernel_head_ref.py on the host cores (one frame)."""
    from oracle import kernel_head_ref, synth
    from polyphonicformer_b200.kernel_head import KernelHeadTail
    H, W = args.height // 8, args.width // 8
    HW = H * W
    sd = synth.synth_kernel_head_state(0)
    tail = KernelHeadTail(sd, dev)
    g = torch.Generator(device='cpu').manual_seed(7)
    sets = [torch.relu(torch.randn(3, B, C, HW, generator=g)).to(torch.bfloat16).to(dev) for _ in range(3)]
    for i in range(3):
        tail.forward(sets[i % 3], H, W)
    torch.cuda.synchronize()
    n = 20
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        tail.forward(sets[i % 3], H, W)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    # algorithmic bytes per frame: three bf16 maps in, two bf16 maps out, 111 + 19 + 1 fp32 prediction maps out
    alg = B * HW * (3 * C * 2 + 2 * C * 2 + (N_KERNELS + synth.NUM_CLASSES + 1) * 4)
    res = dict(ms_per_call=ms, frames_per_s=B / ms * 1e3, batch=B, launches=tail.last_launches,
               algorithmic_MB=alg / 1e6, achieved_GBps=alg / ms / 1e6, frac_of_measured_hbm_peak=alg / ms / 1e6 / peak_gbs,
               what='pf_kernel_head (1x1 convs + GN + ReLU + prediction heads) + pf_mask_pool + pf_init_proposals, '
                    '%d frames of %dx%d maps; intermediate conv output (fp32, %.0f MB written + read) not counted as '
                    'algorithmic' % (B, H, W, B * 3 * C * HW * 4 / 1e6))
    del sets
    # the neck in front of it: SemanticFPNWrapper's 3x3 conv + GN + ReLU pyramid
    # (pf_semantic_fpn, semantic_fpn.py:198-219) + conv_pred / aux_convs (pf_fpn_pred, :221-229) from the four FPN levels
    from polyphonicformer_b200.kernel_head import FpnPred, SemanticFpnPyramid
    fsd = synth.synth_semantic_fpn_state(0)
    pyr, pred = SemanticFpnPyramid(fsd, dev), FpnPred(fsd, dev)
    levels = [torch.randn(B, C, 2 * H >> i, 2 * W >> i, generator=g).to(dev) for i in range(4)]

    def neck():
        fused_b, _, hw = pyr.forward(levels)
        pred.forward(fused_b, hw=hw)

    for _ in range(2):
        neck()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        neck()
    b.record()
    torch.cuda.synchronize()
    ms_neck = a.elapsed_time(b) / 5
    flops = B * 2 * 9 * C * C * (4 * HW + 2 * HW // 4 + HW // 16) + B * 3 * 2 * C * C * HW
    res['semantic_fpn'] = dict(ms_per_call=ms_neck, frames_per_s=B / ms_neck * 1e3, batch=B, launches=pyr.last_launches + pred.last_launches,
                               algorithmic_GFLOP=flops / 1e9, achieved_TFLOPs=flops / ms_neck / 1e9,
                               issued_TFLOPs_3mma_split=3 * flops / ms_neck / 1e9,
                               conv_kernel_bound='tensor pipe: ncu 86 % active, 1.43 PFLOP/s of issued MMAs in the 128x256 convolutions '
                                                 '(profiles/r2_ncu_full_neck.raw.csv) against %.0f TFLOP/s measured sustained bf16' % peaks()['tf'],
                               what='pf_semantic_fpn (7 x [3x3 conv + GN32 + ReLU], x2 steps, level sum) + pf_fpn_pred on the four '
                                    'FPN levels of %d frames (%dx%d .. %dx%d); split-bf16 MMAs issue 3x the algorithmic FLOP'
                                    % (B, 2 * H, 2 * W, H // 4, W // 4))
    del levels
    if cpu:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        from oracle import semantic_fpn_ref
        lv1 = synth.synth_fpn_inputs(1, H, W, 0)
        with torch.no_grad():
            t0 = time.perf_counter()
            semantic_fpn_ref.semantic_fpn_forward(fsd, lv1)
            res['semantic_fpn']['cpu_port_ms_per_frame'] = 1e3 * (time.perf_counter() - t0)
        maps = [torch.relu(torch.randn(1, C, H, W, generator=g)) for _ in range(3)]
        with torch.no_grad():
            kernel_head_ref.decode_init_proposals(sd, maps)
            t0 = time.perf_counter()
            kernel_head_ref.decode_init_proposals(sd, maps)
            res['cpu_port_ms_per_frame'] = 1e3 * (time.perf_counter() - t0)
            res['cpu_cores'] = cores
    return res


def kernel_breakdown(args, eng, lib, feats, mask, prop, dprop, buf, B, N, H, W, dev):
    from polyphonicformer_b200 import _cabi
    from polyphonicformer_b200.decoder import _ptr, _stream_ptr
    HW, HWp = H * W, feats.shape[-1]
    words = (HW + 31) // 32
    S = lib.pf_pool_splits(B, 2, HW)
    NROT = 4    # rotating buffer sets: 4 x (134 MB features + 58 MB logits ...) > 126 MB L2, so back-to-back launches
    #             of the same kernel never find their inputs in L2 ("inputs larger than L2")
    featsR = [feats] + [feats.clone() for _ in range(NROT - 1)]
    maskR = [mask] + [mask.clone() for _ in range(NROT - 1)]
    bitsR = [torch.empty((B, words, 128), dtype=torch.int32, device=dev) for _ in range(NROT)]
    logitsR = [buf['logits']] + [torch.empty_like(buf['logits']) for _ in range(NROT - 1)]
    scaledR = [buf['scaled']] + [torch.empty_like(buf['scaled']) for _ in range(NROT - 1)]
    partial = torch.empty((2 * B, S, N, C), dtype=torch.float32, device=dev)
    cntp = torch.empty((2 * B, S, N), dtype=torch.float32, device=dev)
    kern = (torch.randn((2 * B, 2, N, C), dtype=torch.float32, device=dev) * 0.1).to(torch.bfloat16)
    kbias = torch.zeros((2, B, N), dtype=torch.float32, device=dev)
    wsb = lib.pf_update_workspace_bytes(B, N, 2048)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    obj, dep = prop.clone(), dprop.clone()
    obj_o, dep_o = torch.empty_like(obj), torch.empty_like(dep)
    cls = torch.empty((B, N, NUM_CLASSES), dtype=torch.float32, device=dev)
    st = _stream_ptr()
    reps = max(8, min(args.steps, 40))
    sx = 2   # bytes / feature element (bf16)
    for i in range(NROT):
        _cabi.call('pf_binarise', _ptr(maskR[i]), _ptr(bitsR[i]), B, N, HW, st)
    # (name, launches per step, bound, algorithmic bytes, algorithmic flops, launcher(i) on buffer set i)
    specs = [
        ('binarise', 1, 'hbm', B * N * HW * 4 + B * words * 128 * 4, 0,
         lambda i: _cabi.call('pf_binarise', _ptr(maskR[i]), _ptr(bitsR[i]), B, N, HW, st)),
        # strict SURVEY 8(d) count: features + mask bits in, pooled [2B, N, C] out; the split-K partials this implementation also
        # writes (2 * B * S * N * C * 4 bytes) are traffic, not algorithmic bytes
        ('mask_pool', STAGES, 'hbm', 2 * B * C * HW * sx + B * words * 128 * 4 + 2 * B * N * C * 4,
         2 * 2 * B * N * C * HW,
         lambda i: _cabi.call('pf_mask_pool', _ptr(featsR[i]), _ptr(bitsR[i]), _ptr(partial), _ptr(cntp), B, N, HW, HWp,
                              2, S, st)),
        ('kernel_update (small-N block, %s)' % ('fused cluster kernel, 1 launch' if args.fused else '12 per-layer launches'),
         STAGES, 'latency', 0, 0,
         lambda i: _cabi.call('pf_kernel_update', ctypes.byref(eng.stages[i % STAGES].struct), _ptr(partial), _ptr(cntp),
                              S, _ptr(obj), _ptr(dep), _ptr(obj_o), _ptr(dep_o), _ptr(cls), None, _ptr(kern),
                              _ptr(kbias), _ptr(ws), wsb, B, N, 0, st)),
        ('mask_einsum (bits only, mask branch)', 0 if args.all_stage_outputs else STAGES - 1, 'hbm',
         B * C * HW * sx + B * N * C * 4 + B * words * 128 * 4, 2 * B * N * C * HW,
         lambda i: _cabi.call('pf_mask_einsum', _ptr(featsR[i]), _ptr(kern), _ptr(kbias), None, _ptr(bitsR[i]), B, N,
                              HW, HWp, B, st)),
        ('mask_einsum (fp32 logits, both branches)', STAGES if args.all_stage_outputs else 1, 'hbm',
         2 * B * C * HW * sx + 2 * B * N * C * 4 + 2 * B * N * HW * 4, 2 * 2 * B * N * C * HW,
         lambda i: _cabi.call('pf_mask_einsum', _ptr(featsR[i]), _ptr(kern), _ptr(kbias), _ptr(logitsR[i]), None, B, N,
                              HW, HWp, 2 * B, st)),
        ('upsample2x', 1, 'hbm', 2 * B * N * HW * 4 * 5, 0,
         lambda i: _cabi.call('pf_upsample2x', _ptr(logitsR[i]), _ptr(scaledR[i]), 2 * B * N, H, W, st)),
    ]
    l2buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []
    for name, calls, bound, nbytes, flops, fn in specs:
        if calls == 0:
            continue
        for i in range(NROT):
            fn(i)
        torch.cuda.synchronize()
        # (a) one launch at a time, L2 flushed before it: includes the full launch ramp / teardown of a cold kernel
        cold = 0.0
        for r in range(min(reps, 10)):
            l2buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(r % NROT)
            b.record()
            torch.cuda.synchronize()
            cold += a.elapsed_time(b)
        cold /= min(reps, 10)
        # (b) `reps` launches back to back over the rotating buffer sets (inputs larger than L2), as inside the step
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for r in range(reps):
            fn(r % NROT)
        b.record()
        torch.cuda.synchronize()
        warm = a.elapsed_time(b) / reps
        ms = cold if bound == 'latency' else warm
        out.append(dict(name=name, calls_per_step=calls, bound=bound, ms=ms, ms_single_launch_l2_flushed=cold,
                        alg_bytes=nbytes, alg_flops=flops,
                        timing=('one launch, L2 flushed' if bound == 'latency' else
                                '%d launches back to back over %d rotating buffer sets (inputs > L2)' % (reps, NROT))))
    # measured DRAM traffic per launch from the committed ncu --set full capture of the same shapes (profiles/)
    ncu_name = {'binarise': ('binarise_kernel', 1), 'mask_pool': ('pool_kernel', 1),
                'mask_einsum (bits only, mask branch)': ('einsum_kernel<0>', 1),
                'mask_einsum (fp32 logits, both branches)': ('einsum_kernel<1>', 2),   # the step launches it per branch
                'upsample2x': ('upsample2x_kernel', 2)}     # the step launches it once per branch
    traffic = {}
    tpath = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
    if os.path.exists(tpath) and (B, H, W) == (4, 128, 256):
        traffic = json.load(open(tpath))['kernels']
    for k in out:
        nm = ncu_name.get(k['name'])
        t = traffic.get(nm[0]) if nm else None
        k['traffic'] = (t['dram_read_bytes'] + t['dram_write_bytes']) * nm[1] if t else None
        if k['bound'] == 'latency':
            # what the block must move whatever its schedule: the stage's weights (bf16 hi + lo) once from L2/HBM per
            # launch; 3-MMA split of 2 * rows * params FLOP per unit
            wbytes = 2 * 4.02e6 * 2 * 2 / 2   # 4.02 M params per stage (both branches), hi + lo bf16 planes
            k['weight_bytes'] = wbytes
            k['weight_stream_GBps'] = wbytes / (k['ms'] * 1e-3) / 1e9
            k['tensor_TFLOPs_issued'] = 3 * 2 * 128 * 4.02e6 * B / (k['ms'] * 1e-3) / 1e12
            k['note'] = ('a chain of ten dependent layers on 128 rows per image: latency-bound, not roofline-bound; the '
                         'stage\'s weights (%.1f MB of bf16 hi + lo) streamed once per launch = %.0f GB/s, issued MMA work '
                         '(3-way split, rows padded to 128) = %.1f TFLOP/s' %
                         (wbytes / 1e6, k['weight_stream_GBps'], k['tensor_TFLOPs_issued']))
    return out


def _time_pipeline(pipe, pin_in, pin_out, steps, dev, world, barrier):
    """Submit `steps` batches back to back; device time from the first H2D to the last D2H, max over ranks."""
    import torch.distributed as dist
    depth = len(pin_in)
    for i in range(3):
        pipe.submit(pin_in[i % depth], pin_out[i % depth])
    pipe.drain()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(pipe.s_in)
    for i in range(steps):
        pipe.submit(pin_in[i % depth], pin_out[i % depth])
    b.record(pipe.s_out)
    pipe.drain()
    barrier()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / steps
    h2d, d2h = pipe.h2d_bytes(), pipe.d2h_bytes()
    # the ceiling the host side sets: the SAME copies (same pinned buffers, same two streams, all ranks at once) with no
    # kernel in between -- what the step would cost if the device work were free
    barrier()
    s = pipe.slots[0]
    src = [pin_in[0][k] for k in ('x', 'd', 'mask')]
    dst = [s['x'], s['d'], s['mask']]
    outs = [(v, torch.empty(v.shape, dtype=v.dtype, device=dev)) for v in pin_out[0].values()]
    c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    c0.record(pipe.s_in)
    pipe.s_out.wait_event(c0)
    for _ in range(steps):
        with torch.cuda.stream(pipe.s_in):
            for a_, b_ in zip(dst, src):
                a_.copy_(b_, non_blocking=True)
        with torch.cuda.stream(pipe.s_out):
            for h_, d_ in outs:
                h_.copy_(d_, non_blocking=True)
    c1.record(pipe.s_in)
    c2.record(pipe.s_out)
    pipe.drain()
    barrier()
    tc = torch.tensor([max(c0.elapsed_time(c1), c0.elapsed_time(c2))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    ms_copy = tc.item() / steps
    return dict(value=world * pipe.B * steps / (t.item() / 1e3), unit='frames/s', h2d_bytes_per_step=h2d,
                d2h_bytes_per_step=d2h, steps=steps, ms_per_step=ms,
                # the two copy directions run on their own streams: each one's rate if it alone filled the step
                h2d_GBps_per_rank=h2d / ms / 1e6, d2h_GBps_per_rank=d2h / ms / 1e6,
                host_GBps_all_ranks=world * (h2d + d2h) / ms / 1e6,
                copies_only=dict(ms_per_step=ms_copy, frames_per_s=world * pipe.B / (ms_copy / 1e3),
                                 host_GBps_all_ranks=world * (h2d + d2h) / ms_copy / 1e6,
                                 what='the same host<->device copies with no kernels, all ranks at once: the ceiling the host '
                                      'link (PCIe / host DRAM) sets for this step'),
                bound='host copies (PCIe / host DRAM): the step runs at %.0f %% of its copies-only ceiling' % (100 * ms_copy / ms))


def run_e2e(args, eng, hin, B, N, H, W, dev, world, barrier):
    """The headline e2e: the call a user of the reference makes, KernelUpdateIterHead.simple_test
    (kernel_update.py:282-354), over pinned HOST buffers through polyphonicformer_b200.decoder.PanopticPipeline --
    every step copies its inputs (x_feats, depth_feats bf16, mask logits, proposal kernels, KernelHead's depth
    prediction) host->device, runs the 3-stage decoder and the batched pf_panoptic (which samples the decoder's stride-8
    logits directly: the x2 / x4 up-sampled maps never exist), and copies panoptic int32 + depth_final + depth_basic
    (12 B per pixel) + the segment records back to pinned HOST memory.  It does MORE than the metric's decoder-only
    workload (the reference arm stops at the logits); `e2e_logits` is the decoder-only variant."""
    from polyphonicformer_b200.decoder import PanopticPipeline
    depth = 2
    g = torch.Generator().manual_seed(77)
    pin_in = []
    for _ in range(depth):
        d = {k: v.clone().pin_memory() for k, v in hin.items()}
        d['depth_pred'] = torch.randn(B, 1, H, W, generator=g).pin_memory()
        pin_in.append(d)
    H0, W0 = 8 * H, 8 * W
    pin_out = [dict(panoptic=torch.empty((B, H0, W0), dtype=torch.int32).pin_memory(),
                    depth_final=torch.empty((B, H0, W0)).pin_memory(), depth_basic=torch.empty((B, H0, W0)).pin_memory(),
                    segments=torch.empty((B, 128, 24), dtype=torch.uint8).pin_memory(),
                    nseg=torch.empty(B, dtype=torch.int32).pin_memory()) for _ in range(depth)]
    pipe = PanopticPipeline(eng, B, N, H, W, depth=depth)
    out = _time_pipeline(pipe, pin_in, pin_out, max(4, min(args.steps, 20)), dev, world, barrier)
    out['api'] = ('PanopticPipeline = KernelUpdateIterHead.simple_test over host buffers: H2D | 3-stage decode + batched '
                  'pf_panoptic on the stride-8 logits | D2H of panoptic + depth_final + depth_basic + segment records, '
                  'neighbouring steps overlapped on 3 streams, 2 device slots')
    return out


def run_e2e_logits(args, eng, hin, B, N, H, W, dev, world, barrier):
    """Decoder-only variant (the metric's own workload, HostPipeline): results = cls scores + the x2-upsampled fp32 mask
    and depth logits, 466 MB per step of 4 frames -- bound by the read-back of those logits."""
    from polyphonicformer_b200.decoder import HostPipeline
    depth = 2
    pin_in = [{k: v.clone().pin_memory() for k, v in hin.items()} for _ in range(depth)]
    pin_out = [dict(cls=torch.empty((B, N, NUM_CLASSES), dtype=torch.float32).pin_memory(),
                    scaled=torch.empty((2, B, N, 2 * H, 2 * W), dtype=torch.float32).pin_memory())
               for _ in range(depth)]
    pipe = HostPipeline(eng, B, N, H, W, upsample=True, depth=depth)
    out = _time_pipeline(pipe, pin_in, pin_out, max(4, min(args.steps, 12)), dev, world, barrier)
    out['api'] = 'HostPipeline: H2D | decode | D2H of cls + scaled_mask_preds + scaled_depth_preds on 3 streams'
    return out


def postprocess_timing(args, dev, cpu=True):
    """NOT part of the headline metric (SURVEY.md section 8d excludes post-processing): pf_panoptic per frame -- the
    reference's get_panoptic (kernel_update.py:421-535) -- on hand-constructed predictions at the frame size, device
    time with CUDA events incl. the D2H of the three result maps; next to oracle/panoptic_ref.py on the host cores."""
    import json as _json
    from types import SimpleNamespace
    from oracle import panoptic_ref, synth
    from polyphonicformer_b200 import postprocess
    from polyphonicformer_b200.registry import to_config
    h, w = args.height // 4, args.width // 4
    cfg = to_config(_json.load(open(os.path.join(ROOT, 'tests', 'golden', 'roi_head_cfg.json')))['test_cfg'])
    roi = SimpleNamespace(num_proposals=synth.N_PROPOSALS, num_thing_classes=synth.NUM_THING, merge_joint=True)
    last = SimpleNamespace(depth_act_mode='sigmoid', num_classes=synth.NUM_CLASSES)
    meta = dict(img_shape=(4 * h, 4 * w, 3), ori_shape=(4 * h, 4 * w, 3), pad_shape=(4 * h, 4 * w, 3), scale_factor=1.0,
                flip=False, batch_input_shape=(4 * h, 4 * w))
    inp = synth.synth_panoptic_inputs(h, w, 0)
    d = {k: v.to(dev) for k, v in inp.items()}
    call = lambda: postprocess.get_panoptic(roi, last, d['cls_scores'], d['mask_preds'], cfg, meta, d['depth_preds'],
                                            d['depth_init'])
    for _ in range(3):
        out = call()
    torch.cuda.synchronize()
    n = 10
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        call()
    b.record()
    torch.cuda.synchronize()
    res = dict(ms_per_frame=a.elapsed_time(b) / n, segments=len(out[2][1]),
               what='pf_panoptic on one %dx%d frame (111 kernels at %dx%d), incl. D2H of panoptic + 2 depth maps' %
                    (4 * h, 4 * w, h, w))
    if cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            t0 = time.perf_counter()
            panoptic_ref.get_panoptic(roi, last, inp['cls_scores'], inp['mask_preds'], cfg, meta, inp['depth_preds'],
                                      inp['depth_init'])
            res['cpu_port_ms_per_frame'] = 1e3 * (time.perf_counter() - t0)
            res['cpu_cores'] = os.cpu_count() or 1
    return res


def cpu_baseline(args):
    from oracle import decoder_ref as ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd, _ = synth_state()
    H, W = args.height // 8, args.width // 8
    inp = host_inputs(1, H, W, 0)
    x, d = inp['x'].float(), inp['d'].float()
    prop, dprop = inp['prop'].reshape(1, N_KERNELS, C, 1, 1), inp['dprop'].reshape(1, N_KERNELS, C, 1, 1)
    with torch.no_grad():
        ref.decoder_forward(sd, x, prop, inp['mask'], d, dprop)
        n, t0 = 0, time.perf_counter()
        while n < 3 or (time.perf_counter() - t0 < 10 and n < 20):
            ref.decoder_forward(sd, x, prop, inp['mask'], d, dprop)
            n += 1
        dt = time.perf_counter() - t0
    return dict(value=n / dt, unit='frames/s', cores=cores, kind='port',
                sample='%d frames of %dx%d, one frame per call, fp32 oracle/decoder_ref.py on %d torch threads' %
                       (n, args.height, args.width, cores))


def library_baseline(args, B, dev):
    """"the same oracle on the B200 via PyTorch as the stronger library-kernels-on-the-same-box
    baseline" -- oracle/decoder_ref.py (the reference's algorithm as written: fp32 feat_transform convs, einsum
    pooling, per-image conv2d, nn.MultiheadAttention, ...) on the SAME GPU through cuBLAS / cuDNN, the same B frames
    per step, inputs resident, CUDA events.  Reported next to `value`; never on the product path."""
    from oracle import decoder_ref as ref
    sd, _ = synth_state()
    sd = {k: v.to(dev) for k, v in sd.items()}
    H, W = args.height // 8, args.width // 8
    inp = host_inputs(B, H, W, 0)
    x, d = inp['x'].float().to(dev), inp['d'].float().to(dev)
    prop = inp['prop'].reshape(B, N_KERNELS, C, 1, 1).to(dev)
    dprop = inp['dprop'].reshape(B, N_KERNELS, C, 1, 1).to(dev)
    mask = inp['mask'].to(dev)
    out = {}
    for tf32 in (False, True):
        old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            with torch.no_grad():
                for _ in range(3):
                    ref.decoder_forward(sd, x, prop, mask, d, dprop)
                torch.cuda.synchronize()
                n = 10
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(n):
                    ref.decoder_forward(sd, x, prop, mask, d, dprop)
                b.record()
                torch.cuda.synchronize()
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        ms = a.elapsed_time(b) / n
        out['tf32' if tf32 else 'fp32'] = dict(ms_per_step=ms, value=B / ms * 1e3)
    return dict(value=out['fp32']['value'], unit='frames/s', ms_per_step=out['fp32']['ms_per_step'],
                tf32_value=out['tf32']['value'], tf32_ms_per_step=out['tf32']['ms_per_step'], batch=B,
                kind='oracle/decoder_ref.py through PyTorch eager (cuBLAS / cuDNN) on the same GPU, fp32 (TF32 off; '
                     'tf32_* = TF32 allowed, which does not meet the 1e-3 gate), inputs resident')


def emit(line):
    """The ONE JSON line goes to the process's original stdout (see main: fd 1 is pointed at stderr meanwhile)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + '\n').encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    # libraries (NCCL prints its version banner on stdout when NCCL_DEBUG is set on the box) must not pollute the
    # single JSON line: keep a private copy of stdout, point fd 1 at stderr for everything else
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
