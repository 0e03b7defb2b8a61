"""GPU parity tests of the LIVE drop-in call and of the configs round 1 left untested ("untested configs"):

* the call DispRCNN3D._forward_eval makes (disprcnn3d.py:266-284 -> stackhourglass.py:106-174): image crops [R,3,224,224]
  through feature_extraction AND the 3-D stack, against the disparity the UNMODIFIED reference PSMNet.forward produced;
* aligned boxes -> fused ROIAlign crops of both views -> PSMNet, end to end against the oracle chain;
* the host-buffer entry point (the e2e path of bench.py) in the split-precision mode at the live and benchmark shapes;
* default-initialised weights;
* N-rank sharded + gathered == unsharded, bit for bit (needs >= 2 GPUs).
Tolerance everywhere: 1e-3 px abs against the reference's fp32 forward (north_star), unless stated.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import idispnet_oracle as O
import recipe
from helpers import load_case, load_psm_case, load_raw_case, make_full_psmnet, make_psmnet

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-3


@pytest.fixture(scope='module')
def lib(built_lib):
    return built_lib


@pytest.fixture(autouse=True)
def _no_tf32():
    """Any torch/cuDNN op these tests still reach must run in fp32 (PyTorch's conv default is TF32)."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize('prec', ['auto', 'fp32'])
def test_whole_psmnet_on_image_crops_matches_reference_forward(lib, prec):
    case, g, sd, L, R = load_psm_case('psm_live')
    m = make_full_psmnet(case, sd, prec)
    assert m.effective_precision(56, 56) == ('fp16x2' if prec == 'auto' else 'fp32')
    with torch.no_grad():
        pred = m({'left': L.cuda(), 'right': R.cuda()})          # as DispRCNN3D calls it (disprcnn3d.py:273)
        pred_seq = m((L.cuda(), R.cuda()))                        # 2-sequence form (stackhourglass.py:110-111)
        fl = m.feature_extraction(L.cuda())
    assert tuple(pred.shape) == (case['R'], case['size'], case['size']) and torch.equal(pred, pred_seq)
    e = np.abs(pred.cpu().numpy() - g['pred'])
    ef = np.abs(fl.cpu().numpy() - g['fea_left']).max()
    e64 = np.abs(pred.cpu().numpy() - g['pred_f64']).max()
    print(f'\n[psm_live] {prec}: |disp - ref_fp32| max {e.max():.3e} mean {e.mean():.3e}; vs float64 arbiter {e64:.3e} '
          f'(reference itself {float(g["ref_f32_vs_f64_maxabs"][0]):.3e}); |features - ref| max {ef:.3e}')
    assert e.max() < TOL
    assert ef < 1e-3


def test_boxes_to_disparity_end_to_end_vs_oracle(lib):
    """disprcnn3d.py:113-159 + :266-284: boxes of a padded 2-image batch -> aligned crop rectangles -> ROIAlign + normalise of
    both views -> PSMNet.  Product: crop_stereo_rois + PSMNet('auto').  Checker: the oracle's chain on the CPU."""
    from disprcnn_b200.layers.roi_align import crop_stereo_rois
    case, g, sd, _, _ = load_psm_case('psm_live')
    Hd, Wd = 120, 400
    base = recipe.make_images(2, Hd, Wd + 16, 77)
    base = torch.nn.functional.avg_pool2d(base, 5, 1, 2)
    iml, imr = base[..., 8:8 + Wd].contiguous(), base[..., 0:Wd].contiguous()     # right view = left shifted by 8 px
    lbs = [[[30.2, 10.7, 150.9, 100.1]], [[200.5, 5.0, 395.2, 118.9]]]
    rbs = [[[22.4, 10.7, 140.0, 100.1]], [[190.1, 5.0, 380.7, 118.9]]]
    sizes = [(Wd, Hd), (Wd - 20, Hd - 6)]                                           # second image smaller than the padded tensor
    want_l, want_r = O.align_stereo_boxes(lbs, rbs, [s[0] for s in sizes], [s[1] for s in sizes])
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    with torch.no_grad():
        cl = torch.from_numpy(O.crop_and_transform_roi_img(iml.numpy(), np.asarray(want_l, np.float32), 224))
        cr = torch.from_numpy(O.crop_and_transform_roi_img(imr.numpy(), np.asarray(want_r, np.float32), 224))
        want = O.psmnet_forward(cl, cr, sd, case['mindisp'], case['maxdisp']).numpy()
    m = make_full_psmnet(case, sd, 'auto')
    LB = torch.tensor([b for im in lbs for b in im]).cuda()
    RB = torch.tensor([b for im in rbs for b in im]).cuda()
    IDX = torch.tensor([0, 1]).cuda()
    with torch.no_grad():
        gl, gr, x1s, x1ps, x2s, x2ps = crop_stereo_rois(iml.cuda(), imr.cuda(), LB, RB, IDX, 224, image_sizes=sizes)
        got = m({'left': gl, 'right': gr}).cpu().numpy()
    assert np.array_equal(gl.cpu().numpy(), cl.numpy()) and np.array_equal(gr.cpu().numpy(), cr.numpy())   # crops bit-exact
    assert x2s.tolist() == [r[3] for r in want_l] and x1ps.tolist() == [r[1] for r in want_r]
    e = np.abs(got - want)
    print(f'\n[boxes -> disparity] max |d| {e.max():.3e} mean {e.mean():.3e} (disp range {want.min():.1f}..{want.max():.1f})')
    assert e.max() < TOL


@pytest.mark.parametrize('name,prec', [('live', 'fp16x2'), ('full', 'fp16x2'), ('live', 'fp32')])
def test_host_buffer_entry_point_parity(lib, name, prec):
    """idisp_plan_forward_host -- the call bench.py's e2e number goes through -- against the reference's own output, and
    bit-identical to the device-pointer entry."""
    from disprcnn_b200 import _lib
    case, g, sd, L, R = load_case(name)
    m = make_psmnet(case, sd, prec)
    B, Hf, Wf = case['B'], case['Hf'], case['Wf']
    with torch.no_grad():
        dev = m.forward_features(L.cuda(), R.cuda()).cpu()
    Lp, Rp = L.pin_memory(), R.pin_memory()
    out = torch.full((B, 4 * Hf, 4 * Wf), float('nan')).pin_memory()
    for _ in range(2):   # second call reuses the plan-owned staging
        _lib.check(lib.idisp_plan_forward_host(m._plan, _lib.ptr(Lp), _lib.ptr(Rp), B, Hf, Wf, 4 * Hf, 4 * Wf, _lib.ptr(out), _lib.stream_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(out, dev)
    e = np.abs(out.numpy() - g['pred_up']).max()
    print(f'\n[{name}] {prec} through idisp_plan_forward_host: max |disp - ref_fp32| {e:.3e}')
    assert e < TOL


def test_pipelined_host_entry_point_matches_one_batch_at_a_time(lib):
    """idisp_plan_forward_host_async / idisp_plan_host_wait (the call bench.py's e2e number goes through): five batches in
    flight back to back -- two staging slots, so slots are reused while earlier copies may still be running -- each result
    bit-identical to the device-pointer entry on the same inputs, whatever order the host waits in."""
    import ctypes
    from disprcnn_b200 import _lib
    case, g, sd, L, R = load_case('live')
    m = make_psmnet(case, sd, 'fp16x2')
    B, Hf, Wf = case['B'], case['Hf'], case['Wf']
    scales = [1.0, 0.5, 2.0, 0.25, 1.5]
    ins = [((L * a).contiguous().pin_memory(), (R * a).contiguous().pin_memory()) for a in scales]
    with torch.no_grad():
        want = [m.forward_features(a.cuda(), b.cuda()).cpu() for a, b in ins]
    outs = [torch.full((B, 4 * Hf, 4 * Wf), float('nan')).pin_memory() for _ in scales]
    tickets = []
    for (a, b), o in zip(ins, outs):
        t = ctypes.c_ulonglong()
        _lib.check(lib.idisp_plan_forward_host_async(m._plan, _lib.ptr(a), _lib.ptr(b), B, Hf, Wf, 4 * Hf, 4 * Wf, _lib.ptr(o),
                                                     _lib.stream_ptr(), ctypes.byref(t)))
        tickets.append(t.value)
    assert tickets == list(range(tickets[0], tickets[0] + len(scales)))
    for k in (2, 0, 4, 1, 3):
        _lib.check(lib.idisp_plan_host_wait(m._plan, tickets[k]))
        assert torch.equal(outs[k], want[k]), f'batch {k}'
    assert np.abs(outs[0].numpy() - g['pred_up']).max() < TOL
    with pytest.raises(RuntimeError):
        _lib.check(lib.idisp_plan_host_wait(m._plan, tickets[-1] + 1))   # never issued
    # a larger batch re-allocates the double-buffered staging (drains what is in flight first); tickets keep counting
    L2, R2 = torch.cat([L, L * 0.5], 0).contiguous().pin_memory(), torch.cat([R, R * 0.5], 0).contiguous().pin_memory()
    with torch.no_grad():
        want2 = m.forward_features(L2.cuda(), R2.cuda()).cpu()
    out2 = torch.full((2 * B, 4 * Hf, 4 * Wf), float('nan')).pin_memory()
    t = ctypes.c_ulonglong()
    _lib.check(lib.idisp_plan_forward_host_async(m._plan, _lib.ptr(L2), _lib.ptr(R2), 2 * B, Hf, Wf, 4 * Hf, 4 * Wf, _lib.ptr(out2),
                                                 _lib.stream_ptr(), ctypes.byref(t)))
    assert t.value == tickets[-1] + 1
    _lib.check(lib.idisp_plan_host_wait(m._plan, t.value))
    assert torch.equal(out2, want2)
    torch.cuda.synchronize()


@pytest.mark.parametrize('name', ['raw_tiny', 'raw_live'])
def test_default_initialised_weights(lib, name):
    """Raw default initialisation (stackhourglass.py:90-104): logits of std ~15, where the reference's OWN fp32 forward is
    up to 2.5e-3 px from its float64 twin.  Bar: no further from the float64 arbiter than the reference is, plus 1e-3."""
    case, g, sd, L, R = load_raw_case(name)
    ref64 = float(g['ref_f32_vs_f64_maxabs'][0])
    for prec in ('fp32', 'fp16x2', 'auto'):
        m = make_psmnet(case, sd, prec)
        with torch.no_grad():
            up = m.forward_features(L.cuda(), R.cuda()).cpu().numpy()
            gen = m((L.cuda(), R.cuda())).cpu().numpy()
        e32, e64 = np.abs(up - g['pred_up']).max(), np.abs(up - g['pred_up_f64']).max()
        eg = np.abs(gen - g['pred_genuine_f64']).max()
        print(f'\n[{name}] {prec}: |disp - ref_fp32| {e32:.3e}, |disp - ref_fp64| {e64:.3e} (genuine {eg:.3e}); reference fp32-vs-fp64 {ref64:.3e}')
        assert e64 < ref64 + TOL and eg < ref64 + TOL
        assert e32 < 2 * ref64 + TOL


_SHARD_SCRIPT = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'tests')); sys.path.insert(0, os.path.join(%(root)r, 'tests', 'golden'))
import torch, torch.distributed as dist
from helpers import load_case, make_psmnet
from disprcnn_b200.parallel import sharded_forward
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
case, g, sd, L, R = load_case('live')
L, R = torch.cat([L, L.flip(0), L * 0.5]), torch.cat([R, R.flip(0), R * 0.5])      # B = 6 (B = 5 below: ragged shards)
for prec in ('fp16x2', 'fp32'):
    m = make_psmnet(case, sd, prec, device=f'cuda:{rank}')
    for B in (6, 5):
        Ld, Rd = L[:B].cuda(), R[:B].cuda()
        with torch.no_grad():
            whole = m.forward_features(Ld, Rd)                 # unsharded, on this rank
            got = sharded_forward(m, Ld, Rd)                   # this rank's shard + all-gather
        assert got.shape == whole.shape, (got.shape, whole.shape)
        assert torch.equal(got, whole), f'rank {rank} {prec} B={B}: sharded != unsharded, max |d| {(got - whole).abs().max().item()}'
dist.barrier()
if rank == 0:
    print('SHARD_OK')
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')
def test_two_rank_sharded_forward_equals_unsharded_bit_for_bit(lib, tmp_path):
    """parallel.sharded_forward over 2 NCCL ranks (equal and ragged shards) == the unsharded forward on every rank."""
    script = tmp_path / 'shard_check.py'
    script.write_text(_SHARD_SCRIPT % {'root': ROOT})
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
                        '--master-port', '29731', str(script)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and 'SHARD_OK' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def _graph_stats(m):
    import ctypes
    from disprcnn_b200 import _lib
    c, r = ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(_lib.load().idisp_plan_graph_stats(m._plan, ctypes.byref(c), ctypes.byref(r)))
    return c.value, r.value


def test_cuda_graph_replay_is_bit_identical_to_eager_launches(lib, monkeypatch):
    """The conv section is captured once per (B, Hf, Wf, workspace) and replayed: same bits as the eager launch sequence,
    for a second batch size as well, and again after the weights change (the captured launches hold weight pointers)."""
    case, g, sd, L, R = load_case('tiny')
    Lc, Rc = L.cuda(), R.cuda()
    monkeypatch.setenv('IDISP_NO_GRAPH', '1')     # read when a plan is created, i.e. at a model's first forward
    eager = make_psmnet(case, sd, 'fp16x2')
    with torch.no_grad():
        want2, want1 = eager.forward_features(Lc, Rc), eager.forward_features(Lc[:1], Rc[:1])
    monkeypatch.delenv('IDISP_NO_GRAPH')
    m = make_psmnet(case, sd, 'fp16x2')
    with torch.no_grad():
        assert _graph_stats(eager) == (0, 0)
        a = m.forward_features(Lc, Rc)           # captures
        b = m.forward_features(Lc, Rc)           # replays
        c = m.forward_features(Lc[:1], Rc[:1])   # second shape: its own graph
        d = m.forward_features(Lc, Rc)
    assert torch.equal(a, want2) and torch.equal(b, want2) and torch.equal(d, want2) and torch.equal(c, want1)
    caps, reps = _graph_stats(m)
    assert caps == 2 and reps == 2, (caps, reps)
    with torch.no_grad():
        m.classif3[2].weight.mul_(0.5)           # in-place edit -> plan re-finalised -> graphs dropped
        eager.classif3[2].weight.mul_(0.5)
        e = m.forward_features(Lc, Rc)
        assert torch.equal(e, eager.forward_features(Lc, Rc)) and not torch.equal(e, want2)


@pytest.mark.parametrize('name', ['live', 'tiny'])
def test_input_side_copy_is_bit_identical_to_the_producer_written_parity_copy(lib, monkeypatch, name):
    """The parity-layout copy of out_k that the next hourglass's stride-2 conv reads: written by classif{k}.0 from its shared-memory
    input stages (default) vs by conv6's own epilogue (IDISP_NO_SIDE_COPY=1, read when a plan is created).  Pure data movement, so
    the disparity must be bit-identical -- and within the parity tolerance of the reference."""
    case, g, sd, L, R = load_case(name)
    Lc, Rc = L.cuda(), R.cuda()
    with torch.no_grad():
        base = make_psmnet(case, sd, 'fp16x2').forward_features(Lc, Rc)
        monkeypatch.setenv('IDISP_NO_SIDE_COPY', '1')
        a = make_psmnet(case, sd, 'fp16x2').forward_features(Lc, Rc)
        monkeypatch.delenv('IDISP_NO_SIDE_COPY')
    assert torch.equal(a, base)
    assert np.abs(base.cpu().numpy() - g['pred_up']).max() < TOL


def test_depth_split_tail_columns_match_whole_columns_bit_for_bit(lib):
    """With more tile columns than CTAs and a partial last round, the per-step kernels split the last round's columns in depth
    (conv3d_tc.cu, Items).  Eleven ROI pairs at the KITTI shape: 308 full-resolution columns over 148 CTAs -> 12 tail columns in 4
    chunks (stride-1 kernels); 88 half-resolution columns over 74 CTAs per channel slice -> 14 tail columns in 2 chunks (the stride-2
    32->64 conv and the K-split 64->64 convs).  A chunk accumulates each of its planes from the same input planes in the same order,
    so every ROI must equal its single-ROI run (no split) bit for bit."""
    case, g, sd, L, R = load_case('live')
    m = make_psmnet(case, sd, 'fp16x2')
    Ls, Rs = [L, L.flip(-1) * 0.5, R * 1.5, L.flip(-2), R.flip(-2) * 0.75, L[:1] * 1.25], [R, R.flip(-1) * 0.5, L * 1.5, R.flip(-2), L.flip(-2) * 0.75, R[:1] * 1.25]
    LB, RB = torch.cat(Ls, 0).cuda(), torch.cat(Rs, 0).cuda()
    assert LB.shape[0] == 11
    with torch.no_grad():
        full = m.forward_features(LB, RB)
        for i in range(11):
            assert torch.equal(m.forward_features(LB[i:i + 1], RB[i:i + 1]), full[i:i + 1]), i
    assert np.abs(full[:2].cpu().numpy() - g['pred_up']).max() < TOL


def test_module_survives_deepcopy_and_pickle_after_a_forward(lib, tmp_path):
    import copy
    case, g, sd, L, R = load_case('tiny')
    m = make_psmnet(case, sd, 'auto')
    with torch.no_grad():
        want = m.forward_features(L.cuda(), R.cuda())
        m2 = copy.deepcopy(m)
        torch.save(m, tmp_path / 'm.pth')
        m3 = torch.load(tmp_path / 'm.pth', weights_only=False)
        assert m2._plan is None and m3._plans == {}
        assert torch.equal(m2.forward_features(L.cuda(), R.cuda()), want)
        assert torch.equal(m3.forward_features(L.cuda(), R.cuda()), want)
    del m, m2, m3   # three independent plans, each destroyed once


def test_two_streams_do_not_share_a_workspace(lib):
    case, g, sd, L, R = load_case('tiny')
    m = make_psmnet(case, sd, 'fp16x2')
    Lc, Rc = L.cuda(), R.cuda()
    with torch.no_grad():
        want = m.forward_features(Lc, Rc)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        torch.cuda.synchronize()
        outs = []
        for _ in range(3):
            for s in (s1, s2):
                with torch.cuda.stream(s):
                    outs.append(m.forward_features(Lc, Rc))
        torch.cuda.synchronize()
    assert len(m._workspaces) == 3 and all(torch.equal(o, want) for o in outs)


def test_deferred_range_check_reports_at_the_next_call(lib):
    case, g, sd, L, R = load_case('tiny')
    m = make_psmnet(case, sd, 'auto')
    m.check_range = 'deferred'
    f32 = make_psmnet(case, sd, 'fp32')
    Lb, Rb = (L * 3e5).cuda(), (R * 3e5).cuda()   # beyond the IEEE-half range
    with torch.no_grad():
        ok = m.forward_features(L.cuda(), R.cuda())
        assert not m.range_exceeded()
        m.forward_features(Lb, Rb)                  # no sync, no warning yet
        with pytest.warns(UserWarning, match='call #2'):
            again = m.forward_features(L.cuda(), R.cuda())
        assert m._sticky_fp32 and torch.equal(again, f32.forward_features(L.cuda(), R.cuda()))
    assert (ok - again).abs().max().item() < TOL


def test_feature_extraction_native_kernels_match_reference(lib):
    """submodule.py:60-139 inside libidisp (csrc/feature2d.cu) against the features the UNMODIFIED reference produced for the
    same crops (psm_live golden), against the oracle at a non-square size, and -- as a cross-check -- the torch modules."""
    case, g, sd, L, R = load_psm_case('psm_live')
    m = make_full_psmnet(case, sd, 'auto')
    fe = m.feature_extraction
    assert fe.native and fe.precision == 'auto'
    for prec, tol in (('fp32', 2e-4), ('auto', 5e-4)):   # FFMA kernels everywhere / stride-1 3x3 convs on tcgen05 in split precision
        fe.precision = prec
        with torch.no_grad():
            fl, fr = fe(L.cuda()), fe(R.cuda())
            both = fe(torch.cat([L, R]).cuda())
        el, er = np.abs(fl.cpu().numpy() - g['fea_left']).max(), np.abs(fr.cpu().numpy() - g['fea_right']).max()
        e64 = np.abs(fl.cpu().numpy() - g['fea_left_f64']).max()
        print(f'\n[extractor {prec}] |features - ref_fp32| max {el:.3e} / {er:.3e} (|ref| max {np.abs(g["fea_left"]).max():.2f}); vs float64 {e64:.3e}')
        assert el < tol and er < tol
        assert torch.equal(both[:2], fl) and torch.equal(both[2:], fr)          # images are independent
    # other sizes: H/4 x W/4 = 58 x 66 -> SPP pools 1x1, 1x2, 3x4, 7x8 (floor), bilinear upsample from each
    g2 = torch.Generator().manual_seed(3)
    x = torch.randn(1, 3, 232, 264, generator=g2)
    fsd = {k[len('feature_extraction.'):]: v for k, v in sd.items() if k.startswith('feature_extraction.')}
    with torch.no_grad():
        want = O.feature_extraction(x, fsd, prefix='').numpy()
        got = fe(x.cuda()).cpu().numpy()
    # (unnormalised noise input: features up to ~30; the split-precision tensor-core convs sit ~3e-5 relative from the fp32 sum)
    assert got.shape == (1, 32, 58, 66) and np.abs(got - want).max() < 5e-5 * np.abs(want).max(), np.abs(got - want).max()
    fe.precision = 'fp32'
    with torch.no_grad():
        got32 = fe(x.cuda()).cpu().numpy()
    fe.precision = 'auto'
    assert np.abs(got32 - want).max() < 2e-5 * np.abs(want).max()
    fe.native = False
    with torch.no_grad():
        via_torch = fe(x.cuda()).cpu().numpy()
    fe.native = True
    assert np.abs(via_torch - want).max() < 1e-3   # (cuDNN picks its own algorithms: Winograd variants sit a few 1e-4 from the direct sum)
    with pytest.raises(RuntimeError, match='56x56'):
        fe(torch.zeros(1, 3, 128, 128).cuda())                             # feature map smaller than branch1's pool
    with pytest.raises(RuntimeError, match='no CPU path'):
        fe(torch.zeros(1, 3, 224, 224))
    assert fe(torch.zeros(0, 3, 224, 224).cuda()).shape == (0, 32, 56, 56)


def test_roi_disparity_handoff_kernels(lib):
    """csrc/roi_paste.cu against the reference-executed fixture and the oracle (disprcnn3d.py:161-190, point_rcnn.py:113-136).
    Floating point (bilinear weights, one division): 1e-4 abs on disparities of magnitude up to ~220, 1e-3 relative on depths < 1000."""
    from disprcnn_b200.layers.roi_disparity import paste_roi_disparity, roi_depth_maps
    case = recipe.PASTE_CASES['paste_small']
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'paste_small.npz'))
    disp, lbs, rbs, masks, fub = recipe.make_paste_inputs(case)
    H, W = case['H'], case['W']
    LB = torch.tensor([b for im in lbs for b in im]).cuda()
    RB = torch.tensor([b for im in rbs for b in im]).cuda()
    counts = [len(im) for im in lbs]
    got = paste_roi_disparity(disp.cuda(), LB, RB, counts, H, W, masks.cuda()).cpu().numpy()
    e = np.abs(got - g['disparity_maps']).max()
    depth = roi_depth_maps(disp.cuda(), LB, RB, fub.cuda(), H, W).cpu().numpy()
    rel = np.abs(depth - g['depth_maps']) / np.maximum(np.abs(g['depth_maps']), 1e-3)
    print(f'\n[paste] disparity maps max |d| {e:.3e} (max {g["disparity_maps"].max():.1f}); depth maps max rel {rel.max():.3e}')
    assert got.shape == g['disparity_maps'].shape and e < 1e-4          # values up to ~220: 1.4e-7 relative
    # depth = fu*b / (disp + 1e-6) amplifies the disparity's rounding where |disp| is small: compare where the depth is physical
    phys = np.abs(g['depth_maps']) < 1e3
    assert (depth == 0).sum() == (g['depth_maps'] == 0).sum() and rel[phys].max() < 1e-3
    # no masks; bool masks; one image without ROIs; nothing at all
    nomask = paste_roi_disparity(disp.cuda(), LB, RB, counts, H, W).cpu()
    want = O.roi_disp_postprocess(disp, lbs, rbs, None, H, W)
    assert (nomask - want).abs().max().item() < 1e-4
    assert torch.equal(paste_roi_disparity(disp.cuda(), LB, RB, counts, H, W, masks.bool().cuda()).cpu(), torch.from_numpy(got))
    assert paste_roi_disparity(disp[:0].cuda(), LB[:0], RB[:0], [], H, W).shape == (0, H, W)
    assert float(paste_roi_disparity(disp[:0].cuda(), LB[:0], RB[:0], [0, 0], H, W).abs().max()) == 0.0
    assert roi_depth_maps(disp[:0].cuda(), LB[:0], RB[:0], fub[:0].cuda(), H, W).shape == (0, H, W)
    with pytest.raises(RuntimeError):
        paste_roi_disparity(disp.cuda(), LB, RB, [1, 1], H, W)
    # random boxes at the KITTI image size with the live 224 x 224 ROI maps
    gen = torch.Generator().manual_seed(9)
    H2, W2, R2 = 375, 1242, 7
    x1 = torch.rand(R2, generator=gen) * 900; y1 = torch.rand(R2, generator=gen) * 200
    w = 20 + torch.rand(R2, generator=gen) * 300; h = 20 + torch.rand(R2, generator=gen) * 150
    off = torch.rand(R2, generator=gen) * 60
    lb = torch.stack([x1, y1, (x1 + w).clamp(max=W2 - 1), (y1 + h).clamp(max=H2 - 1)], 1)
    rb = torch.stack([(x1 - off).clamp(min=0), y1, (x1 - off + w * 1.1).clamp(max=W2 - 1), (y1 + h).clamp(max=H2 - 1)], 1)
    d2 = torch.randn(R2, 224, 224, generator=gen) * 10
    m2 = (torch.rand(R2, H2, W2, generator=gen) > 0.5)
    counts2 = [3, 4]
    want2 = O.roi_disp_postprocess(d2, [lb[:3].tolist(), lb[3:].tolist()], [rb[:3].tolist(), rb[3:].tolist()], m2, H2, W2)
    got2 = paste_roi_disparity(d2.cuda(), lb.cuda(), rb.cuda(), counts2, H2, W2, m2.cuda()).cpu()
    assert (got2 - want2).abs().max().item() < 1e-4
