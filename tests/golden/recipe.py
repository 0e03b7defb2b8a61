"""Deterministic recipes for test inputs and weights (shared by the golden generator and the tests).

Nothing here touches the reference or the oracle: tensors are pure functions of
(shape, seed) through ``torch.Generator`` on the CPU, so the GPU box regenerates exactly
what the authoring container used.  Every golden file also stores a checksum of each
regenerated tensor; a mismatch fails the test loudly instead of comparing garbage.
"""
import math
import zlib

import numpy as np
import torch


def _gen(seed, key):
    g = torch.Generator()
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
    return g


def make_state_dict(shapes, seed, classif_damp=0.1):
    """shapes: {state_dict key: shape}.  'Trained-like' weights:

    He-normal conv kernels, BN affine parameters jittered around (1, 0), identity BN
    running statistics (the generator calibrates them and stores the calibrated values),
    final classifier kernels damped so the soft-argmin is smooth.
    """
    sd = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        g = _gen(seed, key)
        if key.endswith('num_batches_tracked'):
            t = torch.zeros(shape, dtype=torch.int64)
        elif key.endswith('running_mean'):
            t = torch.zeros(shape)
        elif key.endswith('running_var'):
            t = torch.ones(shape)
        elif len(shape) >= 4:  # conv / deconv kernel
            ksz = int(np.prod(shape[2:]))
            transposed = ('.conv5.' in key) or ('.conv6.' in key)
            cout = shape[1] if transposed else shape[0]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / (ksz * cout))
            if key.startswith('classif') and key.endswith('.2.weight'):
                t = t * classif_damp
        elif key.endswith('.weight'):  # BN gamma
            t = 0.8 + 0.4 * torch.rand(shape, generator=g)
        elif key.endswith('.bias'):  # BN beta
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            raise KeyError(key)
        sd[key] = t
    return sd


def make_features(B, C, Hf, Wf, seed, relu=True):
    gl, gr = _gen(seed, 'left'), _gen(seed, 'right')
    L = torch.randn(B, C, Hf, Wf, generator=gl)
    R = torch.randn(B, C, Hf, Wf, generator=gr)
    if relu:
        L, R = L.relu(), R.relu()
    return L.contiguous(), R.contiguous()


def make_images(B, H, W, seed):
    g = _gen(seed, 'images')
    return torch.rand(B, 3, H, W, generator=g).contiguous()


def checksum(t):
    a = np.ascontiguousarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t)
    return np.array([zlib.crc32(a.tobytes())], dtype=np.int64)


# 3-D stack layer table: state_dict prefix -> (kind, cin, cout); cin of the first layer is 2C.
def stack3d_shapes(C):
    shapes = {}

    def convbn(p, cin, cout, transposed=False):
        shapes[p + '.0.weight'] = (cin, cout, 3, 3, 3) if transposed else (cout, cin, 3, 3, 3)
        for k in ('weight', 'bias', 'running_mean', 'running_var'):
            shapes[p + '.1.' + k] = (cout,)
        shapes[p + '.1.num_batches_tracked'] = ()

    convbn('dres0.0', 2 * C, 32)
    convbn('dres0.2', 32, 32)
    convbn('dres1.0', 32, 32)
    convbn('dres1.2', 32, 32)
    for h in ('dres2', 'dres3', 'dres4'):
        convbn(h + '.conv1.0', 32, 64)
        convbn(h + '.conv2', 64, 64)
        convbn(h + '.conv3.0', 64, 64)
        convbn(h + '.conv4.0', 64, 64)
        convbn(h + '.conv5', 64, 64, True)
        convbn(h + '.conv6', 64, 32, True)
    for c in ('classif1', 'classif2', 'classif3'):
        convbn(c + '.0', 32, 32)
        shapes[c + '.2.weight'] = (1, 32, 3, 3, 3)
    return shapes


CASES = {
    # name: B, C, Hf, Wf, mindisp, maxdisp, seed
    'tiny': dict(B=2, C=32, Hf=16, Wf=16, mindisp=-16, maxdisp=16, seed=11),
    'tiny_pos': dict(B=1, C=32, Hf=12, Wf=20, mindisp=0, maxdisp=32, seed=12),
    'c1': dict(B=1, C=16, Hf=64, Wf=64, mindisp=-48, maxdisp=48, seed=13),
    # the shape tools/test_net.py really runs (KITTI configs: 224x224 crops -> 56x56x32ch features, D=24 -> 224x224), 2 ROI pairs
    'live': dict(B=2, C=32, Hf=56, Wf=56, mindisp=-48, maxdisp=48, seed=15),
    # one ROI pair of BASELINE.json configs[1] (the benchmark shape): 112x112x32ch, D=48 -> 448x448
    'full': dict(B=1, C=32, Hf=112, Wf=112, mindisp=-96, maxdisp=96, seed=14),
}

ROI_CASES = {
    # name: (N, C, H, W, pooled_h, pooled_w, spatial_scale, sampling_ratio, rois)
    'kitti_int': dict(N=2, C=3, H=94, W=310, ph=56, pw=56, scale=1.0, sr=0, seed=21, rois=[
        [0, 10, 5, 66, 61], [0, 100, 20, 290, 93], [1, 0, 0, 309, 93], [1, 200, 40, 256, 68],
        [0, 30, 10, 30 + 224, 10 + 80], [1, 5, 60, 12, 70]]),
    'frac_sr2': dict(N=2, C=8, H=50, W=76, ph=7, pw=7, scale=0.25, sr=2, seed=22, rois=[
        [0, 12.3, 7.9, 180.2, 150.5], [1, 0.0, 0.0, 303.9, 199.9], [0, 250.1, 100.7, 320.0, 220.0],
        [1, 40.0, 40.0, 40.4, 40.2], [0, -20.5, -9.0, 60.0, 50.0]]),
    'border': dict(N=1, C=4, H=33, W=47, ph=14, pw=9, scale=0.5, sr=0, seed=23, rois=[
        [0, 0, 0, 93, 65], [0, 80, 50, 120, 90], [0, 92, 64, 94, 66], [0, -10, -10, 4, 4]]),
}
