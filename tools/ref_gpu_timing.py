"""Reference GPU path on this box: the iDispNet stack in eager PyTorch (cuDNN), timed like SURVEY.md 8(d) asks.

This is a MEASUREMENT of the baseline, not part of the product and not the parity oracle: a self-contained torch.nn
restatement of the layer table in SURVEY.md Appendix A (stackhourglass.py:54-174 semantics: concat cost volume, dres0/1,
three hourglasses, three classifier heads with running sums, trilinear upsample, softmax, disparity regression), random
weights (timing and fp32-vs-tf32 deviation only).  Variants: cuDNN with allow_tf32 False / True (PyTorch's default for
convolutions is True), cost volume built "as written" (CPU zeros + H2D, stackhourglass.py:117) or on the device.

usage: python tools/ref_gpu_timing.py [--batch 32] > profiles/r01_reference_gpu.json
"""
import argparse
import json
import time

import torch
import torch.nn as nn
import torch.nn.functional as F


def cbn(cin, cout, stride=1):
    return nn.Sequential(nn.Conv3d(cin, cout, 3, stride, 1, bias=False), nn.BatchNorm3d(cout))


def dbn(cin, cout):
    return nn.Sequential(nn.ConvTranspose3d(cin, cout, 3, 2, 1, 1, bias=False), nn.BatchNorm3d(cout))


class Hourglass(nn.Module):
    def __init__(self):
        super().__init__()
        self.c1, self.c2, self.c3, self.c4 = cbn(32, 64, 2), cbn(64, 64), cbn(64, 64, 2), cbn(64, 64)
        self.c5, self.c6 = dbn(64, 64), dbn(64, 32)

    def forward(self, x, presqu, postsqu):
        out = F.relu(self.c1(x))
        pre = self.c2(out)
        pre = F.relu(pre + postsqu) if postsqu is not None else F.relu(pre)
        out = F.relu(self.c4(F.relu(self.c3(pre))))
        post = F.relu(self.c5(out) + (presqu if presqu is not None else pre))
        return self.c6(post), pre, post


class Stack(nn.Module):
    def __init__(self, C, mindisp, maxdisp):
        super().__init__()
        self.mind, self.maxd = mindisp, maxdisp
        self.dres0 = nn.Sequential(cbn(2 * C, 32), nn.ReLU(), cbn(32, 32), nn.ReLU())
        self.dres1 = nn.Sequential(cbn(32, 32), nn.ReLU(), cbn(32, 32))
        self.hg = nn.ModuleList([Hourglass() for _ in range(3)])
        self.cls = nn.ModuleList([nn.Sequential(cbn(32, 32), nn.ReLU(), nn.Conv3d(32, 1, 3, 1, 1, bias=False)) for _ in range(3)])

    def cost_volume(self, L, R, on_cpu):
        B, C, H, W = L.shape
        D = (self.maxd - self.mind) // 4
        cost = torch.zeros(B, 2 * C, D, H, W, device='cpu' if on_cpu else L.device)
        if on_cpu:
            cost = cost.to(L.device)
        for k, i in enumerate(range(self.mind // 4, self.maxd // 4)):
            if i > 0:
                cost[:, :C, k, :, i:] = L[:, :, :, i:]
                cost[:, C:, k, :, i:] = R[:, :, :, :-i]
            elif i == 0:
                cost[:, :C, k] = L
                cost[:, C:, k] = R
            else:
                cost[:, :C, k, :, :i] = L[:, :, :, :i]
                cost[:, C:, k, :, :i] = R[:, :, :, -i:]
        return cost

    def forward(self, L, R, H, W, cv_on_cpu=False):
        cost = self.cost_volume(L, R, cv_on_cpu)
        c0 = self.dres0(cost)
        c0 = self.dres1(c0) + c0
        o1, pre1, post1 = self.hg[0](c0, None, None)
        o1 = o1 + c0
        o2, _, post2 = self.hg[1](o1, pre1, post1)
        o2 = o2 + c0
        o3, _, _ = self.hg[2](o2, pre1, post2)
        o3 = o3 + c0
        k1 = self.cls[0](o1)
        k2 = self.cls[1](o2) + k1
        k3 = self.cls[2](o3) + k2
        up = F.interpolate(k3, [self.maxd - self.mind, H, W], mode='trilinear', align_corners=True).squeeze(1)
        p = F.softmax(up, dim=1)
        disp = torch.arange(self.mind, self.maxd, device=p.device, dtype=p.dtype).view(1, -1, 1, 1)
        return (p * disp).sum(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--chunk', type=int, default=8, help='ROIs per forward (the eager path materialises 3 x 154 MB per ROI at the output)')
    ap.add_argument('--steps', type=int, default=3)
    a = ap.parse_args()
    print(json.dumps(measure(a.batch, a.chunk, a.steps), indent=1))


def measure(batch=32, chunk=8, steps=3, as_written=True, device=None):
    """Time the eager torch + cuDNN stack on `device`; returns the dict this tool prints (bench.py calls this for its
    `reference_gpu` entry)."""
    class _A:
        pass
    a = _A()
    a.batch, a.chunk, a.steps = batch, chunk, steps
    dev = torch.device('cuda:0') if device is None else device
    torch.manual_seed(0)
    C, Hf, Wf, mind, maxd = 32, 112, 112, -96, 96
    m = Stack(C, mind, maxd)
    with torch.no_grad():
        for k in m.cls:
            k[2].weight.mul_(0.1)
    m = m.to(dev)
    for mod in m.modules():  # calibrate the BatchNorm statistics on the synthetic features (as tests/golden does): without it the
        if isinstance(mod, nn.BatchNorm3d):  # logits are flat and the tf32-vs-fp32 deviation below would say nothing
            mod.momentum = None
    m.train()
    with torch.no_grad():
        g = torch.Generator(device='cpu').manual_seed(1)
        m(torch.randn(2, C, Hf, Wf, generator=g).relu().to(dev), torch.randn(2, C, Hf, Wf, generator=g).relu().to(dev), Hf, Wf)
    m.eval()
    L = torch.randn(a.batch, C, Hf, Wf, device=dev).relu()
    R = torch.randn(a.batch, C, Hf, Wf, device=dev).relu()
    torch.backends.cudnn.benchmark = True
    out = {'workload': f'B={a.batch} ROI pairs 112x112x32ch, D=48 -> 448x448, eager PyTorch {torch.__version__} / cuDNN {torch.backends.cudnn.version()}, '
                       f'chunks of {a.chunk} ROIs', 'device': torch.cuda.get_device_name(0), 'variants': {}}

    def run(tf32, cv_cpu):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        res = []
        with torch.no_grad():
            for it in range(2 + a.steps):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                outs = [m(L[i:i + a.chunk], R[i:i + a.chunk], 4 * Hf, 4 * Wf, cv_cpu) for i in range(0, a.batch, a.chunk)]
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    res.append((e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
        dev_ms = sorted(r[0] for r in res)[len(res) // 2]
        wall_ms = sorted(r[1] for r in res)[len(res) // 2]
        return torch.cat(outs), dev_ms, wall_ms

    d32, ms, wall = run(False, False)
    out['variants']['fp32 (allow_tf32=False), device-side cost volume'] = {'ms_per_batch': ms, 'rois_per_s': a.batch / ms * 1e3, 'wall_ms': wall}
    dtf, ms, wall = run(True, False)
    diff = (dtf - d32).abs()
    out['variants']['tf32 convs (PyTorch default allow_tf32=True), device-side cost volume'] = {
        'ms_per_batch': ms, 'rois_per_s': a.batch / ms * 1e3, 'wall_ms': wall,
        'disparity_vs_fp32_px': {'max': diff.max().item(), 'mean': diff.mean().item()}}
    if as_written:
        _, ms, wall = run(True, True)
        out['variants']['tf32 convs, cost volume as written (CPU zeros + H2D, stackhourglass.py:117)'] = {
            'ms_per_batch': ms, 'rois_per_s': a.batch / wall * 1e3, 'wall_ms': wall, 'note': 'rois_per_s from wall clock (host work inside)'}
    out['fp32_rois_per_s'] = out['variants']['fp32 (allow_tf32=False), device-side cost volume']['rois_per_s']
    out['tf32_rois_per_s'] = out['variants']['tf32 convs (PyTorch default allow_tf32=True), device-side cost volume']['rois_per_s']
    return out


if __name__ == '__main__':
    main()
