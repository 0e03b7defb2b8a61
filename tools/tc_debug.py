"""Structured diagnosis of the tcgen05 conv kernel through the C-ABI (idisp_conv3d, bf16 precision).

Each experiment isolates one thing the kernel could get wrong (K order of A/B cores, N order, tap
shift direction, halo/zero padding, ring wrap, cout halves, epilogue) and prints a JSON line with
the error against torch's conv3d on the same bf16-rounded operands.  Run one experiment per process
(a device trap must not take the others down):  python tools/tc_debug.py <exp> [cin cout D H W]
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from disprcnn_b200 import _lib  # noqa: E402


def run(x, w, bias=None, res=None, relu=0, kind=0, oshape=None):
    lib = _lib.load()
    B, cin, D, H, W = x.shape
    cout = w.shape[1] if kind == 2 else w.shape[0]
    y = torch.full(oshape if oshape is not None else (B, cout, D, H, W), float('nan'), device='cuda')
    xc, wc = x.cuda(), w.cuda()
    bc = bias.cuda() if bias is not None else None
    rc = res.cuda() if res is not None else None
    _lib.check(lib.idisp_conv3d(_lib.ptr(xc), B, cin, D, H, W, _lib.ptr(wc), cout, kind, None, _lib.ptr(bc), _lib.ptr(rc),
                                relu, 1, _lib.ptr(y), _lib.stream_ptr()))
    torch.cuda.synchronize()
    return y.cpu()


def report(name, y, want, extra=None):
    err = (y - want).abs()
    err[torch.isnan(err)] = 1e9
    out = {'exp': name, 'max_err': float(err.max()), 'ref_absmax': float(want.abs().max()), 'nan': int(torch.isnan(y).sum())}
    if err.max() > 1e-2 * max(1.0, float(want.abs().max())):
        bad = err > 1e-2 * max(1.0, float(want.abs().max()))
        out['bad_frac'] = float(bad.float().mean())
        out['bad_by_cout'] = bad.float().mean(dim=(0, 2, 3, 4)).tolist()
        out['bad_by_d'] = bad.float().mean(dim=(0, 1, 3, 4)).tolist()
        out['bad_by_h'] = bad.float().mean(dim=(0, 1, 2, 4)).tolist()
        out['bad_by_w'] = bad.float().mean(dim=(0, 1, 2, 3)).tolist()
        idx = bad.nonzero()[:6].tolist()
        out['first_bad'] = [(i, float(y[tuple(i)]), float(want[tuple(i)])) for i in idx]
    if extra:
        out.update(extra)
    print(json.dumps(out))


def main():
    exp = sys.argv[1]
    cin, cout, D, H, W = (int(a) for a in sys.argv[2:7]) if len(sys.argv) >= 7 else (32, 32, 4, 16, 8)
    g = torch.Generator().manual_seed(0)
    B = 1
    x = torch.randn(B, cin, D, H, W, generator=g).bfloat16().float()
    w = torch.zeros(cout, cin, 3, 3, 3)
    if exp == 'center_identity':      # y[co] == x[ci=co]: K order of A/B cores, N order, row mapping
        for c in range(min(cin, cout)):
            w[c, c, 1, 1, 1] = 1.0
    elif exp == 'center_onechan':     # only ci=3 -> co=5
        w[5, 3, 1, 1, 1] = 1.0
    elif exp.startswith('tap_'):      # tap_kd_kh_kw: pure shift
        kd, kh, kw = (int(a) for a in exp.split('_')[1:])
        for c in range(min(cin, cout)):
            w[c, c, kd, kh, kw] = 1.0
    elif exp == 'random':
        w = (torch.randn(cout, cin, 3, 3, 3, generator=g) * (2.0 / (27 * cout)) ** 0.5).bfloat16().float()
    elif exp in ('s2', 'dec', 'to1'):
        kind = {'s2': 1, 'dec': 2, 'to1': 0}[exp]
        if exp == 'to1':
            cout = 1
        wshape = (cin, cout, 3, 3, 3) if kind == 2 else (cout, cin, 3, 3, 3)
        w = (torch.randn(wshape, generator=g) * (2.0 / (27 * max(cout, 8))) ** 0.5).bfloat16().float()
        if kind == 1:
            want = F.conv3d(x, w, None, 2, 1)
        elif kind == 2:
            want = F.conv_transpose3d(x, w, None, 2, 1, 1)
        else:
            want = F.conv3d(x, w, None, 1, 1)
        if exp == 'to1':
            res = torch.randn(want.shape, generator=g)
            want = want + res
            report(exp, run(x, w, None, res, 0, kind, tuple(want.shape)), want, {'shape': [cin, cout, D, H, W]})
        else:
            bias = torch.randn(cout, generator=g)
            res = torch.randn(want.shape, generator=g).bfloat16().float()
            want = F.relu(want + bias.view(1, -1, 1, 1, 1) + res)
            report(exp, run(x, w, bias, res, 1, kind, tuple(want.shape)), want, {'shape': [cin, cout, D, H, W]})
        return
    elif exp == 'random_epi':
        w = (torch.randn(cout, cin, 3, 3, 3, generator=g) * (2.0 / (27 * cout)) ** 0.5).bfloat16().float()
        bias = torch.randn(cout, generator=g)
        res = torch.randn(B, cout, D, H, W, generator=g).bfloat16().float()
        want = F.relu(F.conv3d(x, w, None, 1, 1) + bias.view(1, -1, 1, 1, 1) + res)
        report(exp, run(x, w, bias, res, 1), want, {'shape': [cin, cout, D, H, W]})
        return
    else:
        raise SystemExit('unknown experiment ' + exp)
    want = F.conv3d(x, w, None, 1, 1)
    report(exp, run(x, w), want, {'shape': [cin, cout, D, H, W], 'nostack': os.environ.get('IDISP_TC_NOSTACK', '0')})


if __name__ == '__main__':
    main()
