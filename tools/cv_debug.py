"""Compare the tensor-core path's fused cost volume with the oracle's (bf16-rounded), plane by plane."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import idispnet_oracle as O
from disprcnn_b200 import _lib
lib = _lib.load()
for (B, C, Hf, Wf, mind, maxd) in [(1, 32, 16, 16, -16, 16), (2, 32, 20, 40, -48, 48), (1, 16, 16, 24, 0, 32)]:
    g = torch.Generator().manual_seed(1)
    L = torch.randn(B, C, Hf, Wf, generator=g).bfloat16().float(); R = torch.randn(B, C, Hf, Wf, generator=g).bfloat16().float()
    D = (maxd - mind) // 4
    out = torch.full((B, 2 * C, D, Hf, Wf), float('nan'), device='cuda')
    Lc, Rc = L.cuda(), R.cuda()
    _lib.check(lib.idisp_debug_fused_cost_volume(_lib.ptr(Lc), _lib.ptr(Rc), B, C, Hf, Wf, mind, maxd, _lib.ptr(out), _lib.stream_ptr()))
    want = O.cost_volume(L, R, mind, maxd)
    tap = int(os.environ.get('IDISP_DEBUG_TAP', '13'))
    if tap != 13:
        import torch.nn.functional as F
        k = torch.zeros(2 * C, 1, 3, 3, 3); k[:, 0].view(2 * C, 27)[:, tap] = 1.0
        want = F.conv3d(want, k, None, 1, 1, 1, 2 * C)
    err = (out.cpu() - want).abs()
    print('case', (B, C, Hf, Wf, mind, maxd), 'max err', float(err.max()), 'nan', int(torch.isnan(out).sum()))
    if err.max() > 0:
        bad = err > 0
        print('  bad by channel half: L', float(bad[:, :C].float().mean()), 'R', float(bad[:, C:].float().mean()))
        print('  bad by plane:', [round(float(bad[:, :, k].float().mean()), 3) for k in range(D)])
        print('  bad by x (L half):', [round(float(bad[:, :C, :, :, x].float().mean()), 2) for x in range(Wf)])
        print('  bad by x (R half):', [round(float(bad[:, C:, :, :, x].float().mean()), 2) for x in range(Wf)])
        idx = bad.nonzero()[:5].tolist()
        print('  first bad:', [(i, float(out.cpu()[tuple(i)]), float(want[tuple(i)])) for i in idx])
