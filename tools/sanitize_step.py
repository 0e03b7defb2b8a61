"""A small pass over every kernel family for compute-sanitizer (tools/sanitize.sh): the 3-D stack at a tiny shape in the split-precision,
one-word and fp32 modes, the whole PSMNet.forward (native extractor) on one 224x224 crop pair, ROIAlign crops and the hand-off kernels."""
import os
import sys

os.environ.setdefault('IDISP_NO_GRAPH', '1')
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disprcnn_b200.layers.roi_align import crop_stereo_rois  # noqa: E402
from disprcnn_b200.layers.roi_disparity import paste_roi_disparity, roi_depth_maps  # noqa: E402
from disprcnn_b200.modeling.psmnet.stackhourglass import PSMNet  # noqa: E402

torch.manual_seed(0)
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(1)
with torch.no_grad():
    for prec in ('fp16x2', 'fp16', 'bf16', 'fp32'):
        m = PSMNet(16, -16, precision=prec)
        m.feature_extraction = nn.Identity()
        m = m.to(dev).eval()
        L = torch.randn(3, 32, 16, 24, generator=g).relu().to(dev)
        R = torch.randn(3, 32, 16, 24, generator=g).relu().to(dev)
        out = m.forward_features(L, R)
        torch.cuda.synchronize()
        print(prec, 'stack ok', tuple(out.shape), bool(torch.isfinite(out).all()))
    m = PSMNet(48, -48).to(dev).eval()
    il = torch.randn(1, 3, 224, 224, generator=g).to(dev)
    ir = torch.randn(1, 3, 224, 224, generator=g).to(dev)
    out = m({'left': il, 'right': ir})
    torch.cuda.synchronize()
    print('psmnet ok', tuple(out.shape))
    im = torch.rand(2, 3, 120, 300, generator=g).to(dev)
    lb = torch.tensor([[10., 8., 90., 70.], [100., 20., 260., 110.]], device=dev)
    rb = torch.tensor([[4., 8., 84., 70.], [90., 20., 250., 110.]], device=dev)
    idx = torch.tensor([0, 1], device=dev)
    crops = crop_stereo_rois(im, im, lb, rb, idx, 64)
    disp = torch.randn(2, 64, 64, generator=g).to(dev) * 5
    paste_roi_disparity(disp, lb, rb, [1, 1], 120, 300)
    roi_depth_maps(disp, lb, rb, torch.full((2,), 380.0, device=dev), 120, 300)
    torch.cuda.synchronize()
    print('side kernels ok')
print('done')
