"""One forward of the hot path inside an NVTX range 'profiled', for ncu (--nvtx --nvtx-include "profiled/"):
   python tools/ncu_step.py stack [B]      BASELINE configs[1] stack forward (cost volume + 28 conv layers + soft-argmin), B ROI pairs
   python tools/ncu_step.py live [R]       whole PSMNet.forward on R crop pairs (extractor + stack at the KITTI shape)
   python tools/ncu_step.py side           ROIAlign crop, stereo ROI prep, cost volume (NCDHW), disparity paste / depth kernels
CUDA graphs are disabled so that every kernel is its own launch."""
import os
import sys

os.environ.setdefault('IDISP_NO_GRAPH', '1')
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disprcnn_b200.modeling.psmnet.stackhourglass import PSMNet  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'stack'
n = int(sys.argv[2]) if len(sys.argv) > 2 else (32 if what == 'stack' else 8)
torch.manual_seed(0)
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(1)
with torch.no_grad():
    if what == 'stack':
        m = PSMNet(96, -96, precision='fp16x2')
        m.feature_extraction = nn.Identity()
        for c in (m.classif1, m.classif2, m.classif3):
            c[2].weight.mul_(0.1)
        m = m.to(dev).eval()
        L = torch.randn(n, 32, 112, 112, generator=g).relu().to(dev)
        R = torch.randn(n, 32, 112, 112, generator=g).relu().to(dev)
        fn = lambda: m.forward_features(L, R)
    elif what == 'live':
        m = PSMNet(48, -48)
        for c in (m.classif1, m.classif2, m.classif3):
            c[2].weight.mul_(0.1)
        m = m.to(dev)
        fe = m.feature_extraction
        for mod in fe.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.momentum = None
        fe.train()
        for _ in range(2):
            f = fe(torch.randn(4, 3, 224, 224, generator=g).to(dev))
        fe.lastconv[2].weight.mul_(1.0 / float(f.std()))
        m = m.eval()
        il = torch.randn(n, 3, 224, 224, generator=g).to(dev)
        ir = torch.randn(n, 3, 224, 224, generator=g).to(dev)
        fn = lambda: m({'left': il, 'right': ir})
    else:
        from disprcnn_b200 import _lib
        from disprcnn_b200.layers.roi_align import crop_stereo_rois
        from disprcnn_b200.layers.roi_disparity import paste_roi_disparity, roi_depth_maps
        lib = _lib.load()
        im = torch.rand(8, 3, 375, 1242, generator=g).to(dev)
        x1 = torch.randint(0, 800, (32,), generator=g).float()
        y1 = torch.randint(0, 150, (32,), generator=g).float()
        w = torch.randint(60, 400, (32,), generator=g).float()
        h = torch.randint(60, 200, (32,), generator=g).float()
        lb = torch.stack([x1, y1, x1 + w, y1 + h], 1).to(dev)
        rb = torch.stack([(x1 - 20).clamp(min=0), y1, (x1 - 20).clamp(min=0) + w, y1 + h], 1).to(dev)
        idx = (torch.arange(32) // 4).to(dev)
        disp = torch.randn(32, 224, 224, generator=g).to(dev) * 10
        fub = torch.full((32,), 380.0, device=dev)
        Lf = torch.randn(8, 32, 112, 112, generator=g).to(dev)
        Rf = torch.randn(8, 32, 112, 112, generator=g).to(dev)
        cost = torch.empty(8, 64, 48, 112, 112, device=dev)

        def fn():
            crop_stereo_rois(im, im, lb, rb, idx, 224)
            _lib.check(lib.idisp_cost_volume(_lib.ptr(Lf), _lib.ptr(Rf), 8, 32, 112, 112, -96, 96, _lib.ptr(cost), _lib.stream_ptr()))
            paste_roi_disparity(disp, lb, rb, [4] * 8, 375, 1242)
            roi_depth_maps(disp, lb, rb, fub, 375, 1242)
    fn()
    fn()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push('profiled')
    fn()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print('done')
