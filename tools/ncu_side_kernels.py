"""Drive every non-tensor-core kernel of the path once at a meaningful size, for an `ncu --set full` capture:
ROIAlign (+ fused normalise), the standalone cost-volume kernels, the fp32 FFMA conv kernels, the 32->1 head, soft-argmin.

usage (on the GPU box): ncu --set full --clock-control none -c 80 -o /tmp/side python tools/ncu_side_kernels.py
"""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from disprcnn_b200 import _lib  # noqa: E402
from disprcnn_b200.layers.roi_align import crop_and_transform_roi_img, roi_align  # noqa: E402
from disprcnn_b200.modeling.psmnet.stackhourglass import PSMNet  # noqa: E402

lib = _lib.load()
dev = torch.device('cuda:0')
torch.manual_seed(0)
# ROIAlign: the live call (8 images 3x375x1242, 32 boxes -> 224x224, fused mean/std) and a 32-channel feature variant
im = torch.rand(8, 3, 375, 1242, device=dev)
g = torch.Generator().manual_seed(0)
boxes = []
for i in range(32):
    x1 = int(torch.randint(0, 800, (1,), generator=g)); y1 = int(torch.randint(0, 150, (1,), generator=g))
    w = int(torch.randint(60, 400, (1,), generator=g)); h = int(torch.randint(60, 200, (1,), generator=g))
    boxes.append([i % 8, x1, y1, min(x1 + w, 1241), min(y1 + h, 374)])
crop_and_transform_roi_img(im, boxes, 224)
fea = torch.randn(8, 32, 94, 311, device=dev)
roi_align(fea, torch.tensor(boxes, dtype=torch.float32, device=dev), (112, 112), 0.25, 2)
# standalone cost volume (NCDHW test hook: reference layout) at the benchmark shape, 4 ROI pairs
B, C, Hf, Wf, mind, maxd = 4, 32, 112, 112, -96, 96
L = torch.randn(B, C, Hf, Wf, device=dev).relu(); R = torch.randn(B, C, Hf, Wf, device=dev).relu()
D = (maxd - mind) // 4
cost = torch.empty(B, 2 * C, D, Hf, Wf, device=dev)
_lib.check(lib.idisp_cost_volume(_lib.ptr(L), _lib.ptr(R), B, C, Hf, Wf, mind, maxd, _lib.ptr(cost), _lib.stream_ptr()))
# fp32 FFMA mode: blocked cost volume, 27 SIMT conv launches, SIMT 32->1 heads, soft-argmin
m = PSMNet(maxd, mind, precision='fp32')
m.feature_extraction = nn.Identity()
m = m.to(dev).eval()
with torch.no_grad():
    m.forward_features(L, R)
torch.cuda.synchronize()
print('ok')
