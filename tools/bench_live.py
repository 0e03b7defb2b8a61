"""Latency of the reference's LIVE shape (KITTI configs: 224x224 crops -> C=32 features 56x56, D=24, R = 1..15 ROIs per
image) -- the call tools/test_net.py makes once per image.  Prints one JSON line."""
import json, os, sys
import torch
import torch.nn as nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from disprcnn_b200.modeling.psmnet.stackhourglass import PSMNet

out = {}
for prec in ('fp16x2', 'fp16', 'bf16', 'fp32'):
    torch.manual_seed(0)
    m = PSMNet(48, -48, precision=prec)
    m.feature_extraction = nn.Identity()
    m = m.cuda().eval()
    res = {}
    for R in (1, 4, 8, 15):
        L = torch.randn(R, 32, 56, 56, device='cuda').relu(); Rr = torch.randn(R, 32, 56, 56, device='cuda').relu()
        with torch.no_grad():
            for _ in range(5):
                m.forward_features(L, Rr)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            n = 30
            for _ in range(n):
                m.forward_features(L, Rr)
            e1.record(); torch.cuda.synchronize()
        res[R] = round(e0.elapsed_time(e1) / n, 4)
    out[prec] = res
print(json.dumps({'live_shape_ms_per_forward': out, 'shape': 'C32 56x56 D24 -> 224x224, R ROIs'}))
