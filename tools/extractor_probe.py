"""Run the native feature extractor a few times (for an ncu launch list): python tools/extractor_probe.py [R]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disprcnn_b200.modeling.psmnet.submodule import feature_extraction  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
fe = feature_extraction().cuda().eval()
x = torch.randn(R, 3, 224, 224, device='cuda')
with torch.no_grad():
    for _ in range(2):
        fe(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fe(x)
    torch.cuda.synchronize()
    print(f'native extractor (tensor-core split precision), {R} images: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms per call')
    fe.precision = 'fp32'
    for _ in range(2):
        fe(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fe(x)
    torch.cuda.synchronize()
    print(f'native extractor (fp32 FFMA kernels), {R} images: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms per call')
    fe.native = False
    torch.backends.cudnn.allow_tf32 = False
    for _ in range(3):
        fe(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fe(x)
    torch.cuda.synchronize()
    print(f'torch/cuDNN fp32 extractor, {R} images: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms per call')
    torch.backends.cudnn.allow_tf32 = True
    for _ in range(3):
        fe(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fe(x)
    torch.cuda.synchronize()
    print(f'torch/cuDNN tf32 extractor, {R} images: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms per call')
