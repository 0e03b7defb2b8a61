"""Building blocks of iDispNet -- API mirror of disprcnn/modeling/psmnet/submodule.py.

``convbn_3d`` (submodule.py:19-22) and ``disparityregression`` (:51-57) are the hot-path names;
``feature_extraction`` (:60-139) is the adjacent 2-D extractor: a module tree with
the reference's parameter names so reference checkpoints load; in eval mode on a GPU its forward runs inside
libidisp (csrc/feature2d.cu through ``idisp_extractor_forward``: no cuDNN / ATen kernels), in training mode the
torch modules run (training is outside the B200 path).

The 3-D modules built here are PARAMETER HOLDERS with the reference's ``state_dict`` layout; in
eval mode on a GPU their arithmetic is executed by libidisp (see stackhourglass.PSMNet).
"""
import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _lib


def convbn(in_planes, out_planes, kernel_size, stride, pad, dilation):
    conv = nn.Conv2d(in_planes, out_planes, kernel_size, stride, dilation if dilation > 1 else pad, dilation,
                     bias=False)
    return nn.Sequential(conv, nn.BatchNorm2d(out_planes))


def convbn_3d(in_planes, out_planes, kernel_size, stride, pad):
    conv = nn.Conv3d(in_planes, out_planes, kernel_size, stride, pad, bias=False)
    return nn.Sequential(conv, nn.BatchNorm3d(out_planes))


def disparityregression(x, maxdisp, mindisp=0):
    """Expectation over disparity of a probability volume [B, maxdisp-mindisp, H, W] (submodule.py:51-57).

    Kept for API compatibility (training scripts call it on softmax outputs); the eval path of
    PSMNet never materialises ``x`` -- it uses the fused ``soft_argmin`` below."""
    assert x.shape[1] == int(maxdisp - mindisp)
    disp = torch.arange(mindisp, maxdisp, dtype=x.dtype, device=x.device).view(1, -1, 1, 1)
    return (x * disp).sum(1)


def soft_argmin(logits, mindisp, maxdisp, H, W):
    """Fused F.interpolate(trilinear, align_corners) + softmax + disparityregression
    (stackhourglass.py:169-172 + submodule.py:51-57) through ``idisp_softargmin``.
    logits [B,D,Hf,Wf] or [B,1,D,Hf,Wf] f32 CUDA -> [B,H,W]."""
    _lib.require_cuda(logits)
    if logits.dim() == 5:
        logits = logits.squeeze(1)
    logits = logits.contiguous().float()
    B, D, Hf, Wf = logits.shape
    out = torch.empty((B, H, W), dtype=torch.float32, device=logits.device)
    with torch.cuda.device(logits.device):
        _lib.check(_lib.load().idisp_softargmin(_lib.ptr(logits), B, D, Hf, Wf, int(mindisp), int(maxdisp), int(H),
                                               int(W), _lib.ptr(out), _lib.stream_ptr()))
    return out


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride, downsample, pad, dilation):
        super().__init__()
        self.conv1 = nn.Sequential(convbn(inplanes, planes, 3, stride, pad, dilation), nn.ReLU(inplace=True))
        self.conv2 = convbn(planes, planes, 3, 1, pad, dilation)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = self.conv2(self.conv1(x))
        return y + (x if self.downsample is None else self.downsample(x))


class feature_extraction(nn.Module):
    """2-D CNN + SPP, 3 -> 32 channels at 1/4 resolution (parameter names of submodule.py:60-110)."""

    _SPP = ((56, 'branch1'), (32, 'branch2'), (16, 'branch3'), (8, 'branch4'))

    def __init__(self):
        super().__init__()
        self.inplanes = 32
        relu = lambda: nn.ReLU(inplace=True)
        self.firstconv = nn.Sequential(convbn(3, 32, 3, 2, 1, 1), relu(), convbn(32, 32, 3, 1, 1, 1), relu(),
                                       convbn(32, 32, 3, 1, 1, 1), relu())
        self.layer1 = self._stage(32, 3, 1, 1, 1)
        self.layer2 = self._stage(64, 16, 2, 1, 1)
        self.layer3 = self._stage(128, 3, 1, 1, 1)
        self.layer4 = self._stage(128, 3, 1, 1, 2)
        for k, name in self._SPP:
            setattr(self, name, nn.Sequential(nn.AvgPool2d((k, k), stride=(k, k)), convbn(128, 32, 1, 1, 0, 1), relu()))
        self.lastconv = nn.Sequential(convbn(320, 128, 3, 1, 1, 1), relu(),
                                      nn.Conv2d(128, 32, kernel_size=1, padding=0, stride=1, bias=False))
        self.native = True   # eval + CUDA: run inside libidisp (False: the torch modules, e.g. for A/B timing against cuDNN)
        # 'auto' (default): the stride-1 3x3 convs on the tensor cores in split precision ('fp16x2', fp32-grade); a batch whose
        # activations leave the IEEE-half range is redone by the fp32 FFMA kernels ('fp32' selects those outright)
        self.precision = 'auto'
        self.check_range = True
        self._reset_runtime_state()

    def _reset_runtime_state(self):
        self._handles = {}      # precision -> [idisp_extractor_t*, weights key it was finalised with]
        self._handle = None     # handle of the most recent forward
        self._tensors = None
        self._workspaces = {}   # (device index, stream) -> uint8 arena

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in ('_handles', '_handle', '_tensors', '_workspaces'):
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._reset_runtime_state()

    def __del__(self):
        try:
            for slot in getattr(self, '_handles', {}).values():
                if slot[0] is not None:
                    _lib.load().idisp_extractor_destroy(slot[0])
                    slot[0] = None
            self._handle = None
        except Exception:
            pass

    def _items(self):
        cur = self._tensors
        if cur is not None and all(getattr(owner, name) is t for (_, owner, name, t) in cur):
            return cur
        items = []
        for mod_name, mod in self.named_modules():
            for name, t in list(mod._parameters.items()) + list(mod._buffers.items()):
                if t is None or name == 'num_batches_tracked':
                    continue
                items.append(((mod_name + '.' if mod_name else '') + name, mod, name, t))
        self._tensors = items
        return items

    def _ensure_handle(self, device, precision):
        items = self._items()
        key = (str(device), tuple((t._version, t.data_ptr()) for (_, _, _, t) in items))
        slot = self._handles.setdefault(precision, [None, None])
        if slot[0] is not None and key == slot[1]:
            self._handle = slot[0]
            return slot[0]
        lib = _lib.load()
        if slot[0] is None:
            h = ctypes.c_void_p()
            _lib.check(lib.idisp_extractor_create(ctypes.byref(h)))
            _lib.check(lib.idisp_extractor_set_precision(h, _lib.PREC_FP16X2 if precision == 'fp16x2' else _lib.PREC_FP32))
            slot[0] = h
        for k, _, _, t in items:
            host = t.detach().to('cpu', torch.float32).contiguous()
            _lib.check(lib.idisp_extractor_set_tensor(slot[0], k.encode(), _lib.ptr(host), host.numel()))
        with torch.cuda.device(device):
            _lib.check(lib.idisp_extractor_finalize(slot[0], _lib.stream_ptr()))
        slot[1] = key
        self._handle = slot[0]
        return slot[0]

    def forward_native(self, x):
        """submodule.py:112-139 inside libidisp: [B,3,H,W] f32 CUDA -> [B,32,H/4,W/4]."""
        _lib.require_cuda(x)
        x = x.contiguous().float()
        B, C, H, W = x.shape
        if C != 3:
            raise RuntimeError(f'feature_extraction: 3-channel images expected, got {C}')
        Hq, Wq = ((H - 1) // 2 + 1 - 1) // 2 + 1, ((W - 1) // 2 + 1 - 1) // 2 + 1
        out = torch.empty((B, 32, Hq, Wq), dtype=torch.float32, device=x.device)
        if B == 0:
            return out
        if self.precision not in ('auto', 'fp16x2', 'fp32'):
            raise ValueError("feature_extraction.precision must be 'auto', 'fp16x2' or 'fp32'")
        lib = _lib.load()

        def run(precision):
            h = self._ensure_handle(x.device, precision)
            need = lib.idisp_extractor_workspace_bytes(h, B, H, W)
            wkey = (x.device.index, torch.cuda.current_stream().cuda_stream)
            ws = self._workspaces.get(wkey)
            if ws is None or ws.numel() < need:
                self._workspaces.pop(wkey, None)
                ws = self._workspaces[wkey] = torch.empty(need, dtype=torch.uint8, device=x.device)
            _lib.check(lib.idisp_extractor_forward(h, _lib.ptr(x), B, H, W, _lib.ptr(ws), ws.numel(), _lib.ptr(out), _lib.stream_ptr()))
            return h

        with torch.cuda.device(x.device):
            h = run('fp32' if self.precision == 'fp32' else 'fp16x2')
            if self.precision == 'auto' and self.check_range:
                flag = ctypes.c_int(0)
                _lib.check(lib.idisp_extractor_range_exceeded(h, ctypes.byref(flag), _lib.stream_ptr()))
                if flag.value:
                    import warnings
                    warnings.warn('feature_extraction: activations left the fp16 range of the split-precision mode; batch recomputed in fp32')
                    run('fp32')
        return out

    def _stage(self, planes, blocks, stride, pad, dilation):
        down = None
        if stride != 1 or self.inplanes != planes:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes, kernel_size=1, stride=stride, bias=False),
                                 nn.BatchNorm2d(planes))
        mods = [BasicBlock(self.inplanes, planes, stride, down, pad, dilation)]
        self.inplanes = planes
        mods += [BasicBlock(planes, planes, 1, None, pad, dilation) for _ in range(1, blocks)]
        return nn.Sequential(*mods)

    def forward(self, x):
        if not self.training and self.native:
            return self.forward_native(x)   # raises for CPU tensors: no CPU path in eval mode
        raw = self.layer2(self.layer1(self.firstconv(x)))
        skip = self.layer4(self.layer3(raw))
        size = skip.shape[-2:]
        pyramid = [F.interpolate(getattr(self, name)(skip), size, mode='bilinear', align_corners=True)
                   for _, name in self._SPP]
        # concat order of submodule.py:134-135: raw, skip, branch4, branch3, branch2, branch1
        return self.lastconv(torch.cat([raw, skip] + pyramid[::-1], 1))
