"""iDispNet (PSMNet stacked hourglass) -- API mirror of disprcnn/modeling/psmnet/stackhourglass.py.

Same class names, constructor signature (stackhourglass.py:55-58), ``forward(inputs)`` contract
(:106-111: dict with 'left'/'right' or a 2-sequence of [B,3,H,W] crops) and ``state_dict`` keys
(:63-88), so ``DispRCNN3D`` (disprcnn3d.py:21-33,247,273) and ``train_idispnet_fa.py:49-61`` can
construct it and load reference checkpoints unchanged.

What differs is where the arithmetic happens: in eval mode the cost volume (:115-128), the 28
3-D conv layers (:130-144) and the upsample + softmax + regression (:169-174) run inside
libidisp (hand-written sm_100a kernels) through ``idisp_plan_forward``; the ``nn.Module`` tree
below only owns the parameters.  No CPU path, no cuDNN on the 3-D stack.  Training mode is out
of scope for this path and raises.
"""
import ctypes
import warnings
import math

import torch
import torch.nn as nn

from ... import _lib
from .submodule import convbn_3d, feature_extraction

PRECISIONS = {'fp32': _lib.PREC_FP32, 'bf16': _lib.PREC_BF16, 'fp16': _lib.PREC_FP16, 'fp16x2': _lib.PREC_FP16X2}


def _deconvbn_3d(cin, cout):
    return nn.Sequential(nn.ConvTranspose3d(cin, cout, kernel_size=3, padding=1, output_padding=1, stride=2, bias=False),
                         nn.BatchNorm3d(cout))


class hourglass(nn.Module):
    """Parameter holder for one hourglass (stackhourglass.py:7-30); wiring lives in libidisp (plan.cu)."""

    def __init__(self, inplanes):
        super().__init__()
        c = inplanes
        self.conv1 = nn.Sequential(convbn_3d(c, 2 * c, 3, 2, 1), nn.ReLU(inplace=True))
        self.conv2 = convbn_3d(2 * c, 2 * c, 3, 1, 1)
        self.conv3 = nn.Sequential(convbn_3d(2 * c, 2 * c, 3, 2, 1), nn.ReLU(inplace=True))
        self.conv4 = nn.Sequential(convbn_3d(2 * c, 2 * c, 3, 1, 1), nn.ReLU(inplace=True))
        self.conv5 = _deconvbn_3d(2 * c, 2 * c)
        self.conv6 = _deconvbn_3d(2 * c, c)

    def forward(self, x, presqu, postsqu):
        raise RuntimeError('hourglass is executed inside libidisp; call PSMNet.forward')


def _classifier():
    return nn.Sequential(convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True),
                         nn.Conv3d(32, 1, kernel_size=3, padding=1, stride=1, bias=False))


class PSMNet(nn.Module):
    def __init__(self, maxdisp, mindisp=0, input_size=224, is_module=False, feature_level=1,
                 single_modal_weight_average=False, conv_layers=(), use_disparity_regression=True,
                 feature_channels=32, precision='auto'):
        """Positional signature of the reference (stackhourglass.py:55-58); two keyword-only extras with
        reference-compatible defaults: ``feature_channels`` (C of the per-view features; dres0.0 takes 2C)
        and ``precision``: 'auto' (default) = the parity-grade mode the shape allows -- 'fp16x2' (split-precision tensor-core
        kernels) when the tensor-core path covers it, else 'fp32' (CUDA-core FFMA); both are within 1e-3 px of the reference
        forward.  'bf16' / 'fp16' select the faster one-word tensor-core modes (0.02-0.4 px) explicitly."""
        super().__init__()
        if precision != 'auto' and precision not in PRECISIONS:
            raise ValueError(f"precision must be 'auto' or one of {sorted(PRECISIONS)}")
        self.maxdisp, self.mindisp = maxdisp, mindisp
        self.feature_channels, self.precision = feature_channels, precision
        self.feature_extraction = feature_extraction()
        relu = lambda: nn.ReLU(inplace=True)
        self.dres0 = nn.Sequential(convbn_3d(2 * feature_channels, 32, 3, 1, 1), relu(), convbn_3d(32, 32, 3, 1, 1), relu())
        self.dres1 = nn.Sequential(convbn_3d(32, 32, 3, 1, 1), relu(), convbn_3d(32, 32, 3, 1, 1))
        self.dres2, self.dres3, self.dres4 = hourglass(32), hourglass(32), hourglass(32)
        self.classif1, self.classif2, self.classif3 = _classifier(), _classifier(), _classifier()
        self._init_like_reference()
        self.check_range = True   # 'auto' only: ask the plan whether the fp16 range was left (one 4-byte D2H + sync per call)
        self._plans = {}      # effective precision -> [plan handle, weights key]
        self._plan = None     # the plan of the most recent forward
        self._workspace = None

    def _init_like_reference(self):
        # stackhourglass.py:90-104: He-normal for Conv2d/Conv3d (not the transposed convs), BN -> (1, 0)
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Conv3d)):
                n = m.out_channels * math.prod(m.kernel_size)
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    # ---- plan management ------------------------------------------------------------------
    def _stack_items(self):
        return [(k, v) for k, v in self.state_dict().items()
                if not k.startswith('feature_extraction.') and not k.endswith('num_batches_tracked')]

    def effective_precision(self, Hf, Wf):
        """The mode a forward at this feature size runs in: ``precision`` itself, or for 'auto' the split-precision
        tensor-core mode wherever those kernels apply (C in {16, 32}; D = (maxdisp-mindisp)/4 a multiple of 4, at most
        64; Hf, Wf multiples of 4), otherwise the fp32 FFMA mode."""
        if self.precision != 'auto':
            return self.precision
        D = (self.maxdisp - self.mindisp) // 4
        ok = self.feature_channels in (16, 32) and D % 4 == 0 and 4 <= D <= 64 and Hf % 4 == 0 and Wf % 4 == 0
        return 'fp16x2' if ok else 'fp32'

    def _ensure_plan(self, device, precision=None):
        precision = precision or (self.precision if self.precision != 'auto' else 'fp32')
        items = self._stack_items()
        key = (str(device), tuple((k, v._version, v.data_ptr()) for k, v in items))
        slot = self._plans.setdefault(precision, [None, None])
        if slot[0] is not None and key == slot[1]:
            self._plan = slot[0]
            return slot[0]
        lib = _lib.load()
        if slot[0] is None:
            h = ctypes.c_void_p()
            _lib.check(lib.idisp_plan_create(self.feature_channels, int(self.mindisp), int(self.maxdisp),
                                             PRECISIONS[precision], ctypes.byref(h)))
            slot[0] = h
        for k, v in items:
            host = v.detach().to('cpu', torch.float32).contiguous()
            _lib.check(lib.idisp_plan_set_tensor(slot[0], k.encode(), _lib.ptr(host), host.numel()))
        with torch.cuda.device(device):
            _lib.check(lib.idisp_plan_finalize(slot[0], _lib.stream_ptr()))
        slot[1] = key
        self._plan = slot[0]
        return slot[0]

    def __del__(self):
        try:
            for slot in getattr(self, '_plans', {}).values():
                if slot[0] is not None:
                    _lib.load().idisp_plan_destroy(slot[0])
                    slot[0] = None
            self._plan = None
        except Exception:
            pass

    # ---- forward --------------------------------------------------------------------------
    def forward_features(self, left_fea, right_fea, H=None, W=None):
        """stackhourglass.py:115-174 from the per-view features [B,C,Hf,Wf] -> disparity [B,H,W]
        (H,W default to 4Hf,4Wf, the crop size the features came from)."""
        if self.training:
            raise RuntimeError('PSMNet (B200 path) is inference-only: call .eval() '
                               '(training of iDispNet is out of scope for this path)')
        _lib.require_cuda(left_fea, right_fea)
        left_fea, right_fea = left_fea.contiguous().float(), right_fea.contiguous().float()
        B, C, Hf, Wf = left_fea.shape
        if right_fea.shape != left_fea.shape or C != self.feature_channels:
            raise RuntimeError(f'PSMNet: feature shapes {tuple(left_fea.shape)} / {tuple(right_fea.shape)} '
                               f'do not match feature_channels={self.feature_channels}')
        H = 4 * Hf if H is None else H
        W = 4 * Wf if W is None else W
        out = torch.empty((B, H, W), dtype=torch.float32, device=left_fea.device)
        if B == 0:
            return out
        lib = _lib.load()

        def run(precision):
            plan = self._ensure_plan(left_fea.device, precision)
            need = lib.idisp_plan_workspace_bytes(plan, B, Hf, Wf)
            ws = self._workspace
            if ws is None or ws.numel() < need or ws.device != left_fea.device:
                self._workspace = None
                ws = self._workspace = torch.empty(need, dtype=torch.uint8, device=left_fea.device)
            _lib.check(lib.idisp_plan_forward(plan, _lib.ptr(left_fea), _lib.ptr(right_fea), B, Hf, Wf, H, W,
                                              _lib.ptr(ws), ws.numel(), _lib.ptr(out), _lib.stream_ptr()))

        with torch.cuda.device(left_fea.device):
            precision = self.effective_precision(Hf, Wf)
            run(precision)
            if self.precision == 'auto' and precision == 'fp16x2' and self.check_range and self.range_exceeded():
                # the hi words are IEEE halves: an activation beyond 65504 overflows them (and ReLU turns the resulting NaN
                # into 0: finite but wrong).  Not seen with BatchNorm-ed iDispNet weights, but 'auto' promises a parity-grade
                # answer, so such a batch is redone by the fp32 FFMA kernels.
                warnings.warn('PSMNet: activations left the fp16 range of the split-precision mode; batch recomputed in fp32')
                run('fp32')
        return out

    def range_exceeded(self):
        """fp16-word modes: did the most recent forward see a value outside the IEEE-half range (idisp_plan_range_exceeded)?
        Synchronises the current stream."""
        if self._plan is None:
            return False
        flag = ctypes.c_int(0)
        with torch.cuda.device(self._workspace.device):
            _lib.check(_lib.load().idisp_plan_range_exceeded(self._plan, ctypes.byref(flag), _lib.stream_ptr()))
        return bool(flag.value)

    def last_logits(self, B, Hf, Wf):
        """cost3 [B,D,Hf,Wf] of the most recent forward (debug/test hook)."""
        D = (self.maxdisp - self.mindisp) // 4
        dev = self._workspace.device
        out = torch.empty((B, D, Hf, Wf), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().idisp_plan_get_logits(self._plan, _lib.ptr(out), _lib.stream_ptr()))
        return out

    def forward(self, inputs):
        if isinstance(inputs, dict):
            left, right = inputs['left'], inputs['right']
        elif len(inputs) == 2:
            left, right = inputs
        else:
            raise RuntimeError("PSMNet.forward expects {'left','right'} or a 2-sequence")
        _, _, H, W = left.shape
        if isinstance(self.feature_extraction, nn.Identity):  # feature-input configs (BASELINE configs 1-3)
            return self.forward_features(left, right, H, W)
        return self.forward_features(self.feature_extraction(left), self.feature_extraction(right), H, W)
