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
        # 'auto' only: ask the plan whether the fp16 range was left.  True = before returning (one 4-byte D2H + stream sync per
        # call: the result handed back is always parity-grade); 'deferred' = no sync -- the flag of call N is read at call N+1
        # (or by range_exceeded()); if it was set, a warning names the affected call and the module switches to fp32 for good;
        # False = never.
        self.check_range = True
        self._reset_runtime_state()

    def _reset_runtime_state(self):
        """Everything that is NOT a parameter: native plan handles, workspaces, caches (dropped by pickling / deepcopy)."""
        self._plans = {}        # effective precision -> [plan handle, weights key]
        self._plan = None       # the plan of the most recent forward
        self._workspaces = {}   # (device index, stream handle) -> uint8 arena; one per stream, so two streams never share scratch
        self._workspace = None  # the arena of the most recent forward
        self._stack_tensors = None
        self._pending_range = None   # 'deferred' range check: (plan, call number) of a forward whose flag has not been read
        self._calls = 0
        self._sticky_fp32 = False

    def __getstate__(self):
        # ctypes plan handles cannot be pickled / deep-copied (and a copied handle would be destroyed twice): a copy starts
        # without native state and builds its own plan on first use
        state = dict(self.__dict__)
        for k in ('_plans', '_plan', '_workspaces', '_workspace', '_stack_tensors', '_pending_range'):
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._reset_runtime_state()

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
        """(key, tensor) of every 3-D-stack parameter / buffer the plan consumes.  The list of tensor OBJECTS is cached (a
        state_dict() walk over 514 entries costs ~2.5 ms of host time per forward); `_weights_key` notices both in-place
        edits (`_version`) and re-assignment / .to() (`data_ptr`, identity)."""
        cur = self._stack_tensors
        if cur is not None:
            ok = True
            for (k, owner, name, t) in cur:
                if getattr(owner, name) is not t:   # parameter object replaced (load_state_dict(assign=True), .to(), ...)
                    ok = False
                    break
            if ok:
                return [(k, t) for (k, _, _, t) in cur]
        items = []
        for mod_name, mod in self.named_modules():
            if mod_name.startswith('feature_extraction'):
                continue
            for name, t in list(mod._parameters.items()) + list(mod._buffers.items()):
                if t is None or name == 'num_batches_tracked':
                    continue
                items.append(((mod_name + '.' if mod_name else '') + name, mod, name, t))
        self._stack_tensors = items
        return [(k, t) for (k, _, _, t) in items]

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
        key = (str(device), tuple((v._version, v.data_ptr()) for _, v in items))
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
            wkey = (left_fea.device.index, torch.cuda.current_stream().cuda_stream)
            ws = self._workspaces.get(wkey)
            if ws is None or ws.numel() < need:
                self._workspaces.pop(wkey, None)
                ws = self._workspaces[wkey] = torch.empty(need, dtype=torch.uint8, device=left_fea.device)
            self._workspace = ws
            _lib.check(lib.idisp_plan_forward(plan, _lib.ptr(left_fea), _lib.ptr(right_fea), B, Hf, Wf, H, W,
                                              _lib.ptr(ws), ws.numel(), _lib.ptr(out), _lib.stream_ptr()))

        with torch.cuda.device(left_fea.device):
            self._calls += 1
            if self._pending_range is not None:   # 'deferred' check of the previous call (its work is long finished)
                self._read_pending_range()
            precision = self.effective_precision(Hf, Wf)
            if self.precision == 'auto' and self._sticky_fp32:
                precision = 'fp32'
            try:
                run(precision)
            except _lib.Unsupported:
                # the C side knows more than effective_precision() (tensor-map limits, IDISP_TC_DISABLE, ...): 'auto' promised
                # a parity-grade answer, so a shape the tensor-core mode rejects goes to the fp32 FFMA kernels
                if self.precision != 'auto' or precision == 'fp32':
                    raise
                precision = 'fp32'
                run('fp32')
            if self.precision == 'auto' and precision == 'fp16x2' and self.check_range:
                if self.check_range == 'deferred':
                    self._pending_range = (self._plan, self._calls)
                elif self.range_exceeded():
                    # the hi words are IEEE halves: an activation beyond 65504 overflows them (and ReLU turns the resulting NaN
                    # into 0: finite but wrong).  Not seen with BatchNorm-ed iDispNet weights, but 'auto' promises a parity-grade
                    # answer, so such a batch is redone by the fp32 FFMA kernels.
                    warnings.warn('PSMNet: activations left the fp16 range of the split-precision mode; batch recomputed in fp32')
                    run('fp32')
        return out

    def _read_pending_range(self):
        plan, call = self._pending_range
        self._pending_range = None
        flag = ctypes.c_int(0)
        _lib.check(_lib.load().idisp_plan_range_exceeded(plan, ctypes.byref(flag), _lib.stream_ptr()))
        if flag.value:
            self._sticky_fp32 = True
            warnings.warn(f'PSMNet: forward call #{call} left the fp16 range of the split-precision mode (its result is not '
                          f"parity-grade); check_range='deferred' switches this module to the fp32 kernels from now on")
        return bool(flag.value)

    def range_exceeded(self):
        """fp16-word modes: did the most recent forward see a value outside the IEEE-half range (idisp_plan_range_exceeded)?
        Synchronises the current stream."""
        if self._plan is None:
            return False
        if self._pending_range is not None:
            with torch.cuda.device(self._workspace.device):
                return self._read_pending_range()
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
        if not self.training and getattr(self.feature_extraction, 'native', False) and left.shape == right.shape:
            # both views in ONE extractor pass (weights are shared, stackhourglass.py:112-113): twice the parallelism per launch
            fea = self.feature_extraction(torch.cat([left, right], 0))
            nb = left.shape[0]
            return self.forward_features(fea[:nb], fea[nb:], H, W)
        return self.forward_features(self.feature_extraction(left), self.feature_extraction(right), H, W)
