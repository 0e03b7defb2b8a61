"""ROIAlign operator -- same names, arguments and error behaviour as disprcnn/layers/roi_align.py:13-73.

``roi_align(input, roi, output_size, spatial_scale, sampling_ratio)`` and
``ROIAlign(output_size, spatial_scale, sampling_ratio).forward(input, rois, spatial_scale=None)``
call ``idisp_roi_align_forward`` (include/idisp.h) on the current CUDA stream.  The output is a
fresh tensor, inputs are made contiguous inside (reference: ROIAlign_cuda.cu:271,286,294).
Backward raises ``RuntimeError`` exactly like the reference's non-CUDA build
(csrc/ROIAlign.h:44): this is the inference path.
"""
import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from .. import _lib


def roi_align_forward(input, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio, mean=None, std=None):
    """Functional twin of ``disprcnn._C.roi_align_forward`` (csrc/vision.cpp:9) with an optional
    fused per-channel ``(x - mean) / std`` (disprcnn3d.py:47-49)."""
    _lib.require_cuda(input, rois, mean, std)
    if input.dtype != torch.float32 or rois.dtype != torch.float32:
        raise RuntimeError('roi_align: float32 tensors expected')
    if rois.numel() == 0:
        # no detections: DispRCNN3D.crop_and_transform_roi_img hands over torch.as_tensor([]) of shape [0]
        # (disprcnn3d.py:44-46,147-148); the reference only reads rois.size(0) and returns an empty
        # [0,C,ph,pw] tensor (ROIAlign_cuda.cu:271,278-281), which the caller then handles (:149-158)
        return torch.empty((0, input.size(1), pooled_h, pooled_w), dtype=input.dtype, device=input.device)
    if rois.dim() != 2 or rois.size(1) != 5:
        raise RuntimeError('roi_align: rois must be [R,5] (batch_idx,x1,y1,x2,y2)')
    if input.device != rois.device:
        raise RuntimeError('roi_align: input and rois must be on the same device')
    input = input.contiguous()
    rois = rois.contiguous()
    N, C, H, W = input.shape
    R = rois.size(0)
    out = torch.empty((R, C, pooled_h, pooled_w), dtype=input.dtype, device=input.device)
    if out.numel() == 0:
        return out
    if mean is not None:
        mean, std = mean.contiguous().float(), std.contiguous().float()
    with torch.cuda.device(input.device):
        _lib.check(_lib.load().idisp_roi_align_forward(
            _lib.ptr(input), N, C, H, W, _lib.ptr(rois), R, float(spatial_scale), int(pooled_h), int(pooled_w),
            int(sampling_ratio), _lib.ptr(mean), _lib.ptr(std), _lib.ptr(out), _lib.stream_ptr()))
    return out


class _ROIAlign(Function):
    @staticmethod
    def forward(ctx, input, roi, output_size, spatial_scale, sampling_ratio):
        ctx.save_for_backward(roi)
        ctx.output_size = _pair(output_size)
        ctx.spatial_scale = spatial_scale
        ctx.sampling_ratio = sampling_ratio
        ctx.input_shape = input.size()
        oh, ow = _pair(output_size)
        return roi_align_forward(input, roi, spatial_scale, oh, ow, sampling_ratio)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        raise RuntimeError('roi_align backward: not implemented on the B200 inference path '
                           '(reference: csrc/ROIAlign.h:44 raises for its non-CUDA build)')


roi_align = _ROIAlign.apply


class ROIAlign(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio

    def forward(self, input, rois, spatial_scale=None):
        if spatial_scale is None:
            spatial_scale = self.spatial_scale
        return roi_align(input, rois, self.output_size, spatial_scale, self.sampling_ratio)

    def __repr__(self):
        return (f'{self.__class__.__name__}(output_size={self.output_size}, '
                f'spatial_scale={self.spatial_scale}, sampling_ratio={self.sampling_ratio})')


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def crop_and_transform_roi_img(im, rois, resolution=224):
    """Fused form of DispRCNN3D.crop_and_transform_roi_img (disprcnn3d.py:44-50):
    ROIAlign((res,res), 1.0, 0) on the raw image + ImageNet normalisation in ONE kernel."""
    rois = torch.as_tensor(rois, dtype=torch.float32, device=im.device).reshape(-1, 5)
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32, device=im.device)
    std = torch.tensor(IMAGENET_STD, dtype=torch.float32, device=im.device)
    return roi_align_forward(im, rois, 1.0, resolution, resolution, 0, mean, std)


def prepare_stereo_rois(left_boxes, right_boxes, image_index, width, height):
    """Device-side form of the box alignment loop in ``DispRCNN3D.prepare_psmnet_input_and_target`` (disprcnn3d.py:126-146):
    ``left_boxes`` / ``right_boxes`` [R,4] CUDA f32 (the concatenated ``BoxList.bbox`` of a batch), ``image_index`` [R] (which
    image each box belongs to).  ``width`` / ``height``: the size the reference clamps with, ``left_result[i].width/.height``
    -- the UNPADDED BoxList size of the image a box belongs to (:136-141), NOT the padded ImageList tensor: one int for all
    images, or a per-image sequence / int tensor indexed by ``image_index``.
    Returns ``(rois_left, rois_right, x1s, x1ps, x2s, x2ps)``: the [R,5] crop rectangles for
    ``crop_and_transform_roi_img`` and the four int64 column tensors the reference keeps -- with no ``.tolist()`` host sync."""
    _lib.require_cuda(left_boxes, right_boxes, image_index)
    lb = left_boxes.reshape(-1, 4).contiguous().float()
    rb = right_boxes.reshape(-1, 4).contiguous().float()
    if lb.shape != rb.shape:
        raise RuntimeError(f'prepare_stereo_rois: {tuple(lb.shape)} left boxes vs {tuple(rb.shape)} right boxes')
    idx = image_index.reshape(-1).to(torch.int32).contiguous()
    R = lb.size(0)
    if idx.numel() != R:
        raise RuntimeError('prepare_stereo_rois: one image index per box expected')
    rl = torch.empty((R, 5), dtype=torch.float32, device=lb.device)
    rr = torch.empty((R, 5), dtype=torch.float32, device=lb.device)
    xs = torch.empty((4, R), dtype=torch.int64, device=lb.device)
    per_image = isinstance(width, torch.Tensor) or hasattr(width, '__len__')
    wh, n_images, w0, h0 = None, 0, 0, 0
    if per_image:
        wt = torch.as_tensor(width, dtype=torch.int32).reshape(-1)
        ht = torch.as_tensor(height, dtype=torch.int32).reshape(-1)
        if wt.numel() != ht.numel() or wt.numel() == 0:
            raise RuntimeError('prepare_stereo_rois: per-image widths and heights must have the same, non-zero length')
        wh = torch.stack([wt.to(lb.device), ht.to(lb.device)], 1).contiguous()
        n_images = wh.size(0)
    else:
        w0, h0 = int(width), int(height)
    with torch.cuda.device(lb.device):
        _lib.check(_lib.load().idisp_stereo_rois(_lib.ptr(lb), _lib.ptr(rb), _lib.ptr(idx), R, w0, h0, _lib.ptr(wh), n_images,
                                                 _lib.ptr(rl), _lib.ptr(rr), _lib.ptr(xs), _lib.stream_ptr()))
    return rl, rr, xs[0], xs[1], xs[2], xs[3]


def crop_stereo_rois(left_images, right_images, left_boxes, right_boxes, image_index, resolution=224, image_sizes=None):
    """Eval branch of ``prepare_psmnet_input_and_target`` (disprcnn3d.py:113-159) without the host loop: aligned boxes on the
    device, then the fused ROIAlign + ImageNet normalisation of both views.  ``image_sizes`` = per-image ``(width, height)`` of
    the BoxLists (``[b.size for b in left_result]``): the reference clamps every box with the unpadded size of ITS image
    (:136-141) while ``left_images`` is the padded ImageList tensor; ``None`` means an unpadded batch (all images as large as
    the tensor).  Returns ``(left_roi_images, right_roi_images, x1s, x1ps, x2s, x2ps)``."""
    if image_sizes is None:
        H, W = left_images.shape[-2:]
    else:
        sizes = torch.as_tensor(image_sizes, dtype=torch.int32).reshape(-1, 2)
        W, H = sizes[:, 0], sizes[:, 1]
    rl, rr, x1s, x1ps, x2s, x2ps = prepare_stereo_rois(left_boxes, right_boxes, image_index, W, H)
    return (crop_and_transform_roi_img(left_images, rl, resolution), crop_and_transform_roi_img(right_images, rr, resolution),
            x1s, x1ps, x2s, x2ps)
