"""Hand-off of the per-ROI disparity maps iDispNet returns.

Two reference call sites, one fused kernel each (csrc/roi_paste.cu), both without the per-ROI Python loop (``.tolist()`` syncs,
one image-sized ``torch.zeros`` + ``interpolate`` + slice-assign per ROI):

* ``paste_roi_disparity``  <->  ``DispRCNN3D.roi_disp_postprocess`` (disprcnn/modeling/detector/disprcnn3d.py:161-190): resize each
  ROI's [S,S] disparity to its integer-expanded box (``DisparityMap.resize / crop``, structures/disparity.py:39-78), add x1 - x1p,
  clamp at 0, mask, and take the per-image maximum over the ROIs -> [N,H,W].
* ``roi_depth_maps``  <->  the depth part of ``PointRCNN.process_input`` (modeling/pointnet_module/point_rcnn/lib/net/
  point_rcnn.py:113-136): the same resize + shift, depth = fu*baseline / (disp + 1e-6) inside the box, 0 elsewhere -> [R,H,W].
"""
import torch

from .. import _lib


def _boxes(left_boxes, right_boxes):
    lb = left_boxes.reshape(-1, 4).contiguous().float()
    rb = right_boxes.reshape(-1, 4).contiguous().float()
    if lb.shape != rb.shape:
        raise RuntimeError(f'{tuple(lb.shape)} left boxes vs {tuple(rb.shape)} right boxes')
    return lb, rb


def paste_roi_disparity(roi_disp, left_boxes, right_boxes, rois_per_image, height, width, masks=None):
    """roi_disp [R,S,S] f32 CUDA (ROIs grouped by image, as ``torch.split(output, [len(a) for a in left_result])`` assumes,
    disprcnn3d.py:162); left_boxes / right_boxes [R,4]; rois_per_image: list of ints (sum = R); masks: optional [R,H,W] bool /
    uint8 (``masker(...)`` output, :166).  Returns the per-image disparity maps [N,H,W] (the tensor ``lr.add_map('disparity', .)``
    receives, :188)."""
    _lib.require_cuda(roi_disp, left_boxes, right_boxes, masks)
    roi_disp = roi_disp.contiguous().float()
    R, S = roi_disp.shape[0], roi_disp.shape[-1]
    if roi_disp.dim() != 3 or roi_disp.shape[1] != S:
        raise RuntimeError('paste_roi_disparity: roi_disp must be [R,S,S]')
    lb, rb = _boxes(left_boxes, right_boxes)
    counts = [int(c) for c in rois_per_image]
    if sum(counts) != R or lb.shape[0] != R:
        raise RuntimeError(f'paste_roi_disparity: {R} ROI maps, {lb.shape[0]} boxes, rois_per_image sums to {sum(counts)}')
    N = len(counts)
    starts = torch.tensor([0] + list(torch.tensor(counts, dtype=torch.int64).cumsum(0).tolist()) if N else [0], dtype=torch.int32, device=roi_disp.device)
    out = torch.empty((N, height, width), dtype=torch.float32, device=roi_disp.device)
    if N == 0:
        return out
    m = None
    if masks is not None:
        m = masks.reshape(R, height, width).to(torch.uint8).contiguous()
    with torch.cuda.device(roi_disp.device):
        _lib.check(_lib.load().idisp_roi_disparity_paste(_lib.ptr(roi_disp), R, S, _lib.ptr(lb), _lib.ptr(rb), _lib.ptr(starts), N, _lib.ptr(m),
                                                         int(height), int(width), _lib.ptr(out), _lib.stream_ptr()))
    return out


def roi_depth_maps(roi_disp, left_boxes, right_boxes, fu_baseline, height, width):
    """roi_disp [R,S,S], boxes [R,4], fu_baseline [R] (``calib.stereo_fuxbaseline`` of the image each ROI belongs to) ->
    per-ROI image-sized depth maps [R,H,W], zero outside the ROI's box (point_rcnn.py:124-134)."""
    _lib.require_cuda(roi_disp, left_boxes, right_boxes, fu_baseline)
    roi_disp = roi_disp.contiguous().float()
    R, S = roi_disp.shape[0], roi_disp.shape[-1]
    lb, rb = _boxes(left_boxes, right_boxes)
    fub = fu_baseline.reshape(-1).contiguous().float()
    if lb.shape[0] != R or fub.numel() != R:
        raise RuntimeError('roi_depth_maps: one box pair and one fu*baseline per ROI expected')
    out = torch.empty((R, height, width), dtype=torch.float32, device=roi_disp.device)
    if R == 0:
        return out
    with torch.cuda.device(roi_disp.device):
        _lib.check(_lib.load().idisp_roi_depth_paste(_lib.ptr(roi_disp), R, S, _lib.ptr(lb), _lib.ptr(rb), _lib.ptr(fub), int(height), int(width),
                                                     _lib.ptr(out), _lib.stream_ptr()))
    return out
