"""Stand-in for the reference's pybind extension ``disprcnn._C`` (csrc/vision.cpp:7-15; built by setup.py:46-54).

``disprcnn/layers/__init__.py:4-20`` imports nms / roi_align / roi_pool / sigmoid_focal_loss, and each of those modules does
``from disprcnn import _C`` at import time (layers/nms.py:3,8; roi_align.py:8; roi_pool.py:8; sigmoid_focal_loss.py:6).  The
reference's CUDA sources do not build against torch 2.x (THC headers), so ``disprcnn_b200.install()`` registers THIS module
under that name: ``from disprcnn.layers import ROIAlign, nms, ...`` then imports unchanged.

Only the hot-path entry point has an implementation here:
  roi_align_forward(input, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio) -> Tensor   (csrc/ROIAlign.h:11-25)
calls ``idisp_roi_align_forward`` (include/idisp.h) on the current stream.  Every other name exists -- so the import-time
attribute reads succeed -- and raises ``RuntimeError`` when CALLED: roi_align_backward like the reference's own non-CUDA build
(csrc/ROIAlign.h:44), the 2-D detector's ops (nms, roi_pool_*, sigmoid_focalloss_*) because they are not on the B200 path
(the iDispNet stage of tools/test_net.py runs on offline 2-D predictions and never calls them).
"""
from .layers.roi_align import roi_align_forward as _roi_align_forward


def roi_align_forward(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio):
    return _roi_align_forward(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio)


def _not_on_path(name, why):
    def stub(*args, **kwargs):
        raise RuntimeError(f'disprcnn._C.{name}: {why}')
    stub.__name__ = name
    stub.__doc__ = f'Placeholder for csrc/vision.cpp `{name}`: raises RuntimeError when called ({why}).'
    return stub


roi_align_backward = _not_on_path(
    'roi_align_backward', 'not implemented on the B200 inference path (the reference raises the same way for its '
    'non-CUDA build, csrc/ROIAlign.h:44)')
_DETECTOR = 'not on the B200 hot path (2-D detector op; the iDispNet stage runs on offline 2-D predictions)'
nms = _not_on_path('nms', _DETECTOR)
roi_pool_forward = _not_on_path('roi_pool_forward', _DETECTOR)
roi_pool_backward = _not_on_path('roi_pool_backward', _DETECTOR)
sigmoid_focalloss_forward = _not_on_path('sigmoid_focalloss_forward', _DETECTOR)
sigmoid_focalloss_backward = _not_on_path('sigmoid_focalloss_backward', _DETECTOR)

__all__ = ['nms', 'roi_align_forward', 'roi_align_backward', 'roi_pool_forward', 'roi_pool_backward',
           'sigmoid_focalloss_forward', 'sigmoid_focalloss_backward']
