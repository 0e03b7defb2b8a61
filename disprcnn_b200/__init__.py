"""disprcnn_b200 -- B200-native (sm_100a) implementation of Disp R-CNN's instance-disparity hot path.

Public surface mirrors the reference's operator/model API for this path:
  disprcnn_b200.layers.ROIAlign / roi_align            <-> disprcnn.layers (layers/roi_align.py:49-73)
  disprcnn_b200.modeling.psmnet.stackhourglass.PSMNet  <-> disprcnn.modeling.psmnet.stackhourglass.PSMNet
  disprcnn_b200.modeling.psmnet.submodule.{convbn_3d, disparityregression, feature_extraction}
``install()`` aliases these modules under the reference's import names so that
``tools/test_net.py`` runs unchanged.
"""
import sys

__version__ = '0.1.0'


def install(inference_only=None):
    """Route the reference's import paths for the hot path to this package.

    After ``disprcnn_b200.install()``, ``from disprcnn.layers import ROIAlign`` (inside the
    reference tree) and ``from disprcnn.modeling.psmnet.stackhourglass import PSMNet`` resolve to
    the B200 implementations, and ``disprcnn._C`` (the reference's pybind extension, csrc/vision.cpp:7-15, which does
    not build against torch 2.x) resolves to ``disprcnn_b200._C`` so that ``disprcnn/layers/__init__.py:4-20`` imports.
    Only the hot-path modules are replaced; everything else in the reference keeps importing its own code.

    The replacements are INFERENCE-ONLY (ROIAlign backward and PSMNet's training mode raise): a process that also trains the
    detector or iDispNet through the reference modules must not call this.  Pass ``inference_only=True`` to acknowledge;
    without it a warning says so once.
    """
    import importlib
    import warnings
    if inference_only is None:
        warnings.warn('disprcnn_b200.install(): the B200 ROIAlign / PSMNet are inference-only (backward and training mode '
                      'raise); pass inference_only=True to acknowledge', stacklevel=2)
    elif not inference_only:
        raise RuntimeError('disprcnn_b200.install(inference_only=False): there is no training path to install')
    _ra = importlib.import_module(__name__ + '.layers.roi_align')  # (the package re-exports a function of the same name)
    _c = importlib.import_module(__name__ + '._C')
    from .modeling.psmnet import stackhourglass as _sh, submodule as _sm
    sys.modules['disprcnn._C'] = _c
    pkg = sys.modules.get('disprcnn')
    if pkg is not None:
        pkg._C = _c
    sys.modules['disprcnn.layers.roi_align'] = _ra
    sys.modules['disprcnn.modeling.psmnet.stackhourglass'] = _sh
    sys.modules['disprcnn.modeling.psmnet.submodule'] = _sm
    layers = sys.modules.get('disprcnn.layers')
    if layers is not None:  # already imported: patch the re-exported names (layers/__init__.py:10-11)
        layers.ROIAlign = _ra.ROIAlign
        layers.roi_align = _ra.roi_align
