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


def install():
    """Route the reference's import paths for the hot path to this package.

    After ``disprcnn_b200.install()``, ``from disprcnn.layers import ROIAlign`` (inside the
    reference tree) and ``from disprcnn.modeling.psmnet.stackhourglass import PSMNet`` resolve to
    the B200 implementations.  Only the hot-path modules are replaced; everything else in the
    reference keeps importing its own code.
    """
    import importlib
    _ra = importlib.import_module(__name__ + '.layers.roi_align')  # (the package re-exports a function of the same name)
    from .modeling.psmnet import stackhourglass as _sh, submodule as _sm
    sys.modules['disprcnn.layers.roi_align'] = _ra
    sys.modules['disprcnn.modeling.psmnet.stackhourglass'] = _sh
    sys.modules['disprcnn.modeling.psmnet.submodule'] = _sm
    layers = sys.modules.get('disprcnn.layers')
    if layers is not None:  # already imported: patch the re-exported names (layers/__init__.py:10-11)
        layers.ROIAlign = _ra.ROIAlign
        layers.roi_align = _ra.roi_align
