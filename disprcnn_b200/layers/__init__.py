"""Operator layer of the hot path (mirror of disprcnn/layers/__init__.py:10-11 for ROIAlign)."""
from .roi_align import ROIAlign, roi_align

__all__ = ['roi_align', 'ROIAlign']
