from .nms import nms
from .roi_align import ROIAlign, roi_align
from .roi_pool import ROIPool, roi_pool

__all__ = ["nms", "roi_align", "ROIAlign", "roi_pool", "ROIPool"]
