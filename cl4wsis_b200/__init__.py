"""cl4wsis_b200 — B200-native (sm_100a) pseudo-label hot path of CL4WSIS.

Drop-in mirrors of the reference API for this path (same names and signatures):

    cl4wsis_b200.wss.modules.PAMR                          wss/modules.py:122-152
    cl4wsis_b200.wss.utils.peak_extract                    wss/utils.py:3-25
    cl4wsis_b200.modules.utils.find_instance_center        modules/utils.py:463-502
    cl4wsis_b200.modules.utils.group_pixels                modules/utils.py:505-542
    cl4wsis_b200.modules.utils.get_instance_segmentation   modules/utils.py:545-606
    cl4wsis_b200.modules.utils.refine_label_generation     modules/utils.py:257-385
    cl4wsis_b200.modules.utils.refine_label_generation_with_point   modules/utils.py:388-460
    cl4wsis_b200.dataset.utils.get_ins_map                 dataset/utils.py:795-902
    cl4wsis_b200.wss.single_stage.phase1_pseudo_labels     train.py:372-385 (+ pseudo_gtmask, wss/single_stage.py:18-40)
    cl4wsis_b200.wss.utils.cam_peaks / cam_normalize / smoothing   train.py:426-436, wss/modules.py:425-434
    cl4wsis_b200.wss.modules.LocalAffinity[Abs|Copy], LocalStDev   wss/modules.py:17-119

All compute happens in hand-written CUDA kernels behind the C ABI in
``include/cl4wsis_b200.h`` (``libcl4wsis_b200.so``); there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .wss.modules import PAMR  # noqa: F401
from .wss.utils import peak_extract  # noqa: F401
from .modules.utils import (find_instance_center, get_instance_segmentation, group_pixels,  # noqa: F401
                            refine_label_generation)
from .pipeline import HostPseudoLabelPipeline, PseudoLabelStep  # noqa: F401

__version__ = "0.1.0"


def patch_reference(wss_modules=None, wss_utils=None, modules_utils=None):
    """Swap the reference's hot-path symbols for this package's in already-imported
    reference modules.  Pass the reference's ``wss.modules``,
    ``wss.utils`` and/or ``modules.utils`` module objects."""
    from .modules import utils as mu
    from .wss import modules as wm
    from .wss import utils as wu
    if wss_modules is not None:
        wss_modules.PAMR = wm.PAMR
    if wss_utils is not None:
        wss_utils.peak_extract = wu.peak_extract
        wss_utils.smoothing = wu.smoothing
    if modules_utils is not None:
        modules_utils.find_instance_center = mu.find_instance_center
        modules_utils.group_pixels = mu.group_pixels
        modules_utils.get_instance_segmentation = mu.get_instance_segmentation
        modules_utils.refine_label_generation = mu.refine_label_generation
        modules_utils.refine_label_generation_with_point = mu.refine_label_generation_with_point
