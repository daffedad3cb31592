"""CPU parity oracle for the CL4WSIS pseudo-label hot path (TEST INFRASTRUCTURE ONLY).

Importable only from tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` / ``torch_gpu_baseline`` legs (``oracle.torch_ref``: the same path restated
with stock PyTorch ops, the second oracle that runs on CUDA tensors).  The product package
``cl4wsis_b200`` never imports this.
"""
from .oracle import (  # noqa: F401
    build, pamr, pamr_weights, local_affinity, local_stdev, peak_extract, find_instance_center, group_pixels,
    get_instance_segmentation, cluster_peaks, resize_bilinear_ac, num_threads, set_num_threads,
)
from . import labelgen  # noqa: F401,E402  (numpy/OpenCV restatement of the callers: smoothing, label generation)
