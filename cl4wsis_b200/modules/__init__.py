"""Mirror of the reference's ``modules.utils`` post-processing functions for the hot path."""
from .utils import find_instance_center, get_instance_segmentation, group_pixels  # noqa: F401
