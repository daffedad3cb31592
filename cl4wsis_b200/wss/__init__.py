"""Mirror of the reference's ``wss`` package for the hot path (PAMR, peak_extract)."""
from .modules import PAMR, LocalAffinity, LocalAffinityAbs, LocalAffinityCopy, LocalStDev  # noqa: F401
from .utils import peak_extract  # noqa: F401
