"""Mirror of the reference's ``utils`` package for the one function on the PAMR path (``denorm``)."""
from .utils import denorm  # noqa: F401
