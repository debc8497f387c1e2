"""Mirror of the reference's ``utils`` package for the hot path (bbox_utils, train_utils)."""
from . import bbox_utils, train_utils  # noqa: F401
