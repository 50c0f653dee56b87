"""Import shim for `skimage` (bsuite/utils/wrappers.py:26); only ImageObservation calls it."""
from skimage import transform  # noqa: F401
