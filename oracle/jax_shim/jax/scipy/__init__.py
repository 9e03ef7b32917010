from . import sparse, ndimage  # noqa
