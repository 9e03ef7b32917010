from . import linalg  # noqa
