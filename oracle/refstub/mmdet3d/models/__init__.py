from . import builder  # noqa: F401
