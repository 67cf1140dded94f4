from . import adam  # noqa: F401
