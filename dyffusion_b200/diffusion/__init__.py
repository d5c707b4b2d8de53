from .dyffusion import DYffusion, InterpolatorHandle  # noqa: F401
