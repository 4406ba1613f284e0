from .base import Opacities  # noqa: F401
