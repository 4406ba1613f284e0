from .base import StellarModel, Radial1DGeometry, Composition  # noqa: F401
