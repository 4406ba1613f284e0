from .base import RadiationField, create_stellar_radiation_field  # noqa: F401
