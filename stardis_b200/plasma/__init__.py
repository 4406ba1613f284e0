from .columnar import ColumnarLines  # noqa: F401
from .synthetic import SyntheticPlasma, create_synthetic_plasma  # noqa: F401
