from .marcs import read_marcs_model, MARCSModel  # noqa: F401
