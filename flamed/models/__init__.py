from .flamed import Flamed  # noqa: F401
