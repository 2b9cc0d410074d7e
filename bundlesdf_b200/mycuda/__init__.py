"""Drop-in for the reference's `mycuda` package (mycuda/setup.py:18-41 builds `common` and `gridencoder`):
same module names and function signatures, backed by libnof_sm100.so instead of pybind extensions."""
from . import common, gridencoder  # noqa: F401
