"""Mirror of ``femflow.solvers.mpm`` for the hot path: ``mls_mpm``, ``three_d``,
``two_d``, ``particle``, ``utils``."""
from . import three_d, two_d  # noqa: F401
