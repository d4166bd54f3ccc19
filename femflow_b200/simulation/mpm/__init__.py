from .simulation import MPMSimulation  # noqa: F401
