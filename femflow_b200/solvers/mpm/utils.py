"""Host-side helpers of the MPM path (reference solvers/mpm/utils.py, numerics/fem.py)."""


def Ev_to_mu(E: float, v: float) -> float:
    """numerics/fem.py:1-2"""
    return E / (2 * (1 + v))


def Ev_to_lambda(E: float, v: float) -> float:
    """numerics/fem.py:5-6"""
    return E * v / ((1 + v) * (1 - 2 * v))


def constant_hardening(mu_0: float, lambda_0: float, e: float):
    """solvers/mpm/utils.py:7-24: the hardening coefficient is a plain multiplier."""
    return mu_0 * e, lambda_0 * e
