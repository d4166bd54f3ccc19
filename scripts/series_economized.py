"""Economised series for the fp32 stress (femflow_b200/csrc/mpm_math.cuh).  LEFT form (the one the kernels use):
(F - R) F^T = h(G) = G p(G) with G = F F^T - I and p(x) = (1 + x - sqrt(1 + x)) / x fitted on [-r_t, r_t] by Chebyshev
interpolation (near-minimax) instead of truncating its Taylor series.  Prints the tier table that mpm_math.cuh embeds
(stress_coef / stress_tier_r, TIERS_LEFT below); the CPU suite regenerates it and compares
(tests/test_kernel_math_host.py::test_economised_series_table_is_the_generated_one).  The right form
M = I - (I+G)^(-1/2) = G q(G) with G = F^T F - I, which round 1 shipped as a truncated Taylor series and needs two
more products with F, is kept here for the comparison only.

Tiers are chosen so that the uniform error of q on the tier's interval, WITH the coefficients rounded to fp32,
stays below 5e-8 (relative 1e-7 on M, whose leading coefficient is 1/2)."""
import numpy as np
from numpy.polynomial import chebyshev as Ch, polynomial as P

# (upper bound of ||G||_F for the tier, degree of q)
TIERS = ((0.007, 2), (0.025, 3), (0.069, 4), (0.109, 5), (0.15, 6))
# The LEFT form: (F - R) F^T = B - B^(1/2) with B = F F^T = I + G_B, i.e. the stress
# term is the matrix function h(G_B) = G_B p(G_B), p(x) = (1 + x - sqrt(1 + x)) / x = 1 - 1 / (sqrt(1 + x) + 1), of the
# left Cauchy-Green strain -- no product with F at all, and the square-root series converges faster than the
# inverse square root's: degree 4 up to ||G||_F = 0.112.
TIERS_LEFT = ((0.0136, 2), (0.0436, 3), (0.112, 4), (0.15, 5))


def q_exact(x):
    x = np.asarray(x, dtype=np.float64)
    out = np.full_like(x, 0.5)
    nz = np.abs(x) > 1e-9
    out[nz] = (1 - (1 + x[nz]) ** -0.5) / x[nz]
    return out


def p_exact(x):
    x = np.asarray(x, dtype=np.float64)
    return 1 - 1 / (np.sqrt(1 + x) + 1)


def coefficients(r, deg, fn=q_exact):
    """Power-basis coefficients (c_0 .. c_deg, as float32) of the Chebyshev interpolant of fn on [-r, r]."""
    c = Ch.chebinterpolate(lambda t: fn(t * r), deg)
    return (Ch.cheb2poly(c) / r ** np.arange(deg + 1)).astype(np.float32)


def table(left=False):
    return [(r, deg, coefficients(r, deg, p_exact if left else q_exact)) for r, deg in (TIERS_LEFT if left else TIERS)]


def uniform_error(r, coeffs, fn=q_exact):
    xs = np.linspace(-r, r, 20001)
    return float(np.max(np.abs(P.polyval(xs, coeffs.astype(np.float64)) - fn(xs))) / 0.5)


def worst_matrix_error(r, coeffs, n=1500, seed=0, left=False):
    """As scripts/series_degree.py: symmetric G of Frobenius norm r (every third one rank one: spectral radius = r),
    Horner in float32 as the kernel evaluates it, against an eigendecomposition (left: of 1 + x - sqrt(1 + x))."""
    rng = np.random.default_rng(seed)
    worst = 0.0
    for t in range(n):
        a = rng.uniform(-1, 1, (3, 3))
        g = (a + a.T) / 2
        if t % 3 == 0:
            v = rng.normal(size=3)
            v /= np.linalg.norm(v)
            g = np.outer(v, v) * rng.choice([-1, 1])
        g = g / np.linalg.norm(g) * r
        w, q = np.linalg.eigh(g)
        exact = q @ np.diag(1 + w - np.sqrt(1 + w) if left else 1 - 1 / np.sqrt(1 + w)) @ q.T
        g32 = g.astype(np.float32)
        acc = coeffs[-1] * g32 + coeffs[-2] * np.eye(3, dtype=np.float32)
        for c in coeffs[-3::-1]:
            acc = (g32 @ acc).astype(np.float32) + c * np.eye(3, dtype=np.float32)
        series = (g32 @ acc).astype(np.float64)
        worst = max(worst, np.abs(series - exact).max() / np.abs(exact).max())
    return worst


if __name__ == "__main__":
    for left in (False, True):
        print("// left form: p(x) = (1 + x - sqrt(1 + x)) / x" if left else "// right form: q(x) = (1 - (1 + x)^(-1/2)) / x")
        for r, deg, c in table(left):
            print(f"// ||G||_F < {r}: degree {deg}; uniform error {uniform_error(r, c, p_exact if left else q_exact):.1e}, "
                  f"worst matrix error in fp32 {worst_matrix_error(r, c, 600, left=left):.1e}")
            print("  {" + ", ".join(f"{float(v)!r}f" for v in c) + "},")
