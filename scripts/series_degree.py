"""Truncation error of the series M = I - (I+G)^(-1/2) = sum_k c_k G^k used by the fp32 stress
(femflow_b200/csrc/mpm_math.cuh, fixed_corotated_affine3_f32) as a function of the Frobenius
norm r of G, against an eigendecomposition.  Justifies the degree thresholds 3 / 5 / 8."""
import numpy as np

C = [0.5, -0.375, 0.3125, -0.2734375, 0.24609375, -0.2255859375, 0.20947265625, -0.196380615234375]


def worst_error(r, degree, n=3000, seed=0):
    rng = np.random.default_rng(seed)
    worst = 0.0
    for t in range(n):
        a = rng.uniform(-1, 1, (3, 3))
        g = (a + a.T) / 2
        if t % 3 == 0:      # rank one: spectral radius == Frobenius norm, the worst case of the bound
            v = rng.normal(size=3)
            v /= np.linalg.norm(v)
            g = np.outer(v, v) * rng.choice([-1, 1])
        g = g / np.linalg.norm(g) * r
        w, q = np.linalg.eigh(g)
        exact = q @ np.diag(1 - 1 / np.sqrt(1 + w)) @ q.T
        series = sum(C[k] * np.linalg.matrix_power(g, k + 1) for k in range(degree))
        worst = max(worst, np.abs(series - exact).max() / np.abs(exact).max())
    return worst


if __name__ == "__main__":
    for degree, r in ((3, 0.005), (5, 0.04), (8, 0.15)):
        print(f"degree {degree}  r < {r}:  worst relative error {worst_error(r, degree):.2e}")
