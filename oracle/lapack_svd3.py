"""TEST INFRASTRUCTURE (oracle): a scalar restatement of the operation sequence LAPACK's DGESDD takes for a
3x3 matrix -- the routine behind ``np.linalg.svd`` in femflow/solvers/mpm/three_d/g2p.py:49.

Why it exists: the reference's 3D snow branch computes ``U @ diag(clip(sig)) @ Vh.T`` (g2p.py:55 -- ``Vh`` transposed
once more, not ``Vh``), which is NOT invariant under the sign freedom (u_i, v_i) -> (-u_i, -v_i) of an SVD.  Its result
is therefore a function of the signs LAPACK happens to return, and those are a function of the operation sequence:
Householder bidiagonalisation (DGEBD2), implicit-shift / zero-shift QR sweeps on the bidiagonal (DBDSQR through
DBDSDC -> DLASDQ for n <= 25), "make singular values positive" by negating rows of VT, sort, back-transformation
(DORMBR).  LAPACK is a third-party dependency absent from /root/reference (numpy links OpenBLAS 0.3.30 = LAPACK 3.11
in the build container); this file restates its published algorithm and is pinned against ``np.linalg.svd`` itself
(tests/test_lapack_svd3.py: signs of every singular-vector pair on 10^5 matrices) and through the reference's outputs
``F_out`` / ``Jp_out`` of tests/golden/snow3d.npz.

DLARTG is the LAPACK >= 3.10 version (c >= 0, r carries the sign of f).
"""
from __future__ import annotations

import math

import numpy as np

EPS = 2.0 ** -53          # DLAMCH('Epsilon') (rounding unit, as LAPACK defines it)
UNFL = 2.2250738585072014e-308
_RTMIN = math.sqrt(UNFL)
_RTMAX = math.sqrt(1.0 / UNFL / 2)


def _sign(a, b):
    """Fortran SIGN(a, b)."""
    return abs(a) if (b > 0 or (b == 0 and math.copysign(1.0, b) > 0)) else -abs(a)


def dlartg(f, g):
    """LAPACK 3.10+ la_lartg: returns c, s, r with c >= 0."""
    if g == 0.0:
        return 1.0, 0.0, f
    if f == 0.0:
        return 0.0, _sign(1.0, g), abs(g)
    d = math.sqrt(f * f + g * g)
    c = abs(f) / d
    r = _sign(d, f)
    return c, g / r, r


def dlarfg(alpha, x):
    """Householder reflector H = I - tau [1; v][1; v]^T with H [alpha; x] = [beta; 0]."""
    xnorm = math.sqrt(sum(t * t for t in x))
    if len(x) == 0 or xnorm == 0.0:
        return alpha, [0.0] * len(x), 0.0
    beta = -_sign(math.hypot(alpha, xnorm), alpha)
    tau = (beta - alpha) / beta
    scale = 1.0 / (alpha - beta)
    return beta, [t * scale for t in x], tau


def dlas2(f, g, h):
    fa, ga, ha = abs(f), abs(g), abs(h)
    fhmn, fhmx = min(fa, ha), max(fa, ha)
    if fhmn == 0.0:
        return 0.0
    if ga < fhmx:
        a_s = 1.0 + fhmn / fhmx
        at = (fhmx - fhmn) / fhmx
        au = (ga / fhmx) ** 2
        c = 2.0 / (math.sqrt(a_s * a_s + au) + math.sqrt(at * at + au))
        return fhmn * c
    au = fhmx / ga
    if au == 0.0:
        return (fhmn * fhmx) / ga
    a_s = 1.0 + fhmn / fhmx
    at = (fhmx - fhmn) / fhmx
    c = 1.0 / (math.sqrt(1.0 + (a_s * au) ** 2) + math.sqrt(1.0 + (at * au) ** 2))
    return 2.0 * ((fhmn * c) * au)


def dlasv2(f, g, h):
    """SVD of [[f, g], [0, h]]: returns ssmin, ssmax, snr, csr, snl, csl."""
    ft, fa, ht, ha = f, abs(f), h, abs(h)
    pmax = 1
    swap = ha > fa
    if swap:
        pmax = 3
        ft, ht = ht, ft
        fa, ha = ha, fa
    gt, ga = g, abs(g)
    if ga == 0.0:
        ssmin, ssmax, clt, crt, slt, srt = ha, fa, 1.0, 1.0, 0.0, 0.0
    else:
        gasmal = True
        if ga > fa:
            pmax = 2
            if fa / ga < EPS:
                gasmal = False
                ssmax = ga
                ssmin = fa / (ga / ha) if ha > 1.0 else (fa / ga) * ha
                clt, slt, srt, crt = 1.0, ht / gt, 1.0, ft / gt
        if gasmal:
            d = fa - ha
            l = 1.0 if d == fa else d / fa
            m = gt / ft
            t = 2.0 - l
            mm, tt = m * m, t * t
            s = math.sqrt(tt + mm)
            r = abs(m) if l == 0.0 else math.sqrt(l * l + mm)
            a = 0.5 * (s + r)
            ssmin, ssmax = ha / a, fa * a
            if mm == 0.0:
                if l == 0.0:
                    t = _sign(2.0, ft) * _sign(1.0, gt)
                else:
                    t = gt / _sign(d, ft) + m / t
            else:
                t = (m / (s + t) + m / (r + l)) * (1.0 + a)
            l = math.sqrt(t * t + 4.0)
            crt, srt = 2.0 / l, t / l
            clt = (crt + srt * m) / a
            slt = (ht / ft) * srt / a
    if swap:
        csl, snl, csr, snr = srt, crt, slt, clt
    else:
        csl, snl, csr, snr = clt, slt, crt, srt
    if pmax == 1:
        tsign = _sign(1.0, csr) * _sign(1.0, csl) * _sign(1.0, f)
    elif pmax == 2:
        tsign = _sign(1.0, snr) * _sign(1.0, csl) * _sign(1.0, g)
    else:
        tsign = _sign(1.0, snr) * _sign(1.0, snl) * _sign(1.0, h)
    ssmax = _sign(ssmax, tsign)
    ssmin = _sign(ssmin, tsign * _sign(1.0, f) * _sign(1.0, h))
    return ssmin, ssmax, snr, csr, snl, csl


def _rot_rows(vt, i, j, c, s):
    """DROT / one step of DLASR('L','V',.): rows i, j of vt."""
    for k in range(3):
        a, b = vt[i][k], vt[j][k]
        vt[i][k] = c * a + s * b
        vt[j][k] = c * b - s * a


def _rot_cols(u, i, j, c, s):
    for k in range(3):
        a, b = u[k][i], u[k][j]
        u[k][i] = c * a + s * b
        u[k][j] = c * b - s * a


def dbdsqr3(d, e, vt, u):
    """DBDSQR('U', n=3, ncvt=3, nru=3) in place on d[3], e[2], vt (rows rotate), u (columns rotate)."""
    n = 3
    tolmul = max(10.0, min(100.0, EPS ** -0.125))
    tol = tolmul * EPS
    smax = max(max(abs(t) for t in d), max(abs(t) for t in e))
    sminoa = abs(d[0])
    if sminoa != 0.0:
        mu = sminoa
        for i in range(1, n):
            mu = abs(d[i]) * (mu / (mu + abs(e[i - 1])))
            sminoa = min(sminoa, mu)
            if sminoa == 0.0:
                break
    sminoa /= math.sqrt(float(n))
    thresh = max(tol * sminoa, 6 * (n * (n * UNFL)))
    maxit = 6 * n * n
    it = 0
    oldll, oldm = -1, -1
    idir = 0
    m = n                     # 1-based as in the Fortran text; D(i) = d[i-1]
    D = lambda i: d[i - 1]
    E = lambda i: e[i - 1]
    while m > 1:
        if it > maxit:
            raise RuntimeError("dbdsqr3: no convergence")
        smax = abs(D(m))
        split = False
        ll = 0
        for lll in range(1, m):
            ll = m - lll
            abss, abse = abs(D(ll)), abs(E(ll))
            if abse <= thresh:
                split = True
                break
            smax = max(smax, abss, abse)
        if split:
            e[ll - 1] = 0.0
            if ll == m - 1:
                m -= 1
                continue
        else:
            ll = 0
        ll += 1
        if ll == m - 1:
            sigmn, sigmx, sinr, cosr, sinl, cosl = dlasv2(D(m - 1), E(m - 1), D(m))
            d[m - 2], e[m - 2], d[m - 1] = sigmx, 0.0, sigmn
            _rot_rows(vt, m - 2, m - 1, cosr, sinr)
            _rot_cols(u, m - 2, m - 1, cosl, sinl)
            m -= 2
            continue
        if ll > oldm or m < oldll:
            idir = 1 if abs(D(ll)) >= abs(D(m)) else 2
        sminl = 0.0
        if idir == 1:
            if abs(E(m - 1)) <= abs(tol) * abs(D(m)):
                e[m - 2] = 0.0
                continue
            mu = abs(D(ll))
            sminl = mu
            conv = False
            for lll in range(ll, m):
                if abs(E(lll)) <= tol * mu:
                    e[lll - 1] = 0.0
                    conv = True
                    break
                mu = abs(D(lll + 1)) * (mu / (mu + abs(E(lll))))
                sminl = min(sminl, mu)
            if conv:
                continue
        else:
            if abs(E(ll)) <= abs(tol) * abs(D(ll)):
                e[ll - 1] = 0.0
                continue
            mu = abs(D(m))
            sminl = mu
            conv = False
            for lll in range(m - 1, ll - 1, -1):
                if abs(E(lll)) <= tol * mu:
                    e[lll - 1] = 0.0
                    conv = True
                    break
                mu = abs(D(lll)) * (mu / (mu + abs(E(lll))))
                sminl = min(sminl, mu)
            if conv:
                continue
        oldll, oldm = ll, m
        if n * tol * (sminl / smax) <= max(EPS, 0.01 * tol):
            shift = 0.0
        else:
            if idir == 1:
                sll = abs(D(ll))
                shift = dlas2(D(m - 1), E(m - 1), D(m))
            else:
                sll = abs(D(m))
                shift = dlas2(D(ll), E(ll), D(ll + 1))
            if sll > 0.0 and (shift / sll) ** 2 < EPS:
                shift = 0.0
        it += m - ll
        if shift == 0.0:
            if idir == 1:
                cs, oldcs, oldsn = 1.0, 1.0, 0.0
                rots = []
                for i in range(ll, m):
                    cs, sn, r = dlartg(D(i) * cs, E(i))
                    if i > ll:
                        e[i - 2] = oldsn * r
                    oldcs, oldsn, d[i - 1] = dlartg(oldcs * r, D(i + 1) * sn)
                    rots.append((i, cs, sn, oldcs, oldsn))
                h = D(m) * cs
                d[m - 1] = h * oldcs
                e[m - 2] = h * oldsn
                for (i, c1, s1, c2, s2) in rots:          # DLASR(.,'V','F'): pairs (i, i+1) in forward order
                    _rot_rows(vt, i - 1, i, c1, s1)
                for (i, c1, s1, c2, s2) in rots:
                    _rot_cols(u, i - 1, i, c2, s2)
                if abs(E(m - 1)) <= thresh:
                    e[m - 2] = 0.0
            else:
                cs, oldcs, oldsn = 1.0, 1.0, 0.0
                rots = []
                for i in range(m, ll, -1):
                    cs, sn, r = dlartg(D(i) * cs, E(i - 1))
                    if i < m:
                        e[i - 1] = oldsn * r
                    oldcs, oldsn, d[i - 1] = dlartg(oldcs * r, D(i - 1) * sn)
                    rots.append((i, cs, -sn, oldcs, -oldsn))
                h = D(ll) * cs
                d[ll - 1] = h * oldcs
                e[ll - 1] = h * oldsn
                for (i, c1, s1, c2, s2) in rots:          # DLASR(.,'V','B'): pairs (i-1, i) from the bottom up
                    _rot_rows(vt, i - 2, i - 1, c2, s2)
                for (i, c1, s1, c2, s2) in rots:
                    _rot_cols(u, i - 2, i - 1, c1, s1)
                if abs(E(ll)) <= thresh:
                    e[ll - 1] = 0.0
        else:
            if idir == 1:
                f = (abs(D(ll)) - shift) * (_sign(1.0, D(ll)) + shift / D(ll))
                g = E(ll)
                rots = []
                for i in range(ll, m):
                    cosr, sinr, r = dlartg(f, g)
                    if i > ll:
                        e[i - 2] = r
                    f = cosr * D(i) + sinr * E(i)
                    e[i - 1] = cosr * E(i) - sinr * D(i)
                    g = sinr * D(i + 1)
                    d[i] = cosr * D(i + 1)
                    cosl, sinl, r = dlartg(f, g)
                    d[i - 1] = r
                    f = cosl * E(i) + sinl * D(i + 1)
                    d[i] = cosl * D(i + 1) - sinl * E(i)
                    if i < m - 1:
                        g = sinl * E(i + 1)
                        e[i] = cosl * E(i + 1)
                    rots.append((i, cosr, sinr, cosl, sinl))
                e[m - 2] = f
                for (i, c1, s1, c2, s2) in rots:
                    _rot_rows(vt, i - 1, i, c1, s1)
                for (i, c1, s1, c2, s2) in rots:
                    _rot_cols(u, i - 1, i, c2, s2)
                if abs(E(m - 1)) <= thresh:
                    e[m - 2] = 0.0
            else:
                f = (abs(D(m)) - shift) * (_sign(1.0, D(m)) + shift / D(m))
                g = E(m - 1)
                rots = []
                for i in range(m, ll, -1):
                    cosr, sinr, r = dlartg(f, g)
                    if i < m:
                        e[i - 1] = r
                    f = cosr * D(i) + sinr * E(i - 1)
                    e[i - 2] = cosr * E(i - 1) - sinr * D(i)
                    g = sinr * D(i - 1)
                    d[i - 2] = cosr * D(i - 1)
                    cosl, sinl, r = dlartg(f, g)
                    d[i - 1] = r
                    f = cosl * E(i - 1) + sinl * D(i - 1)
                    d[i - 2] = cosl * D(i - 1) - sinl * E(i - 1)
                    if i > ll + 1:
                        g = sinl * E(i - 2)
                        e[i - 3] = cosl * E(i - 2)
                    rots.append((i, cosr, -sinr, cosl, -sinl))
                e[ll - 1] = f
                if abs(E(ll)) <= thresh:
                    e[ll - 1] = 0.0
                for (i, c1, s1, c2, s2) in rots:
                    _rot_rows(vt, i - 2, i - 1, c2, s2)
                for (i, c1, s1, c2, s2) in rots:
                    _rot_cols(u, i - 2, i - 1, c1, s1)
    # make the singular values positive (rows of VT change sign), then sort into decreasing order
    for i in range(n):
        if d[i] < 0.0:
            d[i] = -d[i]
            vt[i] = [-t for t in vt[i]]
    for i in range(1, n):
        isub, smin = 1, d[0]
        for j in range(2, n + 2 - i):
            if d[j - 1] <= smin:
                isub, smin = j, d[j - 1]
        last = n + 1 - i
        if isub != last:
            d[isub - 1], d[last - 1] = d[last - 1], d[isub - 1]
            vt[isub - 1], vt[last - 1] = vt[last - 1], vt[isub - 1]
            for k in range(3):
                u[k][isub - 1], u[k][last - 1] = u[k][last - 1], u[k][isub - 1]


def svd3(a):
    """U, sig, Vh of a 3x3 matrix the way DGESDD (JOBZ='A', path 5: DGEBD2 -> DBDSDC -> DORMBR) returns them."""
    a = [[float(a[i][j]) for j in range(3)] for i in range(3)]
    # DGEBD2, m = n = 3: H1 (column 1), G1 (row 1, columns 2:3), H2 (column 2, rows 2:3); H3, G2 are identities
    d1, v1, tauq1 = dlarfg(a[0][0], [a[1][0], a[2][0]])
    h1 = [1.0] + v1
    for j in (1, 2):
        w = sum(h1[i] * a[i][j] for i in range(3))
        for i in range(3):
            a[i][j] -= tauq1 * h1[i] * w
    e1, g1v, taup1 = dlarfg(a[0][1], [a[0][2]])
    g1 = [1.0] + g1v
    for i in (1, 2):
        w = a[i][1] * g1[0] + a[i][2] * g1[1]
        a[i][1] -= taup1 * w * g1[0]
        a[i][2] -= taup1 * w * g1[1]
    d2, v2, tauq2 = dlarfg(a[1][1], [a[2][1]])
    h2 = [1.0] + v2
    w = h2[0] * a[1][2] + h2[1] * a[2][2]
    a[1][2] -= tauq2 * h2[0] * w
    a[2][2] -= tauq2 * h2[1] * w
    e2, d3 = a[1][2], a[2][2]
    d, e = [d1, d2, d3], [e1, e2]
    u = [[1.0 if i == j else 0.0 for j in range(3)] for i in range(3)]
    vt = [[1.0 if i == j else 0.0 for j in range(3)] for i in range(3)]
    dbdsqr3(d, e, vt, u)
    # DBDSDC's own selection sort (a no-op after DBDSQR's, kept for the record): decreasing order already.
    # DORMBR('Q','L','N'): U <- H1 H2 U
    for j in range(3):
        w = h2[0] * u[1][j] + h2[1] * u[2][j]
        u[1][j] -= tauq2 * h2[0] * w
        u[2][j] -= tauq2 * h2[1] * w
    for j in range(3):
        w = sum(h1[i] * u[i][j] for i in range(3))
        for i in range(3):
            u[i][j] -= tauq1 * h1[i] * w
    # DORMBR('P','R','T'): VT <- VT G1 (columns 2:3)
    for i in range(3):
        w = vt[i][1] * g1[0] + vt[i][2] * g1[1]
        vt[i][1] -= taup1 * w * g1[0]
        vt[i][2] -= taup1 * w * g1[1]
    return np.array(u), np.array(d), np.array(vt)


def snow_return_map_3d(F, Jp):
    """three_d/g2p.py:48-58 on one 3x3 matrix with the singular vectors of svd3()."""
    U, sig, Vh = svd3(F)
    sig = np.clip(sig, 1.0 - 2.5e-2, 1.0 + 7.5e-3)
    old_J = np.linalg.det(np.asarray(F, dtype=np.float64))
    Fn = U @ np.diag(sig) @ Vh.T
    det = np.linalg.det(Fn) + 1e-10
    return Fn, float(np.clip(Jp * old_J / det, 0.6, 20.0))
