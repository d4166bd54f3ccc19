"""A CPU stand-in for libfemflow_mpm.so, for TESTS of the Python host code above the C ABI.

``FakeLib`` implements the entry points of include/femflow_mpm.h in Python over the same raw pointers (ctypes
structs, data_ptr() of torch tensors -- here CPU tensors), with the NumPy oracle as the arithmetic.  It lets the
real ``MpmSolver`` / ``CudaSlab`` / ``SlabDriver`` code -- buffer binding, material tables, the reordering
ping-pong, id carrying, leaver extraction, migration payloads, slab rebuilds -- run on CPU under gloo, where the
round's single-GPU test box can never run the multi-GPU plumbing.  Test infrastructure only: it lives in tests/,
nothing in femflow_b200/ can import it, and it is not a fallback (the product path raises without CUDA).

3D; slabs are cut along x (cfg.origin[0], cfg.n[0]).  The oracle works on cubic grids, so a (res_x, res_y, res_z)
domain is embedded in a cube of the largest extent for the particle phases, and the grid update -- whose walls sit at
different indices per axis then -- is restated here for a box (identical to the oracle's on a cube: asserted in the test).
"""
import contextlib
import ctypes as C

import numpy as np

from femflow_b200 import _native as N
from oracle import mpm_oracle as O

_CT = {4: C.c_float, 8: C.c_double}


def _view(ptr, count, ctype):
    if not ptr or count == 0:
        return None
    return np.ctypeslib.as_array(C.cast(int(ptr), C.POINTER(ctype)), (int(count),))


def _val(a):
    """Python value of a ctypes scalar / plain int / None."""
    return getattr(a, "value", a)


class _Handle:
    pass


class FakeLib:
    def __init__(self):
        self.handles, self.next_id, self.err = {}, 1, b""

    # -- bookkeeping ------------------------------------------------------------------------------
    def ffmpm_abi_version(self):
        return N.ABI_VERSION

    def ffmpm_last_error(self):
        return self.err

    def _fail(self, code, msg):
        self.err = msg.encode()
        return code

    @staticmethod
    def _layout(cfg, capacity):
        es = 8 if cfg.dtype == N.FFMPM_F64 else 4
        nodes = cfg.n[0] * cfg.n[1] * cfg.n[2]
        grid_off = 256
        bin_off = grid_off + ((nodes * 4 * es + 255) // 256) * 256
        return es, nodes, grid_off, bin_off, bin_off + 3 * 4 * max(int(capacity), 1) + 4096

    def ffmpm_workspace_bytes(self, cfg_ref, capacity):
        cfg = cfg_ref._obj
        if cfg.dim not in (2, 3) or not cfg.dt > 0:
            return self._fail(N.FFMPM_E_INVALID, "bad config")
        return self._layout(cfg, _val(capacity))[4]

    def ffmpm_create(self, cfg_ref, device, out_ref):
        cfg = cfg_ref._obj
        h = _Handle()
        h.cfg = N.FfMpmConfig.from_buffer_copy(cfg)
        h.ws = None
        h.st = [None, None]
        h.live, h.n, h.launches, h.n_oob = 0, 0, 0, 0
        h.table = None
        h.colliders = None
        h.own = (-2 ** 31, 2 ** 31 - 1)
        h.slack = 0
        self.handles[self.next_id] = h
        out_ref._obj.value = self.next_id
        self.next_id += 1
        return N.FFMPM_OK

    def _h(self, h):
        return self.handles[_val(h)]

    def ffmpm_destroy(self, h):
        self.handles.pop(_val(h), None)

    def ffmpm_set_workspace(self, h, ptr, nbytes):
        h = self._h(h)
        h.ws, h.ws_bytes = int(_val(ptr)), int(_val(nbytes))
        h.es, h.nodes, h.grid_off, h.bin_off, _ = self._layout(h.cfg, 0)
        return N.FFMPM_OK

    def ffmpm_bind_state(self, h, cur_ref, alt_ref, n):
        h = self._h(h)
        h.st[0] = N.FfMpmState.from_buffer_copy(cur_ref._obj)
        h.st[1] = N.FfMpmState.from_buffer_copy(alt_ref._obj) if alt_ref is not None else None
        n = int(_val(n))
        if n < 0 or n > h.st[0].stride:
            return self._fail(N.FFMPM_E_INVALID, "n must be in [0, stride]")
        mats = sum(bool(getattr(h.st[0], k)) for k in ("mass", "mu0", "lam0"))
        if mats not in (0, 3):
            return self._fail(N.FFMPM_E_INVALID, "state is missing a required plane")
        h.live, h.n = 0, n
        return N.FFMPM_OK

    def ffmpm_live_buffer(self, h):
        return self._h(h).live

    def ffmpm_num_particles(self, h):
        return self._h(h).n

    def ffmpm_set_num_particles(self, h, n):
        self._h(h).n = int(_val(n))
        return N.FFMPM_OK

    def ffmpm_set_materials(self, h, mass, mu0, lam0, count):
        h = self._h(h)
        count = int(_val(count))
        if count == 0:
            h.table = None
            return N.FFMPM_OK
        dt = np.float64 if h.cfg.dtype == N.FFMPM_F64 else np.float32
        h.table = np.stack([np.ctypeslib.as_array(p, (count,)).astype(dt).astype(np.float64) for p in (mass, mu0, lam0)])
        return N.FFMPM_OK

    def ffmpm_set_colliders(self, h, points, normals, count):
        self._h(h).colliders = None if not _val(count) else (np.ctypeslib.as_array(points, (_val(count) * 3,)).reshape(-1, 3).copy(),
                                                            np.ctypeslib.as_array(normals, (_val(count) * 3,)).reshape(-1, 3).copy())
        return N.FFMPM_OK

    def ffmpm_set_owned_range(self, h, lo, hi):
        self._h(h).own = (int(_val(lo)), int(_val(hi)))
        return N.FFMPM_OK

    def ffmpm_set_owned_slack(self, h, slack):
        self._h(h).slack = int(_val(slack))
        return N.FFMPM_OK

    def ffmpm_leaver_count_ptr(self, h, out_ref):
        out_ref._obj.value = self._h(h).ws + 64
        return N.FFMPM_OK

    def ffmpm_grid_ptr(self, h, out_ref):
        h = self._h(h)
        out_ref._obj.value = h.ws + h.grid_off
        return N.FFMPM_OK

    ffmpm_grid_view = ffmpm_grid_ptr

    def ffmpm_bin_ptrs(self, h, keys, perm, off, n_cells):
        return self._fail(N.FFMPM_E_STATE, "the CPU stand-in keeps no binning")

    def ffmpm_launch_count(self, h):
        return self._h(h).launches

    def ffmpm_poll_error(self, h, stream, code_ref, n_ref):
        h = self._h(h)
        n, h.n_oob = h.n_oob, 0
        code_ref._obj.value, n_ref._obj.value = 0, n
        return self._fail(N.FFMPM_E_OOB, "particle stencil left the grid") if n else N.FFMPM_OK

    # -- state access -----------------------------------------------------------------------------
    def _planes(self, h, which):
        """SoA planes of buffer `which` as NumPy views onto the caller's memory."""
        s, ct, st, d = h.st[which], _CT[h.es], h.st[which].stride, h.cfg.dim
        out = {k: _view(getattr(s, k), rows * st, ct).reshape(rows, st) for k, rows in (("x", d), ("v", d), ("C", d * d), ("F", d * d))}
        for k in ("mass", "mu0", "lam0", "Jp"):
            out[k] = _view(getattr(s, k), st, ct)
        out["id"] = _view(s.id, st, C.c_int32)
        out["material"] = _view(s.material, st, C.c_uint8)
        return out

    def _materials(self, h, p, n):
        if p["mass"] is not None:
            return tuple(p[k][:n].astype(np.float64) for k in ("mass", "mu0", "lam0"))
        if h.table is not None:
            row = p["material"][:n].astype(np.int64) if (p["material"] is not None and h.table.shape[1] > 1) else np.zeros(n, np.int64)
            return tuple(h.table[i][row] for i in range(3))
        c = h.cfg
        return tuple(np.full(n, v) for v in (c.mass, c.mu_0, c.lambda_0))

    def _grid(self, h):
        return _view(h.ws + h.grid_off, h.nodes * 4, _CT[h.es]).reshape(h.cfg.n[0], h.cfg.n[1], h.cfg.n[2], 4)

    def _embed(self, h):
        """Local grid planes inside zeroed cubic oracle arrays (side = the largest global extent + 1)."""
        c = h.cfg
        G = max(c.res[0], c.res[1], c.res[2]) + 1
        g, o = self._grid(h), c.origin[0]
        gv, gm = np.zeros((G, G, G, 3)), np.zeros((G, G, G, 1))
        gv[o:o + c.n[0], :c.n[1], :c.n[2]] = g[..., :3]
        gm[o:o + c.n[0], :c.n[1], :c.n[2]] = g[..., 3:]
        return gv, gm, (slice(o, o + c.n[0]), slice(0, c.n[1]), slice(0, c.n[2]))

    @staticmethod
    def box_grid_op(res, dx, dt, gravity, gv, gm):
        """three_d/grid_op.py:5-47 on a (res_x, res_y, res_z) box: as oracle.grid_op_3d, walls per axis."""
        m = gm[..., 0]
        act = m > 0
        gv[act] /= m[act][:, None]
        gv[act, 1] += dt * gravity
        va = dx * 0.9 / dt
        gv[act] = np.clip(gv[act], -va, va)
        for d in range(3):
            idx = np.arange(gv.shape[d])
            wall = (idx < 1) | (idx >= res[d] - 1)
            sel = [slice(None)] * 3
            sel[d] = wall
            gv[tuple(sel) + (d,)] = 0

    def _inside(self, h, x):
        """Particles whose stencil stays inside the LOCAL grid (the kernels flag and skip the others)."""
        base, _ = O.base_and_fx(x, h.cfg.inv_dx)
        b = base - np.array([h.cfg.origin[0], 0, 0])
        with np.errstate(invalid="ignore"):
            return ~np.isnan(x).any(1) & (b >= 0).all(1) & (b + 2 < np.array(list(h.cfg.n))).all(1), base

    # -- phases -----------------------------------------------------------------------------------
    def ffmpm_clear_grid(self, h, stream):
        self._grid(self._h(h))[...] = 0
        return N.FFMPM_OK

    def ffmpm_bin(self, h, stream):
        return N.FFMPM_OK

    def ffmpm_p2g(self, h, stream):
        h = self._h(h)
        h.launches += 1
        if h.n == 0:
            return N.FFMPM_OK
        if h.cfg.dim == 2:
            return self._p2g_2d(h)
        p, n = self._planes(h, h.live), h.n
        x = p["x"][:, :n].T.astype(np.float64)
        ok, _ = self._inside(h, x)
        h.n_oob += int((~ok).sum())
        mass, mu, lam = (a[ok] for a in self._materials(h, p, n))
        gv, gm, sl = self._embed(h)
        c = h.cfg
        jp = p["Jp"][:n].astype(np.float64)[ok].reshape(-1, 1) if p["Jp"] is not None else np.ones((int(ok.sum()), 1))
        O.p2g_3d(c.inv_dx, c.hardening, c.dx, c.dt, c.volume, gv, gm, x[ok], mass, mu, lam, p["v"][:, :n].T.astype(np.float64)[ok],
                 p["F"][:, :n].T.reshape(n, 3, 3).astype(np.float64)[ok], p["C"][:, :n].T.reshape(n, 3, 3).astype(np.float64)[ok],
                 jp, "snow" if c.model == N.FFMPM_SNOW else "neo_hookean")
        g = self._grid(h)
        g[..., :3], g[..., 3:] = gv[sl], gm[sl]
        return N.FFMPM_OK

    # -- 2D (two_d/{p2g,grid_op,g2p}.py): grid nodes are {mom_x, mom_y, mass, -}, no slabs, no reordering ---------
    def _state_2d(self, h):
        p, n = self._planes(h, h.live), h.n
        return (p, n, p["x"][:, :n].T.astype(np.float64), p["v"][:, :n].T.astype(np.float64),
                p["F"][:, :n].T.reshape(n, 2, 2).astype(np.float64), p["C"][:, :n].T.reshape(n, 2, 2).astype(np.float64),
                p["Jp"][:n].astype(np.float64).reshape(n, 1))

    def _p2g_2d(self, h):
        c, g = h.cfg, self._grid(h)
        p, n, x, v, F, Cm, Jp = self._state_2d(h)
        base, _ = O.base_and_fx(x, c.inv_dx)
        ok = (base >= 0).all(1) & (base + 2 < np.array([c.n[0], c.n[1]])).all(1)
        h.n_oob += int((~ok).sum())
        gv, gm = g[:, :, 0, :2].astype(np.float64), g[:, :, 0, 2:3].astype(np.float64)
        O.p2g_2d(c.inv_dx, c.hardening, c.mu_0, c.lambda_0, c.mass, c.dx, c.dt, c.volume, gv, gm, x[ok], v[ok], F[ok], Cm[ok], Jp[ok],
                 "snow" if c.model == N.FFMPM_SNOW else "neo_hookean")
        g[:, :, 0, :2], g[:, :, 0, 2:3] = gv, gm
        return N.FFMPM_OK

    def _grid_op_2d(self, h):
        c, g = h.cfg, self._grid(h)
        gv, gm = g[:, :, 0, :2].astype(np.float64), g[:, :, 0, 2:3].astype(np.float64)
        O.grid_op_2d(c.res[0], c.dt, c.gravity, gv, gm)
        g[:, :, 0, :2] = gv
        return N.FFMPM_OK

    def _g2p_2d(self, h):
        c, g = h.cfg, self._grid(h)
        p, n, x, v, F, Cm, Jp = self._state_2d(h)
        O.g2p_2d(c.inv_dx, c.dt, g[:, :, 0, :2].astype(np.float64), x, v, F, Cm, Jp, "snow" if c.model == N.FFMPM_SNOW else "neo_hookean")
        p["x"][:, :n], p["v"][:, :n] = x.T, v.T
        p["F"][:, :n], p["C"][:, :n] = F.reshape(n, 4).T, Cm.reshape(n, 4).T
        p["Jp"][:n] = Jp[:, 0]
        return N.FFMPM_OK

    def ffmpm_grid_op_halo(self, h, lo, planes_lo, hi, planes_hi, stream):
        h = self._h(h)
        h.launches += 1
        if h.cfg.dim == 2:
            return self._grid_op_2d(h)
        g = self._grid(h)
        plane = h.cfg.n[1] * h.cfg.n[2] * 4
        planes_lo, planes_hi = int(_val(planes_lo) or 0), int(_val(planes_hi) or 0)
        if planes_lo:
            g[:planes_lo] += _view(_val(lo), planes_lo * plane, _CT[h.es]).reshape(planes_lo, *g.shape[1:])
        if planes_hi:
            g[g.shape[0] - planes_hi:] += _view(_val(hi), planes_hi * plane, _CT[h.es]).reshape(planes_hi, *g.shape[1:])
        gv, gm, sl = self._embed(h)
        c = h.cfg
        self.box_grid_op(list(c.res), c.dx, c.dt, c.gravity, gv, gm)
        if h.colliders is not None:
            O.check_collision_points(h.colliders[0], h.colliders[1], gv.shape[0] - 1, c.dx, gv)
        g[..., :3] = gv[sl]
        return N.FFMPM_OK

    def ffmpm_grid_op(self, h, stream):
        return self.ffmpm_grid_op_halo(h, None, 0, None, 0, stream)

    def ffmpm_collide(self, h, stream):
        return N.FFMPM_OK

    def ffmpm_g2p(self, h, stream):
        h = self._h(h)
        h.launches += 1
        if h.n == 0:
            return N.FFMPM_OK
        if h.cfg.dim == 2:
            return self._g2p_2d(h)
        snow = h.cfg.model == N.FFMPM_SNOW
        p, n = self._planes(h, h.live), h.n
        x = p["x"][:, :n].T.astype(np.float64)
        v = p["v"][:, :n].T.astype(np.float64)
        F = p["F"][:, :n].T.reshape(n, 3, 3).astype(np.float64)
        Cm = p["C"][:, :n].T.reshape(n, 3, 3).astype(np.float64)
        ok, base = self._inside(h, x)
        h.n_oob += int((~ok).sum())
        gv, _, _ = self._embed(h)
        xo, vo, Fo, Co = x[ok], v[ok], F[ok], Cm[ok]
        jpo = p["Jp"][:n][ok].astype(np.float64)[:, None] if snow else np.ones((len(xo), 1))
        O.g2p_3d(h.cfg.inv_dx, h.cfg.dt, gv, xo, vo, Fo, Co, jpo, "snow" if snow else "neo_hookean")
        x[ok], v[ok], F[ok], Cm[ok] = xo, vo, Fo, Co
        if snow:     # three_d/g2p.py:58, before the planes are carried into the other buffer
            p["Jp"][:n][ok] = jpo[:, 0]
        if h.st[1] is not None:
            # the reordering G2P: cell-sorted into the other buffer, out-of-grid particles last, planes carried along
            G = gv.shape[0]
            key = np.where(ok, O.cell_keys(np.where(ok[:, None], base, 0), G), np.iinfo(np.int64).max)
            order = np.argsort(key, kind="stable")
            q = self._planes(h, h.live ^ 1)
            for k in ("mass", "mu0", "lam0", "Jp", "id", "material"):
                if p[k] is not None:
                    q[k][:n] = p[k][:n][order]
            h.live ^= 1
        else:
            order, q = np.arange(n), p
        q["x"][:, :n] = x[order].T
        q["v"][:, :n] = v[order].T
        q["F"][:, :n] = F[order].reshape(n, 9).T
        q["C"][:, :n] = Cm[order].reshape(n, 9).T
        nb, _ = O.base_and_fx(x[ok], h.cfg.inv_dx)
        leavers = int(((nb[:, 0] < h.own[0]) | (nb[:, 0] >= h.own[1])).sum())
        urgent = int(((h.own[0] - nb[:, 0] > h.slack) | (nb[:, 0] - h.own[1] >= h.slack)).sum())
        _view(h.ws + 64, 2, C.c_int32)[:] = (leavers, urgent)
        return N.FFMPM_OK

    def ffmpm_scatter(self, h, stream):
        self.ffmpm_clear_grid(h, stream)
        return self.ffmpm_p2g(h, stream)

    def ffmpm_gather(self, h, stream):
        return self.ffmpm_g2p(h, stream)

    def ffmpm_substep(self, h, n, stream):
        for _ in range(int(_val(n))):
            self.ffmpm_scatter(h, stream)
            self.ffmpm_grid_op(h, stream)
            self.ffmpm_gather(h, stream)
        return N.FFMPM_OK

    # -- slab migration (csrc/mpm_migrate.cuh): same message layout, same back-fill rule -----------
    def ffmpm_migrate_rows(self, h):
        h = self._h(h)
        return 28 + (1 if h.st[0].Jp else 0)

    def _mig_row(self, h, p, r):
        """(array, row index or None) behind payload row r of buffer planes p."""
        if r < 3: return p["x"], r
        if r < 6: return p["v"], r - 3
        if r < 15: return p["C"], r - 6
        if r < 24: return p["F"], r - 15
        if r == 24: return (p["mass"] if p["mass"] is not None else p["material"]), None
        if r == 25: return p["mu0"], None
        if r == 26: return p["lam0"], None
        if r == 27: return p["id"], None
        return p["Jp"], None

    def _mig_get(self, h, p, r, idx):
        a, row = self._mig_row(h, p, r)
        dt = np.float64 if h.es == 8 else np.float32
        if a is None:
            return np.zeros(len(idx), dt)
        v = a[row][idx] if row is not None else a[idx]
        if r == 27:
            return v.astype(np.int32).view(np.float32) if h.es == 4 else v.astype(np.float64)
        return v.astype(dt)

    def _mig_put(self, h, p, r, idx, vals):
        a, row = self._mig_row(h, p, r)
        if a is None:
            return
        if r == 27:
            vals = np.ascontiguousarray(vals, np.float32).view(np.int32) if h.es == 4 else vals.astype(np.int32)
        if row is not None:
            a[row][idx] = vals
        else:
            a[idx] = vals.astype(a.dtype)

    def ffmpm_migrate_pack(self, h, out_lo, out_hi, cap, stream):
        h = self._h(h)
        cap, rows = int(_val(cap)), 28 + (1 if h.st[0].Jp else 0)
        if h.st[h.live].id in (None, 0):
            return self._fail(N.FFMPM_E_STATE, "slab migration needs the id plane")
        p, n = self._planes(h, h.live), h.n
        x0 = p["x"][0, :n].astype(np.float64)
        gbx = (x0 * h.cfg.inv_dx - 0.5).astype(np.int64)
        ct = _CT[h.es]
        sides = []
        for ptr, sel in ((_val(out_lo), gbx < h.own[0]), (_val(out_hi), gbx >= h.own[1])):
            idx = np.nonzero(sel)[0] if ptr else np.zeros(0, np.int64)
            over = max(0, len(idx) - cap)
            idx = idx[:cap]
            if ptr:
                box = _view(ptr, (rows + 1) * cap, ct).reshape(rows + 1, cap)
                box[0, 0] = len(idx)
                for r in range(rows):
                    box[1 + r, :len(idx)] = self._mig_get(h, p, r, idx)
            sides.append((idx, over))
        gone = np.concatenate([sides[0][0], sides[1][0]])
        keep = n - len(gone)
        hole = np.zeros(n, bool)
        hole[gone] = True
        lo_holes = np.nonzero(hole[:keep])[0]
        movers = keep + np.nonzero(~hole[keep:n])[0]
        assert len(lo_holes) == len(movers)
        for r in range(rows):
            self._mig_put(h, p, r, lo_holes, self._mig_get(h, p, r, movers))
        h.mig = dict(out_lo=len(sides[0][0]), out_hi=len(sides[1][0]), keep=keep, overflow=sides[0][1] + sides[1][1])
        h.launches += 4
        return N.FFMPM_OK

    def ffmpm_migrate_unpack(self, h, in_lo, in_hi, cap, record, stream):
        h = self._h(h)
        cap, rows = int(_val(cap)), 28 + (1 if h.st[0].Jp else 0)
        p = self._planes(h, h.live)
        ct = _CT[h.es]
        at, counts = h.mig["keep"], []
        for ptr in (_val(in_lo), _val(in_hi)):
            k = 0
            if ptr:
                box = _view(ptr, (rows + 1) * cap, ct).reshape(rows + 1, cap)
                k = min(int(box[0, 0]), cap)
                if at + k > h.st[h.live].stride:
                    h.mig["overflow"] += (1 << 24)
                    k = 0
                idx = np.arange(at, at + k)
                for r in range(rows):
                    self._mig_put(h, p, r, idx, box[1 + r, :k])
                at += k
            counts.append(k)
        rec = _view(_val(record), 6, C.c_int32)
        rec[:] = [h.mig["out_lo"], h.mig["out_hi"], counts[0], counts[1], at, h.mig["overflow"]]
        h.launches += 1
        return N.FFMPM_OK

    def ffmpm_snapshot(self, h, coeff, out, stream):
        h = self._h(h)
        p, n = self._planes(h, h.live), h.n
        dst = _view(_val(out), 3 * n, C.c_double).reshape(n, 3)
        ids = p["id"][:n] if p["id"] is not None else np.arange(n)
        dst[ids] = p["x"][:, :n].T.astype(np.float64) / _val(coeff)
        return N.FFMPM_OK

    def ffmpm_export_state(self, h, dst_ref, stream):
        h = self._h(h)
        p, n, d = self._planes(h, h.live), h.n, h.cfg.dim
        saved = h.st[1]
        h.st[1] = N.FfMpmState.from_buffer_copy(dst_ref._obj)
        try:
            q = self._planes(h, 1)
        finally:
            h.st[1] = saved
        ids = p["id"][:n] if p["id"] is not None else np.arange(n)
        for k in ("x", "v", "C", "F"):
            q[k][:, ids] = p[k][:, :n]
        if p["Jp"] is not None and q["Jp"] is not None:
            q["Jp"][ids] = p["Jp"][:n]
        h.launches += 1
        return N.FFMPM_OK

    def ffmpm_debug_red_add4(self, dst, v, count, stream):
        return N.FFMPM_OK


class _Stream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def synchronize(self):
        pass

    def wait_event(self, ev):
        pass


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass

    def synchronize(self):
        pass


def install(monkeypatch=None):
    """Route femflow_b200 through FakeLib and make torch.cuda's stream / event / pinning calls no-ops.
    With a pytest ``monkeypatch`` everything is undone at the end of the test; without one (spawned worker
    processes) the patches last for the life of the process."""
    import torch

    def put(obj, name, value):
        if monkeypatch is not None:
            monkeypatch.setattr(obj, name, value, raising=False)
        else:
            setattr(obj, name, value)
    fake = FakeLib()
    put(N, "_lib", fake)
    put(torch.cuda, "is_available", lambda: True)
    put(torch.cuda, "current_device", lambda: 0)
    put(torch.cuda, "set_device", lambda *a: None)
    put(torch.cuda, "synchronize", lambda *a, **k: None)
    put(torch.cuda, "current_stream", lambda *a, **k: _Stream())
    put(torch.cuda, "Stream", _Stream)
    put(torch.cuda, "Event", _Event)
    put(torch.cuda, "device", lambda *a, **k: contextlib.nullcontext())
    put(torch.cuda, "stream", lambda *a, **k: contextlib.nullcontext())
    put(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    return fake
