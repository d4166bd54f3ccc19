"""State types of the MPM path (reference solvers/mpm/particle.py:10-33)."""
from __future__ import annotations

from typing import Sequence

import numpy as np


class Particle(object):
    """Field-for-field stand-in for the reference's numba jitclass
    (particle.py:10-27): argument order is (pos, force, mass, lambda_, mu)."""

    __slots__ = ("pos", "force", "mass", "lambda_0", "mu_0")

    def __init__(self, pos: np.ndarray, force: float, mass: float, lambda_: float, mu: float):
        self.pos = pos
        self.force = force
        self.mass = mass
        self.lambda_0 = lambda_
        self.mu_0 = mu


class ParticleArray(object):
    """SoA view of a particle list: the fast way to hand particles to the GPU path.
    Behaves like a read-only sequence of :class:`Particle` whose ``pos`` rows alias
    ``self.pos`` (so in-place position updates are visible through both)."""

    def __init__(self, pos: np.ndarray, mass, lambda_0, mu_0, force: float = 0.0):
        self.pos = np.ascontiguousarray(pos, dtype=np.float64)
        n = len(self.pos)
        self.mass = np.ascontiguousarray(np.broadcast_to(np.asarray(mass, dtype=np.float64), (n,)))
        self.lambda_0 = np.ascontiguousarray(np.broadcast_to(np.asarray(lambda_0, dtype=np.float64), (n,)))
        self.mu_0 = np.ascontiguousarray(np.broadcast_to(np.asarray(mu_0, dtype=np.float64), (n,)))
        self.force = force

    def __len__(self):
        return len(self.pos)

    def __getitem__(self, i):
        return Particle(self.pos[i], self.force, float(self.mass[i]), float(self.lambda_0[i]), float(self.mu_0[i]))

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


def particles_to_soa(particles) -> ParticleArray:
    """Accepts a ParticleArray, a numba typed list / Python list of objects with
    ``pos, mass, lambda_0, mu_0`` (the reference's Particle), and returns SoA arrays."""
    if isinstance(particles, ParticleArray):
        return particles
    n = len(particles)
    pos = np.empty((n, 3), dtype=np.float64)
    mass = np.empty(n); lam = np.empty(n); mu = np.empty(n)
    for i, p in enumerate(particles):
        pos[i] = p.pos
        mass[i] = p.mass
        lam[i] = p.lambda_0
        mu[i] = p.mu_0
    return ParticleArray(pos, mass, lam, mu)


def write_back_positions(particles, pos: np.ndarray) -> None:
    """In-place position update, as three_d/g2p.py:45 does with ``particle.pos +=``."""
    if isinstance(particles, ParticleArray):
        particles.pos[:] = pos
        return
    for i, p in enumerate(particles):
        p.pos[:] = pos[i]


def map_particles_to_pos(particles: Sequence, coeff: float) -> np.ndarray:
    """particle.py:30-33: flat f64 vector of ``pos / coeff``."""
    if isinstance(particles, ParticleArray):
        return (particles.pos / coeff).reshape(-1)
    return np.array([p.pos.copy() / coeff for p in particles]).reshape(-1)
