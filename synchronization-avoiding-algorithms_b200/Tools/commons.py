"""Scalars and holders of the time-step path — same names as /root/reference/Tools/commons.py.

Index helpers return int64 numpy arrays where the reference returns Python lists (they are used only as
fancy indices / membership sets by the drivers); everything else has the reference's types.
"""
import numpy as np

from saa_b200 import maps as _maps
from saa_b200 import mesh as _mesh

__all__ = ["linear_ramp", "elasticity", "Time_integration_displacement", "node_to_dof", "Meshsize", "lumping",
           "lumping_to_vec", "dirac", "basis", "np"]


def linear_ramp(t):
    """Load ramp reaching 1 at t = 1 s (commons.py:7-11)."""
    return t if t <= 1 else 1.0


class elasticity:
    """Material + load parameters (commons.py:15-41): Lame constants, density, body force fz, ramp flag R."""

    def __init__(self, lmd, mu, rho, fz, R):
        self.lmd, self.mu, self.rho, self.fz, self.R = lmd, mu, rho, fz, R

    def D(self):
        lmd, mu = self.lmd, self.mu
        d = np.zeros((6, 6))
        d[:3, :3] = lmd
        d[np.arange(3), np.arange(3)] = lmd + 2.0 * mu
        d[np.arange(3, 6), np.arange(3, 6)] = mu
        return d

    def f(self, X, t):
        s = linear_ramp(t) if self.R else None
        fz = -self.fz if s is None else -self.fz * s
        return np.array([[0.0], [fz], [fz]])


class Time_integration_displacement:
    """(tn, dt, d0 = d_n, dn = d_{n-1}) handed to the step (commons.py:47-55)."""

    def __init__(self, tn, dt, d0, dn):
        self.tn, self.dt, self.d0, self.dn = tn, dt, d0, dn

    def tn_plus_1(self):
        return self.tn + self.dt


def node_to_dof(d, ls, P):
    """Interleaved numbering d*g + i, g-major (commons.py:66-71) as an int64 array."""
    return _maps.node_to_dof(d, ls, P)


def Meshsize(Element, Points):
    """2*min_edge/sqrt(24) over the given tets (commons.py:79-90), vectorised, same value."""
    return _mesh.meshsize(np.asarray(Element), np.asarray(Points))


class LumpedCarrier(np.ndarray):
    """Dense-looking handle returned by Global_Assembly_no_bc for the mass matrix when the mesh is too large
    for a dense (3N)^2 array: it carries the row sums so that lumping_to_vec() needs no dense matrix."""
    row_sums = None


def lumping_to_vec(M):
    """Row-sum lumping to a (n,1) vector (commons.py:103-107)."""
    rs = getattr(M, "row_sums", None)
    if rs is not None:
        return rs.copy()
    M = np.asarray(M)
    out = np.zeros((len(M), 1))
    for i in range(len(M)):
        out[i] = np.sum(M[i, :])
    return out


def lumping(M):
    """Row-sum lumping to a diagonal matrix (commons.py:95-99)."""
    return np.diag(lumping_to_vec(M)[:, 0])


def dirac(i, j):
    return 1.0 if i == j else 0.0


def basis(i):
    if i in (0, 1, 2):
        e = np.zeros(3, dtype=int)
        e[i] = 1
        return e
    print('Not a basis!')
