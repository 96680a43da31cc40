"""Small host-side helpers of the drivers (timers, chemical-potential search, DIIS)."""
import functools
import time

import numpy as np


class Timer:
    def __init__(self):
        self.t_init = self.t_prev = time.perf_counter()

    def lap(self):
        now = time.perf_counter()
        dt, self.t_prev = now - self.t_prev, now
        return dt

    def total(self):
        return time.perf_counter() - self.t_init


def record_time(key):
    """Accumulates wall time of a method into ``self._timings[key]`` (auxgf/util/decor.py:5-13)."""
    def deco(fn):
        @functools.wraps(fn)
        def wrapped(self, *args, **kwargs):
            t0 = time.perf_counter()
            try:
                return fn(self, *args, **kwargs)
            finally:
                self._timings[key] = self._timings.get(key, 0.0) + time.perf_counter() - t0
        return wrapped
    return deco


def record_energy(key):
    """Appends the returned energy to ``self._energies[key]`` (auxgf/util/decor.py:16-24)."""
    def deco(fn):
        @functools.wraps(fn)
        def wrapped(self, *args, **kwargs):
            val = fn(self, *args, **kwargs)
            self._energies.setdefault(key, []).append(val)
            return val
        return wrapped
    return deco


def find_chempot(nphys, nelec, h, occupancy=2.0):
    """Aufbau chemical potential of an (extended) Hamiltonian given as (w, v) or a matrix
    (auxgf/util/chempot.py:7-59): fill quasi-particle states in energy order, each carrying
    ``occupancy * |v[:nphys, i]|^2`` physical electrons, stop where the running sum brackets ``nelec``."""
    if isinstance(h, tuple):
        w, v = h
    else:
        w, v = np.linalg.eigh(h)
    occ = occupancy * np.einsum("xi,xi->i", v[:nphys], v[:nphys])
    csum = np.cumsum(occ)
    n = csum.size
    i = n - 1
    for k in range(1, n):
        if csum[k - 1] <= nelec <= csum[k]:
            i = k
            break
    prv = csum[i - 1] if i > 0 else 0.0
    cur = csum[i]
    if abs(prv - nelec) < abs(cur - nelec):
        homo, error = i - 1, nelec - prv
    else:
        homo, error = i, nelec - cur
    lumo = min(homo + 1, n - 1)
    return 0.5 * (w[lumo] + w[homo]), error


class DIIS:
    """Pulay DIIS on successive Fock matrices (error = difference of consecutive iterates), the role
    ``pyscf.lib.diis.DIIS`` plays in auxgf/util/diis.py:8-39.  An accelerator only: the fixed point of
    the Fock loop does not depend on it."""

    def __init__(self, space=8, min_space=1):
        self.space = space
        self.min_space = min_space
        self._x = []
        self._e = []
        self._prev = None

    def update(self, x):
        x = np.asarray(x, dtype=np.float64)
        if self._prev is not None:
            self._x.append(x.copy())
            self._e.append((x - self._prev).ravel())
            if len(self._x) > self.space:
                self._x.pop(0)
                self._e.pop(0)
        self._prev = x.copy()
        n = len(self._x)
        if n <= self.min_space:
            return x
        b = np.zeros((n + 1, n + 1))
        for i in range(n):
            for j in range(i + 1):
                b[i, j] = b[j, i] = self._e[i] @ self._e[j]
        b[n, :n] = b[:n, n] = 1.0
        rhs = np.zeros(n + 1)
        rhs[n] = 1.0
        try:
            c = np.linalg.solve(b, rhs)[:n]
        except np.linalg.LinAlgError:
            return x
        out = sum(ci * xi for ci, xi in zip(c, self._x))
        self._prev = out.copy()
        return out
