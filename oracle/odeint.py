"""Restatement of ``torchdiffeq.odeint`` (TEST INFRASTRUCTURE, not product code).

PARITY UNPINNED.  ``torchdiffeq`` is a third-party dependency of the reference
(``import torchdiffeq`` at models/flow_model.py:11, call site models/flow_model.py:315-324).
The reference pins no version (no requirements/lock file), the package is not vendored
under /root/reference and is not installable in this image (no network, not in the
wheelhouse).  This file restates the published algorithm of torchdiffeq 0.2.x
(``FixedGridODESolver`` for euler / midpoint / rk4, ``RKAdaptiveStepsizeODESolver`` with the
Dormand-Prince tableau for dopri5) from its documented behaviour; there is no golden vector
from the library itself to check it against.  What IS pinned: the velocity function it
drives (see sr_oracle.py) and the call convention of the reference call site.

Semantics restated:
  * ``solution[0] = y0``; output has shape ``(len(t), *y0.shape)``.
  * fixed grid: the grid is ``t`` itself; one step per interval; outputs on grid points are
    the step results exactly (linear interpolation short-circuits at ``t == t1``).
  * dopri5: times are float64 internally, cast to ``y0.dtype`` before reaching ``func``;
    rms error norm over the WHOLE state tensor; FSAL; 4th-order dense output.
"""
from __future__ import annotations

from typing import Callable, List

import torch

Tensor = torch.Tensor


# ---------------------------------------------------------------- fixed grid
def _euler(func, t0, dt, t1, y0):
    return dt * func(t0, y0)


def _midpoint(func, t0, dt, t1, y0):
    half = 0.5 * dt
    k1 = func(t0, y0)
    return dt * func(t0 + half, y0 + k1 * half)


def _rk4_38(func, t0, dt, t1, y0):
    third = 1.0 / 3.0
    k1 = func(t0, y0)
    k2 = func(t0 + dt * third, y0 + dt * k1 * third)
    k3 = func(t0 + dt * (2.0 / 3.0), y0 + dt * (k2 - k1 * third))
    k4 = func(t1, y0 + dt * (k1 - k2 + k3))
    return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125


_FIXED = {"euler": _euler, "midpoint": _midpoint, "rk4": _rk4_38}


def _odeint_fixed(func, y0: Tensor, t: Tensor, step) -> Tensor:
    sol = torch.empty(len(t), *y0.shape, dtype=y0.dtype)
    sol[0] = y0
    y = y0
    for j in range(1, len(t)):
        t0, t1 = t[j - 1], t[j]
        y = y + step(func, t0, t1 - t0, t1, y)
        sol[j] = y
    return sol


# ---------------------------------------------------------------- dopri5
_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
_C_ERR = [
    35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0,
]
_C_MID = [
    6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
    187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2,
]


def _rms(x: Tensor) -> Tensor:
    return x.pow(2).mean().sqrt()


def _initial_step(func, t0, y0, order, rtol, atol, f0):
    scale = atol + y0.abs() * rtol
    d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = torch.tensor(1e-6, dtype=y0.dtype)
    else:
        h0 = 0.01 * d0 / d1
    h0 = h0.abs()
    f1 = func((t0 + h0).to(y0.dtype), y0 + h0 * f0)
    d2 = (_rms((f1 - f0) / scale) / h0).abs()
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6, dtype=y0.dtype), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / float(order + 1))
    return torch.min(100 * h0, h1.abs()).to(torch.float64)


def _combine(y0: Tensor, ks: List[Tensor], coeffs, dt) -> Tensor:
    acc = torch.zeros_like(y0)
    for k, c in zip(ks, coeffs):
        if c != 0:
            acc = acc + k * (c * dt)
    return y0 + acc


class _Dopri5:
    def __init__(self, func, y0, t, rtol, atol, max_num_steps=2 ** 31 - 1):
        self.func, self.rtol, self.atol = func, rtol, atol
        self.t = t.to(torch.float64)
        self.max_num_steps = max_num_steps
        self.nfe = 0
        self.n_accept = self.n_reject = 0
        f0 = self._f(self.t[0], y0)
        dt = _initial_step(self._f, self.t[0], y0, 4, rtol, atol, f0)
        self.y0, self.f0, self.t0, self.t1, self.dt = y0, f0, self.t[0], self.t[0], dt
        self.coeff = [y0] * 5

    def _f(self, t, y):
        self.nfe += 1
        return self.func(t.to(y.dtype), y)

    def _step(self):
        y0, f0, t0, dt = self.y0, self.f0, self.t1, self.dt
        t1 = t0 + dt
        dty = dt.to(y0.dtype)
        t0y, t1y = t0.to(y0.dtype), t1.to(y0.dtype)
        ks = [f0]
        yi = y0
        for a, b in zip(_ALPHA, _BETA):
            ti = t1y if a == 1.0 else t0y + a * dty
            yi = _combine(y0, ks, b, dty)
            ks.append(self._f(ti, yi))
        y1, f1 = yi, ks[-1]                      # FSAL: c_sol == beta[-1], last coeff 0
        err = _combine(torch.zeros_like(y0), ks, _C_ERR, dty)
        tol = self.atol + self.rtol * torch.max(y0.abs(), y1.abs())
        ratio = _rms(err / tol).abs()
        if ratio <= 1:
            y_mid = _combine(y0, ks, _C_MID, dty)
            a_ = 2 * dty * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
            b_ = dty * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
            c_ = dty * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
            self.coeff = [y0, dty * f0, c_, b_, a_]
            self.y0, self.f0, self.t0, self.t1 = y1, f1, t0, t1
            self.n_accept += 1
        else:
            self.t0 = t0
            self.n_reject += 1
        # step-size controller: safety 0.9, ifactor 10, dfactor 0.2, order 5
        if ratio == 0:
            self.dt = dt * 10.0
        else:
            dfac = 1.0 if ratio < 1 else 0.2
            r = ratio.to(torch.float64)
            self.dt = dt * min(10.0, max(0.9 / float(r) ** 0.2, dfac))

    def advance(self, t_next):
        n = 0
        while t_next > self.t1:
            assert n < self.max_num_steps
            self._step()
            n += 1
        x = ((t_next - self.t0) / (self.t1 - self.t0)).to(self.coeff[0].dtype)
        total = self.coeff[0] + x * self.coeff[1]
        xp = x
        for c in self.coeff[2:]:
            xp = xp * x
            total = total + xp * c
        return total


def odeint(func: Callable, y0: Tensor, t: Tensor, method: str = "dopri5",
           atol: float = 1e-4, rtol: float = 1e-4, stats: dict | None = None) -> Tensor:
    """``torchdiffeq.odeint(func, y0, t, method=, atol=, rtol=)`` as called at
    models/flow_model.py:315-324."""
    if method in _FIXED:
        return _odeint_fixed(func, y0, t, _FIXED[method])
    if method != "dopri5":
        raise ValueError(f"unsupported method {method!r}")
    s = _Dopri5(func, y0, t, rtol, atol)
    sol = torch.empty(len(t), *y0.shape, dtype=y0.dtype)
    sol[0] = y0
    for j in range(1, len(t)):
        sol[j] = s.advance(s.t[j])
    if stats is not None:
        stats.update(nfe=s.nfe, accepted=s.n_accept, rejected=s.n_reject)
    return sol
