"""Tsit5 (Tsitouras 2011) with constant step, as ``diffrax.Tsit5`` + ``ConstantStepSize``.

Test infrastructure only.  ref: somax/_src/core/model.py:39-88 (BC before every RHS,
BC on state0, SaveAt(t1=True)); tableau SURVEY App. A.  Stage increments are
``k_i = dt * f(BC(Y_i))`` and ``Y_i = y_n + sum_j a_ij k_j`` (diffrax convention);
``y_{n+1} = Y_7`` and ``k_7`` is reused as the next step's ``k_1`` (FSAL).
"""
from __future__ import annotations

import math

A = [
    [],
    [0.161],
    [-0.008480655492356989, 0.335480655492357],
    [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
    [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
    [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401,
     -0.028269050394068383],
    [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742,
     -3.290069515436081, 2.324710524099774],
]
C = [0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0]


def step_plan(t0: float, t1: float, dt: float):
    """(n_full, dt_last): n_full steps of dt then, if dt does not divide t1-t0, one clipped
    step of dt_last (0.0 if none).  ref: SURVEY App. A (ConstantStepSize + clip to t1)."""
    span = t1 - t0
    n = int(math.floor(span / dt * (1.0 + 1e-12) + 1e-9))
    rem = span - n * dt
    if rem <= 1e-9 * max(abs(dt), 1e-300):
        rem = 0.0
    return n, rem


def tree_axpy(y, coeffs, ks):
    """y + sum_j coeffs[j]*ks[j] over a tuple-of-arrays state."""
    out = []
    for f, yf in enumerate(y):
        acc = yf.copy()
        for cj, kj in zip(coeffs, ks):
            acc = acc + cj * kj[f]
        out.append(acc)
    return tuple(out)


def integrate(rhs, bc, y0, t0, t1, dt, on_step=None):
    """Integrate ``dy/dt = rhs(bc(y))`` from t0 to t1.  ``y0`` is a tuple of arrays.
    Returns the final state (NOT re-projected by bc: the ghost ring drifts exactly as in
    the reference)."""
    y = bc(tuple(a.copy() for a in y0))
    n, rem = step_plan(t0, t1, dt)
    hs = [dt] * n + ([rem] if rem > 0.0 else [])
    f1 = None  # FSAL derivative f(Y_7) of the previous step (not yet scaled by dt)
    for istep, h in enumerate(hs):
        if f1 is None:
            f1 = rhs(bc(y))
        ks = [tuple(h * a for a in f1)]
        fs_last = None
        for s in range(1, 7):
            Y = tree_axpy(y, A[s], ks)
            fs_last = rhs(bc(Y))
            if s < 6:
                ks.append(tuple(h * a for a in fs_last))
        y = Y
        f1 = fs_last
        if on_step is not None:
            on_step(istep + 1, y)
    return y
