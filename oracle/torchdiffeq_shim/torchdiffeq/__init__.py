"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Stand-in for the third-party dependency `torchdiffeq` (unpinned in the reference:
/root/reference/.travis.yml:9 `pip install torchdiffeq`, i.e. the 0.2.x line current
when the reference CI ran).  The package is absent from this image and there is no
network, so this module restates the two published algorithms of `torchdiffeq.odeint`
that the reference reaches:

  * the default adaptive solver `dopri5` (Dormand-Prince 5(4), Shampine dense output,
    Hairer initial step, safety 0.9 / ifactor 10 / dfactor 0.2 step controller, RMS
    norm, "mixed" max-over-components norm for tuple states) -- used by
    /root/reference/src/NeuralODE/nnModule.py:69 with rtol/atol;
  * the fixed-grid solver `method="rk4"` (3/8-rule RK4, `rk4_alt_step_func`) with
    `options={"step_size": h}` -- the fixed-step variant the B200 path implements.

It is NOT the original source.  Agreement with the real package cannot be checked
here ("parity unpinned" for the adaptive step controller details); what the tests
rely on is only that both solvers converge to the exact ODE solution, which is
verified against closed-form solutions (tests/test_oracle_odeint.py, mirroring
/root/reference/tests/test_odeint.py) and against scipy.integrate.solve_ivp.

Only forward numerics are needed: the reference wraps odeint in its own
torch.autograd.Function (nnModule.py:8) and never back-propagates through it.
"""
import math
import torch

__version__ = "0.2-shim"

_A = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_B = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_C_SOL = [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0]
_C_ERR = [
    35 / 384 - 1951 / 21600, 0.0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
    -2187 / 6784 + 12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0,
]
_C_MID = [
    6025192743 / 30085553152 / 2, 0.0, 51252292925 / 65400821598 / 2,
    -2691868925 / 45128329728 / 2, 187940372067 / 1594534317056 / 2,
    -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2,
]


class _Packed:
    """Tuple-of-tensors state flattened to one vector (what odeint does for tuples)."""

    def __init__(self, func, shapes):
        self.func, self.shapes = func, shapes
        self.numels = [int(torch.Size(s).numel()) for s in shapes]

    def unpack(self, y):
        out, o = [], 0
        for s, m in zip(self.shapes, self.numels):
            out.append(y[o:o + m].view(s))
            o += m
        return tuple(out)

    def __call__(self, t, y):
        f = self.func(t, self.unpack(y))
        return torch.cat([fi.reshape(-1) for fi in f])

    def norm(self, y):
        return max(float(c.pow(2).mean().sqrt()) if c.numel() else 0.0 for c in self.unpack(y))


def _rms(y):
    return float(y.pow(2).mean().sqrt())


def _initial_step(func, t0, y0, order, rtol, atol, norm, f0):
    scale = atol + y0.abs() * rtol
    d0, d1 = norm(y0 / scale), norm(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = func(t0 + h0, y0 + h0 * f0)
    d2 = norm((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / (order + 1))
    return min(100 * h0, h1)


def _dopri5(func, y0, t, rtol, atol, norm, max_num_steps=2 ** 31 - 1):
    order, safety, ifactor, dfactor = 5, 0.9, 10.0, 0.2
    t0 = float(t[0])
    f0 = func(t0, y0)
    dt = _initial_step(func, t0, y0, order - 1, rtol, atol, norm, f0)
    # state: y at t_hi, interpolant valid on [t_lo, t_hi]
    y_cur, f_cur, t_lo, t_hi, coeff = y0, f0, t0, t0, None
    out = [y0]
    for tj in t[1:]:
        tj = float(tj)
        nsteps = 0
        while tj > t_hi:
            assert nsteps < max_num_steps
            nsteps += 1
            k = [f_cur]
            for ai, bi in zip(_A, _B):
                yi = y_cur
                for bij, kj in zip(bi, k):
                    if bij != 0.0:
                        yi = yi + (dt * bij) * kj
                k.append(func(t_hi + ai * dt, yi))
            y1 = y_cur
            for c, kj in zip(_C_SOL, k):
                if c != 0.0:
                    y1 = y1 + (dt * c) * kj
            err = sum((dt * c) * kj for c, kj in zip(_C_ERR, k) if c != 0.0)
            tol = atol + rtol * torch.max(y_cur.abs(), y1.abs())
            ratio = norm(err / tol)
            if ratio <= 1:
                y_mid = y_cur + sum((dt * c) * kj for c, kj in zip(_C_MID, k) if c != 0.0)
                fa, fb = k[0], k[-1]
                coeff = (
                    2 * dt * (fb - fa) - 8 * (y1 + y_cur) + 16 * y_mid,
                    dt * (5 * fa - 3 * fb) + 18 * y_cur + 14 * y1 - 32 * y_mid,
                    dt * (fb - 4 * fa) - 11 * y_cur - 5 * y1 + 16 * y_mid,
                    dt * fa,
                    y_cur,
                )
                t_lo, t_hi, y_cur, f_cur = t_hi, t_hi + dt, y1, k[-1]
            if ratio == 0:
                dt = dt * ifactor
            else:
                lo = 1.0 if ratio < 1 else dfactor
                dt = dt * min(ifactor, max(safety / ratio ** (1.0 / order), lo))
        x = (tj - t_lo) / (t_hi - t_lo)
        a, b, c, d, e = coeff
        out.append(e + x * (d + x * (c + x * (b + x * a))))
    return torch.stack(out)


def _rk4_fixed(func, y0, t, step_size):
    out = [y0]
    y = y0
    for ta, tb in zip(t[:-1], t[1:]):
        ta, tb = float(ta), float(tb)
        if step_size is None:
            grid = [ta, tb]
        else:
            niters = int(math.ceil((tb - ta) / step_size + 1))
            grid = [ta + i * step_size for i in range(niters)]
            grid[-1] = tb
        for g0, g1 in zip(grid[:-1], grid[1:]):
            dt = g1 - g0
            k1 = func(g0, y)
            k2 = func(g0 + dt / 3, y + dt * k1 / 3)
            k3 = func(g0 + dt * 2 / 3, y + dt * (k2 - k1 / 3))
            k4 = func(g1, y + dt * (k1 - k2 + k3))
            y = y + (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
        out.append(y)
    return torch.stack(out)


def odeint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None):
    """Integrate dy/dt = func(t, y) from t[0] through t[1:]; returns y at every t.

    y0 may be a tensor or a tuple of tensors (a tuple of stacked solutions is
    returned then).  Decreasing t is handled as in the original by time reversal.
    """
    options = options or {}
    is_tuple = not isinstance(y0, torch.Tensor)
    with torch.no_grad():
        t = torch.as_tensor(t, dtype=torch.float64)
        if is_tuple:
            shapes = [yi.shape for yi in y0]
            packed = _Packed(func, shapes)
            f, y = packed, torch.cat([yi.reshape(-1) for yi in y0])
            norm = packed.norm
        else:
            f, y, norm = func, y0, _rms
        if len(t) > 1 and float(t[0]) > float(t[1]):
            fwd = f
            f = lambda s, yy, _f=fwd: -_f(-s, yy)  # noqa: E731
            t = -t
        tl = [float(ti) for ti in t]
        if method in (None, "dopri5"):
            sol = _dopri5(f, y, tl, rtol, atol, norm)
        elif method == "rk4":
            sol = _rk4_fixed(f, y, tl, options.get("step_size"))
        else:
            raise ValueError("shim implements dopri5 and rk4 only, not %r" % (method,))
        if is_tuple:
            comps = [packed.unpack(s) for s in sol]
            return tuple(torch.stack([c[i] for c in comps]) for i in range(len(shapes)))
        return sol
