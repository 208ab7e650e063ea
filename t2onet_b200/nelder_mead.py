"""Nelder-Mead as a coroutine, so that many independent fits can be advanced in lock-step and their
pending vertices scored by ONE candidate-scoring launch per round.

The reference fits every (state, operator) with ``scipy.optimize.minimize(func, param0,
method='Nelder-Mead')`` (utils/beam_search.py:88).  This restates scipy's ``_minimize_neldermead``
(third-party, unpinned by the reference; restated from scipy 1.x: rho=1, chi=2, psi=sigma=0.5,
nonzdelt=0.05, zdelt=0.00025, xatol=fatol=1e-4, maxiter=maxfev=200*N, evaluation-count guard that
aborts the current iteration) with float64 simplex arithmetic in the same order, so that equal
function values give the same vertex sequence.
"""
import numpy as np


class _MaxFun(Exception):
    pass


class NMResult:
    __slots__ = ('x', 'fun', 'nit', 'nfev', 'status', 'success')

    def __init__(self, x, fun, nit, nfev, status):
        self.x, self.fun, self.nit, self.nfev, self.status = x, fun, nit, nfev, status
        self.success = status == 0


def nelder_mead(x0, xatol=1e-4, fatol=1e-4, stable=False):
    """Generator: yields a float64 vertex to evaluate, expects f(vertex) via send(); returns NMResult.

    stable=True sorts the simplex with a stable argsort (ties keep their order) instead of scipy's np.argsort
    default (an unstable quicksort whose tie order is an implementation detail): that is what the device version
    (csrc/t2o_nm.cu) does, and the setting its cross-check uses."""
    argsort = (lambda a: np.argsort(a, kind='stable')) if stable else np.argsort
    x0 = np.asarray(np.atleast_1d(x0).flatten(), dtype=np.float64)
    rho, chi, psi, sigma = 1, 2, 0.5, 0.5
    nonzdelt, zdelt = 0.05, 0.00025
    N = len(x0)
    sim = np.empty((N + 1, N), dtype=x0.dtype)
    sim[0] = x0
    for k in range(N):
        y = np.array(x0, copy=True)
        if y[k] != 0:
            y[k] = (1 + nonzdelt) * y[k]
        else:
            y[k] = zdelt
        sim[k + 1] = y
    maxiter = N * 200
    maxfun = N * 200
    fsim = np.full((N + 1,), np.inf, dtype=float)
    fcalls = [0]

    def func(x):
        if fcalls[0] >= maxfun:
            raise _MaxFun()
        fcalls[0] += 1
        fx = yield np.copy(x)
        return fx

    try:
        for k in range(N + 1):
            fsim[k] = yield from func(sim[k])
    except _MaxFun:
        pass
    ind = argsort(fsim)
    sim = np.take(sim, ind, 0)
    fsim = np.take(fsim, ind, 0)
    ind = argsort(fsim)
    fsim = np.take(fsim, ind, 0)
    sim = np.take(sim, ind, 0)
    iterations = 1
    while fcalls[0] < maxfun and iterations < maxiter:
        try:
            if (np.max(np.ravel(np.abs(sim[1:] - sim[0]))) <= xatol and
                    np.max(np.abs(fsim[0] - fsim[1:])) <= fatol):
                break
            xbar = np.add.reduce(sim[:-1], 0) / N
            xr = (1 + rho) * xbar - rho * sim[-1]
            fxr = yield from func(xr)
            doshrink = 0
            if fxr < fsim[0]:
                xe = (1 + rho * chi) * xbar - rho * chi * sim[-1]
                fxe = yield from func(xe)
                if fxe < fxr:
                    sim[-1] = xe
                    fsim[-1] = fxe
                else:
                    sim[-1] = xr
                    fsim[-1] = fxr
            else:
                if fxr < fsim[-2]:
                    sim[-1] = xr
                    fsim[-1] = fxr
                else:
                    if fxr < fsim[-1]:
                        xc = (1 + psi * rho) * xbar - psi * rho * sim[-1]
                        fxc = yield from func(xc)
                        if fxc <= fxr:
                            sim[-1] = xc
                            fsim[-1] = fxc
                        else:
                            doshrink = 1
                    else:
                        xcc = (1 - psi) * xbar + psi * sim[-1]
                        fxcc = yield from func(xcc)
                        if fxcc < fsim[-1]:
                            sim[-1] = xcc
                            fsim[-1] = fxcc
                        else:
                            doshrink = 1
                    if doshrink:
                        for j in range(1, N + 1):
                            sim[j] = sim[0] + sigma * (sim[j] - sim[0])
                            fsim[j] = yield from func(sim[j])
            iterations += 1
        except _MaxFun:
            pass
        ind = argsort(fsim)
        sim = np.take(sim, ind, 0)
        fsim = np.take(fsim, ind, 0)
    x = sim[0]
    fval = np.min(fsim)
    status = 0
    if fcalls[0] >= maxfun:
        status = 1
    elif iterations >= maxiter:
        status = 2
    return NMResult(x, fval, iterations, fcalls[0], status)


def run_lockstep(generators, score_batch):
    """Advance generators (dict key -> nelder_mead generator) in lock-step.

    score_batch(keys, points) -> sequence of float function values, one per pending point, is called once
    per round with every active problem's pending vertex.  Returns dict key -> NMResult."""
    pending, results = {}, {}
    for key, gen in generators.items():
        try:
            pending[key] = next(gen)
        except StopIteration as stop:
            results[key] = stop.value
    while pending:
        keys = list(pending.keys())
        values = score_batch(keys, [pending[k] for k in keys])
        for key, fx in zip(keys, values):
            try:
                pending[key] = generators[key].send(fx)
            except StopIteration as stop:
                results[key] = stop.value
                del pending[key]
    return results
