"""CPU emulation of csrc/t2o_nm.cu's state machine (same phases, stable sort, permutation rows) checked against the
host coroutine on synthetic functions: python scripts/dev/nm_emul.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from t2onet_b200.nelder_mead import nelder_mead

INIT, REFLECT, EXPAND, COUT, CIN, SHRINK, DONE = range(7)


class Emu:
    def __init__(self, x0):
        x0 = np.asarray(x0, dtype=np.float64)
        N = len(x0)
        self.N, self.maxfun = N, 200 * N
        self.sim = np.zeros((N + 1, N))
        for k in range(N + 1):
            v = x0.copy()
            if k >= 1:
                v[k - 1] = (1 + 0.05) * x0[k - 1] if x0[k - 1] != 0 else 0.00025
            self.sim[k] = v
        self.f = np.full(N + 1, np.inf)
        self.row = np.arange(N + 1)
        self.fcalls = self.iters = 0
        self.xbar = self.xr = self.pend = None
        self.fxr = 0.0
        self.phase, self.k = INIT, 0
        self.propose(self.sim[0].copy(), INIT, 0)

    def propose(self, x, phase, k):
        if self.fcalls >= self.maxfun:
            return False
        self.fcalls += 1
        self.pend = x.copy()
        self.phase, self.k = phase, k
        return True

    def sort(self):
        key = np.where(np.isnan(self.f), np.inf, self.f)
        order = np.argsort(key, kind='stable')
        self.f, self.row = self.f[order], self.row[order]

    def finish(self):
        self.phase = DONE
        self.x = self.sim[self.row[0]].copy()
        self.fun = self.f[0]
        self.status = 1 if self.fcalls >= self.maxfun else (2 if self.iters >= self.maxfun else 0)

    def begin(self):
        N = self.N
        if not (self.fcalls < self.maxfun and self.iters < self.maxfun):
            return self.finish()
        s0 = self.sim[self.row[0]]
        dx = max(np.max(np.abs(self.sim[self.row[k]] - s0)) for k in range(1, N + 1))
        df = np.max(np.abs(self.f[0] - self.f[1:]))
        if dx <= 1e-4 and df <= 1e-4:
            return self.finish()
        s = self.sim[self.row[0]].copy()
        for k in range(1, N):
            s = s + self.sim[self.row[k]]
        self.xbar = s / N
        self.xr = 2.0 * self.xbar - self.sim[self.row[N]]
        if not self.propose(self.xr, REFLECT, 0):
            self.sort(); self.finish()

    def end(self, aborted):
        if not aborted:
            self.iters += 1
        self.sort(); self.begin()

    def accept(self, x, f):
        self.sim[self.row[self.N]] = x
        self.f[self.N] = f

    def shrink_vertex(self, j):
        b = self.sim[self.row[0]]
        x = b + 0.5 * (self.sim[self.row[j]] - b)
        self.sim[self.row[j]] = x
        return self.propose(x, SHRINK, j)

    def advance(self, fv):
        N, k = self.N, self.k
        f0, fN, fN1 = self.f[0], self.f[N], self.f[max(N - 1, 0)]
        ph = self.phase
        if ph == INIT:
            self.f[k] = fv
            if k < N:
                self.propose(self.sim[k + 1].copy(), INIT, k + 1)
            else:
                self.sort(); self.iters = 1; self.begin()
        elif ph == REFLECT:
            self.fxr = fv
            worst = self.sim[self.row[N]]
            if fv < f0:
                xe = 3.0 * self.xbar - 2.0 * worst
                if not self.propose(xe, EXPAND, 0): self.end(True)
            elif fv < fN1:
                self.accept(self.xr, fv); self.end(False)
            else:
                outside = fv < fN
                xc = 1.5 * self.xbar - 0.5 * worst if outside else 0.5 * self.xbar + 0.5 * worst
                if not self.propose(xc, COUT if outside else CIN, 0): self.end(True)
        elif ph == EXPAND:
            if fv < self.fxr: self.accept(self.pend, fv)
            else: self.accept(self.xr, self.fxr)
            self.end(False)
        elif ph in (COUT, CIN):
            if (fv <= self.fxr) if ph == COUT else (fv < fN):
                self.accept(self.pend, fv); self.end(False)
            elif not self.shrink_vertex(1):
                self.end(True)
        elif ph == SHRINK:
            self.f[k] = fv
            if k < N:
                if not self.shrink_vertex(k + 1): self.end(True)
            else:
                self.end(False)


def run_emu(func, x0):
    e = Emu(x0)
    while e.phase != DONE:
        e.advance(func(e.pend))
    return e


def run_co(func, x0):
    g = nelder_mead(np.asarray(x0, dtype=np.float64))
    try:
        x = next(g)
        while True:
            x = g.send(func(x))
    except StopIteration as s:
        return s.value


if __name__ == '__main__':
    rng = np.random.default_rng(1)
    for N, x0 in ((1, [0.0]), (8, np.ones(8)), (24, np.ones(24)), (8, np.zeros(8))):
        for trial in range(5):
            c = rng.random(N) * 2
            w = rng.random(N) + 0.1
            def func(x, c=c, w=w):
                x32 = np.asarray(x, dtype=np.float32)
                return float(np.float32(np.sum(np.abs(x32 - c.astype(np.float32)) * w.astype(np.float32)) * np.float32(0.01)))
            a, b = run_emu(func, x0), run_co(func, x0)
            ok = np.array_equal(a.x, b.x) and a.fcalls == b.nfev and a.iters == b.nit and a.status == b.status
            print(N, trial, 'OK' if ok else 'MISMATCH', a.fcalls, b.nfev, a.iters, b.nit, a.status, b.status)
