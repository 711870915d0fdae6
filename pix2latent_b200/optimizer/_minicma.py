"""Minimal (mu/mu_w, lambda)-CMA-ES with PyCMA's interface subset used by pix2latent
(``CMAEvolutionStrategy(x0, sigma0, opts)``, ``.sp.popsize``, ``.ask()``, ``.tell(X, f)``,
``.mean``) — Hansen, "The CMA Evolution Strategy: A Tutorial" (2016), default strategy
parameters, default population size 4 + floor(3 ln N) (=> 18 for N=128, 22 for N=512, the sizes
README.md:74 of the reference quotes).

Used ONLY when the real ``cma`` package (requirements.txt:1 of the reference, ``cma>=3.0.3``) is
not importable, as in the offline build image; the search loop is host code outside the
accelerated path (SURVEY.md §2.1 'host-side'). It is not a re-implementation of PyCMA's many
options (no boundary handling, no restarts, no active CMA)."""
import math
import types

import numpy as np


class CMAEvolutionStrategy:
    def __init__(self, x0, sigma0, inopts=None):
        opts = dict(inopts or {})
        self.mean = np.array(x0, dtype=np.float64).ravel().copy()
        N = self.N = self.mean.size
        self.sigma = float(sigma0)
        self.rng = np.random.RandomState(opts.get("seed", None))
        lam = int(opts.get("popsize", 4 + int(3 * math.log(N))))
        mu = lam // 2
        w = math.log((lam + 1) / 2.0) - np.log(np.arange(1, mu + 1))
        self.weights = w / w.sum()
        self.mueff = 1.0 / np.sum(self.weights ** 2)
        self.mu, self.lam = mu, lam
        me = self.mueff
        self.cc = (4 + me / N) / (N + 4 + 2 * me / N)
        self.cs = (me + 2) / (N + me + 5)
        on = float(opts.get("CMA_on", 1))
        self.c1 = on * 2 / ((N + 1.3) ** 2 + me)
        self.cmu = on * min(1 - self.c1, 2 * (me - 2 + 1 / me) / ((N + 2) ** 2 + me))
        self.damps = 1 + 2 * max(0, math.sqrt((me - 1) / (N + 1)) - 1) + self.cs
        self.chiN = math.sqrt(N) * (1 - 1.0 / (4 * N) + 1.0 / (21 * N * N))
        self.pc = np.zeros(N)
        self.ps = np.zeros(N)
        self.C = np.eye(N)
        self.B = np.eye(N)
        self.D = np.ones(N)
        self.invsqrtC = np.eye(N)
        self.countiter = 0
        self.eigen_iter = 0
        self.sp = types.SimpleNamespace(popsize=lam)

    def ask(self, number=None):
        n = self.lam if number is None else int(number)
        z = self.rng.standard_normal((n, self.N))
        y = (z * self.D) @ self.B.T
        return [self.mean + self.sigma * yi for yi in y]

    def tell(self, solutions, function_values):
        X = np.asarray(solutions, dtype=np.float64)
        f = np.asarray(function_values, dtype=np.float64).ravel()
        assert X.shape[0] == f.shape[0] == self.lam, "tell() needs popsize solutions"
        self.countiter += 1
        N = self.N
        order = np.argsort(f)
        Xs = X[order[: self.mu]]
        old = self.mean
        self.mean = self.weights @ Xs
        y = (self.mean - old) / self.sigma
        self.ps = (1 - self.cs) * self.ps + math.sqrt(self.cs * (2 - self.cs) * self.mueff) * (self.invsqrtC @ y)
        hsig = (np.linalg.norm(self.ps) / math.sqrt(1 - (1 - self.cs) ** (2 * self.countiter)) / self.chiN
                < 1.4 + 2.0 / (N + 1))
        self.pc = (1 - self.cc) * self.pc + hsig * math.sqrt(self.cc * (2 - self.cc) * self.mueff) * y
        art = (Xs - old) / self.sigma
        self.C = ((1 - self.c1 - self.cmu) * self.C
                  + self.c1 * (np.outer(self.pc, self.pc) + (1 - hsig) * self.cc * (2 - self.cc) * self.C)
                  + self.cmu * (art.T * self.weights) @ art)
        self.sigma *= math.exp((self.cs / self.damps) * (np.linalg.norm(self.ps) / self.chiN - 1))
        if self.c1 + self.cmu > 0 and self.countiter - self.eigen_iter > 1.0 / (self.c1 + self.cmu) / N / 10.0:
            self.eigen_iter = self.countiter
            self.C = np.triu(self.C) + np.triu(self.C, 1).T
            d, self.B = np.linalg.eigh(self.C)
            self.D = np.sqrt(np.maximum(d, 1e-20))
            self.invsqrtC = (self.B / self.D) @ self.B.T
