"""Minimal stand-in for the subset of ``nevergrad`` pix2latent uses (reference:
pix2latent/optimizer/base_ng_optimizer.py:33,81-83,107-109,168-169): ``ng.p.Array(init=mu)``,
``ng.optimizers.registry[method](parametrization=..., budget=...)``, ``optimizer.ask()`` -> candidate with
``.args[0]``, ``optimizer.tell(candidate, loss)``.

Used ONLY when the real package (requirements.txt:3 of the reference, ``nevergrad>=0.4.0.post3``) is not importable,
as in the offline build image, so that the Nevergrad / hybrid search loops can run and be tested; the search is host
code outside the accelerated path. Two methods: ``CMA`` (the package's default choice in the examples,
examples/invert_biggan_hybrid_nevergrad.py: --ng_method CMA) on top of ``_minicma`` with nevergrad's ask / tell
buffering — any number of asks per update, the distribution moves once ``popsize`` losses have been told — and
``RandomSearch``. It is not a re-implementation of nevergrad."""
import types

import numpy as np

from . import _minicma


class _Candidate:
    def __init__(self, x):
        self.value = x
        self.args = (x,)
        self.kwargs = {}


class _Array:
    def __init__(self, init):
        self.init = np.asarray(init, dtype=np.float64)


class _Base:
    def __init__(self, parametrization, budget=None, num_workers=1):
        self.parametrization = parametrization
        self.budget = budget
        self.num_ask = 0
        self.num_tell = 0

    def provide_recommendation(self):
        return _Candidate(self.recommendation())


class _CMA(_Base):
    seed = None  # class attribute so that tests can fix it

    def __init__(self, parametrization, budget=None, num_workers=1):
        super().__init__(parametrization, budget, num_workers)
        x0 = parametrization.init.ravel()
        opts = {} if self.seed is None else {"seed": self.seed}
        self.shape = parametrization.init.shape
        self.es = _minicma.CMAEvolutionStrategy(x0, 1.0, opts)
        self._queue, self._told_x, self._told_f = [], [], []

    def ask(self):
        if not self._queue:
            self._queue = list(self.es.ask())
        self.num_ask += 1
        return _Candidate(self._queue.pop(0).reshape(self.shape))

    def tell(self, candidate, loss):
        self.num_tell += 1
        self._told_x.append(np.asarray(candidate.args[0], dtype=np.float64).ravel())
        self._told_f.append(float(loss))
        if len(self._told_f) >= self.es.sp.popsize:
            n = self.es.sp.popsize
            self.es.tell(self._told_x[:n], self._told_f[:n])
            self._told_x, self._told_f = self._told_x[n:], self._told_f[n:]
            self._queue = []  # candidates drawn from the old distribution are stale

    def recommendation(self):
        return self.es.mean.reshape(self.shape)


class _RandomSearch(_Base):
    seed = None

    def __init__(self, parametrization, budget=None, num_workers=1):
        super().__init__(parametrization, budget, num_workers)
        self.rng = np.random.RandomState(self.seed)
        self.best, self.best_f = parametrization.init.copy(), np.inf

    def ask(self):
        self.num_ask += 1
        return _Candidate(self.parametrization.init + self.rng.standard_normal(self.parametrization.init.shape))

    def tell(self, candidate, loss):
        self.num_tell += 1
        if loss < self.best_f:
            self.best, self.best_f = np.asarray(candidate.args[0]), float(loss)

    def recommendation(self):
        return self.best


p = types.SimpleNamespace(Array=_Array)
optimizers = types.SimpleNamespace(registry={"CMA": _CMA, "RandomSearch": _RandomSearch})
STAND_IN = True
