"""CPU oracle for the littlemcmc HMC / NUTS hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the reference's algorithm (eigenfoo/littlemcmc v0.2.2), written
as plain functions over explicit state with an *explicit* random source, and with the recursive tree
builder unrolled into the iterative binary-counter stack the CUDA kernels use.
It exists to check the CUDA path; nothing under ``littlemcmc_b200/`` may import it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs use it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified reference (imported from
/root/reference in the build container) and stores its outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays the same legacy-MT19937 seeds through this oracle and requires
identical discrete statistics and <=1e-12 relative agreement on the continuous ones.

Every function cites the reference file:line it follows.  All potentials are float64 (the reference's
float32 default is a quirk we do not reproduce: SURVEY.md A.2-1).

Random source protocol (``rng``): ``rng.normal(size=n) -> ndarray[n]``, ``rng.uniform() -> float``,
``rng.rand() -> float``.  ``numpy.random.RandomState(seed)`` satisfies it and reproduces the
reference's global legacy stream; :class:`TapeRecorder` wraps any source and records what was consumed
per transition so the same numbers can be handed to the CUDA kernels in tape mode; :class:`TapeRNG`
replays such a tape.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

LogpFunc = Callable[[np.ndarray], Tuple[float, np.ndarray]]

NUTS_STAT_NAMES = (
    "depth", "step_size", "tune", "mean_tree_accept", "step_size_bar", "tree_size",
    "diverging", "energy_error", "energy", "max_energy_error", "model_logp",
)  # reference nuts.py:87-101
HMC_STAT_NAMES = (
    "step_size", "n_steps", "tune", "step_size_bar", "accept", "diverging",
    "energy_error", "energy", "path_length", "accepted", "model_logp",
)  # reference hmc.py:36-50


# --------------------------------------------------------------------------------------------------
# random sources
# --------------------------------------------------------------------------------------------------
class TapeRecorder:
    """Wrap a random source and record the numbers consumed, split per transition."""

    def __init__(self, rng):
        self._rng = rng
        self.normals: List[np.ndarray] = []      # one [D] vector per transition
        self.uniforms: List[List[float]] = []    # uniforms/rands consumed in that transition

    def begin_transition(self):
        self.uniforms.append([])

    def normal(self, size):
        v = np.asarray(self._rng.normal(size=size), dtype=np.float64)
        self.normals.append(v.copy())
        return v

    def uniform(self):
        u = float(self._rng.uniform())
        self.uniforms[-1].append(u)
        return u

    def rand(self):
        u = float(self._rng.rand())
        self.uniforms[-1].append(u)
        return u

    def tapes(self, pad_to: Optional[int] = None):
        """-> (normals [T, D], uniforms [T, U] padded with 0.5, n_uniforms [T])."""
        n = np.array([len(u) for u in self.uniforms], dtype=np.int64)
        width = int(max(int(n.max()) if len(n) else 0, 1))
        if pad_to is not None:
            if pad_to < width:
                raise ValueError("pad_to smaller than the longest uniform run")
            width = pad_to
        u = np.full((len(self.uniforms), width), 0.5, dtype=np.float64)
        for t, row in enumerate(self.uniforms):
            u[t, : len(row)] = row
        return np.stack(self.normals), u, n


class TapeRNG:
    """Replay tapes: ``normals[T, D]`` one row per transition, ``uniforms[T, U]`` by sequential counter."""

    def __init__(self, normals: np.ndarray, uniforms: np.ndarray):
        self._normals, self._uniforms = np.asarray(normals), np.asarray(uniforms)
        self._t, self._k = -1, 0

    def begin_transition(self):
        self._t += 1
        self._k = 0

    def normal(self, size):
        row = self._normals[self._t]
        assert row.shape[0] == size
        return row.astype(np.float64, copy=True)

    def uniform(self):
        u = float(self._uniforms[self._t, self._k])
        self._k += 1
        return u

    rand = uniform


def _begin(rng):
    hook = getattr(rng, "begin_transition", None)
    if hook is not None:
        hook()


# --------------------------------------------------------------------------------------------------
# scalar helpers  (reference math.py)
# --------------------------------------------------------------------------------------------------
def logbern(log_p: float, rng) -> bool:
    """Bernoulli trial in log space: ``log(U) < log_p``  (reference math.py:21-25)."""
    if np.isnan(log_p):
        raise FloatingPointError("log_p can't be nan.")
    return bool(np.log(rng.uniform()) < log_p)


def log1mexp(x: float) -> float:
    """log(1 - exp(-x)) with the 0.683 switch  (reference math.py:28-35)."""
    with np.errstate(all="ignore"):
        if x < 0.683:
            return float(np.log(-np.expm1(-x)))
        return float(np.log1p(-np.exp(-x)))


def logdiffexp(a: float, b: float) -> float:
    """log(exp(a) - exp(b))  (reference math.py:38-40)."""
    return a + log1mexp(a - b)


# --------------------------------------------------------------------------------------------------
# diagonal potentials  (reference quadpotential.py:148-387)
# --------------------------------------------------------------------------------------------------
@dataclass
class Welford:
    """Running weighted mean / sum of squares  (reference quadpotential.py:294-340, weight always 1)."""
    mean: np.ndarray
    raw_var: np.ndarray
    w_sum: float

    @staticmethod
    def fresh(n: int) -> "Welford":
        return Welford(np.zeros(n), np.zeros(n), 0.0)                     # :305-313 with weight 0

    @staticmethod
    def seeded(mean, variance, weight) -> "Welford":
        w = float(weight)
        return Welford(np.array(mean, dtype="d"), np.array(variance, dtype="d") * w, w)   # :308-315

    def add_sample(self, x):
        # :322-330 with weight == 1
        self.w_sum += 1
        prop = 1 / self.w_sum
        old_diff = x - self.mean
        self.mean = self.mean + prop * old_diff
        new_diff = x - self.mean
        self.raw_var = self.raw_var + 1 * old_diff * new_diff

    def variance(self):
        if self.w_sum == 0:
            raise ValueError("Can not compute variance without samples.")   # :333-334
        return self.raw_var / self.w_sum                                    # :335-338


class DiagPotential:
    """``QuadPotentialDiag`` (static, ``adapt=False``) or ``QuadPotentialDiagAdapt`` in float64.

    reference quadpotential.py:346-387 (static) and :148-245 (adaptive).
    """

    def __init__(self, n, *, var=None, initial_mean=None, initial_weight=0.0, adapt=True,
                 adaptation_window=101, adaptation_window_multiplier=1.0):
        self.n, self.adapt = int(n), bool(adapt)
        if var is None:                                        # :178-180
            var, initial_weight = np.ones(n), 1
        self._initial_diag = np.array(var, dtype="d")
        self._initial_mean = np.zeros(n) if initial_mean is None else np.array(initial_mean, dtype="d")
        self._initial_weight = float(initial_weight)
        self._initial_window = int(adaptation_window)
        self.adaptation_window_multiplier = float(adaptation_window_multiplier)
        self.reset()

    def reset(self):                                           # :195-204
        self.var = self._initial_diag.copy()
        self.stds = np.sqrt(self._initial_diag)
        self.inv_stds = 1.0 / self.stds
        self.fg = Welford.seeded(self._initial_mean, self._initial_diag, self._initial_weight)
        self.bg = Welford.fresh(self.n)
        self.n_samples = 0
        # the reference mutates adaptation_window in place and never restores it on reset(); with the
        # default multiplier of 1 that is unobservable, and we keep the same (non-restoring) behaviour
        if not hasattr(self, "adaptation_window"):
            self.adaptation_window = self._initial_window

    def velocity(self, p):                                     # :206-208 / :367-372
        return self.var * p

    def random(self, rng):                                     # :221-224 / :374-376
        return self.inv_stds * rng.normal(size=self.n)

    def update(self, sample, tune):                            # :231-245
        if not (tune and self.adapt):
            return
        self.fg.add_sample(sample)
        self.bg.add_sample(sample)
        self.var = self.fg.variance()                          # :226-229
        self.stds = np.sqrt(self.var)
        self.inv_stds = 1 / self.stds
        if self.n_samples > 0 and self.n_samples % self.adaptation_window == 0:
            self.fg = self.bg
            self.bg = Welford.fresh(self.n)
            self.adaptation_window = int(self.adaptation_window * self.adaptation_window_multiplier)
        self.n_samples += 1


# --------------------------------------------------------------------------------------------------
# dense potentials  (reference quadpotential.py:390-615), float64 throughout
# --------------------------------------------------------------------------------------------------
class FullPotential:
    """``QuadPotentialFull``: static dense covariance (reference quadpotential.py:430-468)."""
    adapt = False

    def __init__(self, cov):
        import scipy.linalg
        self.cov = np.array(cov, dtype="d")                                  # :445
        self.chol = scipy.linalg.cholesky(self.cov, lower=True)               # :446
        self.n = len(self.cov)

    def velocity(self, p):                                                    # :449-451
        return np.dot(self.cov, p)

    def random(self, rng):                                                    # :453-456
        import scipy.linalg
        vals = rng.normal(size=self.n)
        return scipy.linalg.solve_triangular(self.chol.T, vals)

    def update(self, sample, tune):
        pass

    def reset(self):                                                          # base class: no-op (:138-140)
        pass


class FullInvPotential(FullPotential):
    """``QuadPotentialFullInv``: static dense inverse covariance A (reference quadpotential.py:390-427)."""

    def __init__(self, A):
        import scipy.linalg
        self.A = np.array(A, dtype="d")
        self.L = scipy.linalg.cholesky(self.A, lower=True)                    # :405
        self.n = len(self.A)

    def velocity(self, p):                                                    # :407-412
        import scipy.linalg
        return scipy.linalg.cho_solve((self.L, True), p)

    def random(self, rng):                                                    # :414-417
        return np.dot(self.L, rng.normal(size=self.n))


class WelfordCov:
    """``_WeightedCovariance`` (reference quadpotential.py:573-615)."""

    def __init__(self, n, initial_mean=None, initial_covariance=None, initial_weight=0.0):
        self.n_samples = float(initial_weight)
        self.mean = np.zeros(n) if initial_mean is None else np.array(initial_mean, dtype="d")
        self.raw_cov = np.eye(n) if initial_covariance is None else np.array(initial_covariance, dtype="d")
        self.raw_cov = self.raw_cov * self.n_samples                          # :600

    def add_sample(self, x, weight=1):                                        # :607-613
        x = np.asarray(x, dtype="d")
        self.n_samples += 1
        old_diff = x - self.mean
        self.mean = self.mean + old_diff / self.n_samples
        new_diff = x - self.mean
        self.raw_cov = self.raw_cov + weight * new_diff[:, None] * old_diff[None, :]

    def current_covariance(self):                                             # :615-621
        if self.n_samples == 0:
            raise ValueError("Can not compute covariance without samples.")
        return self.raw_cov / (self.n_samples - 1)


class FullAdaptPotential(FullPotential):
    """``QuadPotentialFullAdapt`` (reference quadpotential.py:471-570).  The reference never resets this potential
    between chains (base-class ``reset`` is a no-op, :138-140); ``reset`` here does nothing either, so a fresh object
    per chain is what gives every chain the same initial mass matrix."""
    adapt = True

    def __init__(self, n, initial_mean, initial_cov=None, initial_weight=0, adaptation_window=101,
                 adaptation_window_multiplier=2, update_window=1):
        import scipy.linalg
        if initial_cov is None:                                               # :500-502
            initial_cov, initial_weight = np.eye(n), 1
        self.n = int(n)
        self.cov = np.array(initial_cov, dtype="d")
        self.chol = scipy.linalg.cholesky(self.cov, lower=True)
        self.chol_error = None
        self.fg = WelfordCov(self.n, initial_mean, initial_cov, initial_weight)
        self.bg = WelfordCov(self.n)
        self.n_samples = 0
        self.adaptation_window = int(adaptation_window)
        self.adaptation_window_multiplier = float(adaptation_window_multiplier)
        self.update_window = int(update_window)
        self.previous_update = 0

    def update(self, sample, tune):                                           # :528-554
        import scipy.linalg
        if not tune:
            return
        delta = self.n_samples - self.previous_update
        self.fg.add_sample(sample, weight=1)
        self.bg.add_sample(sample, weight=1)
        if (delta + 1) % self.update_window == 0:                             # :540-541 -> :520-526
            self.cov = self.fg.current_covariance()
            try:
                self.chol = scipy.linalg.cholesky(self.cov, lower=True)
            except (scipy.linalg.LinAlgError, ValueError) as error:
                self.chol_error = error
        if delta >= self.adaptation_window:                                   # :545-552
            self.fg = self.bg
            self.bg = WelfordCov(self.n)
            self.previous_update = self.n_samples
            self.adaptation_window = int(self.adaptation_window * self.adaptation_window_multiplier)
        self.n_samples += 1


# --------------------------------------------------------------------------------------------------
# dual averaging  (reference step_sizes.py:23-99)
# --------------------------------------------------------------------------------------------------
class DualAverage:
    def __init__(self, initial_step, target=0.8, gamma=0.05, k=0.75, t0=10):
        self.initial_step, self.target, self.gamma, self.k, self.t0 = initial_step, target, gamma, k, t0
        self.reset()

    def reset(self):                                           # :49-56
        self.log_step = np.log(self.initial_step)
        self.log_bar = self.log_step
        self.hbar = 0.0
        self.count = 1
        self.mu = np.log(10 * self.initial_step)

    def current(self, tune):                                   # :58-69
        return np.exp(self.log_step) if tune else np.exp(self.log_bar)

    def update(self, accept_stat, tune):                       # :71-92
        if not tune:
            return
        count, k, t0 = self.count, self.k, self.t0
        w = 1.0 / (count + t0)
        self.hbar = (1 - w) * self.hbar + w * (self.target - accept_stat)
        self.log_step = self.mu - self.hbar * np.sqrt(count) / self.gamma
        mk = count ** -k
        self.log_bar = mk * self.log_step + (1 - mk) * self.log_bar
        self.count += 1

    def stats(self):                                           # :94-99
        return {"step_size": np.exp(self.log_step), "step_size_bar": np.exp(self.log_bar)}


# --------------------------------------------------------------------------------------------------
# leapfrog  (reference integration.py)
# --------------------------------------------------------------------------------------------------
@dataclass
class State:                                                   # integration.py:25
    q: np.ndarray
    p: np.ndarray
    v: np.ndarray
    q_grad: np.ndarray
    energy: float
    model_logp: float


def _scalar(x) -> float:
    return float(np.asarray(x).reshape(-1)[0]) if np.ndim(x) else float(x)


def compute_state(f: LogpFunc, pot: DiagPotential, q, p) -> State:
    """reference integration.py:52-66."""
    logp, dlogp = f(q)
    logp = _scalar(logp)
    v = pot.velocity(p)
    kinetic = 0.5 * p.dot(v)                                   # quadpotential.py:210-214
    return State(q, p, v, np.asarray(dlogp, dtype="d"), kinetic - logp, logp)


def leapfrog(f: LogpFunc, pot: DiagPotential, epsilon: float, s: State) -> State:
    """One leapfrog step, reference integration.py:100-121 (epsilon may be negative)."""
    dt = 0.5 * epsilon
    p_half = s.p + dt * s.q_grad                               # :108
    v_half = pot.velocity(p_half)                              # :111
    q_new = s.q + epsilon * v_half                             # :112
    logp, grad_new = f(q_new)                                  # :115
    logp = _scalar(logp)
    grad_new = np.asarray(grad_new, dtype="d")
    p_new = p_half + dt * grad_new                             # :116
    v_new = pot.velocity(p_new)                                # :118 -> quadpotential.py:216-219
    kinetic = 0.5 * np.dot(p_new, v_new)
    return State(q_new, p_new, v_new, grad_new, kinetic - logp, logp)


# --------------------------------------------------------------------------------------------------
# NUTS transition, iterative form  (reference nuts.py:204-224, 251-435)
# --------------------------------------------------------------------------------------------------
@dataclass
class _Node:
    """Summary of a (sub)tree: the `Subtree` namedtuple of nuts.py:246-248 minus n_proposals."""
    left: State
    right: State
    p_sum: np.ndarray
    prop: State            # proposal: q, q_grad, energy, logp are read from it (nuts.py:243)
    log_size: float
    lwas: float            # log_weighted_accept_sum


def nuts_transition(f: LogpFunc, pot: DiagPotential, start: State, step_size: float, Emax: float,
                    max_treedepth: int, rng) -> Tuple[State, Dict[str, float], bool]:
    """One NUTS trajectory from ``start``; returns (proposal state, stats, diverging)."""
    E0 = start.energy
    left = right = start                                       # nuts.py:275
    prop = start                                               # :276
    log_size, lwas = 0.0, -np.inf                              # :278-279
    p_sum = start.p.copy()                                     # :282
    depth, n_prop, max_dE = 0, 0, 0.0
    diverging = turning = False

    for _ in range(max_treedepth):                             # :212
        direction = 1 if logbern(np.log(0.5), rng) else -1     # :213
        z = right if direction > 0 else left                   # :297 / :306
        eps = direction * step_size
        stack: List[_Node] = []
        fail = None
        n_leaves = 0
        for i in range(1 << depth):                            # leaves of _build_subtree(:377) in order
            z = leapfrog(f, pot, eps, z)                       # :347
            dE = z.energy - E0                                 # :352
            if np.isnan(dE):
                dE = np.inf                                    # :353-354
            if np.abs(dE) > np.abs(max_dE):
                max_dE = dE                                    # :356-357
            n_leaves += 1
            if not (np.abs(dE) < Emax):                        # :358, else-branch :370-375
                fail = "diverge"
                break
            lpaw = -dE + min(0.0, -dE)                         # :363
            cur = _Node(z, z, z.p, z, -dE, lpaw)               # :364-368
            j, lvl = i, 0
            while j & 1:                                       # one merge per trailing 1-bit == post-order
                t1, t2 = stack.pop(), cur
                ps = t1.p_sum + t2.p_sum                                                   # :390
                turn = (ps.dot(t1.left.v) <= 0) or (ps.dot(t2.right.v) <= 0)               # :391
                if lvl > 0:                                                                # :393
                    ps1 = t1.p_sum + t2.left.p                                             # :394
                    turn1 = (ps1.dot(t1.left.v) <= 0) or (ps1.dot(t2.left.v) <= 0)         # :395
                    ps2 = t1.right.p + t2.p_sum                                            # :396
                    turn2 = (ps2.dot(t1.right.v) <= 0) or (ps2.dot(t2.right.v) <= 0)       # :397
                    turn = turn or turn1 or turn2
                nls = np.logaddexp(t1.log_size, t2.log_size)                               # :400
                nlw = np.logaddexp(t1.lwas, t2.lwas)                                       # :401-403
                pr = t2.prop if logbern(t2.log_size - nls, rng) else t1.prop               # :404-407
                cur = _Node(t1.left, t2.right, ps, pr, float(nls), float(nlw))
                if turn:
                    fail = "turn"
                    break
                j >>= 1
                lvl += 1
            if fail:
                break
            stack.append(cur)
        depth += 1                                             # :315
        n_prop += n_leaves                                     # :316
        if fail:                                               # :318-319 -> :216-217
            diverging, turning = fail == "diverge", fail == "turn"
            break
        (T,) = stack
        old_left, old_right = left, right
        if direction > 0:
            right = T.right                                    # :304
        else:
            left = T.right                                     # :313
        if logbern(T.log_size - log_size, rng):                # :321-323
            prop = T.prop
        log_size = float(np.logaddexp(log_size, T.log_size))   # :325
        lwas = float(np.logaddexp(lwas, T.lwas))               # :326-328
        p_sum = p_sum + T.p_sum                                # :329 (in place in the reference)
        turn = (p_sum.dot(left.v) <= 0) or (p_sum.dot(right.v) <= 0)                       # :333-335
        if direction > 0:                                      # :300-303; leftmost_p_sum ALIASES the updated p_sum
            lm_b, lm_e, rm_b, rm_e = old_left, old_right, T.left, T.right
            lm_ps, rm_ps = p_sum, T.p_sum
        else:                                                  # :309-312; rightmost_p_sum aliases it
            lm_b, lm_e, rm_b, rm_e = T.right, T.left, old_left, old_right
            lm_ps, rm_ps = T.p_sum, p_sum
        ps1 = lm_ps + rm_b.p                                                               # :336
        turn1 = (ps1.dot(lm_b.v) <= 0) or (ps1.dot(rm_b.v) <= 0)                           # :337
        ps2 = lm_e.p + rm_ps                                                               # :338
        turn2 = (ps2.dot(lm_e.v) <= 0) or (ps2.dot(rm_e.v) <= 0)                           # :339
        if turn or turn1 or turn2:                                                         # :340
            turning = True
            break
    reached_max = not (diverging or turning)                   # the for/else of nuts.py:212-220 (counted at :219-220)

    mean_tree_accept = 0.0                                     # :280
    if log_size > 0:                                           # :421-425
        mean_tree_accept = float(np.exp(lwas - logdiffexp(log_size, 0.0)))
    stats = {
        "depth": depth,
        "mean_tree_accept": mean_tree_accept,
        "energy_error": prop.energy - E0,
        "energy": prop.energy,
        "tree_size": n_prop,
        "max_energy_error": max_dE,
        "model_logp": prop.model_logp,
        "reached_max_treedepth": reached_max,                  # not a reference statistic: NUTS._reached_max_treedepth += 1
    }                                                          # :427-435
    return prop, stats, diverging


# --------------------------------------------------------------------------------------------------
# HMC transition  (reference hmc.py:140-182)
# --------------------------------------------------------------------------------------------------
def hmc_transition(f: LogpFunc, pot: DiagPotential, start: State, step_size: float, Emax: float,
                   path_length_max: float, max_steps: int, rng) -> Tuple[State, Dict[str, float], bool]:
    path_length = rng.rand() * path_length_max                 # :141
    n_steps = max(1, int(path_length / step_size))             # :142
    n_steps = min(max_steps, n_steps)                          # :143
    state = start
    diverging = False
    for _ in range(n_steps):                                   # :149-150
        state = leapfrog(f, pot, step_size, state)
    if not np.isfinite(state.energy):                          # :154-155
        diverging = True
    energy_change = start.energy - state.energy                # :156
    if np.isnan(energy_change):
        energy_change = -np.inf                                # :157-158
    if np.abs(energy_change) > Emax:                           # :159-162
        diverging = True
    with np.errstate(over="ignore"):
        accept_stat = min(1, np.exp(energy_change))            # :164
    if diverging or rng.rand() >= accept_stat:                 # :166 (short-circuit: no draw when diverging)
        end, accepted = start, False
    else:
        end, accepted = state, True
    stats = {
        "path_length": path_length, "n_steps": n_steps, "accept": accept_stat,
        "energy_error": energy_change, "energy": state.energy, "accepted": accepted,
        "model_logp": state.model_logp,
    }                                                          # :173-181
    return end, stats, diverging


# --------------------------------------------------------------------------------------------------
# one chain: BaseHMC._astep + sampling._iter_sample
# --------------------------------------------------------------------------------------------------
@dataclass
class Sampler:
    """State of one chain's step method (reference base_hmc.py:28-131, nuts.py:103-202, hmc.py:52-138)."""
    f: LogpFunc
    ndim: int
    pot: DiagPotential
    kind: str = "nuts"
    target_accept: float = 0.8
    Emax: float = 1000.0
    adapt_step_size: bool = True
    step_scale: float = 0.25
    gamma: float = 0.05
    k: float = 0.75
    t0: int = 10
    max_treedepth: int = 10
    early_max_treedepth: int = 8
    path_length: float = 2.0
    max_steps: int = 1024
    tune: bool = True
    iter_count: int = 0
    step_rand: Optional[Callable[[float], float]] = None       # base_hmc.py:43,154-155
    reached_max_treedepth: int = 0                             # nuts.py:202,219-220
    step_adapt: DualAverage = field(init=False)

    def __post_init__(self):
        self.step_size = self.step_scale / (self.ndim ** 0.25)                         # base_hmc.py:102
        self.step_adapt = DualAverage(self.step_size, self.target_accept, self.gamma, self.k, self.t0)

    def reset_tuning(self):                                                            # base_hmc.py:192-200
        self.step_adapt.reset()
        self.tune = True
        self.pot.reset()

    def astep(self, q0: np.ndarray, rng):
        """One transition, reference base_hmc.py:140-190.  Returns (q_new, stats dict)."""
        _begin(rng)
        p0 = self.pot.random(rng)                                                      # :142
        start = compute_state(self.f, self.pot, q0, p0)                                # :143
        if not np.isfinite(start.energy):                                              # :145-148
            raise ValueError("Bad initial energy: {}. The model might be misspecified.".format(start.energy))
        adapt_step = self.tune and self.adapt_step_size                                # :151
        step_size = self.step_adapt.current(adapt_step)                                # :152
        if self.step_rand is not None:
            step_size = self.step_rand(step_size)                                      # :154-155
        self.step_size = step_size
        if self.kind == "nuts":
            early = self.tune and self.iter_count < 200                                # nuts.py:205-208
            depth_cap = self.early_max_treedepth if early else self.max_treedepth
            end, st, diverging = nuts_transition(self.f, self.pot, start, step_size, self.Emax, depth_cap, rng)
            accept_stat = st["mean_tree_accept"]
            if st["reached_max_treedepth"] and not self.tune:                          # nuts.py:218-220
                self.reached_max_treedepth += 1
        else:
            end, st, diverging = hmc_transition(self.f, self.pot, start, step_size, self.Emax,
                                                self.path_length, self.max_steps, rng)
            accept_stat = st["accept"]
        self.step_adapt.update(accept_stat, adapt_step)                                # :161
        self.pot.update(end.q, self.tune)                                              # :162
        self.iter_count += 1                                                           # :181
        stats = {"tune": self.tune, "diverging": bool(diverging)}                      # :185
        stats.update(st)
        stats.update(self.step_adapt.stats())                                          # :188
        return end.q, stats


def sample_chain(sampler: Sampler, start: np.ndarray, draws: int, tune: int, rng):
    """reference sampling.py:481-521 for one chain.  Returns (trace [T, D], stats {name: [T]})."""
    names = NUTS_STAT_NAMES if sampler.kind == "nuts" else HMC_STAT_NAMES
    T = tune + draws
    q = np.array(start, dtype="d")
    trace = np.zeros((T, sampler.ndim))
    stats = {n: np.zeros(T) for n in names}
    sampler.tune = bool(tune)                                                          # :503
    sampler.reset_tuning()                                                             # :504-505
    for i in range(T):
        if i == 0:
            sampler.iter_count = 0                                                     # :508-509
        if i == tune:
            sampler.tune = False                                                       # :510-511
        q, st = sampler.astep(q, rng)
        trace[i] = q
        for n in names:
            stats[n][i] = st[n]
    return trace, stats


# --------------------------------------------------------------------------------------------------
# the synthetic target densities of BASELINE.json's configs (same arithmetic as the CUDA functors)
# --------------------------------------------------------------------------------------------------
def diag_gaussian(tau: np.ndarray) -> LogpFunc:
    """logp = -1/2 sum tau_i q_i^2 with tau = 1/sigma^2;  g = -(tau*q), logp = 0.5 * q.g"""
    tau = np.asarray(tau, dtype="d")

    def f(q):
        g = -(tau * q)
        return 0.5 * np.dot(q, g), g

    return f


def dense_gaussian(prec: np.ndarray) -> LogpFunc:
    """logp = -1/2 q' P q with a dense precision matrix P;  g = -(P q), logp = 0.5 * q.g"""
    prec = np.asarray(prec, dtype="d")

    def f(q):
        g = -np.dot(prec, q)
        return 0.5 * np.dot(q, g), g

    return f


def neal_funnel(ndim: int, v_scale: float = 3.0) -> LogpFunc:
    """q[0] = v ~ N(0, v_scale^2), q[1:] | v ~ N(0, e^v)  (SURVEY.md section 8d, cfg4).

    logp = -v^2/(2 s^2) - 1/2 e^{-v} S - (n-1)/2 v,  S = sum_{i>=1} q_i^2
    dv   = -v/s^2 + 1/2 e^{-v} S - (n-1)/2 ;  dx_i = -(e^{-v} x_i)
    """
    inv_s2 = 1.0 / (v_scale * v_scale)
    half_nm1 = 0.5 * (ndim - 1)

    def f(q):
        v = q[0]
        x = q[1:]
        S = np.dot(x, x)
        with np.errstate(over="ignore", invalid="ignore"):
            ev = np.exp(-v)
            g = np.empty_like(q)
            g[1:] = -(ev * x)
            hs = 0.5 * ev * S
            g[0] = -(v * inv_s2) + hs - half_nm1
            logp = -(0.5 * v * v * inv_s2) - hs - half_nm1 * v
        return logp, g

    return f


def run_chains(make_f: Callable[[], LogpFunc], ndim: int, kind: str, draws: int, tune: int,
               start: np.ndarray, seeds, *, potential: dict, record: bool = False, **sampler_kw):
    """Run ``len(seeds)`` independent chains the way reference sampling.py:331-399 does (sequentially,
    each reseeding the legacy global stream with its own seed).  Returns stacked arrays; with
    ``record=True`` also the consumed tapes (normals [C,T,D], uniforms [C,T,U], n_uniforms [C,T])."""
    traces, stats_all, tapes = [], [], []
    for s in seeds:
        rng = np.random.RandomState(int(s))
        if record:
            rng = TapeRecorder(rng)
        smp = Sampler(make_f(), ndim, DiagPotential(ndim, **potential), kind=kind, **sampler_kw)
        tr, st = sample_chain(smp, start, draws, tune, rng)
        traces.append(tr)
        stats_all.append(st)
        if record:
            tapes.append(rng)
    trace = np.stack(traces)
    stats = {n: np.stack([s[n] for s in stats_all]) for n in stats_all[0]}
    if not record:
        return trace, stats
    width = max(max(len(u) for u in t.uniforms) for t in tapes)
    parts = [t.tapes(pad_to=max(width, 1)) for t in tapes]
    return trace, stats, tuple(np.stack([p[i] for p in parts]) for i in range(3))
