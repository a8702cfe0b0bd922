"""Device-side chain state and kernel launches (host plumbing over the C ABI; PyTorch tensors own the memory).

One `DeviceChains` holds everything the reference keeps per chain in Python objects -- the current position,
`QuadPotentialDiag(Adapt)`'s arrays and `DualAverageAdaptation`'s scalars -- as row-major [n_chains, ld] / [n_chains, k]
float64 tensors ("each becomes a [C, D] (or [C]) row-major device tensor").
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


_GRAPH_RES = {}   # device -> (side stream, CUDA-graph memory pool) of callback mode
_GRAPH_KEEP = {}  # device -> most recent captured graph (keeps the shared pool in use)
_GRAPH_BRANCH = {}  # device -> extra capture streams (parallel branches of a callback graph)

# launches of this library's kernels enqueued by this process (bench.py reports the count inside its timed region)
LAUNCH_COUNT = {"kernels": 0}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def padded_ld(ndim):
    return int(ndim) + (int(ndim) & 1)


def as_device_rows(x, n_chains, ndim, device):
    """numpy/torch [D] or [C, D] -> float64 device tensor [C, ld] with zero padding."""
    t = torch.as_tensor(np.array(x, dtype="d") if not torch.is_tensor(x) else x, dtype=torch.float64, device=device)
    if t.ndim == 1:
        t = t.unsqueeze(0).expand(n_chains, -1)
    if t.shape != (n_chains, ndim):
        raise ValueError("expected shape (%d, %d) or (%d,), got %s" % (n_chains, ndim, ndim, tuple(t.shape)))
    out = torch.zeros(n_chains, padded_ld(ndim), dtype=torch.float64, device=device)
    out[:, :ndim] = t
    return out


class DeviceChains:
    """State of `n_chains` independent chains on one GPU."""

    def __init__(self, n_chains, ndim, device):
        L.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.LmcError("littlemcmc_b200 runs on CUDA devices only (got %s); there is no CPU fallback" % device)
        self.n_chains, self.ndim, self.ld = int(n_chains), int(ndim), padded_ld(ndim)
        z = lambda *s: torch.zeros(*s, dtype=torch.float64, device=self.device)  # noqa: E731
        self.q = z(self.n_chains, self.ld)
        self.var = z(self.n_chains, self.ld)
        self.mean_fg, self.rawvar_fg = z(self.n_chains, self.ld), z(self.n_chains, self.ld)
        self.mean_bg, self.rawvar_bg = z(self.n_chains, self.ld), z(self.n_chains, self.ld)
        self.adapt = z(self.n_chains, L.ADAPT_STRIDE)
        self.status = torch.zeros(self.n_chains, dtype=torch.int32, device=self.device)
        self._workspace = None

    # -- initialisation from the host-side descriptors (reference reset(): quadpotential.py:195-204, step_sizes.py:49-56)
    def reset_potential(self, var, mean, weight, window):
        C_, D = self.n_chains, self.ndim
        self.var.copy_(as_device_rows(var, C_, D, self.device))
        self.mean_fg.copy_(as_device_rows(mean, C_, D, self.device))
        self.rawvar_fg.copy_(self.var * float(weight))          # raw_var[:] *= w_sum (quadpotential.py:315)
        self.mean_bg.zero_()
        self.rawvar_bg.zero_()
        self.adapt[:, L.ADAPT_W_FG] = float(weight)
        self.adapt[:, L.ADAPT_W_BG] = 0.0
        self.adapt[:, L.ADAPT_NSAMPLES] = 0.0
        self.adapt[:, L.ADAPT_WINDOW] = float(window)

    def reset_step_adapt(self, initial_step):
        ls = float(np.log(initial_step))
        self.adapt[:, L.ADAPT_LOG_STEP] = ls
        self.adapt[:, L.ADAPT_LOG_BAR] = ls
        self.adapt[:, L.ADAPT_HBAR] = 0.0
        self.adapt[:, L.ADAPT_COUNT] = 1.0
        self.adapt[:, L.ADAPT_MU] = float(np.log(10 * initial_step))

    def set_position(self, q):
        self.q.copy_(as_device_rows(q, self.n_chains, self.ndim, self.device))

    def workspace(self, nbytes):
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._workspace

    def rows(self, lo, hi):
        """The chains [lo, hi) as a DeviceChains sharing this one's memory (own workspace)."""
        key = (int(lo), int(hi))
        views = self.__dict__.setdefault("_row_views", {})
        if key not in views:
            v = object.__new__(DeviceChains)
            v.device, v.ndim, v.ld, v.n_chains = self.device, self.ndim, self.ld, key[1] - key[0]
            for name in ("q", "var", "mean_fg", "rawvar_fg", "mean_bg", "rawvar_bg", "adapt", "status"):
                setattr(v, name, getattr(self, name)[key[0]:key[1]])
            v._workspace = None
            views[key] = v
        return views[key]


class FusedTarget:
    """A built-in target density evaluated inside the kernels (include/lmc_b200.h: lmc_target)."""

    def __init__(self, kind, ndim, tau=None, v_scale=3.0):
        self.kind, self.ndim, self.v_scale = int(kind), int(ndim), float(v_scale)
        self.tau_host = None if tau is None else np.ascontiguousarray(tau, dtype=np.float64)
        self._tau_dev = {}

    def c_struct(self, device):
        t = L.Target()
        t.kind, t.v_scale = self.kind, self.v_scale
        if self.kind == L.TARGET_DIAG_GAUSSIAN:
            key = str(device)
            if key not in self._tau_dev:
                buf = torch.zeros(padded_ld(self.ndim), dtype=torch.float64, device=device)
                buf[: self.ndim] = torch.as_tensor(self.tau_host, device=device)
                self._tau_dev[key] = buf
            t.tau = self._tau_dev[key].data_ptr()
        return t


class UserFusedTarget:
    """A target density written by the user as CUDA C++ (protocol: csrc/lmc_device.cuh, "built-in target densities"),
    compiled into the fused sampler kernel at run time (include/lmc_b200.h: lmc_user_kernel_build, NVRTC for sm_100a).
    `source` defines `struct <type_name>` whose only data member is `const double* params`; `params` (host array,
    float64) is uploaded once per device and its address handed to the kernel."""

    def __init__(self, source, type_name, ndim, params=None):
        self.source, self.type_name, self.ndim = str(source), str(type_name), int(ndim)
        self.params_host = np.zeros(2) if params is None else np.ascontiguousarray(params, dtype=np.float64).ravel()
        self._params_dev, self._kernels = {}, {}

    @staticmethod
    def cache_dir():
        import os
        import tempfile
        d = os.environ.get("LMC_USER_CACHE") or os.path.join(os.path.expanduser("~"), ".cache", "littlemcmc_b200")
        try:
            os.makedirs(d, exist_ok=True)
            return d
        except OSError:
            return tempfile.mkdtemp(prefix="littlemcmc_b200_")

    def kernel(self, kind, chunk, tape):
        """Handle of the sampler kernel specialised for (this target, kind, chunk, RNG mode): compiled on first use,
        the cubin cached on disk under a hash of the source, the specialisation and the library's kernel headers."""
        import hashlib
        import os
        key = (int(kind), int(chunk), bool(tape))
        if key in self._kernels:
            return self._kernels[key]
        lib = L.load()
        pkg = os.path.dirname(os.path.abspath(__file__))
        dirs = [os.path.join(pkg, "csrc"), os.path.join(os.path.dirname(pkg), "include")]
        h = hashlib.sha256(repr((self.source, self.type_name, self.ndim, key, L.ABI_VERSION)).encode())
        for d in dirs:
            for f in sorted(os.listdir(d)):
                if f.endswith((".cuh", ".h")):
                    h.update(open(os.path.join(d, f), "rb").read())
        cache = os.path.join(self.cache_dir(), h.hexdigest()[:32] + ".cubin")
        arr = (C.c_char_p * len(dirs))(*[d.encode() for d in dirs])
        handle = C.c_void_p()
        rc = lib.lmc_user_kernel_build(self.source.encode(), self.type_name.encode(), int(kind), self.ndim, int(chunk),
                                       int(bool(tape)), arr, len(dirs), cache.encode(), C.byref(handle))
        if rc != L.OK:
            log = lib.lmc_user_kernel_log().decode(errors="replace")
            raise L.LmcError("compiling the user target %r failed (%s):\n%s"
                             % (self.type_name, L._ERR_NAMES.get(rc, rc), log or lib.lmc_last_error().decode()))
        self._kernels[key] = handle
        return handle

    def target_bytes(self, device):
        key = str(device)
        if key not in self._params_dev:
            self._params_dev[key] = torch.as_tensor(self.params_host, dtype=torch.float64, device=device)
        return C.c_void_p(self._params_dev[key].data_ptr())


def _fill_base(a, kind, chains, *, n_trans, iter0, n_tune, params, seeds, tapes, trace, stats, knobs, stream,
               step_size_override=None):
    """Fill an lmc_sampler_args (everything but target / workspace).  Returns the tensors that must outlive the launch."""
    dev = chains.device
    Cn, D = chains.n_chains, chains.ndim
    a.abi_version, a.n_chains, a.ndim, a.ld = L.ABI_VERSION, Cn, D, chains.ld
    a.q, a.var = chains.q.data_ptr(), chains.var.data_ptr()
    a.adapt_mass, a.adapt_step_size = int(params["adapt_mass"]), int(params["adapt_step_size"])
    a.mean_fg, a.rawvar_fg = chains.mean_fg.data_ptr(), chains.rawvar_fg.data_ptr()
    a.mean_bg, a.rawvar_bg = chains.mean_bg.data_ptr(), chains.rawvar_bg.data_ptr()
    a.adapt = chains.adapt.data_ptr()
    a.window_multiplier = float(params.get("window_multiplier", 1.0))
    a.target_accept, a.gamma = float(params["target_accept"]), float(params["gamma"])
    a.k, a.t0 = float(params["k"]), float(params["t0"])
    a.iter0, a.n_tune, a.n_trans = int(iter0), int(n_tune), int(n_trans)
    a.Emax = float(params["Emax"])
    a.max_treedepth = int(params.get("max_treedepth", 10))
    a.early_max_treedepth = int(params.get("early_max_treedepth", 8))
    a.path_length, a.max_steps = float(params.get("path_length", 2.0)), int(params.get("max_steps", 1024))
    keep = []
    if tapes is not None:
        normals, uniforms = tapes
        normals = torch.as_tensor(normals, dtype=torch.float64, device=dev).contiguous()
        uniforms = torch.as_tensor(uniforms, dtype=torch.float64, device=dev).contiguous()
        assert normals.shape == (Cn, n_trans, D), normals.shape
        assert uniforms.shape[:2] == (Cn, n_trans), uniforms.shape
        a.rng.mode, a.rng.normals, a.rng.uniforms = L.RNG_TAPE, normals.data_ptr(), uniforms.data_ptr()
        a.rng.u_stride = uniforms.shape[2]
        keep += [normals, uniforms]
    else:
        if seeds is None:
            raise ValueError("either per-chain seeds or tapes are required")
        a.rng.mode, a.rng.seeds = L.RNG_PHILOX, seeds.data_ptr()
        keep.append(seeds)
    assert trace.stride(2) == 1
    a.trace, a.trace_chain_stride, a.trace_draw_stride = trace.data_ptr(), trace.stride(0), trace.stride(1)
    a.stats, a.status = stats.data_ptr(), chains.status.data_ptr()
    if step_size_override is not None:   # BaseHMC.step_rand result for this call's transitions (base_hmc.py:154-155)
        step_size_override = torch.as_tensor(step_size_override, dtype=torch.float64, device=dev).contiguous()
        assert step_size_override.shape == (Cn,), step_size_override.shape
        a.step_size_override = step_size_override.data_ptr()
        keep.append(step_size_override)
    knobs = knobs or {}
    a.tune_group = int(knobs.get("group", 0))
    a.tune_smem_vecs = int(knobs.get("smem_vecs", -1))
    a.tune_max_slots = int(knobs.get("max_slots", 0))
    a.tune_chunk = int(knobs.get("chunk", 0))
    a.stream = (stream or torch.cuda.current_stream(dev)).cuda_stream
    return keep


def _alloc_outputs(chains, n_trans, trace, stats):
    dev, Cn, D = chains.device, chains.n_chains, chains.ndim
    if trace is None:
        trace = torch.empty(Cn, n_trans, D, dtype=torch.float64, device=dev)
    if stats is None:
        stats = torch.empty(Cn, n_trans, L.NSTATS, dtype=torch.float64, device=dev)
    return trace, stats


def run_transitions(kind, chains, target, *, n_trans, iter0, n_tune, params, seeds=None, tapes=None,
                    trace=None, stats=None, knobs=None, stream=None, events=None, step_size_override=None,
                    trace_skip=0, progress=None, progress_block=0):
    """Enqueue `n_trans` transitions of every chain (lmc_nuts_sample / lmc_hmc_sample).  Returns (trace, stats)
    device tensors [C, n_trans, D] and [C, n_trans, NSTATS].  Asynchronous on the current CUDA stream.
    `events`: optional pair of torch.cuda.Event recorded on the launching stream immediately around the library call
    (bench.py times the launch with them, so host-side argument marshalling is not inside the bracket).
    `trace_skip` / `progress` / `progress_block`: lmc_sampler_args of the same names (one launch for a whole run: the
    first `trace_skip` transitions are not kept, `trace` holds the rest, and the int32 device counters `progress` tell the
    caller which blocks of kept draws are final)."""
    lib = L.load()
    dev = chains.device
    Cn, D = chains.n_chains, chains.ndim
    trace, stats = _alloc_outputs(chains, n_trans, trace, stats)
    if n_trans == 0 or Cn == 0:      # nothing to do (empty tensors have no device pointer to hand to the library)
        return trace, stats
    a = L.SamplerArgs()
    with torch.cuda.device(dev):
        keep = _fill_base(a, kind, chains, n_trans=n_trans, iter0=iter0, n_tune=n_tune, params=params, seeds=seeds,
                          tapes=tapes, trace=trace, stats=stats, knobs=knobs, stream=stream,
                          step_size_override=step_size_override)
        a.trace_skip, a.progress_block = int(trace_skip), int(progress_block)
        if progress is not None:
            assert progress.dtype == torch.int32 and progress.is_cuda
            a.progress = progress.data_ptr()
            keep.append(progress)
        user = isinstance(target, UserFusedTarget)
        if user:
            a.tune_group = 0                  # the kernel a user target is compiled into is the library's default choice
            handle = target.kernel(kind, a.tune_chunk, tapes is not None)
            tptr = target.target_bytes(dev)   # struct { const double* params; } passed by value
        else:
            a.target = target.c_struct(dev)
        nbytes = lib.lmc_workspace_bytes(kind, Cn, D, max(a.max_treedepth, a.early_max_treedepth), a.tune_group)
        if nbytes < 0:
            L.check(int(nbytes), "lmc_workspace_bytes")
        ws = chains.workspace(nbytes)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        fn = lib.lmc_nuts_sample if kind == L.KIND_NUTS else lib.lmc_hmc_sample
        launch_stream = stream or torch.cuda.current_stream(dev)
        if events is not None:
            events[0].record(launch_stream)
        rc = lib.lmc_user_sample(handle, C.byref(a), C.byref(tptr)) if user else fn(C.byref(a))
        if events is not None:
            events[1].record(launch_stream)
        L.check(rc, "lmc_user_sample" if user else ("lmc_nuts_sample" if kind == L.KIND_NUTS else "lmc_hmc_sample"))
        LAUNCH_COUNT["kernels"] += 2      # sched_init_kernel + sampler_kernel
    for t in keep:  # tensors referenced by the enqueued kernel must outlive it on this stream
        t.record_stream(torch.cuda.current_stream(dev)) if t.is_cuda else None
    return trace, stats


# ---- callback mode -----------------------------------------------------------------------------------------------------
def evaluate_callback(f, q):
    """Evaluate a user callback for every chain.  `q`: device tensor [C, D] float64.  -> (logp [C], grad [C, D]) on
    the device.  A `targets.TorchBatched` callable gets the whole batch (one torch op); any other callable is the
    reference's per-chain NumPy contract `f(q[D]) -> (logp, dlogp[D])` (base_hmc.py:34) and is looped over on the host."""
    from .targets import TorchBatched
    Cn, D = q.shape
    if isinstance(f, TorchBatched):
        logp, grad = f(q)
        if logp.shape != (Cn,) and logp.numel() == Cn:
            logp = logp.reshape(Cn)
        if logp.shape != (Cn,) or grad.shape != (Cn, D):
            raise ValueError("batched callback must return (logp[%d], grad[%d, %d]); got %s, %s"
                             % (Cn, Cn, D, tuple(logp.shape), tuple(grad.shape)))
        return logp.to(torch.float64), grad.to(torch.float64)
    qh = q.cpu().numpy()
    logps, gh = np.empty(Cn), np.empty((Cn, D))
    for c in range(Cn):
        lp, gr = f(qh[c])
        logps[c] = float(np.asarray(lp, dtype="d").reshape(-1)[0])   # 0-d or shape-(1,) logp (tests/test_utils.py:19-28)
        gh[c] = np.asarray(gr, dtype="d").reshape(D)
    return torch.as_tensor(logps, device=q.device), torch.as_tensor(gh, device=q.device)


class CallbackRun:
    """One callback-mode run: `n_trans` transitions of every chain driven by lmc_callback_begin / lmc_callback_advance
    around a user gradient callback.  The loop body (callback + advance) can be captured in a CUDA graph."""

    def __init__(self, kind, chains, callback, *, n_trans, iter0, n_tune, params, seeds=None, tapes=None, trace=None,
                 stats=None, stream=None, step_size_override=None, n_running=None):
        self.lib = L.load()
        self.kind, self.chains, self.callback = kind, chains, callback
        dev, Cn, D = chains.device, chains.n_chains, chains.ndim
        self.trace, self.stats = _alloc_outputs(chains, n_trans, trace, stats)
        self.q_eval = torch.zeros(Cn, chains.ld, dtype=torch.float64, device=dev)
        self.g_eval = torch.zeros(Cn, chains.ld, dtype=torch.float64, device=dev)
        self.logp_eval = torch.zeros(Cn, dtype=torch.float64, device=dev)
        # chains that still need gradient evaluations; several runs over disjoint chains may share one counter
        self.n_running = torch.zeros(1, dtype=torch.int32, device=dev) if n_running is None else n_running
        self.n_running_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.c = L.CallbackArgs()
        with torch.cuda.device(dev):
            self.keep = _fill_base(self.c.base, kind, chains, n_trans=n_trans, iter0=iter0, n_tune=n_tune, params=params,
                                   seeds=seeds, tapes=tapes, trace=self.trace, stats=self.stats, knobs=None, stream=stream,
                                   step_size_override=step_size_override)
            nbytes = self.lib.lmc_callback_state_bytes(kind, Cn, D, max(self.c.base.max_treedepth,
                                                                               self.c.base.early_max_treedepth))
            if nbytes < 0:
                L.check(int(nbytes), "lmc_callback_state_bytes")
            self.machine = chains.workspace(nbytes)
        c = self.c
        c.q_eval, c.g_eval, c.logp_eval = self.q_eval.data_ptr(), self.g_eval.data_ptr(), self.logp_eval.data_ptr()
        c.machine, c.machine_bytes, c.n_running = self.machine.data_ptr(), self.machine.numel(), self.n_running.data_ptr()
        per = (1 << max(c.base.max_treedepth, c.base.early_max_treedepth)) + 1 if kind == L.KIND_NUTS else c.base.max_steps + 1
        self.max_iters = int(n_trans) * per + 1
        self.n_evals = 0

    def _stream_ptr(self):
        return torch.cuda.current_stream(self.chains.device).cuda_stream

    def begin(self):
        self.c.base.stream = self._stream_ptr()
        L.check(self.lib.lmc_callback_begin(self.kind, C.byref(self.c)), "lmc_callback_begin")
        LAUNCH_COUNT["kernels"] += 1

    def iteration(self):
        """callback at q_eval, then advance every chain (one gradient evaluation per chain)."""
        D = self.chains.ndim
        logp, grad = evaluate_callback(self.callback, self.q_eval[:, :D])
        c = self.c
        # hand the callback's own output buffers to the kernel when their layout allows it (rows of D = ld doubles,
        # 16-byte aligned): saves two copy kernels per gradient evaluation in this launch-latency-bound mode
        if (D == self.chains.ld and grad.is_contiguous() and grad.dtype == torch.float64 and grad.data_ptr() % 16 == 0
                and grad.device == self.q_eval.device):
            c.g_eval = grad.data_ptr()
        else:
            self.g_eval[:, :D].copy_(grad)
            c.g_eval = self.g_eval.data_ptr()
        if logp.is_contiguous() and logp.dtype == torch.float64 and logp.device == self.q_eval.device:
            c.logp_eval = logp.data_ptr()
        else:
            self.logp_eval.copy_(logp)
            c.logp_eval = self.logp_eval.data_ptr()
        self._held = (logp, grad)           # keep the buffers alive until the next evaluation replaces them
        self.c.base.stream = self._stream_ptr()
        L.check(self.lib.lmc_callback_advance(self.kind, C.byref(self.c)), "lmc_callback_advance")
        self.n_evals += 1
        LAUNCH_COUNT["kernels"] += 1

    def run(self, cuda_graph=False, iters_per_graph=None, poll=4):
        """`cuda_graph`: False = eager host loop; True / "device" = ONE graph launch whose WHILE conditional node loops
        on the device until every chain has finished (lmc_callback_loop_*); "replay" = graphs of `iters_per_graph`
        iterations replayed from the host with one synchronisation per replay."""
        dev = self.chains.device
        with torch.cuda.device(dev):
            self.begin()
            if cuda_graph in (True, "device"):
                self._run_device_loop(iters_per_graph or 8)
            elif cuda_graph == "replay":
                self._run_graphed(iters_per_graph or 8)
            elif cuda_graph:
                raise ValueError("cuda_graph must be False, True, 'device' or 'replay'")
            else:
                ev, it = None, 0
                while it < self.max_iters:
                    self.iteration()
                    it += 1
                    if it % poll == 0:
                        # non-blocking poll: look at the counter copied `poll` iterations ago
                        if ev is not None and ev.query() and int(self.n_running_host[0]) == 0:
                            break
                        if ev is None or ev.query():
                            self.n_running_host.copy_(self.n_running, non_blocking=True)
                            ev = torch.cuda.Event()
                            ev.record()
                if int(self.n_running.item()) != 0:
                    raise L.LmcError("callback mode: chains still running after %d gradient evaluations" % it)
        return self.trace, self.stats

    def _capture_iterations(self, n_iters, keep_graph):
        """Capture `n_iters` x (callback + advance) on torch's capture stream.  -> torch.cuda.CUDAGraph"""
        return _capture_runs([self], n_iters, keep_graph)

    def _run_device_loop(self, iters_per_body):
        """The whole run as one graph launch: WHILE(n_running > 0) { callback; advance; } on the device."""
        _device_loop([self], iters_per_body)

    def _run_graphed(self, iters_per_graph):
        graph = self._capture_iterations(iters_per_graph, keep_graph=False)
        it = 0
        while it < self.max_iters:
            graph.replay()
            it += iters_per_graph
            self.n_evals += iters_per_graph
            LAUNCH_COUNT["kernels"] += iters_per_graph
            if int(self.n_running.item()) == 0:   # one sync per replay (iters_per_graph gradient evaluations)
                return
        raise L.LmcError("callback mode: chains still running after %d gradient evaluations" % it)


def _capture_runs(runs, n_iters, keep_graph):
    """Capture `n_iters` x (callback + advance) of every run, the runs on PARALLEL branches of one graph (forked streams):
    the kernels of callback mode are tiny, so independent batches of chains overlap on the GPU."""
    dev = runs[0].chains.device
    # one side stream (+ branch streams) and ONE graph memory pool per device, shared by every capture of this process:
    # a fresh pool per graph means cudaMalloc at capture and a synchronising cudaFree when the graph dies, every call
    key = str(dev)
    if key not in _GRAPH_RES:
        _GRAPH_RES[key] = (torch.cuda.Stream(device=dev), torch.cuda.graph_pool_handle())
    side, pool = _GRAPH_RES[key]
    branches = _GRAPH_BRANCH.setdefault(key, [])
    while len(branches) < len(runs) - 1:
        branches.append(torch.cuda.Stream(device=dev))
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):       # warm the callback up outside capture (lazy init, autotune, allocations)
        for r in runs:
            evaluate_callback(r.callback, r.q_eval[:, :r.chains.ndim])
    torch.cuda.current_stream(dev).wait_stream(side)
    graph = torch.cuda.CUDAGraph(keep_graph=True) if keep_graph else torch.cuda.CUDAGraph()
    before = [r.n_evals for r in runs]
    # capture_begin / capture_end directly: the torch.cuda.graph context manager also runs gc.collect(),
    # empty_cache() and a device synchronisation on entry (tens of milliseconds per run, measured)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        graph.capture_begin(pool=pool)
        try:
            for k, r in enumerate(runs):
                st = side if k == 0 else branches[k - 1]
                if k:
                    st.wait_stream(side)                     # fork
                with torch.cuda.stream(st):
                    for _ in range(n_iters):
                        r.iteration()
            for k in range(1, len(runs)):
                side.wait_stream(branches[k - 1])            # join
        finally:
            graph.capture_end()
    torch.cuda.current_stream(dev).wait_stream(side)
    for r, b in zip(runs, before):
        r.n_evals = b
    # the shared pool must never lose its last user between two captures (the allocator retires a pool whose use
    # count drops to zero and asserts if it is handed out again): keep the newest graph of this device alive
    _GRAPH_KEEP[key] = graph
    return graph


def _device_loop(runs, iters_per_body):
    """All runs (disjoint chains, one shared n_running counter) as ONE graph launch: a WHILE conditional node whose body
    is `iters_per_body` x (callback; advance) per run on parallel branches, looping on the device until every chain of
    every run has finished (include/lmc_b200.h: lmc_callback_loop_*)."""
    lib, dev = runs[0].lib, runs[0].chains.device
    n_running = runs[0].n_running
    assert all(r.n_running is n_running for r in runs)
    graph = _capture_runs(runs, iters_per_body, keep_graph=True)
    iters = torch.zeros(1, dtype=torch.int32, device=dev)
    loop = C.c_void_p()
    max_bodies = (max(r.max_iters for r in runs) + iters_per_body - 1) // iters_per_body
    L.check(lib.lmc_callback_loop_create(C.c_void_p(int(graph.raw_cuda_graph())), _ptr(n_running), _ptr(iters), max_bodies,
                                         C.byref(loop)), "lmc_callback_loop_create")
    try:
        stream = torch.cuda.current_stream(dev)
        L.check(lib.lmc_callback_loop_launch(loop, C.c_void_p(stream.cuda_stream)), "lmc_callback_loop_launch")
        left, done = int(n_running.item()), int(iters.item())       # the one synchronisation of the run
    finally:
        torch.cuda.synchronize(dev)
        lib.lmc_callback_loop_destroy(loop)
    for r in runs:
        r.n_evals += done * iters_per_body
    LAUNCH_COUNT["kernels"] += done * (iters_per_body * len(runs) + 1)   # advance kernels + the loop-condition kernel
    if left != 0:
        raise L.LmcError("callback mode: chains still running after %d gradient evaluations" % (done * iters_per_body))


def run_transitions_callback(kind, chains, callback, *, n_trans, iter0, n_tune, params, seeds=None, tapes=None,
                             trace=None, stats=None, cuda_graph=False, step_size_override=None, split=None):
    """Callback-mode counterpart of run_transitions (synchronous: returns when every chain has finished).
    `split` (device-driven loop only, default 1): number of batches the chains are cut into, each with its own callback
    evaluation and advance kernel on a parallel branch of the graph.  Measured SLOWER than one batch on B200 (1024 x 100:
    4.4e7 -> 2.8e7 leapfrog/s with 2 batches; 8192 x 50: 2.2e7 -> 1.2e7 with 4): the nodes of a WHILE body execute one
    after the other whatever the graph's shape, so the cost is per node, and splitting doubles the nodes.  Kept as an
    option for callbacks whose kernels are large enough to fill the GPU only together."""
    from .targets import TorchBatched
    Cn = chains.n_chains
    if cuda_graph in (True, "device") and isinstance(callback, TorchBatched) and n_trans > 0 and Cn > 0:
        K = int(split) if split else 1
        K = max(1, min(K, Cn))
        if K > 1:
            dev = chains.device
            trace, stats = _alloc_outputs(chains, n_trans, trace, stats)
            n_running = torch.zeros(1, dtype=torch.int32, device=dev)
            bounds = np.linspace(0, Cn, K + 1).astype(int)
            cut = lambda x, lo, hi: None if x is None else x[lo:hi]    # noqa: E731
            if tapes is not None:
                tapes = tuple(torch.as_tensor(t, dtype=torch.float64, device=dev) for t in tapes)
            if step_size_override is not None:
                step_size_override = torch.as_tensor(step_size_override, dtype=torch.float64, device=dev)
            runs = []
            for lo, hi in zip(bounds[:-1], bounds[1:]):
                runs.append(CallbackRun(kind, chains.rows(lo, hi), callback, n_trans=n_trans, iter0=iter0, n_tune=n_tune,
                                        params=params, seeds=cut(seeds, lo, hi),
                                        tapes=None if tapes is None else (tapes[0][lo:hi], tapes[1][lo:hi]),
                                        trace=trace[lo:hi], stats=stats[lo:hi],
                                        step_size_override=cut(step_size_override, lo, hi), n_running=n_running))
            with torch.cuda.device(dev):
                for r in runs:
                    r.begin()
                n_running.fill_(Cn)          # every begin wrote its own batch size: the shared counter is their sum
                _device_loop(runs, 8)
            return trace, stats
    run = CallbackRun(kind, chains, callback, n_trans=n_trans, iter0=iter0, n_tune=n_tune, params=params, seeds=seeds,
                      tapes=tapes, trace=trace, stats=stats, step_size_override=step_size_override)
    run.run(cuda_graph=cuda_graph)
    return run.trace, run.stats


# ---- dense-mass mode ---------------------------------------------------------------------------------------------------
class DenseRun:
    """`n_trans` transitions of every chain with a DENSE potential (quadpotential_dense.py), driven by lmc_dense_begin /
    lmc_dense_advance: after every advance each chain says which batched result it waits for (gradient, velocity,
    momentum draw, mass-matrix update) and the host computes exactly those, for exactly those chains."""

    def __init__(self, kind, chains, callback, potential, *, n_trans, iter0, n_tune, params, seeds=None, tapes=None,
                 trace=None, stats=None, stream=None, step_size_override=None):
        self.lib = L.load()
        self.kind, self.chains, self.callback, self.potential = kind, chains, callback, potential
        dev, Cn = chains.device, chains.n_chains
        ld = chains.ld
        self.trace, self.stats = _alloc_outputs(chains, n_trans, trace, stats)
        z = lambda *s: torch.zeros(*s, dtype=torch.float64, device=dev)  # noqa: E731
        self.q_eval, self.g_eval, self.logp_eval = z(Cn, ld), z(Cn, ld), z(Cn)
        self.x_eval, self.v_eval = z(Cn, 2, ld), z(Cn, 2, ld)
        self.n_eval, self.p0_eval = z(Cn, ld), z(Cn, ld)
        self.need = torch.zeros(Cn, dtype=torch.int32, device=dev)
        self.need_host = torch.zeros(Cn, dtype=torch.int32).pin_memory()
        self._need_stage = torch.zeros(Cn, dtype=torch.int32).pin_memory()   # need + HOLD marks on their way back
        # index lists travel through pinned staging (a pageable H2D copy blocks the host for tens of microseconds)
        self._idx_pinned = [torch.zeros(Cn, dtype=torch.int64).pin_memory() for _ in range(4)]
        self._idx_dev = [torch.zeros(Cn, dtype=torch.int64, device=dev) for _ in range(4)]
        self._idx_slot = 0
        self.n_running = torch.zeros(1, dtype=torch.int32, device=dev)
        self.c = L.DenseArgs()
        with torch.cuda.device(dev):
            self.keep = _fill_base(self.c.base, kind, chains, n_trans=n_trans, iter0=iter0, n_tune=n_tune, params=params,
                                   seeds=seeds, tapes=tapes, trace=self.trace, stats=self.stats, knobs=None, stream=stream,
                                   step_size_override=step_size_override)
            self.c.base.adapt_mass = int(bool(getattr(potential, "_adaptive", False)))
            nbytes = self.lib.lmc_dense_state_bytes(kind, Cn, chains.ndim, max(self.c.base.max_treedepth,
                                                                                   self.c.base.early_max_treedepth))
            if nbytes < 0:
                L.check(int(nbytes), "lmc_dense_state_bytes")
            self.machine = chains.workspace(nbytes)
        c = self.c
        c.q_eval, c.g_eval, c.logp_eval = self.q_eval.data_ptr(), self.g_eval.data_ptr(), self.logp_eval.data_ptr()
        c.x_eval, c.v_eval = self.x_eval.data_ptr(), self.v_eval.data_ptr()
        c.n_eval, c.p0_eval, c.need = self.n_eval.data_ptr(), self.p0_eval.data_ptr(), self.need.data_ptr()
        c.machine, c.machine_bytes, c.n_running = self.machine.data_ptr(), self.machine.numel(), self.n_running.data_ptr()
        per = (1 << max(c.base.max_treedepth, c.base.early_max_treedepth)) + 2 if kind == L.KIND_NUTS else c.base.max_steps + 2
        self.max_iters = 2 * int(n_trans) * per + 4
        self.n_grad_evals = self.n_vel_evals = 0

    def _subset(self, mask_host, all_above=1.0):
        """-> None (= every chain) when at least the fraction `all_above` of the chains is selected, else an int64
        device index tensor.  Potentials whose operations are one library GEMM over a matrix shared by all chains
        set `_all_rows_above` < 1: serving rows nobody asked for is cheaper than gathering / scattering the rest
        (a chain only reads a result buffer in the phase in which it asked for it)."""
        if mask_host.all() or mask_host.mean() >= all_above:
            return None
        idx = np.nonzero(mask_host)[0]
        k = self._idx_slot
        self._idx_slot = (k + 1) % 4        # four lists per iteration at most; stream order protects their reuse
        self._idx_pinned[k].numpy()[:idx.size] = idx
        out = self._idx_dev[k][:idx.size]
        out.copy_(self._idx_pinned[k][:idx.size], non_blocking=True)
        return out

    def run(self):
        dev, D = self.chains.device, self.chains.ndim
        pot = self.potential
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            self.c.base.stream = stream.cuda_stream
            L.check(self.lib.lmc_dense_begin(self.kind, C.byref(self.c)), "lmc_dense_begin")
            LAUNCH_COUNT["kernels"] += 1
            for _ in range(self.max_iters):
                self.need_host.copy_(self.need, non_blocking=True)
                stream.synchronize()                      # the host must know who needs what: one sync per iteration
                need = self.need_host.numpy()
                if not need.any():
                    break
                m_upd = (need & L.NEED_UPDATE) != 0
                if m_upd.any():                           # potential.update first: the momentum draw uses the new matrix
                    # An adapted dense matrix is refactored after every tuning sample (quadpotential.py:520-526), and a
                    # batched Cholesky costs the same ~10 ms for 8 matrices as for 256: chains that ask for their update
                    # are HELD (LMC_NEED_HOLD: the kernel leaves them where they are) until a quarter of the chains
                    # still running wait for one, or nobody is left in the middle of a trajectory.
                    busy = ((need & (L.NEED_GRAD | L.NEED_VEL)) != 0) & ~m_upd
                    batch = float(getattr(pot, "_update_batch_fraction", 0.0))
                    if busy.any() and m_upd.sum() < batch * ((need & ~L.NEED_HOLD) != 0).sum():
                        if (need[m_upd] & L.NEED_HOLD).all():
                            pass                          # all of them are marked already
                        else:
                            self._need_stage.numpy()[:] = need
                            self._need_stage.numpy()[m_upd] |= L.NEED_HOLD
                            self.need.copy_(self._need_stage, non_blocking=True)
                        need = need & ~np.where(m_upd, L.NEED_GRAD | L.NEED_MOM | L.NEED_UPDATE, 0).astype(need.dtype)
                    else:
                        pot._update_rows(torch.as_tensor(np.nonzero(m_upd)[0], device=dev), self.chains.q)
                        if (need[m_upd] & L.NEED_HOLD).any():
                            self._need_stage.numpy()[:] = need & ~np.where(m_upd, L.NEED_HOLD, 0).astype(need.dtype)
                            self.need.copy_(self._need_stage, non_blocking=True)
                m_mom = (need & L.NEED_MOM) != 0
                m_grad, m_vel = (need & L.NEED_GRAD) != 0, (need & L.NEED_VEL) != 0
                self._idx_slot = 0
                lib_all = float(getattr(pot, "_all_rows_above", 1.0))
                if m_mom.any():
                    pot._momentum_rows(self._subset(m_mom, lib_all), self.n_eval, self.p0_eval)
                if m_grad.any():
                    idx = self._subset(m_grad)
                    qs = self.q_eval[:, :D] if idx is None else self.q_eval[idx][:, :D]
                    logp, grad = evaluate_callback(self.callback, qs)
                    if idx is None:
                        self.g_eval[:, :D] = grad
                        self.logp_eval.copy_(logp)
                    else:
                        self.g_eval[idx, :D] = grad
                        self.logp_eval[idx] = logp
                    self.n_grad_evals += 1
                if m_vel.any():
                    pot._velocity_rows(self._subset(m_vel, lib_all), self.x_eval, self.v_eval)
                    self.n_vel_evals += 1
                self.c.base.stream = stream.cuda_stream
                L.check(self.lib.lmc_dense_advance(self.kind, C.byref(self.c)), "lmc_dense_advance")
                LAUNCH_COUNT["kernels"] += 1
            else:
                raise L.LmcError("dense mode: chains still running after %d iterations" % self.max_iters)
        return self.trace, self.stats


def run_transitions_dense(kind, chains, callback, potential, **kw):
    run = DenseRun(kind, chains, callback, potential, **kw)
    run.run()
    return run.trace, run.stats


def rng_fill(seeds, ndim, iter0, n_trans, u_stride):
    """The numbers PHILOX mode consumes, as tapes (lmc_rng_fill)."""
    lib = L.load()
    dev = seeds.device
    Cn = seeds.shape[0]
    normals = torch.empty(Cn, n_trans, ndim, dtype=torch.float64, device=dev)
    uniforms = torch.empty(Cn, n_trans, u_stride, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.lmc_rng_fill(_ptr(seeds), Cn, ndim, int(iter0), int(n_trans), int(u_stride), _ptr(normals),
                                 _ptr(uniforms), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "lmc_rng_fill")
    return normals, uniforms


def seeds_tensor(seeds, device):
    """Per-chain integer seeds -> uint64 keys on the device (stored as int64 bit patterns)."""
    arr = np.asarray(seeds, dtype=np.uint64).astype(np.int64)
    return torch.as_tensor(arr, device=device)
