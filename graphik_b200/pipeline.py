"""Pipelined batches with deferred stragglers: the throughput mode of the engine.

`BatchIK.solve` / `RiemannianSolver.solve_batch` return when the slowest goal of the batch is
done -- for UR10 that is one of the ~0.4 % of goals that run into maxiter = 3000 outer iterations
(~240 k tCG iterations against a median of 5 k), so a 4096-goal batch holds a B200 for ~170 ms of
which ~30 ms are work.  `IKStream` keeps the device busy instead:

  * batches go round-robin to a few *slots* (CUDA stream + work counter + carry queues), so the
    small kernels of one batch (goal distances, bound smoothing + initialisation, joint recovery)
    overlap the trust-region launch of another;
  * the trust-region launch is `gik_rtr_solve_sliced`: a goal that has spent `inner_budget` tCG
    iterations in a launch parks in the slot's carry queue and the slot's NEXT launch resumes it
    before it starts new goals, writing the final values into the tensors of the batch it came
    from.  A launch therefore lasts about as long as its work, not as long as its slowest goal, and
    the stragglers of earlier batches run in the shadow of later ones (longest-first is also the
    makespan-optimal order).  A parked goal follows bit for bit the trajectory of an unparked one.

A batch is complete when its device-side `pending` counter is back to zero; `result(ticket)`
waits for that (draining the slot if nothing else is going to), re-runs the joint recovery for
batches that had late arrivals, and returns the same dict as `BatchIK.solve`.

The reference has nothing comparable (one pose per call, riemannian_solver.py:220-234); this is
host-side scheduling around the same per-goal algorithm.
"""
import ctypes

from graphik_b200 import _lib
from graphik_b200.engine import BatchIK, _p

STATUS_PENDING = 4


class Ticket:
    """One submitted batch.  `out` holds the device tensors (final once `done`)."""

    def __init__(self, index, slot, launch, B):
        self.index, self.slot, self.launch, self.B = index, slot, launch, B
        self.out = None
        self.pending_at = None     # position of this batch's counter in the slot's counter block
        self.done = False
        self.late = None           # True once known: some goals finished in a later launch
        self.host = None           # pinned host copies when the stream was asked for them


class _Slot:
    def __init__(self, eng, capacity, max_tickets):
        torch = eng.torch
        self.stream = torch.cuda.Stream(device=eng.device)
        self.counter = torch.zeros(1, dtype=torch.int32, device=eng.device)
        nbytes = int(eng.lib.gik_carry_bytes(eng.plan.handle, capacity))
        self.carry = [torch.empty((nbytes + 15) // 16 * 2, dtype=torch.int64, device=eng.device) for _ in range(2)]
        self.flip = 0
        self.workspace = (torch.empty(eng._ws_bytes // 8, dtype=torch.float64, device=eng.device)
                          if eng._ws_bytes else None)
        # one pending counter per outstanding batch of this slot, snapshotted to pinned memory after every launch
        self.pending = torch.zeros(max_tickets, dtype=torch.int32, device=eng.device)
        self.snap = [torch.zeros(max_tickets, dtype=torch.int32).pin_memory() for _ in range(2)]
        self.snap_event = [None, None]
        self.snap_launch = [-1, -1]
        self.free_counters = list(range(max_tickets))
        self.launches = 0
        self.tickets = []          # outstanding (not yet finalised) batches, oldest first
        self.host_ring, self.host_next = {}, {}


class IKStream:
    """See the module docstring.  Typical use:

        stream = RiemannianSolver(graph).stream()
        tickets = [stream.submit(T) for T in batches]     # asynchronous
        results = [stream.result(t) for t in tickets]     # or: stream.drain()
    """

    HOST_RING = 8

    def __init__(self, engine: BatchIK, slots=1, inner_budget=None, carry_capacity=None, to_host=False,
                 max_outstanding=64, record_events=False):
        self.eng = engine
        self.torch = engine.torch
        self.lib = engine.lib
        if isinstance(engine.opts, _lib.CgOpts):
            raise _lib.GikError("IKStream drives the sliced trust-region solve; solver='ConjugateGradient' runs through "
                                "solve_batch / solve_with_riemannian")
        N = engine.plan.N
        # tCG iterations a goal must have spent in a launch before it may park once the launch has no new goal left
        # (measured on UR10, 20 x 4096 goals, one slot: 144 k solves/s at 512, 140 k at 1024, 136 k at 2048, 119 k at 4096)
        self.inner_budget = int(inner_budget) if inner_budget is not None else 512
        self.capacity = int(carry_capacity) if carry_capacity is not None else 16384
        self.to_host = bool(to_host)
        self.slots = [_Slot(engine, self.capacity, max_outstanding) for _ in range(max(1, int(slots)))]
        with self.torch.cuda.device(engine.device):
            for sl in self.slots:
                for c in sl.carry:
                    with self.torch.cuda.stream(sl.stream):
                        _lib.check(self.lib.gik_carry_init(engine.plan.handle, _p(c), self.capacity,
                                                           ctypes.c_void_p(sl.stream.cuda_stream)), "gik_carry_init")
        self.n_submitted = 0
        self.launches = 0
        self.record_events = bool(record_events)
        self.launch_events = []    # (start, end) CUDA events around every trust-region launch when record_events

    # ------------------------------------------------------------------ internals
    def _launch(self, sl, g2, Y0, out, pending_ptr, budget):
        """One gik_rtr_solve_sliced on slot `sl` (current stream must be sl.stream)."""
        eng = self.eng
        B = 0 if Y0 is None else Y0.shape[0]
        cin, cout = sl.carry[sl.flip], sl.carry[1 - sl.flip]
        if self.record_events:
            e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            e0.record(sl.stream)
        _lib.check(self.lib.gik_rtr_solve_sliced(
            eng.plan.handle, _p(g2), _p(Y0), B, ctypes.byref(eng.opts),
            _p(out["x"]) if out else None, _p(out["f(x)"]) if out else None, _p(out["gradnorm"]) if out else None,
            _p(out["iterations"]) if out else None, _p(out["status"]) if out else None,
            _p(out["n_inner"]) if out else None, int(budget), _p(cin), _p(cout), pending_ptr,
            _p(sl.counter), ctypes.c_void_p(sl.stream.cuda_stream)), "gik_rtr_solve_sliced")
        if self.record_events:
            e1.record(sl.stream)
            self.launch_events.append((e0, e1))
        sl.flip ^= 1
        sl.launches += 1
        self.launches += 3      # counter memset + queue reset + persistent kernel
        eng.launches += 3
        # snapshot of the slot's pending counters, readable by the host once the event has passed
        k = sl.launches & 1
        sl.snap[k].copy_(sl.pending, non_blocking=True)
        ev = self.torch.cuda.Event()
        ev.record(sl.stream)
        sl.snap_event[k], sl.snap_launch[k] = ev, sl.launches

    def _finalise(self, tk):
        """All goals of the batch are final: recover joints again if some arrived late, copy to the host."""
        sl = self.slots[tk.slot]
        eng = self.eng
        with self.torch.cuda.stream(sl.stream):
            if tk.late:
                tk.out["q"] = eng.joints(tk.out["x"], tk.out["T_goal"])
            if self.to_host:
                tk.host = self._host_buffers(sl, tk)
                for k, h in tk.host.items():
                    h.copy_(tk.out[k], non_blocking=True)
            tk.ready = self.torch.cuda.Event()
            tk.ready.record(sl.stream)
        tk.done = True
        sl.free_counters.append(tk.pending_at)
        sl.tickets.remove(tk)

    def _host_buffers(self, sl, tk):
        """Pinned host copies of (q, f(x), status) from a small per-slot ring: they stay valid until HOST_RING more
        batches of the same size have been finalised on the slot (pinning memory costs milliseconds)."""
        ring = sl.host_ring.setdefault(tk.B, [])
        if len(ring) < self.HOST_RING:
            ring.append({k: self.torch.empty(tk.out[k].shape, dtype=tk.out[k].dtype).pin_memory()
                         for k in ("q", "f(x)", "status")})
            return ring[-1]
        sl.host_next[tk.B] = (sl.host_next.get(tk.B, -1) + 1) % self.HOST_RING
        return ring[sl.host_next[tk.B]]

    def poll(self):
        """Non-blocking: finalise every batch whose pending counter is known to be back at zero."""
        for sl in self.slots:
            best = None
            for k in (0, 1):
                ev = sl.snap_event[k]
                if ev is not None and ev.query() and (best is None or sl.snap_launch[k] > sl.snap_launch[best]):
                    best = k
            if best is None:
                continue
            snap, upto = sl.snap[best], sl.snap_launch[best]
            for tk in list(sl.tickets):
                if tk.launch > upto:
                    continue
                if int(snap[tk.pending_at]) == 0:
                    # seen complete right after its own launch: nothing of it was ever parked; otherwise goals may
                    # have arrived with a later launch and the joint recovery runs again over the batch
                    tk.late = tk.seen_parked or upto > tk.launch
                    self._finalise(tk)
                else:
                    tk.seen_parked = True

    # ------------------------------------------------------------------ API
    def submit(self, T_goal, Y_init=None) -> Ticket:
        """Enqueue one batch T_goal[B,4,4] (CUDA tensor, pinned/pageable host tensor or array).  Returns at once."""
        eng, torch = self.eng, self.torch
        self.poll()
        sl_index = self.n_submitted % len(self.slots)
        sl = self.slots[sl_index]
        if not sl.free_counters:
            self.drain_slot(sl)
        caller = torch.cuda.current_stream(eng.device)
        with torch.cuda.device(eng.device), torch.cuda.stream(sl.stream):
            sl.stream.wait_stream(caller)          # T_goal may have been produced on the caller's stream
            if isinstance(T_goal, torch.Tensor) and not T_goal.is_cuda:
                T = T_goal.to(eng.device, non_blocking=True).reshape(-1, 4, 4)
            else:
                T = eng._f64(T_goal).reshape(-1, 4, 4)
            B = T.shape[0]
            tk = Ticket(self.n_submitted, sl_index, sl.launches + 1, B)
            tk.pending_at = sl.free_counters.pop()
            tk.seen_parked = False
            sl.pending[tk.pending_at].zero_()
            g2 = eng.goal_distances(T)
            if Y_init is None:
                Y0 = eng._empty(B, eng.plan.N, 3)
                _lib.check(self.lib.gik_bounds_init(eng.plan.handle, _p(g2), B, _p(Y0), _p(sl.workspace),
                                                    ctypes.c_void_p(sl.stream.cuda_stream)), "gik_bounds_init")
                eng.launches += 1
            else:
                Y0 = eng._f64(Y_init).reshape(B, eng.plan.N, 3)
            out = {"x": eng._empty(B, eng.plan.N, 3), "f(x)": eng._empty(B), "gradnorm": eng._empty(B),
                   "iterations": eng._empty(B, dtype=torch.int32), "status": eng._empty(B, dtype=torch.int32),
                   "n_inner": eng._empty(B, dtype=torch.int32)}
            pend = ctypes.c_void_p(sl.pending.data_ptr() + 4 * tk.pending_at)
            self._launch(sl, g2, Y0, out, pend, self.inner_budget)
            out["goal_d2"], out["T_goal"], out["Y_init"] = g2, T, Y0
            out["q"] = eng.joints(out["x"], T)     # final unless goals of this batch were parked (see _finalise)
            tk.out = out
        sl.tickets.append(tk)
        self.n_submitted += 1
        return tk

    def _drain_launch(self, sl):
        """Run the slot's parked goals to their end: one launch without a budget, no new goals."""
        if sl.tickets:
            with self.torch.cuda.device(self.eng.device), self.torch.cuda.stream(sl.stream):
                self._launch(sl, None, None, None, None, 0)
            return True
        return False

    def _drain_collect(self, sl):
        sl.stream.synchronize()
        snap = sl.snap[sl.launches & 1]
        for tk in list(sl.tickets):
            if int(snap[tk.pending_at]) != 0:
                raise _lib.GikError("a drained slot still reports parked goals")
            tk.late = True   # not seen complete before this launch: some of its goals may have arrived with it
            self._finalise(tk)

    def reserve(self, n_batches, B):
        """Prime torch's caching allocator for `n_batches` outstanding batches of B goals (and, for a to_host stream, pin
        its ring of result buffers): without it the first pass over a long stream calls cudaMalloc / cudaHostAlloc from
        submit() and result(), and those wait for the running launch."""
        eng, torch = self.eng, self.torch
        N, n, ng = eng.plan.N, max(eng.plan.n_joints, 1), max(eng.plan.n_goal, 1)
        per_slot = -(-int(n_batches) // len(self.slots))
        with torch.cuda.device(eng.device):
            for sl in self.slots:
                with torch.cuda.stream(sl.stream):
                    hold = []
                    for _ in range(per_slot):
                        hold += [eng._empty(B, 4, 4), eng._empty(B, ng), eng._empty(B, N, 3), eng._empty(B, N, 3),
                                 eng._empty(B), eng._empty(B), eng._empty(B, dtype=torch.int32),
                                 eng._empty(B, dtype=torch.int32), eng._empty(B, dtype=torch.int32), eng._empty(B, n),
                                 eng._empty(B, n)]
                    del hold
                if self.to_host:       # ... and the ring of pinned result buffers (pinning memory costs milliseconds)
                    ring = sl.host_ring.setdefault(int(B), [])
                    while len(ring) < self.HOST_RING:
                        ring.append({"q": torch.empty((B, eng.plan.n_joints), dtype=torch.float64).pin_memory(),
                                     "f(x)": torch.empty((B,), dtype=torch.float64).pin_memory(),
                                     "status": torch.empty((B,), dtype=torch.int32).pin_memory()})
                sl.stream.synchronize()

    def drain_slot(self, sl):
        if self._drain_launch(sl):
            self._drain_collect(sl)

    def drain(self):
        """Complete every submitted batch.  Blocks until the device is done."""
        self.poll()
        launched = [sl for sl in self.slots if self._drain_launch(sl)]
        for sl in launched:
            self._drain_collect(sl)
        for sl in self.slots:
            sl.stream.synchronize()

    def result(self, tk: Ticket, host=False):
        """The finished batch: the dict of `BatchIK.solve` (device tensors), or its pinned host copies
        (`q`, `f(x)`, `status`) with host=True on a stream created with to_host=True.  Blocks."""
        if not tk.done:
            self.slots[tk.slot].stream.synchronize()
            self.poll()
        if not tk.done:
            self.drain_slot(self.slots[tk.slot])
        tk.ready.synchronize()
        return tk.host if host else tk.out

    def stats(self):
        """Device-side queue statistics (synchronises): goals that found the carry queue full."""
        full = 0
        for sl in self.slots:
            sl.stream.synchronize()
            for c in sl.carry:
                full += int(c.view(self.torch.int32)[3])
        return {"queue_full_events": full, "launches": self.launches, "slots": len(self.slots),
                "inner_budget": self.inner_budget, "carry_capacity": self.capacity}
