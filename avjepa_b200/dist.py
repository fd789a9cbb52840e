"""Data-parallel plumbing: one process per GPU, NCCL over NVLink/NVSwitch.

The reference wraps its models in single-process ``nn.DataParallel`` and never synchronises
gradients between ranks; with one visible GPU per process that is a
pass-through.  This build shards the batch across ranks and averages gradients so that N
ranks x B/N clips reproduce the 1-GPU full-batch step:
one all-reduce per flat gradient buffer of the fused optimizer (4 buffers for the AV model,
the largest ~1.2 GB fp32 at ViT-L), issued in bucket-sized chunks so the first chunks are on
the wire while later ones are still being enqueued.
"""
import os

import torch
import torch.distributed as tdist


def init_distributed(port=37123, rank_and_world_size=(None, None)):
    """Returns (world_size, rank).  Reads torchrun's env (RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_ADDR / MASTER_PORT); falls back to world_size 1 like the reference."""
    if tdist.is_available() and tdist.is_initialized():
        return tdist.get_world_size(), tdist.get_rank()
    rank, world = rank_and_world_size
    if rank is None or world is None:
        if 'RANK' in os.environ and 'WORLD_SIZE' in os.environ:
            rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
        elif 'SLURM_NTASKS' in os.environ and 'SLURM_PROCID' in os.environ:        # what the reference reads
            rank, world = int(os.environ['SLURM_PROCID']), int(os.environ['SLURM_NTASKS'])
            os.environ.setdefault('MASTER_ADDR', os.environ.get('HOSTNAME', '127.0.0.1'))
        else:
            return 1, 0
    if world <= 1:
        return 1, 0
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', str(port))
    backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank % max(1, torch.cuda.device_count()))))
    tdist.init_process_group(backend=backend, world_size=world, rank=rank)
    return world, rank


_ACTIVE = None


def active_sync():
    """The GradSync armed for the step whose backward is running (None outside a data-parallel step)."""
    return _ACTIVE


class GradSync(object):
    """Averages the optimizer's flat gradient buffers across ranks.

    Plain mode: after backward, every flat buffer is all-reduced (SUM) in bucket-sized chunks.

    Overlapped mode (default with NCCL + the fused optimizer; ``AVJ_DDP_OVERLAP=0`` turns it off): the
    backward itself reports which gradient ranges are final and their all-reduce starts behind them on a
    side stream while the rest of the backward still computes --

    * predictor ranges when the last predictor backward of the step has been enqueued (the whole context-
      encoder backward is still to come),
    * encoder layers in buckets of ``layers_per_bucket`` while the LAST encoder backward runs: the C schedule
      (``avj_stack_backward``) records one CUDA event per layer, the side stream waits for the event of the
      bucket's lowest layer,
    * whatever is left (patch embedding, final norm, ...) at the end.

    Each range keeps the list of element intervals already reduced, so nothing is reduced twice and
    ``all_reduce`` finishes the complement.
    """

    def __init__(self, world_size, bucket_bytes=64 << 20, group=None, overlap=None, layers_per_bucket=4):
        self.world_size = world_size
        self.bucket_elems = bucket_bytes // 4
        self.group = group
        if overlap is None:
            overlap = os.environ.get('AVJ_DDP_OVERLAP', '1') != '0'
        self.overlap = bool(overlap) and torch.cuda.is_available() and world_size > 1
        self.layers_per_bucket = layers_per_bucket
        # AVJ_DDP_BF16=1: the overlapped all-reduces move bf16 copies of the gradient intervals (half the bytes on the
        # wire and half the HBM / L2 traffic beside the backward); the fp32 flat buffer receives the widened sum.  Off by
        # default: the sum is then rounded to 8 mantissa bits once per element (relative 2^-9), which the fp32 check
        # mode's 1e-4 bar does not allow.
        self.bf16_wire = os.environ.get('AVJ_DDP_BF16', '0') == '1'
        self._wire = {}
        self._side = None
        self._events = None
        self._works = []
        self._done = {}
        self._left = {}
        self._enc = self._opt = None
        # measurement (bench.py): CUDA events around the point where the compute stream has to WAIT for the
        # collectives -- the time between them is the all-reduce time the backward did not hide
        self.measure = False
        self._exposed = []

    def _wait_all(self, works, new_step=True):
        if self.measure and torch.cuda.is_available():
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for w in works:
                w.wait()
            b.record()
            if new_step or not self._exposed:
                self._exposed.append([])
            self._exposed[-1].append((a, b))
        else:
            for w in works:
                w.wait()

    def exposed_ms(self, reset=True):
        """Mean per-step time the compute stream spent blocked on gradient collectives since the last reset
        (synchronises on the recorded events; a step may wait several times, the waits are summed)."""
        if not self._exposed:
            return 0.0
        vals = []
        for pairs in self._exposed:
            t = 0.0
            for a, b in pairs:
                b.synchronize()
                t += a.elapsed_time(b)
            vals.append(t)
        if reset:
            self._exposed = []
        return sum(vals) / len(vals)

    # ------------------------------------------------------------------ plain path
    def all_reduce_flat(self, flats, average=True):
        """SUM every flat buffer across ranks in bucket-sized chunks.  With average=False the
        caller applies the returned 1/world factor itself (the fused optimizer folds it into its
        gradient multiplier).  Returns the factor still owed (1.0 when already applied)."""
        if self.world_size <= 1:
            return 1.0
        works = []
        for g in flats:
            n = g.numel()
            for lo in range(0, n, self.bucket_elems):
                chunk = g[lo:min(n, lo + self.bucket_elems)]
                works.append(tdist.all_reduce(chunk, op=tdist.ReduceOp.SUM, group=self.group, async_op=True))
        self._wait_all(works)
        inv = 1.0 / self.world_size
        if not average:
            return inv
        for g in flats:
            g.mul_(inv)
        return 1.0

    # ------------------------------------------------------------------ overlapped path
    def begin_step(self, optimizer, encoder_backbone, n_encoder_backward, n_predictor_backward):
        """Arms the hooks for one backward.  `encoder_backbone`: the module whose `.blocks` the per-layer
        events refer to; the counts say how many backward calls of each kind this step will make."""
        global _ACTIVE
        if not (self.overlap and hasattr(optimizer, '_ranges')):
            _ACTIVE = None
            return
        optimizer.ensure_built()
        self._opt, self._enc = optimizer, encoder_backbone
        self._left = {'encoder': int(n_encoder_backward), 'predictor': int(n_predictor_backward)}
        self._works = []
        self._records = []
        self._done = {id(r['g']): [] for r in optimizer._ranges if r['g'] is not None}
        if self._side is None:
            self._side = torch.cuda.Stream()
        _ACTIVE = self

    @staticmethod
    def pending_intervals(done, lo, hi):
        """Sub-intervals of [lo, hi) not covered by the (possibly unsorted, non-overlapping) intervals in `done`."""
        todo, cur = [], lo
        for a, b in sorted(done):
            if b <= cur:
                continue
            if a >= hi:
                break
            if a > cur:
                todo.append((cur, min(a, hi)))
            cur = max(cur, b)
        if cur < hi:
            todo.append((cur, hi))
        return todo

    def _reduce(self, g, lo, hi):
        """Async SUM of g[lo:hi] (minus what was already reduced), chunked; records the interval."""
        done = self._done[id(g)]
        todo = self.pending_intervals(done, lo, hi)
        for a, b in todo:
            works = []
            for c0 in range(a, b, self.bucket_elems):
                c1 = min(b, c0 + self.bucket_elems)
                if self.bf16_wire and g.is_cuda:
                    works.append(self._reduce_bf16(g, c0, c1))
                else:
                    works.append(tdist.all_reduce(g[c0:c1], op=tdist.ReduceOp.SUM, group=self.group, async_op=True))
            self._works += works
            self._records.append((g, a, b, works))               # issue order == completion order on the NCCL stream
            done.append((a, b))

    class _EventWork(object):
        """work-like handle: wait() makes the current stream wait for an event."""

        def __init__(self, ev):
            self.ev = ev

        def wait(self):
            torch.cuda.current_stream().wait_event(self.ev)

    def _reduce_bf16(self, g, c0, c1):
        """g[c0:c1] <- widen(SUM over ranks of bf16(g[c0:c1])): narrowing copy, all-reduce and widening copy run on the side
        stream (behind everything the calling stream has enqueued); returns a handle whose wait() orders the caller behind
        the write-back."""
        wire = self._wire.get(id(g))
        if wire is None:
            wire = torch.empty(g.numel(), dtype=torch.bfloat16, device=g.device)
            self._wire[id(g)] = wire
        if self._side is None:
            self._side = torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        if cur != self._side:                       # called from the compute stream: everything enqueued so far comes first
            self._side.wait_stream(cur)
        ev = torch.cuda.Event()
        with torch.cuda.stream(self._side):
            w = wire[c0:c1]
            w.copy_(g[c0:c1])
            tdist.all_reduce(w, op=tdist.ReduceOp.SUM, group=self.group, async_op=True).wait()
            g[c0:c1].copy_(w)
            ev.record(self._side)
        return GradSync._EventWork(ev)

    def _ranges_of(self, kind):
        groups = (0, 2) if kind == 'encoder' else (1, 3)
        return [r for r in self._opt._ranges if r['g'] is not None and r['group'] in groups]

    def _layer_interval(self, r, layers):
        """[lo, hi) of range r covered by the parameters of encoder layers `layers` (contiguous: the flat
        order follows module order), or None."""
        ids = set()
        for i in layers:
            ids.update(id(p) for p in self._enc.blocks[i].parameters())
        lo = hi = None
        for p, o in zip(r['params'], r['offs']):
            if id(p) in ids:
                lo = o if lo is None else min(lo, o)
                hi = max(hi or 0, o + (p.numel() + 7) // 8 * 8)
        return None if lo is None else (lo, min(hi, r['g'].numel()))

    def layer_events_for(self, mod, n_layers):
        """Events for the per-layer hooks of `mod`'s backward -- only for the LAST encoder backward of the step
        (earlier calls still accumulate into the same gradients)."""
        if _ACTIVE is not self or mod is not self._enc or self._left.get('encoder', 0) != 1:
            return None
        if self._events is None or len(self._events) != n_layers:
            self._events = [torch.cuda.Event() for _ in range(n_layers)]
            for e in self._events:
                e.record()                      # materialise the CUDA handle the C schedule records into
        return self._events

    def on_layers_enqueued(self, mod, events):
        """Called right after avj_stack_backward returned (all kernels enqueued, events recorded)."""
        L = len(events)
        step = max(1, self.layers_per_bucket)
        for hi in range(L, 0, -step):
            layers = range(max(0, hi - step), hi)
            self._side.wait_event(events[layers[0]])          # the lowest layer of the bucket finishes last
            with torch.cuda.stream(self._side):
                for r in self._ranges_of('encoder'):
                    iv = self._layer_interval(r, layers)
                    if iv is not None:
                        self._reduce(r['g'], iv[0], iv[1])

    def on_backward_done(self, kind, mod):
        if _ACTIVE is not self:
            return
        self._left[kind] = self._left.get(kind, 0) - 1
        if self._left[kind] == 0:
            for r in self._ranges_of(kind):                   # current stream: everything enqueued so far is ordered before
                self._reduce(r['g'], 0, r['g'].numel())

    def _check_coverage(self, flats):
        for g in flats:                                       # every element exactly once
            iv = sorted(self._done[id(g)])
            assert iv and iv[0][0] == 0 and iv[-1][1] == g.numel() and all(a[1] == b[0] for a, b in zip(iv, iv[1:])), \
                'gradient all-reduce coverage error'

    def can_pipeline(self, optimizer):
        """True when this step's backward ran with the overlap hooks armed for `optimizer` (finish_pipelined is legal)."""
        return _ACTIVE is self and self._opt is optimizer and hasattr(optimizer, 'step_interval')

    def finish_pipelined(self, optimizer):
        """Generator over (flat gradient buffer, lo, hi) in the order the all-reduces were issued: each interval is yielded
        once the compute stream has been made to wait for ITS collective only, so the caller can update that interval
        (optimizer.step_interval) while later intervals are still being reduced.  Gradients are SUMS over ranks; the caller
        owes the factor 1 / world_size.  Every element of every flat buffer is yielded exactly once."""
        global _ACTIVE
        assert self.can_pipeline(optimizer)
        _ACTIVE = None
        flats = optimizer.flat_grads()
        for g in flats:                                       # the complement of what the hooks already started
            self._reduce(g, 0, g.numel())
        self._check_coverage(flats)
        first = True
        for g, a, b, works in self._records:
            self._wait_all(works, new_step=first)
            first = False
            yield g, a, b
        self._works = []
        self._records = []

    # ------------------------------------------------------------------ entry point after backward
    def all_reduce(self, optimizer, average=True):
        global _ACTIVE
        if hasattr(optimizer, 'flat_grads'):
            flats = optimizer.flat_grads()
        else:
            flats = [p.grad for grp in optimizer.param_groups for p in grp['params'] if p.grad is not None]
        if _ACTIVE is not self:
            return self.all_reduce_flat(flats, average=average)
        _ACTIVE = None
        for g in flats:                                       # the complement of what the hooks already started
            self._reduce(g, 0, g.numel())
        self._wait_all(self._works)
        self._works = []
        self._records = []
        self._check_coverage(flats)
        inv = 1.0 / self.world_size
        if not average:
            return inv
        for g in flats:
            g.mul_(inv)
        return 1.0
