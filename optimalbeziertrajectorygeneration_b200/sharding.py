"""Multi-GPU sharding of the constraint path (one process per GPU,
torch.distributed; NCCL over NVLink on the GPUs, gloo in the CPU tests).

The reference has no parallelism at all; the units of this path
are independent, so sharding needs no data-path collective:

  * ``batch``  -- the finite-difference perturbations x + h_k e_k (or any batch
    of optimisation vectors) are dealt out in contiguous blocks; each rank
    evaluates all pairs for its own x's (weak scaling).
  * ``pairs``  -- one x, the lexicographic pair list is cut into ``world``
    contiguous, equally sized ranges (the C-ABI takes [pair_begin, npairs)), so
    the output of rank r is exactly rows [begin_r, end_r) of the full vector.

The only collective is one all-gather of the per-pair minimum (the active-pair
source): after it every rank holds the whole [B_total, P] min-distance matrix.
"""
import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def block_range(total, world, rank):
    """Contiguous balanced block [begin, end) of ``total`` units for ``rank``."""
    base, rem = divmod(int(total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def pair_range(n_curves, world, rank):
    """[pair_begin, pair_end) of the i<j pair list handled by ``rank``."""
    return block_range(n_curves * (n_curves - 1) // 2, world, rank)


def gather_pair_minima(local, mode="batch", total=None):
    """All-gathers the per-pair minima.

    mode 'batch': local is [B_local, P] (equal B_local on all ranks) -> [world*B_local, P]
    mode 'pairs': local is [B, P_local] with possibly unequal P_local; ``total`` = P
                  -> [B, P]   (padded all-gather, then the pad columns are dropped)
    """
    rank, world = world_info()
    if world == 1:
        return local
    if mode == "batch":
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    if mode != "pairs":
        raise ValueError("mode must be 'batch' or 'pairs'")
    B = local.shape[0]
    width = -(-int(total) // world)                      # widest shard
    padded = torch.zeros((B, width), dtype=local.dtype, device=local.device)
    padded[:, :local.shape[1]] = local
    buf = torch.empty((world, B, width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf.view(world * B, width), padded)
    parts = []
    for r in range(world):
        b, e = block_range(total, world, r)
        parts.append(buf[r, :, :e - b])
    return torch.cat(parts, dim=1)


class PairMinimaGatherer:
    """Overlapped form of gather_pair_minima(mode='batch') for a stream of steps: the
    all-gather of step k runs on its own CUDA stream while the kernels of step k+1 run
    on the caller's stream (two local/gathered buffer pairs, events, no host sync).

        g = PairMinimaGatherer(B, P, device)
        for step in ...:
            local = g.local_buffer()          # [B, P] the kernels of this step write into
            ... launch kernels writing `local` on the current stream ...
            gathered = g.gather()             # [world*B, P]; valid after g.wait() / g.finish()
        g.finish()
    """

    def __init__(self, B, P, device, dtype=torch.float64):
        self.rank, self.world = world_info()
        self.local = [torch.empty((B, P), dtype=dtype, device=device) for _ in range(2)]
        self.full = [torch.empty((self.world * B, P), dtype=dtype, device=device) for _ in range(2)] \
            if self.world > 1 else self.local
        # CPU tensors (gloo, host-logic tests): same buffer rotation, synchronous collective
        self.cuda = torch.device(device).type == "cuda"
        self.stream = torch.cuda.Stream(device=device) if (self.world > 1 and self.cuda) else None
        self.done = [None, None]
        self.k = 0

    def local_buffer(self):
        """Buffer for the next step; waits (on the current stream) until the gather that last
        read it has finished."""
        i = self.k & 1
        if self.done[i] is not None and self.cuda:
            torch.cuda.current_stream().wait_event(self.done[i])
        return self.local[i]

    def gather(self):
        i = self.k & 1
        self.k += 1
        if self.world == 1:
            return self.local[i]
        if not self.cuda:
            dist.all_gather_into_tensor(self.full[i], self.local[i])
            return self.full[i]
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            dist.all_gather_into_tensor(self.full[i], self.local[i])
            self.done[i] = torch.cuda.Event()
            self.done[i].record(self.stream)
        return self.full[i]

    def finish(self):
        """Makes the current stream wait for every outstanding gather."""
        for ev in self.done:
            if ev is not None and self.cuda:
                torch.cuda.current_stream().wait_event(ev)


class PeerMinima:
    """The all-gather of the per-pair minima fused into the pair kernel: the gathered
    [world*B, P] matrix of every rank lives in symmetric memory (torch.distributed
    ._symmetric_memory: peer-mapped device buffers over NVLink / NVSwitch), and each rank's
    kernel stores its minima into its block of *every* rank's matrix while it computes
    (8 bytes per pair per peer next to 976 bytes per pair of HBM traffic), so no collective
    kernel follows.  Completion = every rank has finished its kernel: a symmetric-memory
    signal-pad barrier, issued on a *side stream* behind an event, so that it never sits
    between two pair kernels on the launch stream (round 1 issued it there: 30-40 us per
    step of barrier latency + rank skew in front of the next persistent kernel).

    Three matrices rotate.  Step k writes matrix k % 3 (locally and on every peer); before
    a rank launches step k it waits (launch stream) for the barrier of step k-2, which says
    that every rank has finished the kernel of step k-2 and therefore -- consumers run in
    launch-stream order -- finished reading matrix k % 3 from step k-3.  That barrier was
    enqueued a whole kernel earlier on a high-priority stream (it takes an SM slot at the
    kernel boundary, next to the following persistent kernel), so the wait is free.

        pm = PeerMinima(B, P, device)                  # collective (rendezvous)
        local, peers = pm.targets()                    # this step's destinations
        eng.separation(cpts, E, maxSep, out=..., pairmin=local, peer_ptrs=peers)
        gathered = pm.complete()                       # [world*B, P]; valid after pm.wait()
        ...
        pm.wait()                                      # launch stream waits for the last barrier

    ``layout='pairs'`` (strong scaling: one batch of B evaluations, the pair list cut into
    ``world`` contiguous ranges): the gathered matrix is [B, P] and rank r fills columns
    [begin_r, end_r) of every peer's matrix (``targets()`` returns views / addresses offset to
    the range's first column; launch with ``min_pitch=P``)."""

    NBUF = 3

    def __init__(self, B, P, device, group=None, layout="batch"):
        import torch.distributed._symmetric_memory as symm_mem
        self.rank, self.world = world_info()
        if self.world > 8:
            raise ValueError("PeerMinima supports one NVLink domain of up to 8 GPUs")
        if layout not in ("batch", "pairs"):
            raise ValueError("layout must be 'batch' or 'pairs'")
        self.B, self.P, self.layout = int(B), int(P), layout
        group = dist.group.WORLD if group is None else group
        rows = self.world * self.B if layout == "batch" else self.B
        self.rows = rows
        self.buf = symm_mem.empty((self.NBUF, rows, self.P), dtype=torch.float64, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        # NVSwitch multicast (NVLS): one store to the multicast address of the symmetric buffer lands
        # in every rank's copy (this rank's included), replicated inside the switch -- one remote store
        # per tile instead of world-1, and 1x instead of (world-1)x NVLink egress.  multimem.st of an
        # f64 is a plain STG to that address, so the kernel needs no separate code path: the multicast
        # address is handed over as the only "peer".  BEZGPU_PEER_MULTICAST=0 keeps the per-peer stores.
        import os
        self.mc_ptr = 0
        try:
            if os.environ.get("BEZGPU_PEER_MULTICAST", "1") != "0" and self.world > 1:
                self.mc_ptr = int(self.hdl.multicast_ptr or 0)
        except Exception:
            self.mc_ptr = 0
        self.side = torch.cuda.Stream(device=device, priority=-1)
        self.done = [None] * self.NBUF
        self.k = 0
        self.pair_lo, self.pair_hi = pair_range_of(self.P, self.world, self.rank) if layout == "pairs" else (0, self.P)

    def targets(self):
        """(local view, list of peer device addresses) for the current step."""
        i = self.k % self.NBUF
        prev2 = self.done[(self.k - 2) % self.NBUF] if self.k >= 2 else None
        if prev2 is not None:
            torch.cuda.current_stream().wait_event(prev2)
        if self.layout == "batch":
            off = (i * self.rows + self.rank * self.B) * self.P * 8
            local = self.buf[i, self.rank * self.B:(self.rank + 1) * self.B]
        else:
            off = (i * self.rows * self.P + self.pair_lo) * 8
            local = self.buf[i, :, self.pair_lo:]
        if self.mc_ptr:
            peers = [self.mc_ptr + off]
        else:
            peers = [self.ptrs[r] + off for r in range(self.world) if r != self.rank]
        return local, peers

    def complete(self):
        """Issues the barrier of the current step on the side stream (behind an event recorded
        on the launch stream) and rotates.  Returns the gathered matrix of the step; it is
        complete once :meth:`wait` (or the event ``last_event``) has passed."""
        i = self.k % self.NBUF
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            self.hdl.barrier(channel=i)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[i] = ev
        self.last_event = ev
        self.k += 1
        return self.buf[i]

    def wait(self):
        """Launch stream waits until the most recent step is complete on every rank."""
        if getattr(self, "last_event", None) is not None:
            torch.cuda.current_stream().wait_event(self.last_event)


def pair_range_of(P, world, rank):
    return block_range(P, world, rank)
