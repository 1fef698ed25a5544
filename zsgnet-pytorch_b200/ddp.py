"""Data-parallel gradient exchange (the path's only collective; main_dist.py:36-40).

One process per GPU.  Parameters and gradients live in flat arenas ordered by backward completion,
so a gradient bucket is a contiguous slice: as soon as the engine reports a slice final, it is
all-reduced with NCCL (NVLink 5 / NVSwitch) on a side stream while the backward continues.  The
reduction is an AVERAGE (NCCL's ReduceOp.AVG: the 1/N is applied inside the collective, no extra pass
touches the gradients), exactly what DistributedDataParallel leaves in param.grad (main_dist.py:36-40),
so any optimiser -- this repo's FusedAdam or a stock torch.optim.Adam over the arena views -- sees the
same gradients as under the reference.  BatchNorm statistics stay per rank, like the reference's
non-synchronised BatchNorm."""
import torch
import torch.distributed as dist


class GradReducer:
    def __init__(self, store, min_bucket_elems=4 << 20, group=None):
        self.store, self.group = store, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.min_bucket = min_bucket_elems
        self.comm_stream = torch.cuda.Stream() if (self.world > 1 and store.grad_arena.is_cuda) else None
        self._lo = None
        self._works = []
        self.bytes_reduced = 0
        self.calls = 0

    @property
    def grad_scale(self):
        """Factor the optimiser still has to apply to the arena gradients: none, the collective averages."""
        return 1.0

    def on_bucket(self, lo, hi):
        """Engine callback: grad_arena[lo:hi] is final (called in increasing arena order)."""
        if self.world == 1:
            return
        if self._lo is None:
            self._lo = lo
        if hi - self._lo >= self.min_bucket or hi >= self.store.used:
            self._launch(self._lo, hi)
            self._lo = None

    def _launch(self, lo, hi):
        buf = self.store.grad_arena[lo:hi]
        self.bytes_reduced += buf.numel() * 4
        self.calls += 1
        if self.comm_stream is None:                      # gloo / CPU tests: no AVG there, divide after the wait
            self._works.append((dist.all_reduce(buf, group=self.group, async_op=True), buf))
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.comm_stream.wait_event(ev)
        with torch.cuda.stream(self.comm_stream):
            dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)   # NCCL: enqueued on comm_stream, overlaps the backward

    def finish(self):
        """Make the compute stream wait for every outstanding all-reduce (call before the optimiser)."""
        if self.world == 1:
            return
        if self._lo is not None:
            self._launch(self._lo, self.store.used)
            self._lo = None
        for w, buf in self._works:
            w.wait()
            buf.div_(self.world)
        self._works = []
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    def broadcast_buffers(self, net):
        """Rank 0's BatchNorm running statistics to every rank, on the compute stream: what DistributedDataParallel(
        broadcast_buffers=True) does at the start of EVERY forward (main_dist.py:37-40).  Statistics are still computed per
        rank; this only decides whose running averages eval mode and the checkpoint see (rank 0's, as in the reference)."""
        if self.world == 1 or net._bn_f.numel() == 0:
            return
        dist.broadcast(net._bn_f, 0, group=self.group)
        dist.broadcast(net._bn_n, 0, group=self.group)

    def broadcast_state(self, net):
        """Rank 0's parameters and BatchNorm buffers to every rank (DDP does this at construction)."""
        if self.world == 1:
            return
        dist.broadcast(self.store.param_arena, 0, group=self.group)
        dist.broadcast(net._bn_f, 0, group=self.group)
        dist.broadcast(net._bn_n, 0, group=self.group)


def shard_range(n_items, rank, world):
    """Contiguous shard of a global batch: rank r takes [r*B, (r+1)*B)."""
    per = n_items // world
    return rank * per, (rank + 1) * per
