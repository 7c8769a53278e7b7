"""Data-parallel training: one process per GPU, full replica, per-rank micro-batch (BatchNorm1d statistics stay
per rank -- the reference has neither SyncBN nor DDP, so N-GPU parity is defined as N independent reference
micro-batches whose gradients are averaged).

The only exchange step is the gradient all-reduce.  ``TrainEngine`` keeps every parameter gradient in one flat buffer
ordered by backward completion, so the reduction runs on contiguous slices, launched as soon as a slice is final
(NCCL enqueues on its own stream behind the compute stream's work so far) while the backward of the earlier layers
is still running.  Gradients are SUMMED; the 1/world factor is applied inside the AdamW kernel (grad_scale).
The same code runs over gloo on CPU tensors for the host-logic tests."""
import torch.distributed as dist


class StagedGradAllReduce:
    def __init__(self, engine, group=None, bucket_elems=8 * 1024 * 1024):
        self.engine, self.group = engine, group
        self.world = dist.get_world_size(group)
        self.bucket_elems = bucket_elems
        self.pending = []
        self._start = None
        self.launched = []          # (start, end) of every collective of the last step, for tests / logging
        self.time_exposed = False   # bench: CUDA events around the wait in finish() = communication NOT hidden behind the backward
        self._exposed = []

    def on_stage(self, i, start, end):
        """TrainEngine.backward callback: flat_grad[start:end] is final (in stream order)"""
        if i == 0:
            self.pending, self.launched, self._start = [], [], start
        last = i == len(self.engine.stages) - 1
        if end - self._start >= self.bucket_elems or last:
            if self.world > 1:
                self.pending.append(dist.all_reduce(self.engine.flat_grad[self._start:end], op=dist.ReduceOp.SUM,
                                                    group=self.group, async_op=True))
            self.launched.append((self._start, end))
            self._start = end

    def finish(self):
        """make the current stream (CPU: the caller) wait for every collective of this step"""
        ev = None
        if self.time_exposed and self.pending and self.engine.flat_grad.is_cuda:
            import torch
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        for w in self.pending:
            w.wait()
        if ev is not None:
            ev[1].record()
            self._exposed.append(ev)
        self.pending = []

    def exposed_ms(self):
        """mean time per step the compute stream spent waiting for the gradient all-reduces (call after a synchronize)"""
        if not self._exposed:
            return None
        return sum(a.elapsed_time(b) for a, b in self._exposed) / len(self._exposed)
