"""Several batches in flight: run `model(x)` for consecutive batches on alternating CUDA streams.

One forward is a short chain of launches around the tensor-core sweep -- prior, sweep, merge + decode, tail for b_sae --
and the sweep owns every SM while it runs, so on ONE stream the small latency-bound kernels of a batch can only run
before or after it. Rows of different batches are independent, so with two batches in flight the merge
of batch i shares the GPU with the prior kernel of batch i + 1: b_sae 512 -> 32768 at batch 4096 goes from 167 to 144 us
per batch (profiles/r2s2_*). The library needs nothing special for this: every entry point launches on the caller's
current stream and keeps one workspace per stream.

    pipe = StreamPipeline(model, n_streams=2)
    for latent, recon, pol in pipe.map(batches):      # results in order; each is complete when it is yielded
        ...

The reference has no counterpart (single stream, eager ops: training/trainer.py:73-88, scripts/analysis/*); this is the
batch loop of those callers, restated for a GPU that can overlap independent batches.
"""
from __future__ import annotations

from collections import deque

import torch


def _record(out, stream) -> None:
    """Tell the caching allocator that `stream` uses the tensors in `out` (they were allocated on a worker stream)."""
    if isinstance(out, torch.Tensor):
        if out.is_cuda:
            out.record_stream(stream)
    elif isinstance(out, (tuple, list)):
        for o in out:
            _record(o, stream)
    elif isinstance(out, dict):
        for o in out.values():
            _record(o, stream)
    elif hasattr(out, "values") and hasattr(out, "indices"):      # SparseLatents
        _record(out.values, stream)
        _record(out.indices, stream)


class StreamPipeline:
    """Round-robin `fn(batch)` over `n_streams` CUDA streams with `torch.no_grad()`; outputs are handed back in
    submission order once their stream has finished them (the consumer's stream waits on an event, not the host)."""

    def __init__(self, fn, n_streams: int = 2, device=None):
        if n_streams < 1:
            raise ValueError("n_streams must be >= 1")
        if not torch.cuda.is_available():
            raise RuntimeError("StreamPipeline needs a CUDA device (no CPU fallback)")
        self.fn = fn
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
        self._i = 0
        self._primed = False

    def reprime(self) -> None:
        """Call after the model's weights changed: the next batch re-creates the prepared (packed / bf16) weights on its
        stream, and the other streams must not run ahead of that."""
        self._primed = False

    def submit(self, batch):
        """Enqueue fn(batch) on the next stream. -> (outputs, event): the outputs may be used on any stream that has
        waited for the event (`torch.cuda.current_stream().wait_event(event)`), or after `event.synchronize()`."""
        st = self.streams[self._i % len(self.streams)]
        self._i += 1
        st.wait_stream(torch.cuda.current_stream(self.device))      # the batch was produced on the caller's stream
        with torch.cuda.stream(st), torch.no_grad():
            out = self.fn(batch)
            ev = torch.cuda.Event()
            ev.record(st)
        if not self._primed:
            # the first forward builds the modules' prepared weights (cached per weight version) on ITS stream: the other
            # streams wait for it once, so that they never read a packed dictionary that is still being written
            for other in self.streams:
                if other is not st:
                    other.wait_event(ev)
            self._primed = True
        if isinstance(batch, torch.Tensor):
            batch.record_stream(st)
        return out, ev

    def map(self, batches, depth: int | None = None):
        """Generator over fn(batch) for every batch, `depth` (default: n_streams) batches in flight. The caller's
        current stream waits for each result's event before it is yielded."""
        depth = len(self.streams) if depth is None else max(1, depth)
        pending = deque()
        for b in batches:
            pending.append(self.submit(b))
            if len(pending) >= depth:
                yield self._take(pending)
        while pending:
            yield self._take(pending)

    def _take(self, pending):
        out, ev = pending.popleft()
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        _record(out, cur)
        return out
