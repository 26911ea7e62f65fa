"""Host-buffer streaming around the device-resident hot loops.

The reference feeds its models from host arrays batch by batch with a blocking H2D copy in front of and a
blocking D2H copy behind every batch (`full_mse_scores_batched`, 04_vae_thresholding.py:113-124; eval_group,
06_test_full_pipeline.py:338-344; `recon_mse_per_window`, 10_test_hybrid_pipeline.py:240-251).  `HostStream`
is the same loop for chunks of 2^18..2^20 windows with the copies taken off the critical path: chunk i+1's
host-to-device copy and chunk i-1's device-to-host copy run on their own CUDA streams under chunk i's
kernels (two device input buffers, two pinned result sets, events between the three streams, no
host synchronisation inside a chunk).

The dense label scatter (`y_pred[idx] = label; hyb_score_full[idx] = p_struct`, 06_test_full_pipeline.py:336,356,
368-372) is a libshmfast kernel (`ops.scatter_flagged_4dof` / inside `shm_hybrid4dof_score`): the flagged count
never leaves the device and no ATen kernel runs on the chunk path.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Sequence, Tuple  # noqa: F401

import torch

from ._lib import ShmfastError


class HostStream:
    """Double-buffered H2D -> step -> D2H pipeline.

    step(dev_in, i) runs the device work of chunk i on the CURRENT stream and returns {name: device tensor};
    every returned tensor is copied into this chunk's pinned host buffer `name` (declared in `out_specs`).
    `run` yields (i, {name: pinned host tensor}) once chunk i's results have landed.  There are `depth + 1` pinned
    result sets, so the tensors of chunk i stay untouched while chunk i+1 is being fetched (its resume enqueues the
    D2H copy of chunk i+2 into a third set); they are overwritten once chunk i+2 is requested.
    """

    def __init__(self, device: torch.device, in_shape: Sequence[int], out_specs: Dict[str, Tuple[Sequence[int], torch.dtype]],
                 in_dtype: torch.dtype = torch.float32, depth: int = 2):
        if device.type != "cuda":
            raise ShmfastError("HostStream needs a CUDA device: libshmfast has no CPU path")
        if depth < 2:
            raise ShmfastError("depth must be >= 2 (one buffer being filled while the other is read)")
        self.device, self.depth = device, int(depth)
        self.s_in = torch.cuda.Stream(device)
        self.s_out = torch.cuda.Stream(device)
        self.dev_in = [torch.empty(tuple(in_shape), dtype=in_dtype, device=device) for _ in range(depth)]
        self.n_out = self.depth + 1
        self.host_out = [{k: torch.empty(tuple(shape), dtype=dt).pin_memory() for k, (shape, dt) in out_specs.items()} for _ in range(self.n_out)]
        mk = lambda m: [torch.cuda.Event() for _ in range(m)]
        self.h2d_done, self.step_done, self.d2h_done = mk(depth), mk(depth), mk(self.n_out)
        self._keep = [None] * depth                   # device results stay referenced until their D2H copy has run
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _enqueue_h2d(self, i: int, host: torch.Tensor) -> torch.Tensor:
        slot = i % self.depth
        buf = self.dev_in[slot]
        if host.dtype != buf.dtype or host.dim() != buf.dim() or any(a > b for a, b in zip(host.shape, buf.shape)):
            raise ShmfastError(f"chunk {i}: host tensor {tuple(host.shape)} {host.dtype} does not fit the staging buffer {tuple(buf.shape)}")
        if not host.is_pinned():
            raise ShmfastError("host chunks must be pinned (torch.Tensor.pin_memory) for the copy to overlap")
        view = buf[tuple(slice(0, s) for s in host.shape)] if tuple(host.shape) != tuple(buf.shape) else buf
        if i >= self.depth:
            self.s_in.wait_event(self.step_done[slot])          # the buffer's previous chunk has been consumed
        with torch.cuda.stream(self.s_in):
            view.copy_(host, non_blocking=True)
            self.h2d_done[slot].record(self.s_in)
        self.h2d_bytes += host.numel() * host.element_size()
        return view

    def run(self, chunks: Iterable[torch.Tensor], step: Callable[[torch.Tensor, int], Dict[str, torch.Tensor]]):
        compute = torch.cuda.current_stream(self.device)
        it = iter(chunks)
        nxt = next(it, None)
        views = {}
        if nxt is not None:
            views[0] = self._enqueue_h2d(0, nxt)
        i = 0
        while nxt is not None:
            nxt = next(it, None)
            if nxt is not None:
                views[i + 1] = self._enqueue_h2d(i + 1, nxt)     # runs under chunk i's kernels
            slot = i % self.depth
            compute.wait_event(self.h2d_done[slot])
            outs = step(views.pop(i), i)
            self.step_done[slot].record(compute)
            self._keep[slot] = outs
            self.s_out.wait_event(self.step_done[slot])
            oslot = i % self.n_out
            with torch.cuda.stream(self.s_out):
                for k, host in self.host_out[oslot].items():
                    src = outs[k]
                    dst = host[tuple(slice(0, s) for s in src.shape)] if tuple(src.shape) != tuple(host.shape) else host
                    dst.copy_(src, non_blocking=True)
                    self.d2h_bytes += src.numel() * src.element_size()
                self.d2h_done[oslot].record(self.s_out)
            if i >= 1:                                          # chunk i-1's results landed while chunk i was being issued / run
                yield self._deliver(i - 1)
            i += 1
        if i >= 1:
            yield self._deliver(i - 1)

    def _deliver(self, i: int):
        oslot = i % self.n_out
        self.d2h_done[oslot].synchronize()
        self._keep[i % self.depth] = None
        return i, self.host_out[oslot]
