"""One whole DSVI iteration -- zero_grad, fused forward/gradient, gradient all-reduce, Adam -- captured once into a CUDA
graph and replayed.  A step is ~200-900 kernel launches (most of them microseconds long: the
Q x Q factorisations, KL terms and their adjoints), so at small shapes (the shipped simulation data: T=100, D=2, Q=20)
and on the non-scaling tail of the 8-GPU run the step is bound by launch latency and host enqueue time, not by the
GPU; the graph removes both.

Everything that changes from step to step must be DATA, not a kernel argument frozen at capture:
  * the rows (x, y, I) live in static device buffers (``load_rows`` copies a new minibatch into them);
  * the Monte-Carlo noise is counter-based and keyed by a step counter kept on the device
    (``NMGP.use_device_step_counter``), bumped inside the graph -- every replay draws the noise the eager path would;
  * Adam runs with ``capturable=True`` (its step count is a device tensor);
  * the deferred positive-definiteness flag of the step is a static device scalar, checked by ``check()``.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _ops as ops
from . import parallel


class GraphedStep:
    def __init__(self, model, optimizer, x, y, I, n_mc=1, row_gid=None, distributed=False, warmup=3):
        dev = model.device
        if dev.type != "cuda":
            raise RuntimeError("GraphedStep needs a CUDA model")
        for group in optimizer.param_groups:
            if not group.get("capturable", False):
                raise ValueError("the optimizer must be created with capturable=True to be replayed from a CUDA graph")
        self.model, self.opt, self.n_mc, self.distributed = model, optimizer, n_mc, distributed
        self.x, self.y, self.I = x.clone(), y.clone(), I.clone()          # static input buffers
        self.gid = None if row_gid is None else row_gid.clone()
        self.params = list(model.parameters())
        model.step_options = dict(model.step_options, defer_pd_check=True)
        model.use_device_step_counter(True)
        self._pd_flags = []
        self.loss = None
        # warm-up on a side stream (allocator pools, library scratch, lazy optimizer state) before capturing
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(max(1, warmup)):
                self._one_step()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.check()
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self._one_step()
        self.kernels_per_replay = ops.launch_count() - n0
        self.replays = 0

    def _one_step(self):
        self.opt.zero_grad(set_to_none=True)
        loss = self.model.forward_rows(self.x, self.y, self.I, n_mc=self.n_mc, row_gid=self.gid)
        loss.backward()
        flag = self.model._last_pd_info
        if self.distributed:
            tot = parallel.allreduce_loss_and_grads(loss, self.params, pd_info=flag, check="defer")
            self._pd_flags = list(parallel._pending_pd)
            parallel._pending_pd[:] = []
        else:
            tot = loss.detach()
            self._pd_flags = [flag] if flag is not None else []
        self.opt.step()
        self.model.advance_noise_step()
        return tot

    def load_rows(self, x=None, y=None, I=None, row_gid=None):
        """Copy a new minibatch of the SAME shape into the static buffers (host or device tensors)."""
        for dst, src in ((self.x, x), (self.y, y), (self.I, I), (self.gid, row_gid)):
            if src is not None:
                dst.copy_(src, non_blocking=True)

    def step(self):
        """Replay the captured iteration; returns the (static) loss tensor of this replay."""
        self.graph.replay()
        self.replays += 1
        ops._GRAPH_LAUNCHES[0] += self.kernels_per_replay
        return self.loss

    def check(self):
        """Host check of the step's deferred positive-definiteness flag (RuntimeError like torch.cholesky)."""
        for f in self._pd_flags:
            if f is not None and float(f) != 0.0:
                raise RuntimeError("cholesky: a matrix of the step is not positive-definite")
