"""CUDA-graph replay of one ClipLoss fwd+bwd (world_size == 1).

OneProt's shipped batch sizes (N = 2048..6144 per modality pair) make the loss launch-bound: one
fwd+bwd is ~25 kernel launches and ~35 host operations (~0.4 ms of enqueue) for < 0.1 ms of device
work.  ``ClipLoss(graph=True)`` captures the forward and the backward launch sequences once per
(shape, dtype, gradient pattern) into two CUDA graphs over static buffers and replays them: the
host cost of a step drops to two input copies and two graph launches.  Streams and graphs instead
of a tracing compiler: the captured sequence is exactly the eager one of ``clip_loss.py``.

Rules kept from torch's CUDA-graph contract: a warm-up step runs eagerly on a side stream before
capture (lazy initialisation such as cudaFuncSetAttribute happens there), all tensors used inside
the graphs are static, results are cloned out, and ``robust="auto"`` (a host-side decision) is not
capturable.
"""
from __future__ import annotations

import types

import torch


class _Ctx(types.SimpleNamespace):
    """Stand-in for the autograd ctx so that the eager forward / backward bodies can be captured."""

    def set_materialize_grads(self, flag):
        pass

    def mark_non_differentiable(self, *tensors):
        pass


class GraphedStep:
    def __init__(self, fn_cls, A, B, scale_t, cfg, needs):
        self.fn_cls, self.cfg, self.needs = fn_cls, dict(cfg), needs
        dev = A.device
        self.A_s = torch.empty_like(A, memory_format=torch.contiguous_format)
        self.B_s = torch.empty_like(B, memory_format=torch.contiguous_format)
        self.scale_s = torch.empty(1, dtype=torch.float32, device=dev)
        self.g_s = torch.ones((), dtype=torch.float32, device=dev)
        self.A_s.copy_(A.detach()); self.B_s.copy_(B.detach())
        self.scale_s.copy_(scale_t.detach().to(device=dev, dtype=torch.float32).reshape(1))
        self.scale_needs_grad = bool(scale_t.requires_grad)
        self.generation = 0          # bumped by every replayed forward: a backward must belong to the latest one

        # warm-up on a side stream (torch.cuda.graphs contract), then capture
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self._run_eager()
        cur.wait_stream(side)

        self.fwd_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.fwd_graph):
            self.ctx, self.loss_out, self.loss_f32, self.flag = self._forward_body()
        self.bwd_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.bwd_graph, pool=self.fwd_graph.pool()):
            self.gA, self.gB, self.gS = self._backward_body(self.ctx)

    # -- bodies (the eager implementation, run on static buffers) ----------------------------
    def _scale_view(self):
        s = self.scale_s.reshape(())
        return s.requires_grad_(True) if self.scale_needs_grad else s

    def _forward_body(self):
        ctx = _Ctx(needs_input_grad=(self.needs[0], self.needs[1], self.scale_needs_grad, False))
        with torch.no_grad():
            loss_out, loss_f32, flag = self.fn_cls._forward_impl(ctx, self.A_s, self.B_s, self._scale_view(), self.cfg)
        return ctx, loss_out, loss_f32, flag

    def _backward_body(self, ctx):
        with torch.no_grad():
            gA, gB, gS, _ = self.fn_cls._backward_impl(ctx, self.g_s)
        return gA, gB, gS

    def _run_eager(self):
        ctx, *_ = self._forward_body()
        self._backward_body(ctx)

    # -- replay ------------------------------------------------------------------------------
    def forward(self, A, B, scale_t):
        self.A_s.copy_(A.detach()); self.B_s.copy_(B.detach())
        self.scale_s.copy_(scale_t.detach().to(device=self.A_s.device, dtype=torch.float32).reshape(1))
        self.fwd_graph.replay()
        self.generation += 1
        return self.loss_out.clone(), self.loss_f32.clone(), self.flag.clone()

    def backward(self, g, generation):
        if generation != self.generation:
            # the static buffers (operands, saved sums, kept panel) hold a LATER forward's state: replaying would
            # return that call's gradients silently (e.g. loss = m(A1, B1) + m(A2, B2) on one graphed module)
            raise RuntimeError("ClipLoss(graph=True): backward of a forward that is no longer the latest one on this module; "
                               "call backward before the next forward of the same shape, or use graph=False")
        self.g_s.copy_(g.detach().to(dtype=torch.float32).reshape(()))
        self.bwd_graph.replay()
        gA = self.gA.clone() if self.gA is not None else None
        gB = self.gB.clone() if self.gB is not None else None
        gS = self.gS.clone() if self.gS is not None else None
        return gA, gB, gS


class GraphedClipFunction(torch.autograd.Function):
    """autograd bridge: forward / backward replay the two graphs of a GraphedStep."""

    @staticmethod
    def forward(ctx, A, B, scale_t, step: GraphedStep):
        ctx.step = step
        ctx.set_materialize_grads(False)
        loss_out, loss_f32, flag = step.forward(A, B, scale_t)
        ctx.generation = step.generation
        ctx.mark_non_differentiable(loss_f32, flag)
        return loss_out, loss_f32, flag

    @staticmethod
    def backward(ctx, g_loss, _g32, _gflag):
        if g_loss is None:
            return None, None, None, None
        gA, gB, gS = ctx.step.backward(g_loss, ctx.generation)
        need = ctx.needs_input_grad
        return (gA if need[0] else None), (gB if need[1] else None), (gS if need[2] else None), None
