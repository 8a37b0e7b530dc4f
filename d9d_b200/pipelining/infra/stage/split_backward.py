"""Split backward (dI now, dW later) for zero-bubble schedules, on top of plain autograd graphs.

For one (stage, microbatch) the autograd graph hanging off the stage outputs is analysed once:

* *input-dependent* nodes are those from which a stage-input gradient sink is reachable — exactly the nodes the
  input pass has to execute;
* a *frontier* node is an input-dependent node with an edge into the input-independent part of the graph that leads
  to trainable parameters.  The gradient arriving at a frontier node is everything the deferred weight pass needs.

Input pass: pre-hooks on the frontier nodes capture those gradients, then ``autograd.backward(..., inputs=stage
inputs, retain_graph=True)`` runs with only :attr:`GradDirection.inputs` enabled so custom ops skip their dW work.
Weight pass: each frontier node is re-executed alone on its captured gradient (only :attr:`GradDirection.weight`
enabled) and what it emits along its input-independent edges seeds one autograd run over the parameter-only part
of the graph.

Same capability as reference ``d9d/pipelining/infra/stage/splitgrad.py``; organised around one forward DFS with
memoised reachability instead of reverse-graph closures.
"""

from __future__ import annotations

import dataclasses
from collections.abc import Iterable

import torch
from torch import nn
from torch.autograd.graph import GradientEdge, Node

from d9d_b200.core.autograd import GLOBAL_GRAD_CONTEXT, GradDirection


def grad_sink(t: torch.Tensor) -> Node | None:
    """The autograd node that receives the gradient of ``t``: its ``grad_fn`` or, for a leaf, its AccumulateGrad."""
    if not t.requires_grad:
        return None
    if t.grad_fn is not None:
        return t.grad_fn
    return t.view_as(t).grad_fn.next_functions[0][0]  # leaf: reach the AccumulateGrad node through a view


def backward_full(outputs: list[torch.Tensor], output_grads: list[torch.Tensor] | None, inputs: list[torch.Tensor]) -> list[torch.Tensor | None]:
    """Ordinary backward through the stage; returns (and clears) the gradients of the stage inputs."""
    with GLOBAL_GRAD_CONTEXT.with_directions(GradDirection.inputs, GradDirection.weight):
        torch.autograd.backward(tensors=outputs, grad_tensors=output_grads)
    grads = []
    for t in inputs:
        grads.append(t.grad)
        t.grad = None
    return grads


@dataclasses.dataclass
class _Frontier:
    node: Node
    #: indices into ``node.next_functions`` that lead (through input-independent nodes only) to trainable parameters
    weight_edges: list[int]
    captured: tuple[torch.Tensor | None, ...] | None = None


@dataclasses.dataclass
class DeferredWeightBackward:
    """State kept between the input pass and the weight pass of one microbatch."""

    frontier: list[_Frontier]
    params: list[nn.Parameter]
    keep_alive: list[torch.Tensor]  # stage outputs: own the retained graph


def _analyse(outputs: list[torch.Tensor], inputs: list[torch.Tensor], params: Iterable[nn.Parameter]) -> tuple[list[_Frontier], list[nn.Parameter]]:
    input_sinks = {n for t in inputs if (n := grad_sink(t)) is not None}
    param_of: dict[Node, nn.Parameter] = {}
    for p in params:
        if p.requires_grad and (n := grad_sink(p)) is not None:
            param_of[n] = p

    roots = [n for t in outputs if (n := grad_sink(t)) is not None]
    # iterative post-order DFS: children are finalised before their parents
    order: list[Node] = []
    seen: set[Node] = set()
    stack: list[tuple[Node, bool]] = [(r, False) for r in roots]
    while stack:
        node, expanded = stack.pop()
        if expanded:
            order.append(node)
            continue
        if node in seen:
            continue
        seen.add(node)
        stack.append((node, True))
        for child, _ in node.next_functions:
            if child is not None and child not in seen:
                stack.append((child, False))

    depends_on_input: dict[Node, bool] = {}
    leads_to_param: dict[Node, bool] = {}  # only meaningful for input-independent nodes
    for node in order:
        dep = node in input_sinks
        leads = node in param_of
        for child, _ in node.next_functions:
            if child is None:
                continue
            dep = dep or depends_on_input[child]
            leads = leads or (not depends_on_input[child] and leads_to_param[child])
        depends_on_input[node] = dep
        leads_to_param[node] = leads

    frontier: list[_Frontier] = []
    for node in order:
        if not depends_on_input[node]:
            continue
        edges = [i for i, (child, _) in enumerate(node.next_functions)
                 if child is not None and not depends_on_input[child] and leads_to_param[child]]
        if edges:
            frontier.append(_Frontier(node=node, weight_edges=edges))
    reachable = [param_of[n] for n in order if n in param_of]
    return frontier, reachable


def backward_input(outputs: list[torch.Tensor], output_grads: list[torch.Tensor] | None, inputs: list[torch.Tensor],
                   params: Iterable[nn.Parameter]) -> tuple[list[torch.Tensor | None], DeferredWeightBackward]:
    """Input pass: gradients w.r.t. the stage inputs only; everything needed for the weight pass is captured."""
    frontier, reachable = _analyse(outputs, inputs, params)
    handles = []
    for f in frontier:
        def capture(grad_outputs, _f=f):
            _f.captured = tuple(grad_outputs)

        handles.append(f.node.register_prehook(capture))
    if output_grads is None:
        output_grads = [torch.ones_like(o) for o in outputs]
    wanted = [t for t in inputs if t.requires_grad]
    try:
        with GLOBAL_GRAD_CONTEXT.with_directions(GradDirection.inputs):
            torch.autograd.backward(tensors=outputs, grad_tensors=output_grads, inputs=wanted, retain_graph=True)
    finally:
        for h in handles:
            h.remove()
    grads = []
    for t in inputs:
        grads.append(t.grad)
        t.grad = None
    return grads, DeferredWeightBackward(frontier=frontier, params=reachable, keep_alive=list(outputs))


def _run_node(node: Node, grad_outputs: tuple[torch.Tensor | None, ...]):
    """Execute one autograd node in isolation (saved tensors are still alive thanks to ``retain_graph``)."""
    if isinstance(node, torch.autograd.function.BackwardCFunction):
        return node.apply(*grad_outputs)  # python custom Function: calls the user's ``backward``
    return node(*grad_outputs)  # C++ node


def _conform(grad: torch.Tensor, child: Node, input_nr: int) -> torch.Tensor:
    """What the engine's output validation would do between two nodes: reduce broadcast dims, fix dtype."""
    meta = child._input_metadata[input_nr]  # noqa: SLF001
    shape = tuple(meta.shape)
    if tuple(grad.shape) != shape:
        grad = grad.sum_to_size(shape)
    if grad.dtype != meta.dtype and not getattr(meta, "is_nested_tensor", False):
        grad = grad.to(meta.dtype)
    return grad


def backward_weight(deferred: DeferredWeightBackward) -> None:
    """Weight pass.

    Every frontier node is re-executed in isolation on its captured incoming gradient (custom ops only compute their
    dW because just :attr:`GradDirection.weight` is enabled); the gradients it emits along its input-independent
    edges become the roots of ONE autograd run over the parameter-only part of the graph.  One run => every
    parameter accumulates (and fires its post-accumulate hook) exactly once per microbatch, and nothing on the
    input path is traversed twice.
    """
    roots: list[GradientEdge] = []
    root_grads: list[torch.Tensor] = []
    with GLOBAL_GRAD_CONTEXT.with_directions(GradDirection.weight):
        for f in deferred.frontier:
            if f.captured is None:
                continue  # no gradient reached this node during the input pass
            emitted = _run_node(f.node, f.captured)
            if not isinstance(emitted, tuple):
                emitted = (emitted,)
            for i in f.weight_edges:
                grad = emitted[i]
                if grad is None:
                    continue
                child, input_nr = f.node.next_functions[i]
                roots.append(GradientEdge(child, input_nr))
                root_grads.append(_conform(grad, child, input_nr))
            f.captured = None
        if roots and deferred.params:
            torch.autograd.backward(tensors=roots, grad_tensors=root_grads, inputs=deferred.params)
    deferred.frontier = []
    deferred.params = []
    deferred.keep_alive = []
