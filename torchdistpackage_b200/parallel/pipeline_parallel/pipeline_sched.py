"""1F1B (one-forward-one-backward) pipeline schedule with user-supplied stage functions.

API parity with the reference (parallel/pipeline_parallel/pipeline_sched.py:72-269):

    forward_backward(optimizer, fwd_fn, bwd_fn, inputs, num_microbatches=1, forward_only=False,
                     dtype=torch.bfloat16, scatter_gather_tensors=False)
    forward_eval(fwd_fn, inputs, dtype)

``fwd_fn`` receives the previous stage's output (first stage: nothing) followed by this stage's
own ``inputs`` sliced to the current micro-batch; a single argument is passed bare, several as a
list.  ``bwd_fn(output, output_grad)`` may be ``None`` (default: ``output.backward()`` on the last
stage, ``torch.autograd.backward(output, grad)`` elsewhere).  There is no engine object -- the
call replaces ``forward + backward`` in the user's loop, so heterogeneous stages (e.g. CLIP
towers) work.

Fixes relative to the reference: any pipeline depth works (the reference only works for pp=2
because ``tpc.is_first_in_pipeline_group`` is tested without being called during warm-up, :129);
multi-tensor stage boundaries are supported; communication never blocks the host
(see ``comm.py``).  Schedule: warm-up = ``pp_size - pp_rank - 1`` forwards, steady 1F1B,
cool-down backwards.  ``interleave``-free by design (only 1F1B exists in the reference).
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from ...dist.process_topo import tpc
from . import comm


def _forward_step(input_from_prev, mb_index: int, micro_bs: int, fwd_fn: Callable,
                  extra_inputs: List[torch.Tensor]):
    args = []
    if input_from_prev is not None:
        if isinstance(input_from_prev, torch.Tensor):
            args.append(input_from_prev)
        else:
            args.extend(input_from_prev)
    for t in extra_inputs:
        args.append(t[mb_index * micro_bs:(mb_index + 1) * micro_bs])
    return fwd_fn(args[0]) if len(args) == 1 else fwd_fn(args)


def _backward_step(input_obj, output_obj, output_grad, bwd_fn: Optional[Callable]):
    """Run backward for one micro-batch and return d(loss)/d(stage input)."""
    inputs = []
    if input_obj is not None:
        inputs = [input_obj] if isinstance(input_obj, torch.Tensor) else \
            [t for t in input_obj if t is not None]
        for t in inputs:
            if t.requires_grad:
                t.retain_grad()
    if bwd_fn is not None:
        bwd_fn(output_obj, output_grad)
    elif output_grad is None:
        output_obj.backward()
    else:
        outs = [output_obj] if isinstance(output_obj, torch.Tensor) else list(output_obj)
        grads = [output_grad] if isinstance(output_grad, torch.Tensor) else list(output_grad)
        # a stage may pass plain data along (e.g. a mask or an untouched input): nothing to
        # differentiate there, the gradient the next stage sent for it is dropped
        pairs = [(o, g) for o, g in zip(outs, grads) if o.requires_grad]
        if pairs:
            torch.autograd.backward(tensors=[o for o, _ in pairs],
                                    grad_tensors=[g for _, g in pairs])
    if input_obj is None:
        return None
    if isinstance(input_obj, torch.Tensor):
        return input_obj.grad
    return [t.grad for t in input_obj]


def _shapes_of(obj):
    if isinstance(obj, torch.Tensor):
        return obj.shape
    return [t.shape for t in obj]


_SHAPE_CACHE: dict = {}      # static_shapes=True: (config key) -> (ft_shapes, bt_shapes)


def clear_shape_cache() -> None:
    _SHAPE_CACHE.clear()


def forward_backward(optimizer, fwd_fn: Callable, bwd_fn: Optional[Callable], inputs,
                     num_microbatches: int = 1, forward_only: bool = False,
                     dtype: torch.dtype = torch.bfloat16, scatter_gather_tensors: bool = False,
                     static_shapes: Optional[bool] = None):
    """Run one mini-batch through this pipeline stage.  Returns the last micro-batch's output
    (the loss on the last stage).

    ``static_shapes=True`` (or ``TDP_PP_STATIC_SHAPES=1``): the caller promises that, for a given
    ``(num_microbatches, mini-batch size, dtype)``, the activations exchanged between stages have
    the same shapes on every call (the usual case in training).  The shape handshake -- a blocking
    metadata message plus a host read-back per call (reference comm.py:26-105 repeats it on every
    call) -- is then done once and cached, so steady-state steps issue no blocking communication
    and no host synchronisation.  All stages must pass the same setting."""
    if static_shapes is None:
        import os
        static_shapes = os.environ.get("TDP_PP_STATIC_SHAPES", "0") == "1"
    pp_size = tpc.get_group_size("pipe")
    pp_rank = tpc.get_group_rank("pipe")
    first = tpc.is_first_in_pipeline_group()
    last = tpc.is_last_in_pipeline_group()
    sg = scatter_gather_tensors

    if isinstance(inputs, torch.Tensor):
        inputs = [inputs]
    elif inputs is None:
        assert not first, "pipeline 1st stage should have valid inputs!"
        inputs = []
    micro_bs = inputs[0].size(0) // num_microbatches if len(inputs) > 0 else 0

    n_warmup = min(pp_size - pp_rank - 1, num_microbatches)
    n_steady = num_microbatches - n_warmup

    input_objs: List = []
    output_objs: List = []
    ft_shapes = None          # shapes arriving from the previous stage (handshake once per call)
    bt_shapes = None          # shapes of the grads coming back = shapes of what we send forward
    need_send_meta = True
    output_obj = None
    cache_key = None
    if static_shapes:
        # every stage derives the same key from information all stages share: the schedule
        # (micro-batch count), the pipe group and the dtype; the first stage's batch size reaches
        # later stages only through the shapes themselves, hence the caller's promise above
        cache_key = (id(tpc.get_group("pipe")), int(num_microbatches), str(dtype), bool(sg),
                     bool(forward_only))
        hit = _SHAPE_CACHE.get(cache_key)
        if hit is not None:
            ft_shapes, bt_shapes = hit
            need_send_meta = False

    if optimizer is not None:
        optimizer.zero_grad()

    def recv_fwd():
        nonlocal ft_shapes
        if first:
            return None
        ft_shapes = comm.recv_obj_meta(ft_shapes)
        return comm.recv_forward(ft_shapes, dtype=dtype, scatter_gather_tensors=sg)

    def note_output(out):
        nonlocal bt_shapes, need_send_meta
        if not last:
            bt_shapes = _shapes_of(out)
            need_send_meta = comm.send_obj_meta(out, need_send_meta)

    # ---- warm-up forwards
    for i in range(n_warmup):
        input_obj = recv_fwd()
        output_obj = _forward_step(input_obj, i, micro_bs, fwd_fn, inputs)
        note_output(output_obj)
        comm.send_forward(output_obj, scatter_gather_tensors=sg)
        if not forward_only:
            input_objs.append(input_obj)
            output_objs.append(output_obj)

    if n_steady > 0:
        input_obj = recv_fwd()

    # ---- steady state: one forward, one backward
    for i in range(n_steady):
        last_iter = i == n_steady - 1
        output_obj = _forward_step(input_obj, i + n_warmup, micro_bs, fwd_fn, inputs)
        note_output(output_obj)
        if forward_only:
            comm.send_forward(output_obj, scatter_gather_tensors=sg)
            if not last_iter:
                input_obj = comm.recv_forward(ft_shapes, dtype=dtype, scatter_gather_tensors=sg)
            continue
        output_grad = comm.send_forward_recv_backward(output_obj, bt_shapes, dtype=dtype,
                                                      scatter_gather_tensors=sg)
        input_objs.append(input_obj)
        output_objs.append(output_obj)
        in_b, out_b = input_objs.pop(0), output_objs.pop(0)
        input_grad = _backward_step(in_b, out_b, output_grad, bwd_fn)
        if last_iter:
            input_obj = None
            comm.send_backward(input_grad, scatter_gather_tensors=sg)
        else:
            input_obj = comm.send_backward_recv_forward(input_grad, ft_shapes, dtype=dtype,
                                                        scatter_gather_tensors=sg)

    # ---- cool-down backwards
    if not forward_only:
        for _ in range(n_warmup):
            in_b, out_b = input_objs.pop(0), output_objs.pop(0)
            output_grad = comm.recv_backward(bt_shapes, dtype=dtype, scatter_gather_tensors=sg)
            input_grad = _backward_step(in_b, out_b, output_grad, bwd_fn)
            comm.send_backward(input_grad, scatter_gather_tensors=sg)
    if cache_key is not None and cache_key not in _SHAPE_CACHE:
        _SHAPE_CACHE[cache_key] = (ft_shapes, bt_shapes)
    return output_obj


def forward_eval(fwd_fn: Callable, inputs, dtype: torch.dtype = torch.bfloat16, **kwargs):
    """Single forward pass through the pipeline (no micro-batching, no backward)."""
    args = []
    if not tpc.is_first_in_pipeline_group():
        shapes = comm.recv_obj_meta(None)
        prev = comm.recv_forward(shapes, dtype=dtype, scatter_gather_tensors=False)
        if isinstance(prev, torch.Tensor):
            args.append(prev)
        else:
            args.extend(prev)
    if isinstance(inputs, torch.Tensor):
        args.append(inputs)
    elif isinstance(inputs, (list, tuple)):
        args.extend(inputs)
    out = fwd_fn(args[0] if len(args) == 1 else args)
    if not tpc.is_last_in_pipeline_group():
        comm.send_obj_meta(out, True)
        comm.send_forward(out, scatter_gather_tensors=False)
    return out
