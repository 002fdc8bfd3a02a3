"""Point-to-point communication between pipeline stages.

Same 10-function surface as the reference (parallel/pipeline_parallel/comm.py:362-595):
``recv_forward``, ``recv_backward``, ``send_forward``, ``send_backward``,
``send_forward_recv_backward``, ``send_backward_recv_forward``, ``send_forward_recv_forward``,
``send_backward_recv_backward``, ``send_forward_backward_recv_forward_backward`` plus the shape
metadata handshake ``send_obj_meta`` / ``recv_obj_meta``.  Peers are the *global* ranks
``tpc.get_prev/next_global_rank('pipe')``.

B200-first differences:
* activations / grads stay on **NCCL p2p** (``batch_isend_irecv``) but are issued on a dedicated
  side stream; ordering against compute is by stream events, the host never blocks and there is
  no device-wide ``torch.cuda.synchronize()`` (the reference synchronises after every exchange,
  comm.py:326-327).  Pure sends (warm-up forward sends, cool-down backward sends) therefore
  overlap the next micro-batch's compute;
* shape metadata travels as ONE packed int64 message instead of ``1 + 2k`` blocking scalar sends
  (comm.py:26-63);
* works on CPU / gloo for tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple, Union

import torch
import torch.distributed as dist

from ...dist.process_topo import tpc

TensorOrList = Union[torch.Tensor, List[torch.Tensor], Tuple[torch.Tensor, ...]]
ShapeOrList = Union[torch.Size, Sequence[int], List[torch.Size]]

_META_LEN = 64          # packed metadata message: [n, nd0, d00, d01, ..., nd1, ...]
_P2P_STREAM = None


def _device() -> torch.device:
    if torch.cuda.is_available() and dist.get_backend() != "gloo":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _p2p_stream():
    global _P2P_STREAM
    if _P2P_STREAM is None:
        _P2P_STREAM = torch.cuda.Stream()
    return _P2P_STREAM


# ------------------------------------------------------------------------------------------
# metadata
# ------------------------------------------------------------------------------------------
def _pack_meta(obj: TensorOrList) -> torch.Tensor:
    tensors = [obj] if isinstance(obj, torch.Tensor) else list(obj)
    vals = [len(tensors)]
    for t in tensors:
        vals.append(t.dim())
        vals.extend(int(s) for s in t.shape)
    if len(vals) > _META_LEN:
        raise ValueError("pipeline metadata message too long")
    vals += [0] * (_META_LEN - len(vals))
    # negative n marks "single tensor (not a list)"
    if isinstance(obj, torch.Tensor):
        vals[0] = -1
    return torch.tensor(vals, dtype=torch.int64, device=_device())


def _unpack_meta(msg: torch.Tensor):
    vals = msg.tolist()
    n = vals[0]
    single = n < 0
    n = 1 if single else n
    shapes, i = [], 1
    for _ in range(n):
        nd = vals[i]
        shapes.append(torch.Size(vals[i + 1:i + 1 + nd]))
        i += 1 + nd
    return shapes[0] if single else shapes


def send_obj_meta(obj: TensorOrList, need_meta: bool = True, next_rank: Optional[int] = None) -> bool:
    """Tell the next stage the shape(s) it is about to receive.  Returns ``False`` so callers can
    write ``need_meta = send_obj_meta(obj, need_meta)`` and send only once (reference idiom)."""
    if need_meta:
        if next_rank is None:
            next_rank = tpc.get_next_global_rank("pipe")
        dist.send(_pack_meta(obj), next_rank)
    return False


def recv_obj_meta(obj_shape=None, prev_rank: Optional[int] = None):
    """Receive the shape(s) of the incoming activation(s) unless already known."""
    if obj_shape is not None:
        return obj_shape
    if prev_rank is None:
        prev_rank = tpc.get_prev_global_rank("pipe")
    msg = torch.empty(_META_LEN, dtype=torch.int64, device=_device())
    dist.recv(msg, prev_rank)
    return _unpack_meta(msg)


# ------------------------------------------------------------------------------------------
# scatter / gather over the tensor-parallel group (send 1/tp of the activation, rebuild by
# all-gather at the receiver -- comm.py:108-155 of the reference)
# ------------------------------------------------------------------------------------------
def _tp_size() -> int:
    return tpc.get_group_size("tensor") if tpc.is_mode_inited("tensor") else 1


def _chunk_numel(shape) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n // _tp_size()


def split_tensor_into_1d_equal_chunks(t: torch.Tensor) -> torch.Tensor:
    tp = _tp_size()
    flat = t.contiguous().view(-1)
    k = flat.numel() // tp
    r = tpc.get_group_rank("tensor") if tp > 1 else 0
    return flat[r * k:(r + 1) * k]


_SG_BUF = {}          # (group id, bytes) -> [SymmBuffer, use counter]: scatter-gather staging


def _symm_gather(chunk: torch.Tensor):
    """All-gather of a pipeline activation chunk over the tensor group on the package's own
    multicast kernel (one ``multimem.st`` fans out to every TP rank): two ping-pong halves of a
    symmetric buffer, the kernel's own barriers order the exchange.  Returns ``None`` when the
    group has no symmetric memory (cross-node TP, gloo, dtype)."""
    if chunk.dtype not in (torch.bfloat16, torch.float32) or (chunk.numel() * chunk.element_size()) % 16:
        return None
    from ...ops._loader import native
    from ...ops.symm import get_symm_group
    if native() is None:
        return None
    grp = tpc.get_group("tensor")
    sg = get_symm_group(grp)
    if not sg.enabled:
        return None
    tp = _tp_size()
    nb = chunk.numel() * chunk.element_size()
    key = (id(grp), nb)
    ent = _SG_BUF.get(key)
    if ent is None:
        ent = _SG_BUF[key] = [sg.alloc(2 * nb * tp), 0]
    buf, use = ent[0], ent[1] & 1
    ent[1] += 1
    off = use * nb * tp
    buf.all_gather(off, nb, chunk.contiguous().view(-1))
    return buf.view(off, (chunk.numel() * tp,), chunk.dtype).clone()


def gather_split_1d_tensor(chunk: torch.Tensor) -> torch.Tensor:
    tp = _tp_size()
    if tp == 1:
        return chunk
    out = torch.empty(chunk.numel() * tp, dtype=chunk.dtype, device=chunk.device)
    if chunk.is_cuda:
        got = _symm_gather(chunk)
        if got is not None:
            return got
        dist.all_gather_into_tensor(out, chunk.contiguous(), group=tpc.get_group("tensor"))
    else:
        parts = [torch.empty_like(chunk) for _ in range(tp)]
        dist.all_gather(parts, chunk.contiguous(), group=tpc.get_group("tensor"))
        out = torch.cat(parts)
    return out


# ------------------------------------------------------------------------------------------
# the exchange primitive
# ------------------------------------------------------------------------------------------
def _as_list(x):
    if x is None:
        return None, False
    if isinstance(x, torch.Tensor):
        return [x], True
    return list(x), False


def _alloc_recv(shapes, dtype, scatter_gather: bool):
    if shapes is None:
        raise AssertionError("receiving without a known shape: call recv_obj_meta first")
    single = isinstance(shapes, torch.Size) or (len(shapes) > 0 and isinstance(shapes[0], int))
    shape_list = [torch.Size(shapes)] if single else [torch.Size(s) for s in shapes]
    bufs = []
    for s in shape_list:
        if scatter_gather and _tp_size() > 1:
            bufs.append(torch.empty(_chunk_numel(s), dtype=dtype, device=_device()))
        else:
            bufs.append(torch.empty(s, dtype=dtype, device=_device()))
    return bufs, shape_list, single


def _communicate(object_send_next=None, object_send_prev=None, recv_prev: bool = False,
                 recv_next: bool = False, recv_prev_shape=None, recv_next_shape=None,
                 prev_rank: Optional[int] = None, next_rank: Optional[int] = None,
                 dtype: torch.dtype = torch.float32, scatter_gather_tensors: bool = False):
    """Exchange tensors with the neighbouring stages.  Returns
    ``(tensor(s)_from_prev, tensor(s)_from_next)`` (``None`` where nothing was requested)."""
    if prev_rank is None and (object_send_prev is not None or recv_prev):
        prev_rank = tpc.get_prev_global_rank("pipe")
    if next_rank is None and (object_send_next is not None or recv_next):
        next_rank = tpc.get_next_global_rank("pipe")

    send_next, _ = _as_list(object_send_next)
    send_prev, _ = _as_list(object_send_prev)
    sg = scatter_gather_tensors and _tp_size() > 1
    if sg:
        if send_next is not None:
            send_next = [split_tensor_into_1d_equal_chunks(t) for t in send_next]
        if send_prev is not None:
            send_prev = [split_tensor_into_1d_equal_chunks(t) for t in send_prev]

    from_prev = from_next = None
    prev_shapes = next_shapes = None
    prev_single = next_single = False
    if recv_prev:
        from_prev, prev_shapes, prev_single = _alloc_recv(recv_prev_shape, dtype, sg)
    if recv_next:
        from_next, next_shapes, next_single = _alloc_recv(recv_next_shape, dtype, sg)

    ops = []
    keep = []
    if send_prev is not None:
        for t in send_prev:
            t = t.contiguous(); keep.append(t)
            ops.append(dist.P2POp(dist.isend, t, prev_rank))
    if from_prev is not None:
        for t in from_prev:
            ops.append(dist.P2POp(dist.irecv, t, prev_rank))
    if from_next is not None:
        for t in from_next:
            ops.append(dist.P2POp(dist.irecv, t, next_rank))
    if send_next is not None:
        for t in send_next:
            t = t.contiguous(); keep.append(t)
            ops.append(dist.P2POp(dist.isend, t, next_rank))

    if ops:
        on_cuda = _device().type == "cuda"
        if on_cuda:
            cur = torch.cuda.current_stream()
            side = _p2p_stream()
            side.wait_stream(cur)                      # producer -> p2p stream (device side)
            with torch.cuda.stream(side):
                for r in dist.batch_isend_irecv(ops):
                    r.wait()                           # stream-level dependency, host continues
            for t in keep:
                t.record_stream(side)
            if from_prev is not None or from_next is not None:
                for t in (from_prev or []) + (from_next or []):
                    t.record_stream(cur)
                cur.wait_stream(side)                  # consumer waits only when data was received
        else:
            for r in dist.batch_isend_irecv(ops):
                r.wait()

    def finish(bufs, shapes, single):
        if bufs is None:
            return None
        outs = []
        for b, s in zip(bufs, shapes):
            if sg:
                b = gather_split_1d_tensor(b).view(s)
            outs.append(b.requires_grad_())
        return outs[0] if single else outs

    return finish(from_prev, prev_shapes, prev_single), finish(from_next, next_shapes, next_single)


# ------------------------------------------------------------------------------------------
# the wrappers (first / last stage short-circuit)
# ------------------------------------------------------------------------------------------
def recv_forward(input_tensor_shape, prev_rank=None, dtype=torch.float32,
                 scatter_gather_tensors=False):
    if tpc.is_first_in_pipeline_group():
        return None
    t, _ = _communicate(recv_prev=True, recv_prev_shape=input_tensor_shape, prev_rank=prev_rank,
                        dtype=dtype, scatter_gather_tensors=scatter_gather_tensors)
    return t


def recv_backward(output_grad_shape, next_rank=None, dtype=torch.float32,
                  scatter_gather_tensors=False):
    if tpc.is_last_in_pipeline_group():
        return None
    _, t = _communicate(recv_next=True, recv_next_shape=output_grad_shape, next_rank=next_rank,
                        dtype=dtype, scatter_gather_tensors=scatter_gather_tensors)
    return t


def send_forward(output_tensor, next_rank=None, scatter_gather_tensors=False):
    if not tpc.is_last_in_pipeline_group():
        _communicate(object_send_next=output_tensor, next_rank=next_rank,
                     scatter_gather_tensors=scatter_gather_tensors)


def send_backward(input_tensor_grad, prev_rank=None, scatter_gather_tensors=False):
    if not tpc.is_first_in_pipeline_group():
        _communicate(object_send_prev=input_tensor_grad, prev_rank=prev_rank,
                     scatter_gather_tensors=scatter_gather_tensors)


def send_forward_recv_backward(output_tensor, output_grad_shape, recv_next=True, next_rank=None,
                               dtype=torch.float32, scatter_gather_tensors=False):
    if tpc.is_last_in_pipeline_group():
        return None
    _, g = _communicate(object_send_next=output_tensor, recv_next=recv_next,
                        recv_next_shape=output_grad_shape, next_rank=next_rank, dtype=dtype,
                        scatter_gather_tensors=scatter_gather_tensors)
    return g


def send_backward_recv_forward(input_tensor_grad, input_tensor_shape, recv_prev=True,
                               prev_rank=None, dtype=torch.float32, scatter_gather_tensors=False):
    if tpc.is_first_in_pipeline_group():
        return None
    t, _ = _communicate(object_send_prev=input_tensor_grad, recv_prev=recv_prev,
                        recv_prev_shape=input_tensor_shape, prev_rank=prev_rank, dtype=dtype,
                        scatter_gather_tensors=scatter_gather_tensors)
    return t


def send_forward_recv_forward(output_tensor, input_tensor_shape, recv_prev=True, prev_rank=None,
                              next_rank=None, dtype=torch.float32, scatter_gather_tensors=False):
    t, _ = _communicate(object_send_next=output_tensor, recv_prev=recv_prev,
                        recv_prev_shape=input_tensor_shape, prev_rank=prev_rank,
                        next_rank=next_rank, dtype=dtype,
                        scatter_gather_tensors=scatter_gather_tensors)
    return t


def send_backward_recv_backward(input_tensor_grad, output_grad_shape, recv_next=True,
                                prev_rank=None, next_rank=None, dtype=torch.float32,
                                scatter_gather_tensors=False):
    _, g = _communicate(object_send_prev=input_tensor_grad, recv_next=recv_next,
                        recv_next_shape=output_grad_shape, prev_rank=prev_rank,
                        next_rank=next_rank, dtype=dtype,
                        scatter_gather_tensors=scatter_gather_tensors)
    return g


def send_forward_backward_recv_forward_backward(output_tensor, input_tensor_grad,
                                                input_tensor_shape, output_grad_shape,
                                                recv_prev=True, recv_next=True, prev_rank=None,
                                                next_rank=None, dtype=torch.float32,
                                                scatter_gather_tensors=False):
    t, g = _communicate(object_send_next=output_tensor, object_send_prev=input_tensor_grad,
                        recv_prev=recv_prev, recv_next=recv_next,
                        recv_prev_shape=input_tensor_shape, recv_next_shape=output_grad_shape,
                        prev_rank=prev_rank, next_rank=next_rank, dtype=dtype,
                        scatter_gather_tensors=scatter_gather_tensors)
    return t, g
