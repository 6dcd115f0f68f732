"""Tensor-level entry points with the reference extension's names and argument order.

`ms_deform_attn_forward` / `ms_deform_attn_backward` mirror the two functions the reference's
compiled module `MultiScaleDeformableAttention` exports (/root/reference/mdqe/models/ops/src/vision.cpp:13-16,
src/ms_deform_attn.h:20-61, src/cuda/ms_deform_attn_cuda.cu:20-153): same positional arguments, same
returned shapes, same failure modes (non-contiguous or CPU tensors and a batch that `im2col_step`
does not divide raise RuntimeError).  Underneath they pass raw device pointers and the current CUDA
stream to the C ABI of libmsda_b200.so.
"""
import torch

from . import _lib

_VALUE_DTYPES = (torch.float32, torch.bfloat16, torch.float64)


def _dtype_code(value, loc, aw, who):
    if value.dtype not in _VALUE_DTYPES:
        raise RuntimeError(f"{who}: unsupported value dtype {value.dtype} (float32, bfloat16, float64)")
    if loc.dtype != aw.dtype:
        raise RuntimeError(f"{who}: sampling_loc ({loc.dtype}) and attn_weight ({aw.dtype}) dtypes differ")
    if value.dtype == torch.float32 and loc.dtype == torch.float32:
        return _lib.MSDA_F32
    if value.dtype == torch.float64 and loc.dtype == torch.float64:
        return _lib.MSDA_F64
    if value.dtype == torch.bfloat16 and loc.dtype == torch.bfloat16:
        return _lib.MSDA_BF16
    if value.dtype == torch.bfloat16 and loc.dtype == torch.float32:
        return _lib.MSDA_BF16_LOC32
    raise RuntimeError(f"{who}: unsupported dtype combination value={value.dtype}, sampling_loc={loc.dtype}")


def _check_inputs(who, tensors):
    for name, t in tensors:
        if not t.is_cuda:
            # the reference dispatches on value.type().is_cuda() and AT_ERRORs otherwise (ms_deform_attn.h:38,60)
            raise RuntimeError(f"{who}: Not implemented on the CPU ({name} must be a CUDA tensor)")
        if not t.is_contiguous():
            raise RuntimeError(f"{who}: {name} tensor has to be contiguous")
    dev = tensors[0][1].device
    for name, t in tensors:
        if t.device != dev:
            raise RuntimeError(f"{who}: {name} is on {t.device}, expected {dev}")


def _dims(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step, who):
    if value.dim() != 4 or sampling_loc.dim() != 6 or attn_weight.dim() != 5:
        raise RuntimeError(f"{who}: expected value[N,S,M,D], sampling_loc[N,Lq,M,L,P,2], attn_weight[N,Lq,M,L,P]")
    N, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    if tuple(sampling_loc.shape) != (N, Lq, M, L, P, 2) or tuple(attn_weight.shape) != (N, Lq, M, L, P):
        raise RuntimeError(f"{who}: inconsistent shapes value={tuple(value.shape)} spatial_shapes={tuple(spatial_shapes.shape)} "
                           f"sampling_loc={tuple(sampling_loc.shape)} attn_weight={tuple(attn_weight.shape)}")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError(f"{who}: spatial_shapes and level_start_index must be int64")
    if tuple(spatial_shapes.shape) != (L, 2) or level_start_index.numel() != L:
        raise RuntimeError(f"{who}: spatial_shapes must be [L,2] and level_start_index [L]")
    step = min(N, int(im2col_step))
    if N > 0 and (step <= 0 or N % step != 0):
        # same contract as ms_deform_attn_cuda.cu:50-52
        raise RuntimeError(f"{who}: batch({N}) must divide im2col_step({step})")
    return N, S, M, D, L, Lq, P


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream_ptr(device):
    """cudaStream_t of torch's current stream on `device` (the raw getter when this torch has it: the Stream object costs microseconds,
    and an eager module step is host-bound -- tools/host_overhead_profile.py)."""
    if _raw_stream is not None and device.index is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream


class _on_device:
    """`with _on_device(dev)` that does nothing when `dev` already is the current device (the usual case; the context
    manager costs two device switches and several Python frames per call)."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if (device.index is None or device.index == torch.cuda.current_device()) else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    """-> Tensor[N, Lq, M*D]; drop-in for MultiScaleDeformableAttention.ms_deform_attn_forward."""
    who = "ms_deform_attn_forward"
    _check_inputs(who, [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                        ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)])
    N, S, M, D, L, Lq, P = _dims(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step, who)
    code = _dtype_code(value, sampling_loc, attn_weight, who)
    lib = _lib.load()
    with _on_device(value.device):
        out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        rc = lib.msda_forward(_stream_ptr(value.device), code, value.data_ptr(), spatial_shapes.data_ptr(),
                              level_start_index.data_ptr(), sampling_loc.data_ptr(), attn_weight.data_ptr(),
                              N, S, M, D, L, Lq, P, out.data_ptr())
    _lib.check(rc, who)
    return out


def packed_supported(value, n_levels, n_points, n_queries):
    """True when the paired-corner bf16 forward (msda_pack_value + msda_forward_packed) implements this configuration."""
    N, S, M, D = value.shape
    return (value.is_cuda and D == 32 and n_levels * n_points == 16 and n_levels <= 32 and value.data_ptr() % 16 == 0
            and N * 2 * S * M < 2 ** 32 and N * n_queries * M < 2 ** 31 // 64)


def pack_value(value, spatial_shapes, level_start_index):
    """value [N,S,M,32] (float32 or bfloat16) -> the paired-corner bf16 layout of csrc/msda_packed.cu (opaque uint8 tensor): the two
    x-neighbours of a bilinear sample share one 128-byte line, so the sampler gathers two lines per sample instead of four."""
    who = "pack_value"
    _check_inputs(who, [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index)])
    if value.dim() != 4 or value.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"{who}: value must be a float32 or bfloat16 [N,S,M,D] tensor")
    N, S, M, D = value.shape
    lib = _lib.load()
    nbytes = lib.msda_packed_value_bytes(N, S, M, D)
    if nbytes == 0:
        raise RuntimeError(f"{who}: the paired-corner layout needs D = 32 (got {D})")
    with _on_device(value.device):
        packed = torch.empty(nbytes, dtype=torch.uint8, device=value.device)
        rc = lib.msda_pack_value(_stream_ptr(value.device), _lib.MSDA_F32 if value.dtype == torch.float32 else _lib.MSDA_BF16,
                                 value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), N, S, M, D,
                                 spatial_shapes.shape[0], packed.data_ptr())
    _lib.check(rc, who)
    return packed


def ms_deform_attn_forward_packed(packed, value_shape, spatial_shapes, level_start_index, sampling_loc, attn_weight):
    """-> bfloat16 Tensor[N, Lq, M*32]: the forward on a tensor produced by ``pack_value`` (same shapes / level starts);
    sampling_loc / attn_weight bfloat16 or float32."""
    who = "ms_deform_attn_forward_packed"
    _check_inputs(who, [("packed", packed), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                        ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)])
    N, S, M, D = value_shape
    if sampling_loc.dim() != 6 or attn_weight.dim() != 5 or sampling_loc.dtype != attn_weight.dtype or \
            sampling_loc.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"{who}: sampling_loc [N,Lq,M,L,P,2] and attn_weight [N,Lq,M,L,P] must both be float32 or bfloat16")
    _, Lq, _, L, P, _ = sampling_loc.shape
    lib = _lib.load()
    if packed.numel() * packed.element_size() < lib.msda_packed_value_bytes(N, S, M, D):
        raise RuntimeError(f"{who}: packed tensor too small for value shape {tuple(value_shape)}")
    code = _lib.MSDA_BF16 if sampling_loc.dtype == torch.bfloat16 else _lib.MSDA_BF16_LOC32
    with _on_device(packed.device):
        out = torch.empty((N, Lq, M * D), dtype=torch.bfloat16, device=packed.device)
        rc = lib.msda_forward_packed(_stream_ptr(packed.device), code, packed.data_ptr(), spatial_shapes.data_ptr(),
                                     level_start_index.data_ptr(), sampling_loc.data_ptr(), attn_weight.data_ptr(),
                                     N, S, M, D, L, Lq, P, out.data_ptr())
    _lib.check(rc, who)
    return out


def new_backward_accumulator(value):
    """Zero-filled tensor the backward accumulates grad_value into: grad_value itself (fp32 / fp64) or the fp32 workspace of
    the bf16 modes.  Allocate it early (MSDeformAttnFunction does so on a side stream during the forward pass) and hand it to
    the backward entries as ``accumulator=`` -- the 20.9 MB zero-fill then leaves the critical path (MSDA_BWD_ACC_ZEROED)."""
    dtype = torch.float32 if value.dtype == torch.bfloat16 else value.dtype
    return torch.zeros(value.shape, dtype=dtype, device=value.device)


def _backward_buffers(value, code, accumulator, who):
    """-> (grad_value, workspace tensor or None, workspace bytes, flags)"""
    lib = _lib.load()
    N, S, M, D = value.shape
    ws_bytes = lib.msda_backward_workspace_bytes(code, N, S, M, D)
    if accumulator is None:
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=value.device) if ws_bytes else None
        return torch.empty_like(value), ws, ws_bytes, 0
    want = torch.float32 if ws_bytes else value.dtype
    if accumulator.dtype != want or accumulator.numel() != value.numel() or accumulator.device != value.device or \
            not accumulator.is_contiguous():
        raise RuntimeError(f"{who}: accumulator must be a contiguous zero-filled {want} tensor shaped like value")
    if ws_bytes:
        return torch.empty_like(value), accumulator, ws_bytes, _lib.BWD_ACC_ZEROED
    return accumulator.view(value.shape), None, 0, _lib.BWD_ACC_ZEROED


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step, accumulator=None):
    """-> [grad_value, grad_sampling_loc, grad_attn_weight]; drop-in for the extension's backward.
    ``accumulator``: optional result of new_backward_accumulator(value) (zero-filled ahead of time)."""
    who = "ms_deform_attn_backward"
    _check_inputs(who, [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                        ("sampling_loc", sampling_loc), ("attn_weight", attn_weight), ("grad_output", grad_output)])
    N, S, M, D, L, Lq, P = _dims(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step, who)
    if grad_output.dtype != value.dtype or grad_output.numel() != N * Lq * M * D:
        raise RuntimeError(f"{who}: grad_output must be {value.dtype} with {N * Lq * M * D} elements")
    code = _dtype_code(value, sampling_loc, attn_weight, who)
    lib = _lib.load()
    with _on_device(value.device):
        grad_value, ws, ws_bytes, flags = _backward_buffers(value, code, accumulator, who)
        grad_loc = torch.empty_like(sampling_loc)
        grad_aw = torch.empty_like(attn_weight)
        rc = lib.msda_backward_grouped_flags(_stream_ptr(value.device), code, value.data_ptr(), spatial_shapes.data_ptr(),
                                             level_start_index.data_ptr(), sampling_loc.data_ptr(), attn_weight.data_ptr(),
                                             grad_output.data_ptr(), N, S, M, D, 1, L, Lq, P, 1.0,
                                             grad_value.data_ptr(), grad_loc.data_ptr(), grad_aw.data_ptr(),
                                             ws.data_ptr() if ws is not None else None, ws_bytes, flags)
    _lib.check(rc, who)
    return [grad_value, grad_loc, grad_aw]


def tc_linear_forward_packed(x, weight, bias, row_mask, spatial_shapes, level_start_index, n_heads):
    """value_proj whose GEMM epilogue writes the sampler's paired-corner bf16 layout directly (tc_linear_forward_packed): x [N,S,in]
    fp32, weight [n_heads*32, in], bias or None, row_mask [N,S] or None -> opaque packed tensor (as pack_value would build from the
    fp32 projection).  Inference only."""
    who = "tc_linear_forward_packed"
    tensors = [("x", x), ("weight", weight), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index)]
    if bias is not None:
        tensors.append(("bias", bias))
    if row_mask is not None:
        tensors.append(("row_mask", row_mask))
    _check_inputs(who, tensors)
    if x.dim() != 3 or x.dtype != torch.float32 or weight.dtype != torch.float32 or tuple(weight.shape) != (n_heads * 32, x.shape[2]):
        raise RuntimeError(f"{who}: x must be fp32 [N,S,in] and weight fp32 [n_heads*32, in]; got {tuple(x.shape)}, {tuple(weight.shape)}")
    N, S, in_f = x.shape
    mask8 = None
    if row_mask is not None:
        if row_mask.numel() != N * S:
            raise RuntimeError(f"{who}: row_mask must have one entry per row of x")
        mask8 = row_mask.view(torch.uint8) if row_mask.dtype == torch.bool else row_mask.to(torch.uint8)
    lib = _lib.load()
    nbytes = lib.msda_packed_value_bytes(N, S, n_heads, 32)
    with _on_device(x.device):
        packed = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        rc = lib.tc_linear_forward_packed(_stream_ptr(x.device), x.data_ptr(), weight.data_ptr(), bias.data_ptr() if bias is not None else None,
                                          mask8.data_ptr() if mask8 is not None else None, N, S, in_f, n_heads,
                                          spatial_shapes.data_ptr(), level_start_index.data_ptr(), spatial_shapes.shape[0], packed.data_ptr())
    _lib.check(rc, who)
    return packed


def ms_deform_attn_fused_forward_packed_joint(packed, value_shape, spatial_shapes, level_start_index, reference_points, qproj, n_points,
                                              grid, mode, offset_scale, out_dtype=torch.float32):
    """The forward on a packed value tensor with the module's softmax / location arithmetic inside the kernel and the raw query
    projection ``qproj`` [N,Lq,row_stride] as input (msda_fused_forward_packed_joint) -> Tensor[N, Lq, M*32] of ``out_dtype``."""
    who = "ms_deform_attn_fused_forward_packed_joint"
    tensors = [("packed", packed), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
               ("reference_points", reference_points), ("qproj", qproj)]
    if grid is not None:
        tensors.append(("grid", grid))
    _check_inputs(who, tensors)
    N, S, M, D = value_shape
    L, P = spatial_shapes.shape[0], int(n_points)
    if qproj.dim() != 3 or qproj.shape[0] != N or qproj.dtype != torch.float32 or qproj.shape[2] < 3 * M * L * P or qproj.shape[2] % 4:
        raise RuntimeError(f"{who}: qproj must be fp32 [N,Lq,row_stride] with row_stride % 4 == 0 and >= 3*M*L*P, got {tuple(qproj.shape)}")
    Lq = qproj.shape[1]
    if reference_points.dim() != 3 or tuple(reference_points.shape[:2]) != (N, Lq) or reference_points.dtype != torch.float32:
        raise RuntimeError(f"{who}: reference_points must be fp32 [N,Lq,R]")
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"{who}: out_dtype must be float32 or bfloat16")
    lib = _lib.load()
    if packed.numel() * packed.element_size() < lib.msda_packed_value_bytes(N, S, M, D):
        raise RuntimeError(f"{who}: packed tensor too small for value shape {tuple(value_shape)}")
    with _on_device(packed.device):
        out = torch.empty((N, Lq, M * D), dtype=out_dtype, device=packed.device)
        rc = lib.msda_fused_forward_packed_joint(_stream_ptr(packed.device), packed.data_ptr(), spatial_shapes.data_ptr(),
                                                 level_start_index.data_ptr(), reference_points.data_ptr(), int(reference_points.shape[2]),
                                                 qproj.data_ptr(), int(qproj.shape[2]), grid.data_ptr() if grid is not None else None,
                                                 int(mode), float(offset_scale), N, S, M, D, L, Lq, P,
                                                 _lib.MSDA_F32 if out_dtype == torch.float32 else _lib.MSDA_BF16, out.data_ptr())
    _lib.check(rc, who)
    return out


def _fast_kernel_limits(value, n_queries, *others):
    """the conditions csrc/msda_launch.cuh (fast_eligible) puts on the fast kernels besides the shape family: 32-bit row
    offsets and pair indices, 16-byte aligned tensors"""
    if value.numel() >= 2 ** 31 or value.data_ptr() % 16:
        return False
    if n_queries is not None and value.shape[0] * n_queries * value.shape[2] >= 2 ** 31 // 64:
        return False
    return all(t is None or t.data_ptr() % 16 == 0 for t in others)


def grouped_supported(value, n_groups, n_levels, n_points, n_queries=None, *others):
    """True when msda_forward_grouped / msda_backward_grouped implement this configuration (fast kernels only);
    mirrors grouped_supported() / fast_eligible() of the library so that callers can fall back instead of failing."""
    return (value.is_cuda and value.dim() == 4 and value.shape[3] in (32, 24) and value.dtype in (torch.float32, torch.bfloat16)
            and n_levels * n_points in (8, 12, 16) and n_groups * n_levels <= 32 and _fast_kernel_limits(value, n_queries, *others))


def _grouped_dims(value, shapes, level_start, loc, aw, who):
    if shapes.dim() != 3 or shapes.shape[2] != 2 or tuple(level_start.shape) != tuple(shapes.shape[:2]):
        raise RuntimeError(f"{who}: expected spatial_shapes[G,L,2] and level_start_index[G,L]")
    if shapes.dtype != torch.int64 or level_start.dtype != torch.int64:
        raise RuntimeError(f"{who}: spatial_shapes and level_start_index must be int64")
    G, L = shapes.shape[0], shapes.shape[1]
    N, S, M, D = value.shape
    Lq, P = loc.shape[1], loc.shape[4]
    if tuple(loc.shape) != (N, Lq, M, L, P, 2) or tuple(aw.shape) != (N, Lq, M, L, P):
        raise RuntimeError(f"{who}: inconsistent shapes value={tuple(value.shape)} sampling_loc={tuple(loc.shape)} "
                           f"attn_weight={tuple(aw.shape)} for L={L}")
    return N, S, M, D, G, L, Lq, P


def ms_deform_attn_grouped_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, scale):
    """scale * sum_g msda(value, spatial_shapes[g], level_start_index[g], sampling_loc, attn_weight) in ONE launch
    (the clip-level decoder attention of ms_deform_attn.py:219-235).  -> Tensor[N, Lq, M*D]."""
    who = "ms_deform_attn_grouped_forward"
    _check_inputs(who, [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                        ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)])
    N, S, M, D, G, L, Lq, P = _grouped_dims(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, who)
    code = _dtype_code(value, sampling_loc, attn_weight, who)
    lib = _lib.load()
    with _on_device(value.device):
        out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        rc = lib.msda_forward_grouped(_stream_ptr(value.device), code, value.data_ptr(), spatial_shapes.data_ptr(),
                                      level_start_index.data_ptr(), sampling_loc.data_ptr(), attn_weight.data_ptr(),
                                      N, S, M, D, G, L, Lq, P, float(scale), out.data_ptr())
    _lib.check(rc, who)
    return out


def ms_deform_attn_grouped_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, scale,
                                    accumulator=None):
    who = "ms_deform_attn_grouped_backward"
    _check_inputs(who, [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                        ("sampling_loc", sampling_loc), ("attn_weight", attn_weight), ("grad_output", grad_output)])
    N, S, M, D, G, L, Lq, P = _grouped_dims(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, who)
    if grad_output.dtype != value.dtype or grad_output.numel() != N * Lq * M * D:
        raise RuntimeError(f"{who}: grad_output must be {value.dtype} with {N * Lq * M * D} elements")
    code = _dtype_code(value, sampling_loc, attn_weight, who)
    lib = _lib.load()
    with _on_device(value.device):
        grad_value, ws, ws_bytes, flags = _backward_buffers(value, code, accumulator, who)
        grad_loc, grad_aw = torch.empty_like(sampling_loc), torch.empty_like(attn_weight)
        rc = lib.msda_backward_grouped_flags(_stream_ptr(value.device), code, value.data_ptr(), spatial_shapes.data_ptr(),
                                             level_start_index.data_ptr(), sampling_loc.data_ptr(), attn_weight.data_ptr(),
                                             grad_output.data_ptr(), N, S, M, D, G, L, Lq, P, float(scale),
                                             grad_value.data_ptr(), grad_loc.data_ptr(), grad_aw.data_ptr(),
                                             ws.data_ptr() if ws is not None else None, ws_bytes, flags)
    _lib.check(rc, who)
    return [grad_value, grad_loc, grad_aw]


def _joint_dims(value, shapes, level_start, ref, qproj, grid, mode, n_points, who):
    if shapes.dim() == 2:
        shapes, level_start = shapes.unsqueeze(0), level_start.unsqueeze(0)
    if shapes.dim() != 3 or shapes.shape[2] != 2 or tuple(level_start.shape) != tuple(shapes.shape[:2]):
        raise RuntimeError(f"{who}: expected spatial_shapes[G,L,2] (or [L,2]) and matching level_start_index")
    G, L = shapes.shape[0], shapes.shape[1]
    N, S, M, D = value.shape
    P = int(n_points)
    if qproj.dim() != 3 or qproj.shape[0] != N or qproj.shape[2] < 3 * M * L * P or qproj.shape[2] % 4 or qproj.dtype != torch.float32:
        raise RuntimeError(f"{who}: qproj must be fp32 [N,Lq,row_stride] with row_stride % 4 == 0 and >= 3*M*L*P, got {tuple(qproj.shape)}")
    Lq = qproj.shape[1]
    if ref.dim() != 3 or tuple(ref.shape[:2]) != (N, Lq):
        raise RuntimeError(f"{who}: reference_points must be [N,Lq,R]")
    if mode == 1 and (grid is None or grid.numel() != M * L * P * 2 or ref.shape[2] != 4):
        raise RuntimeError(f"{who}: mode 1 needs grid[M,L,P,2] and 4-component reference points")
    return shapes, level_start, (N, S, M, D, G, L, Lq, P), int(ref.shape[2])


def ms_deform_attn_fused_forward_joint(value, spatial_shapes, level_start_index, reference_points, qproj, n_points, grid, mode,
                                       offset_scale, scale=1.0):
    """ms_deform_attn_fused_forward with offsets and logits given as column ranges of ONE matrix ``qproj`` [N,Lq,row_stride] -- the
    output of a single Linear layer over the concatenated sampling_offsets / attention_weights weights (msda_fused_forward_joint)."""
    who = "ms_deform_attn_fused_forward_joint"
    tensors = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
               ("reference_points", reference_points), ("qproj", qproj)]
    if grid is not None:
        tensors.append(("grid", grid))
    _check_inputs(who, tensors)
    shapes, starts, (N, S, M, D, G, L, Lq, P), R = _joint_dims(value, spatial_shapes, level_start_index, reference_points, qproj, grid,
                                                               mode, n_points, who)
    lib = _lib.load()
    with _on_device(value.device):
        out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        rc = lib.msda_fused_forward_joint(_stream_ptr(value.device), _lib.MSDA_F32, value.data_ptr(), shapes.data_ptr(), starts.data_ptr(),
                                          reference_points.data_ptr(), R, qproj.data_ptr(), int(qproj.shape[2]),
                                          grid.data_ptr() if grid is not None else None, int(mode), float(offset_scale),
                                          N, S, M, D, G, L, Lq, P, float(scale), out.data_ptr())
    _lib.check(rc, who)
    return out


def ms_deform_attn_fused_backward_joint(value, spatial_shapes, level_start_index, reference_points, qproj, n_points, grid, mode,
                                        offset_scale, grad_output, scale=1.0, accumulator=None):
    """-> (grad_value, grad_qproj); grad_qproj has the layout of qproj (columns beyond 3*M*L*P are zero)."""
    who = "ms_deform_attn_fused_backward_joint"
    tensors = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
               ("reference_points", reference_points), ("qproj", qproj), ("grad_output", grad_output)]
    if grid is not None:
        tensors.append(("grid", grid))
    _check_inputs(who, tensors)
    shapes, starts, (N, S, M, D, G, L, Lq, P), R = _joint_dims(value, spatial_shapes, level_start_index, reference_points, qproj, grid,
                                                               mode, n_points, who)
    lib = _lib.load()
    with _on_device(value.device):
        grad_value, _, _, flags = _backward_buffers(value, _lib.MSDA_F32, accumulator, who)
        exact = qproj.shape[2] == 3 * M * L * P
        grad_qproj = torch.empty_like(qproj) if exact else torch.zeros_like(qproj)
        rc = lib.msda_fused_backward_joint(_stream_ptr(value.device), _lib.MSDA_F32, value.data_ptr(), shapes.data_ptr(), starts.data_ptr(),
                                           reference_points.data_ptr(), R, qproj.data_ptr(), int(qproj.shape[2]),
                                           grid.data_ptr() if grid is not None else None, int(mode), float(offset_scale),
                                           grad_output.data_ptr(), N, S, M, D, G, L, Lq, P, float(scale),
                                           grad_value.data_ptr(), grad_qproj.data_ptr(), flags)
    _lib.check(rc, who)
    return grad_value, grad_qproj


def fused_supported(value, reference_points, n_groups, n_levels, n_points, n_queries=None, *others):
    """True when msda_fused_forward / msda_fused_backward implement this configuration."""
    return (value.is_cuda and value.dim() == 4 and value.dtype == torch.float32 and value.shape[3] in (32, 24)
            and n_levels * n_points in (8, 16) and n_groups * n_levels <= 32
            and reference_points.dtype == torch.float32 and reference_points.shape[-1] in (2, 4)
            and not reference_points.requires_grad and _fast_kernel_limits(value, n_queries, reference_points, *others))


def _fused_dims(value, shapes, level_start, ref, offsets, logits, grid, mode, who):
    if shapes.dim() == 2:
        shapes, level_start = shapes.unsqueeze(0), level_start.unsqueeze(0)
    if shapes.dim() != 3 or shapes.shape[2] != 2 or tuple(level_start.shape) != tuple(shapes.shape[:2]):
        raise RuntimeError(f"{who}: expected spatial_shapes[G,L,2] (or [L,2]) and matching level_start_index")
    G, L = shapes.shape[0], shapes.shape[1]
    N, S, M, D = value.shape
    if offsets.dim() != 6 or offsets.shape[0] != N or offsets.shape[2] != M or offsets.shape[3] != L or offsets.shape[5] != 2:
        raise RuntimeError(f"{who}: offsets must be [N,Lq,M,L,P,2], got {tuple(offsets.shape)}")
    Lq, P = offsets.shape[1], offsets.shape[4]
    if logits.numel() != N * Lq * M * L * P or ref.dim() != 3 or tuple(ref.shape[:2]) != (N, Lq):
        raise RuntimeError(f"{who}: logits must have N*Lq*M*L*P elements and reference_points be [N,Lq,R]")
    if mode == 1 and (grid is None or grid.numel() != M * L * P * 2 or ref.shape[2] != 4):
        raise RuntimeError(f"{who}: mode 1 needs grid[M,L,P,2] and 4-component reference points")
    return shapes, level_start, (N, S, M, D, G, L, Lq, P), int(ref.shape[2])


def ms_deform_attn_fused_forward(value, spatial_shapes, level_start_index, reference_points, offsets, logits, grid, mode,
                                 offset_scale, scale=1.0):
    """Sampler with the module's elementwise tail folded in (softmax over L*P, loc = ref + offsets/offset_scale, or the
    box-scaled grid form of the decoder); see msda_fused_forward in include/msda_b200.h.  -> Tensor[N, Lq, M*D]."""
    who = "ms_deform_attn_fused_forward"
    tensors = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
               ("reference_points", reference_points), ("offsets", offsets), ("logits", logits)]
    if grid is not None:
        tensors.append(("grid", grid))
    _check_inputs(who, tensors)
    shapes, starts, (N, S, M, D, G, L, Lq, P), R = _fused_dims(value, spatial_shapes, level_start_index, reference_points,
                                                               offsets, logits, grid, mode, who)
    lib = _lib.load()
    with _on_device(value.device):
        out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        rc = lib.msda_fused_forward(_stream_ptr(value.device), _lib.MSDA_F32, value.data_ptr(), shapes.data_ptr(), starts.data_ptr(),
                                    reference_points.data_ptr(), R, offsets.data_ptr(), logits.data_ptr(),
                                    grid.data_ptr() if grid is not None else None, int(mode), float(offset_scale),
                                    N, S, M, D, G, L, Lq, P, float(scale), out.data_ptr())
    _lib.check(rc, who)
    return out


def ms_deform_attn_fused_backward(value, spatial_shapes, level_start_index, reference_points, offsets, logits, grid, mode,
                                  offset_scale, grad_output, scale=1.0, accumulator=None):
    who = "ms_deform_attn_fused_backward"
    tensors = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
               ("reference_points", reference_points), ("offsets", offsets), ("logits", logits), ("grad_output", grad_output)]
    if grid is not None:
        tensors.append(("grid", grid))
    _check_inputs(who, tensors)
    shapes, starts, (N, S, M, D, G, L, Lq, P), R = _fused_dims(value, spatial_shapes, level_start_index, reference_points,
                                                               offsets, logits, grid, mode, who)
    lib = _lib.load()
    with _on_device(value.device):
        grad_value, _, _, flags = _backward_buffers(value, _lib.MSDA_F32, accumulator, who)
        grad_offsets, grad_logits = torch.empty_like(offsets), torch.empty_like(logits)
        rc = lib.msda_fused_backward_flags(_stream_ptr(value.device), _lib.MSDA_F32, value.data_ptr(), shapes.data_ptr(), starts.data_ptr(),
                                           reference_points.data_ptr(), R, offsets.data_ptr(), logits.data_ptr(),
                                           grid.data_ptr() if grid is not None else None, int(mode), float(offset_scale),
                                           grad_output.data_ptr(), N, S, M, D, G, L, Lq, P, float(scale),
                                           grad_value.data_ptr(), grad_offsets.data_ptr(), grad_logits.data_ptr(), flags)
    _lib.check(rc, who)
    return grad_value, grad_offsets, grad_logits


_MASK_CODES = {torch.float32: _lib.MSDA_F32, torch.bfloat16: _lib.MSDA_BF16, torch.float16: _lib.MSDA_F16}


def mask_logits_forward(coeff, proto, out_dtype=None):
    """coeff[B,Q,K] x proto[B,K,*plane] -> [B,Q,*plane]  ==  einsum('bqm,bmthw->bqthw')."""
    who = "mask_logits_forward"
    _check_inputs(who, [("coeff", coeff), ("proto", proto)])
    if coeff.dim() != 3 or proto.dim() < 3 or proto.shape[0] != coeff.shape[0] or proto.shape[1] != coeff.shape[2]:
        raise RuntimeError(f"{who}: expected coeff[B,Q,K] and proto[B,K,...], got {tuple(coeff.shape)} {tuple(proto.shape)}")
    if coeff.dtype != proto.dtype or coeff.dtype not in _MASK_CODES:
        raise RuntimeError(f"{who}: coeff/proto must both be float32, bfloat16 or float16")
    out_dtype = out_dtype or coeff.dtype
    if out_dtype not in _MASK_CODES:
        raise RuntimeError(f"{who}: unsupported out_dtype {out_dtype}")
    B, Q, K = coeff.shape
    plane = tuple(proto.shape[2:])
    ncols = 1
    for s in plane:
        ncols *= s
    lib = _lib.load()
    with _on_device(coeff.device):
        out = torch.empty((B, Q) + plane, dtype=out_dtype, device=coeff.device)
        rc = lib.mask_logits_forward(_stream_ptr(coeff.device), _MASK_CODES[coeff.dtype], _MASK_CODES[out_dtype],
                                     coeff.data_ptr(), proto.data_ptr(), B, Q, K, ncols, out.data_ptr())
    _lib.check(rc, who)
    return out


def mask_logits_backward(coeff, proto, grad_out, need_coeff=True, need_proto=True):
    """(grad_coeff, grad_proto) of mask_logits_forward.  The kernels are fp32 (the reference trains the mask head in fp32, AMP
    disabled: configs/R50_coco.yaml:41-42); 16-bit operands -- only met when a caller trains under autocast -- are widened,
    differentiated in fp32 and the gradients rounded back to the operand dtype."""
    who = "mask_logits_backward"
    _check_inputs(who, [("coeff", coeff), ("proto", proto), ("grad_out", grad_out)])
    if coeff.dtype != torch.float32:
        gc, gp = mask_logits_backward(coeff.float(), proto.float(), grad_out.float().contiguous(), need_coeff, need_proto)
        return (gc.to(coeff.dtype) if gc is not None else None), (gp.to(proto.dtype) if gp is not None else None)
    B, Q, K = coeff.shape
    ncols = proto.numel() // max(1, B * K)
    lib = _lib.load()
    with _on_device(coeff.device):
        gc = torch.empty_like(coeff) if need_coeff else None
        gp = torch.empty_like(proto) if need_proto else None
        rc = lib.mask_logits_backward(_stream_ptr(coeff.device), _MASK_CODES[coeff.dtype], coeff.data_ptr(),
                                      proto.data_ptr(), grad_out.data_ptr(), B, Q, K, ncols,
                                      gc.data_ptr() if gc is not None else None,
                                      gp.data_ptr() if gp is not None else None)
    _lib.check(rc, who)
    return gc, gp


def linear_supported(x, weight):
    """The tensor-core Linear needs fp32 CUDA tensors with in/out features that are multiples of 4, 16-byte aligned, and
    fewer than 2^31 - 128 rows (linear_forward_dispatch in csrc/mask_gemm.cu)."""
    return (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and weight.dim() == 2
            and weight.shape[0] % 4 == 0 and weight.shape[1] % 4 == 0 and x.shape[-1] == weight.shape[1]
            and x.data_ptr() % 16 == 0 and weight.data_ptr() % 16 == 0 and x.numel() // max(weight.shape[1], 1) < 2 ** 31 - 128)


def tc_linear_forward(x, weight, bias=None, row_mask=None):
    """y = x @ weight.T + bias with rows where ``row_mask`` is True zeroed -- ``F.linear`` (+ ``masked_fill(mask[..., None], 0)``,
    ms_deform_attn.py:136-138) as one 3xTF32 tensor-core GEMM.  x [..., in], weight [out, in], row_mask [...] bool."""
    who = "tc_linear_forward"
    tensors = [("x", x), ("weight", weight)]
    if bias is not None:
        tensors.append(("bias", bias))
    if row_mask is not None:
        tensors.append(("row_mask", row_mask))
    _check_inputs(who, tensors)
    if x.dtype != torch.float32 or weight.dtype != torch.float32 or (bias is not None and bias.dtype != torch.float32):
        raise RuntimeError(f"{who}: fp32 tensors expected")
    out_f, in_f = weight.shape
    if x.shape[-1] != in_f or (bias is not None and tuple(bias.shape) != (out_f,)):
        raise RuntimeError(f"{who}: shape mismatch x{tuple(x.shape)} weight{tuple(weight.shape)}")
    rows = x.numel() // max(in_f, 1)
    mask8 = None
    if row_mask is not None:
        if row_mask.numel() != rows:
            raise RuntimeError(f"{who}: row_mask must have one entry per row of x")
        mask8 = row_mask.view(torch.uint8) if row_mask.dtype == torch.bool else row_mask.to(torch.uint8)
    lib = _lib.load()
    with _on_device(x.device):
        y = torch.empty(x.shape[:-1] + (out_f,), dtype=torch.float32, device=x.device)
        rc = lib.tc_linear_forward(_stream_ptr(x.device), x.data_ptr(), weight.data_ptr(),
                                   bias.data_ptr() if bias is not None else None,
                                   mask8.data_ptr() if mask8 is not None else None, rows, in_f, out_f, y.data_ptr())
    _lib.check(rc, who)
    return y


def tc_linear_bias_grad(grad_y):
    """grad_bias = grad_y.reshape(-1, out_features).sum(0) in one pass at memory speed (csrc/msda_api.cu linear_bias_grad_kernel)."""
    who = "tc_linear_bias_grad"
    _check_inputs(who, [("grad_y", grad_y)])
    out_f = grad_y.shape[-1]
    rows = grad_y.numel() // max(out_f, 1)
    if out_f % 4 != 0 or grad_y.data_ptr() % 16 != 0 or grad_y.dtype != torch.float32:
        return grad_y.reshape(-1, out_f).sum(0)
    with _on_device(grad_y.device):
        gb = torch.empty(out_f, dtype=torch.float32, device=grad_y.device)
        rc = _lib.load().tc_linear_bias_grad(_stream_ptr(grad_y.device), grad_y.data_ptr(), rows, out_f, gb.data_ptr())
    _lib.check(rc, who)
    return gb


def tc_linear_backward(grad_y, x, weight, need_x=True, need_weight=True, need_bias=False):
    """(grad_x, grad_weight) of y = x @ weight.T; grad_y must already carry the row mask (masked rows zero).  With ``need_bias`` the
    result is (grad_x, grad_weight, grad_bias): the column sums of grad_y come out of the weight-gradient GEMM (tc_linear_backward_bias)."""
    who = "tc_linear_backward"
    _check_inputs(who, [("grad_y", grad_y), ("x", x), ("weight", weight)])
    out_f, in_f = weight.shape
    rows = x.numel() // max(in_f, 1)
    if grad_y.numel() != rows * out_f:
        raise RuntimeError(f"{who}: grad_y{tuple(grad_y.shape)} does not match x{tuple(x.shape)} / weight{tuple(weight.shape)}")
    lib = _lib.load()
    with _on_device(x.device):
        gx = torch.empty_like(x) if need_x else None
        gw = torch.empty_like(weight) if need_weight else None
        if need_bias:
            gb = torch.empty(out_f, dtype=torch.float32, device=x.device)
            rc = lib.tc_linear_backward_bias(_stream_ptr(x.device), grad_y.data_ptr(), x.data_ptr(), weight.data_ptr(), rows, in_f, out_f,
                                             gx.data_ptr() if gx is not None else None, gw.data_ptr() if gw is not None else None, gb.data_ptr())
            _lib.check(rc, who)
            return gx, gw, gb
        rc = lib.tc_linear_backward(_stream_ptr(x.device), grad_y.data_ptr(), x.data_ptr(), weight.data_ptr(), rows, in_f, out_f,
                                    gx.data_ptr() if gx is not None else None, gw.data_ptr() if gw is not None else None)
    _lib.check(rc, who)
    return gx, gw
