"""ctypes binding of libngpde.so (the C ABI in include/ngpde.h).

The product path has no fallback: if the shared library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NGPDE_LIB_PATH") or os.path.join(HERE, "libngpde.so")  # override: developer builds (tools/)

MAX_LAYERS = 8

# enums of include/ngpde.h
ACT = {"identity": 0, "relu": 1, "tanh": 2, "sigmoid": 3, "swish": 4, "gelu": 5, "softplus": 6, "elu": 7,
       "leakyrelu": 8}
AGGR = {"+": 0, "sum": 0, "mean": 1, "max": 2, "min": 3, "*": 4, "prod": 4}
FAMILY = {"explicit_edge_conv": 0, "vmh_conv": 1, "mppde_conv": 2, "gno_conv": 3}
IDX_I32, IDX_I64 = 0, 1
GA = {"rowptr": 0, "src": 1, "dst": 2, "perm": 3, "tptr": 4, "tpos": 5, "units32": 6, "units64": 7, "units128": 8,
      "gcn_colptr": 9, "gcn_rowval": 10, "gcn_tptr": 11, "gcn_tpos": 12}

EXPORTS = [
    "ngpde_version", "ngpde_last_error", "ngpde_graph_create", "ngpde_graph_destroy", "ngpde_graph_array", "ngpde_graph_array_copy",
    "ngpde_graph_num_nodes", "ngpde_graph_num_edges", "ngpde_aggregate", "ngpde_conv_workspace_bytes",
    "ngpde_conv_forward", "ngpde_conv_backward", "ngpde_explicit_edge_conv_forward",
    "ngpde_explicit_edge_conv_backward", "ngpde_vmh_conv_forward", "ngpde_vmh_conv_backward",
    "ngpde_mppde_conv_forward", "ngpde_mppde_conv_backward", "ngpde_gno_conv_forward", "ngpde_gno_conv_backward",
    "ngpde_gcn_workspace_bytes", "ngpde_gcn_conv_forward", "ngpde_gcn_conv_backward", "ngpde_axpy_stages",
    "ngpde_profile_enable", "ngpde_profile_read", "ngpde_set_option", "ngpde_rows_gather", "ngpde_rows_put",
    "ngpde_rows_segment_add", "ngpde_debug_buffer", "ngpde_conv_kernel_paths", "ngpde_debug_gemm",
    "ngpde_partition_create", "ngpde_partition_destroy", "ngpde_partition_array", "ngpde_morton_order",
    "ngpde_comm_unique_id", "ngpde_comm_init", "ngpde_comm_adopt", "ngpde_comm_destroy", "ngpde_halo_create",
    "ngpde_halo_destroy", "ngpde_halo_forward", "ngpde_halo_backward", "ngpde_allreduce_sum", "ngpde_adam_step",
    "ngpde_rprop_step", "ngpde_loss_workspace_bytes", "ngpde_mse_loss", "ngpde_logit_cross_entropy",
    "ngpde_cuda_graph_kernel_nodes", "ngpde_peer_allreduce_sum", "ngpde_edgeconv_ode_workspace_bytes",
    "ngpde_edgeconv_ode_forward", "ngpde_edgeconv_ode_adjoint", "ngpde_conv_state_bytes",
]
PA = {"bounds": 0, "halo_global": 1, "recv_counts": 2, "send_counts": 3, "send_local": 4, "s_local": 5, "t_local": 6,
      "edge_ids": 7, "seg_rows": 8, "seg_ptr": 9, "seg_pos": 10, "peer_recv_offset": 11}
UNIQUE_ID_BYTES = 128


class Mlp(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dims", C.c_int32 * (MAX_LAYERS + 1)), ("act", C.c_int32 * MAX_LAYERS),
                ("has_bias", C.c_int32 * MAX_LAYERS)]


class ConvDesc(C.Structure):
    _fields_ = [("family", C.c_int32), ("aggr", C.c_int32), ("dx", C.c_int32), ("dhs", C.c_int32),
                ("dpos", C.c_int32), ("de", C.c_int32), ("dtheta", C.c_int32), ("gno_in", C.c_int32),
                ("gno_out", C.c_int32), ("phi", Mlp), ("node", Mlp)]


class ConvIO(C.Structure):
    _fields_ = [("x", C.c_void_p), ("snode", C.c_void_p), ("edata", C.c_void_p), ("theta", C.c_void_p),
                ("phi_params", C.c_void_p), ("node_params", C.c_void_p), ("mbar", C.c_void_p), ("y", C.c_void_p),
                ("dy", C.c_void_p), ("dx", C.c_void_p), ("dphi_params", C.c_void_p), ("dnode_params", C.c_void_p),
                ("state", C.c_void_p)]


class GcnDesc(C.Structure):
    _fields_ = [("in_chs", C.c_int32), ("out_chs", C.c_int32), ("act", C.c_int32), ("has_bias", C.c_int32),
                ("add_self_loops", C.c_int32), ("use_edge_weight", C.c_int32)]


class RkTableau(C.Structure):
    _fields_ = [("n_stages", C.c_int32), ("a", (C.c_float * 8) * 8), ("b", C.c_float * 8)]


class NgpdeError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load libngpde.so; raise loudly when it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NgpdeError(f"{LIB_PATH} is missing: run `python neuralgraphpde.jl_b200/build.py` "
                         "(or __graft_entry__.build()). There is no CPU fallback for the message-passing path.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    lib.ngpde_version.restype = C.c_int
    lib.ngpde_last_error.restype = C.c_char_p
    lib.ngpde_graph_create.argtypes = [C.POINTER(vp), i64, i64, vp, vp, i32, i32, i32, i64, vp]
    lib.ngpde_graph_destroy.argtypes = [vp]
    lib.ngpde_graph_array.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(i64)]
    lib.ngpde_graph_array_copy.argtypes = [vp, i32, i32, vp, i64, vp]
    lib.ngpde_graph_num_nodes.argtypes = [vp]
    lib.ngpde_graph_num_nodes.restype = i64
    lib.ngpde_graph_num_edges.argtypes = [vp]
    lib.ngpde_graph_num_edges.restype = i64
    lib.ngpde_aggregate.argtypes = [vp, i32, vp, i32, vp, vp, vp]
    lib.ngpde_conv_workspace_bytes.argtypes = [vp, C.POINTER(ConvDesc), i32]
    lib.ngpde_conv_workspace_bytes.restype = sz
    lib.ngpde_conv_state_bytes.argtypes = [vp, C.POINTER(ConvDesc)]
    lib.ngpde_conv_state_bytes.restype = sz
    conv_sig = [vp, C.POINTER(ConvDesc), C.POINTER(ConvIO), vp, sz, vp]
    for name in ("conv", "explicit_edge_conv", "vmh_conv", "mppde_conv", "gno_conv"):
        getattr(lib, f"ngpde_{name}_forward").argtypes = conv_sig
        getattr(lib, f"ngpde_{name}_backward").argtypes = conv_sig
    lib.ngpde_gcn_workspace_bytes.argtypes = [vp, C.POINTER(GcnDesc), i32]
    lib.ngpde_gcn_workspace_bytes.restype = sz
    lib.ngpde_gcn_conv_forward.argtypes = [vp, C.POINTER(GcnDesc), vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.ngpde_gcn_conv_backward.argtypes = [vp, C.POINTER(GcnDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.ngpde_axpy_stages.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(C.c_float), i32, i64, vp]
    lib.ngpde_rows_gather.argtypes = [vp, vp, i64, i32, vp, vp]
    lib.ngpde_rows_put.argtypes = [vp, vp, vp, vp, i32, i64, i32, vp]
    lib.ngpde_rows_segment_add.argtypes = [vp, vp, vp, vp, vp, i64, i32, vp]
    lib.ngpde_debug_buffer.argtypes = [vp]
    lib.ngpde_set_option.argtypes = [i32, i32]
    lib.ngpde_profile_enable.argtypes = [i32]
    lib.ngpde_profile_read.argtypes = [C.POINTER(C.c_double), C.POINTER(i64)]
    lib.ngpde_conv_kernel_paths.argtypes = [vp, C.POINTER(ConvDesc), C.POINTER(i32)]
    lib.ngpde_debug_gemm.argtypes = [vp, i32, i32, vp, i32, i32, vp, i32, i64, i32, i64, i32, vp, i32, vp]
    f32, cp = C.c_float, C.c_char_p
    lib.ngpde_partition_create.argtypes = [C.POINTER(vp), i64, i64, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.ngpde_partition_destroy.argtypes = [vp]
    lib.ngpde_partition_array.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(i64)]
    lib.ngpde_morton_order.argtypes = [vp, i64, i32, vp]
    lib.ngpde_comm_unique_id.argtypes = [vp, cp]
    lib.ngpde_comm_init.argtypes = [C.POINTER(vp), vp, i32, i32, cp]
    lib.ngpde_comm_adopt.argtypes = [C.POINTER(vp), vp, i32, i32, cp]
    lib.ngpde_comm_destroy.argtypes = [vp]
    lib.ngpde_halo_create.argtypes = [C.POINTER(vp), vp, vp, vp]
    lib.ngpde_halo_destroy.argtypes = [vp]
    lib.ngpde_halo_forward.argtypes = [vp, vp, i32, vp, vp]
    lib.ngpde_halo_backward.argtypes = [vp, vp, i32, vp, vp]
    lib.ngpde_allreduce_sum.argtypes = [vp, vp, i64, vp]
    lib.ngpde_adam_step.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, f32, vp]
    lib.ngpde_rprop_step.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, vp]
    lib.ngpde_loss_workspace_bytes.argtypes = []
    lib.ngpde_loss_workspace_bytes.restype = sz
    lib.ngpde_mse_loss.argtypes = [vp, vp, i64, vp, vp, vp, sz, vp]
    lib.ngpde_logit_cross_entropy.argtypes = [vp, i64, i32, vp, vp, i64, vp, vp, vp, sz, vp]
    lib.ngpde_peer_allreduce_sum.argtypes = [vp, i32, vp, i64, vp]
    lib.ngpde_edgeconv_ode_workspace_bytes.argtypes = [vp, C.POINTER(ConvDesc), C.POINTER(RkTableau)]
    lib.ngpde_edgeconv_ode_workspace_bytes.restype = sz
    lib.ngpde_edgeconv_ode_forward.argtypes = [vp, C.POINTER(ConvDesc), C.POINTER(RkTableau), C.c_float, i32, vp, vp, vp, vp, vp, sz, vp]
    lib.ngpde_edgeconv_ode_adjoint.argtypes = [vp, C.POINTER(ConvDesc), C.POINTER(RkTableau), C.c_float, i32, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.ngpde_cuda_graph_kernel_nodes.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    _lib = lib
    return lib


OPT_TENSOR_CORES = 0
OPT_GNO_FACTORED = 1
OPT_HOIST = 3
OPT_LAYERED = 4
OPT_GNO_LAYERED = 5


def set_option(option: int, value: int) -> None:
    check(load().ngpde_set_option(option, value))


PROF_SLOTS = ("fwd_edge", "fwd_node", "bwd_node", "bwd_edge")


def profile_enable(on: bool) -> None:
    check(load().ngpde_profile_enable(1 if on else 0))


def profile_read() -> dict:
    """{slot: (total_ms, launches)} of the fused kernels since the last read (synchronises their events)."""
    ms = (C.c_double * 4)()
    n = (C.c_int64 * 4)()
    check(load().ngpde_profile_read(ms, n))
    return {k: (ms[i], int(n[i])) for i, k in enumerate(PROF_SLOTS)}


def kernel_paths(graph_handle, desc) -> dict:
    """{slot: 1 (tcgen05 kernels) | 0 (FP32-FFMA engine) | 2 (factored GNOConv: FFMA edge kernel + dense GEMMs) |
    -1 (no such phase)} for a conv descriptor on a graph."""
    out = (C.c_int32 * 4)()
    check(load().ngpde_conv_kernel_paths(graph_handle, C.byref(desc), out))
    return {k: int(out[i]) for i, k in enumerate(PROF_SLOTS)}


def cuda_graph_kernel_nodes(raw_cuda_graph) -> tuple:
    """(kernel nodes, all nodes) of a captured cudaGraph_t (torch.cuda.CUDAGraph(keep_graph=True).raw_cuda_graph())."""
    nk, nn = C.c_int64(), C.c_int64()
    check(load().ngpde_cuda_graph_kernel_nodes(C.c_void_p(int(raw_cuda_graph)), C.byref(nk), C.byref(nn)))
    return int(nk.value), int(nn.value)


def check(rc: int) -> None:
    if rc != 0:
        msg = load().ngpde_last_error().decode("utf-8", "replace")
        raise NgpdeError(f"libngpde error {rc}: {msg}")


def make_mlp(spec) -> Mlp:
    """spec: list of (in, out, activation-name, has_bias)."""
    m = Mlp()
    if len(spec) > MAX_LAYERS:
        raise NgpdeError(f"at most {MAX_LAYERS} Dense layers per MLP are supported, got {len(spec)}")
    m.n_layers = len(spec)
    for i, (din, dout, act, hb) in enumerate(spec):
        if i > 0 and spec[i - 1][1] != din:
            raise NgpdeError(f"DimensionMismatch: layer {i + 1} expects {din} inputs, previous layer emits {spec[i - 1][1]}")
        m.dims[i] = din
        m.dims[i + 1] = dout
        if act not in ACT:
            raise NgpdeError(f"unsupported activation {act!r}; supported: {sorted(ACT)}")
        m.act[i] = ACT[act]
        m.has_bias[i] = 1 if hb else 0
    return m
