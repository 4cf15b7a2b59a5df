"""ctypes binding of include/geoldm_b200.h.  There is no CPU fallback: if the shared library is
missing or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GEOLDM_B200_LIB") or os.path.join(HERE, "csrc", "libgeoldm_b200.so")   # env: A/B builds

MAX_LAYERS, MAX_SUBLAYERS = 16, 4
MMA_FP32_SIMT, MMA_3XTF32, MMA_TF32, MMA_BF16, MMA_3XF16 = 0, 1, 2, 3, 4
MMA_MODES = {"fp32": MMA_FP32_SIMT, "3xtf32": MMA_3XTF32, "tf32": MMA_TF32, "bf16": MMA_BF16,
             "3xf16": MMA_3XF16}

fp = C.c_void_p  # device pointers travel as integers


class EgnnConfig(C.Structure):
    _fields_ = [("hidden_nf", C.c_int), ("n_layers", C.c_int), ("inv_sublayers", C.c_int),
                ("in_node_nf", C.c_int), ("out_node_nf", C.c_int), ("attention", C.c_int), ("tanh", C.c_int),
                ("norm_constant", C.c_float), ("coords_range", C.c_float), ("agg_div", C.c_float),
                ("mma_mode", C.c_int)]


class EdgeMlp(C.Structure):
    _fields_ = [("pq_wt", fp), ("pq_b", fp), ("w_rd", fp), ("w2t", fp), ("b2", fp), ("w_out", fp), ("b_out", fp),
                ("tc_pack", fp), ("tc_pack_pq", fp)]


class Gcl(C.Structure):
    _fields_ = [("edge", EdgeMlp), ("node_w1t", fp), ("node_b1", fp), ("node_w2t", fp), ("node_b2", fp),
                ("tc_pack_node1", fp), ("tc_pack_node2", fp)]


class Block(C.Structure):
    _fields_ = [("gcl", Gcl * MAX_SUBLAYERS), ("equiv", EdgeMlp), ("tc_pack_pq4", fp), ("pq4_b", fp)]


class EgnnWeights(C.Structure):
    _fields_ = [("emb_w", fp), ("emb_b", fp), ("out_w", fp), ("out_b", fp), ("block", Block * MAX_LAYERS)]


class Batch(C.Structure):
    _fields_ = [("n_mol", C.c_int), ("n_node", C.c_int), ("n_edge", C.c_int), ("n_tile", C.c_int),
                ("tile_m", C.c_int), ("mol_off", fp), ("node_mol", fp), ("edge_i", fp), ("edge_j", fp),
                ("tile_row", fp), ("tile_meta", fp)]


_SIGS = {
    "geoldm_abi_version": (C.c_int, []),
    "geoldm_last_error": (C.c_char_p, []),
    "geoldm_has_tcgen05": (C.c_int, []),
    "geoldm_batch_tile_meta": (C.c_int, [C.POINTER(Batch), fp, fp]),
    "geoldm_node_chain": (C.c_int, [C.c_int, fp, fp, C.c_float, fp, fp, fp, fp, fp, fp, C.c_int, fp, fp, fp, C.c_int, fp]),
    "geoldm_egnn_workspace_bytes": (C.c_size_t, [C.POINTER(EgnnConfig), C.c_int, C.c_int]),
    "geoldm_egnn_forward": (C.c_int, [C.POINTER(EgnnConfig), C.POINTER(EgnnWeights), C.POINTER(Batch), fp, fp, fp, fp,
                                      fp, fp, C.c_size_t, fp]),
    "geoldm_dynamics_prep": (C.c_int, [C.POINTER(Batch), fp, fp, C.c_int, fp, fp, fp, fp, C.c_int, C.c_int, fp,
                                       C.c_int, fp, fp]),
    "geoldm_dynamics_finish_a": (C.c_int, [C.POINTER(Batch), fp, fp, fp]),
    "geoldm_dynamics_finish_b": (C.c_int, [C.POINTER(Batch), fp, fp, fp, C.c_int, C.c_int, fp, fp, C.c_int, fp]),
    "geoldm_sampler_update": (C.c_int, [C.POINTER(Batch), C.c_int, fp, fp, fp, fp, fp, C.c_size_t, C.c_int,
                                        C.c_uint64, fp, fp, fp, fp]),
    "geoldm_sampler_advance": (C.c_int, [fp, C.c_int, fp, C.c_int, fp]),
    "geoldm_edge_gcl": (C.c_int, [C.POINTER(EgnnConfig), C.POINTER(EdgeMlp), C.POINTER(Batch), fp, fp, fp, fp, fp]),
    "geoldm_edge_equiv": (C.c_int, [C.POINTER(EgnnConfig), C.POINTER(EdgeMlp), C.POINTER(Batch), fp, fp, fp, fp, fp]),
    "geoldm_edge_dist": (C.c_int, [C.POINTER(Batch), fp, fp, fp, C.c_float, fp]),
    "geoldm_edge_gcl_pre": (C.c_int, [C.POINTER(EgnnConfig), C.POINTER(EdgeMlp), C.POINTER(Batch), fp, C.c_int, fp, fp,
                                      fp, fp]),
    "geoldm_edge_equiv_pre": (C.c_int, [C.POINTER(EgnnConfig), C.POINTER(EdgeMlp), C.POINTER(Batch), fp, C.c_int, fp,
                                        fp, fp, fp, fp]),
    "geoldm_linear": (C.c_int, [fp, C.c_int, fp, C.c_int, C.c_float, fp, fp, fp, C.c_int, fp, C.c_int, C.c_int,
                                C.c_int, fp]),
    "geoldm_tc_pack_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "geoldm_tc_pack": (C.c_int, [C.c_int, fp, C.c_int, C.c_int, fp, fp]),
    "geoldm_tc_pack16_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "geoldm_tc_pack16": (C.c_int, [C.c_int, fp, C.c_int, C.c_int, fp, fp]),
    "geoldm_tc_pack16_t": (C.c_int, [C.c_int, fp, C.c_int, C.c_int, fp, fp]),
    "geoldm_tc_pack16_pair": (C.c_int, [C.c_int, fp, fp, fp, fp]),
    "geoldm_linear_tc_grad": (C.c_int, [C.c_int, fp, C.c_int, fp, fp, C.c_int, fp, C.c_int, fp]),
    "geoldm_linear_tc": (C.c_int, [C.c_int, C.c_int, fp, C.c_int, fp, C.c_int, C.c_float, fp, C.c_int, fp, fp, C.c_int,
                                   fp, C.c_int, fp]),
    "geoldm_tc_selftest": (C.c_int, [C.c_int, C.c_int, fp, fp, fp, C.c_int, C.c_int, fp, fp, fp]),
    "geoldm_gemm_tn": (C.c_int, [fp, C.c_int, fp, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_int, fp]),
    "geoldm_gemm_tn_bias": (C.c_int, [fp, C.c_int, fp, C.c_int, fp, C.c_int, fp, C.c_int, C.c_int, C.c_int, fp]),
    "geoldm_train_edge_act_fwd": (C.c_int, [C.c_int, C.c_int, fp, C.c_int, fp, fp, fp, fp, fp, fp, fp]),
    "geoldm_train_edge_act_bwd": (C.c_int, [C.c_int, C.c_int, fp, C.c_int, fp, fp, fp, fp, fp, fp, fp, fp, fp, fp, fp]),
    "geoldm_train_edge_tail_fwd": (C.c_int, [C.c_int, C.c_int, fp, fp, fp, fp, C.c_int, C.c_int, fp, C.c_float, fp, fp, fp]),
    "geoldm_train_edge_tail_bwd": (C.c_int, [C.c_int, C.c_int, fp, fp, fp, fp, C.c_int, C.c_int, fp, C.c_float, fp, fp, fp,
                                             fp, fp, fp, fp, fp, fp]),
    "geoldm_adamw_ema_step": (C.c_int, [fp, fp, C.c_int, fp, fp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                        C.c_int, C.c_float, fp]),
    "geoldm_optim_chunk": (C.c_int, []),
    "geoldm_train_coord2diff_fwd": (C.c_int, [C.c_int, fp, fp, fp, C.c_float, fp, fp, fp]),
    "geoldm_train_coord2diff_bwd": (C.c_int, [C.c_int, fp, fp, fp, C.c_float, fp, fp, fp, fp]),
    "geoldm_train_coord_step_fwd": (C.c_int, [C.c_int, fp, fp, fp, C.c_int, C.c_float, C.c_float, fp, fp]),
    "geoldm_train_coord_step_bwd": (C.c_int, [C.c_int, fp, fp, fp, C.c_int, C.c_float, C.c_float, fp, fp, fp, fp]),
    "geoldm_train_bwd_blocks": (C.c_int, [C.c_int]),
    "geoldm_stability": (C.c_int, [C.c_int, fp, fp, fp, C.c_int, fp, fp, C.c_int, fp, fp, fp]),
    "geoldm_tc_read_stats": (C.c_int, [C.POINTER(C.c_ulonglong)]),
    "geoldm_tc16_read_stats": (C.c_int, [C.POINTER(C.c_ulonglong)]),
    "geoldm_philox_normal": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, fp, C.c_int, fp]),
    "geoldm_decode": (C.c_int, [C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, fp]),
}
EXPORTS = tuple(_SIGS)

_lib = None


class GeoldmError(RuntimeError):
    pass


def lib():
    """Load (once) csrc/libgeoldm_b200.so; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GeoldmError(f"{LIB_PATH} is missing: run `python -m geoldm_b200.build` "
                              "(there is no CPU/PyTorch fallback for this path)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.geoldm_abi_version() != 4:
            raise GeoldmError("ABI version mismatch between _lib.py and libgeoldm_b200.so")
        _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise GeoldmError(f"{what} failed ({rc}): {lib().geoldm_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())
