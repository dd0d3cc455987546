"""ctypes loader for the product library gkr_b200/libgkr_b200.so (C ABI: include/gkr_b200.h).

The CUDA library is the ONLY compute path of this package: if it is missing or cannot be loaded,
importing anything that computes raises -- there is no CPU fallback and nothing here imports oracle/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# GKR_B200_LIB lets kernel A/B experiments load an alternative build of the same library
SO_PATH = os.environ.get("GKR_B200_LIB") or os.path.join(_HERE, "libgkr_b200.so")
_SRC = os.path.join(_HERE, "csrc")

GKR_N_KERNEL_CLASSES = 11
GKR_COMM_ID_BYTES = 256
KERNEL_CLASS_NAMES = ["gkr_round", "gkr_round_fused", "prod3_round", "prod3_round_fused", "wiring", "eq",
                      "mobius", "line", "other", "gkr_round_tail", "prod3_round_tail"]
STATUS = {0: "GKR_OK", -1: "GKR_ERR_INVALID", -2: "GKR_ERR_CUDA", -3: "GKR_ERR_OOM", -4: "GKR_ERR_RANGE",
          -5: "GKR_ERR_TRANSCRIPT", -6: "GKR_ERR_COMM", -7: "GKR_ERR_INTERNAL"}


class GkrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


class FrT(C.Structure):
    _fields_ = [("l", C.c_uint32 * 8)]


class LayerDesc(C.Structure):
    _fields_ = [("k_out", C.c_uint32), ("k_in", C.c_uint32), ("n_gates", C.c_uint32),
                ("type", C.c_void_p), ("left", C.c_void_p), ("right", C.c_void_p)]


class Job(C.Structure):
    _fields_ = [("n_layers", C.c_uint32), ("layers", C.POINTER(LayerDesc)), ("input_values", C.c_void_p)]


CHALLENGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(FrT), C.c_uint32, C.POINTER(FrT))


class Transcript(C.Structure):
    _fields_ = [("user", C.c_void_p), ("challenge", CHALLENGE_FN)]


class ProofC(C.Structure):
    _fields_ = [("n_layers", C.c_uint32), ("depth", C.c_uint32), ("k", C.POINTER(C.c_uint32)),
                ("n_rounds", C.c_uint64), ("round_off", C.POINTER(C.c_uint64)), ("msg_len", C.POINTER(C.c_uint8)),
                ("msgs", C.c_void_p), ("chal", C.c_void_p), ("q_off", C.POINTER(C.c_uint64)),
                ("q_len", C.POINTER(C.c_uint32)), ("q", C.c_void_p), ("z_off", C.POINTER(C.c_uint64)),
                ("z", C.c_void_p), ("r", C.c_void_p), ("d_len", C.c_uint64), ("d_coef", C.c_void_p),
                ("input_len", C.c_uint64), ("input_coef", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("transcript_seconds", C.c_double), ("wait_seconds", C.c_double)]


class Profile(C.Structure):
    _fields_ = [("launches", C.c_uint64 * GKR_N_KERNEL_CLASSES), ("ms", C.c_double * GKR_N_KERNEL_CLASSES),
                ("algo_bytes", C.c_double * GKR_N_KERNEL_CLASSES)]


# every symbol include/gkr_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "gkr_ctx_create", "gkr_ctx_destroy", "gkr_ctx_stream", "gkr_ctx_sync", "gkr_ctx_set_option", "gkr_last_error", "gkr_version", "gkr_mimc7_multi_hash", "gkr_mimc7_hash", "gkr_mimc7_round_constant",
    "gkr_frontend_compile", "gkr_frontend_compile_sym", "gkr_frontend_n_outputs", "gkr_frontend_output", "gkr_frontend_n_circuits", "gkr_frontend_n_public", "gkr_frontend_circuit", "gkr_frontend_destroy",
    "gkr_circuit_create", "gkr_circuit_destroy", "gkr_witness_create", "gkr_witness_eval", "gkr_witness_layer",
    "gkr_witness_destroy", "gkr_prove", "gkr_proof_free", "gkr_verify", "gkr_sumcheck_prod", "gkr_dev_table_synth", "gkr_dev_table_synth_strided", "gkr_comm_unique_id", "gkr_comm_init", "gkr_comm_init_shared", "gkr_comm_destroy", "gkr_comm_create",
    "gkr_sumcheck_prod_sharded",
    "gkr_dev_table_upload", "gkr_dev_table_download", "gkr_dev_table_free", "gkr_dev_table_eval", "gkr_fr_binop", "gkr_eq_table",
    "gkr_mobius", "gkr_line_restrict", "gkr_ctx_stats", "gkr_ctx_profile", "gkr_bench_field_mul", "gkr_fold_f64_constants", "gkr_selftest",
    "gkr_batch_create", "gkr_batch_load", "gkr_batch_prove", "gkr_batch_threads", "gkr_batch_lanes", "gkr_batch_set_option", "gkr_batch_simd_hash", "gkr_batch_destroy",
    "gkr_prove_many", "gkr_mimc7_multi_hash_many",
]

_LIB = None


def build(force: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if force and os.path.exists(SO_PATH):
        os.remove(SO_PATH)
    subprocess.check_call(["make", "-C", _SRC, "-s", "-j", str(min(8, os.cpu_count() or 1))])
    return SO_PATH


def lib():
    """Load libgkr_b200.so; raises if it is absent (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C gkr_b200/csrc`; gkr_b200 has no CPU fallback")
    L = C.CDLL(SO_PATH)
    L.gkr_last_error.restype = C.c_char_p
    L.gkr_version.restype = C.c_char_p
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.gkr_ctx_create.argtypes = [i32, C.POINTER(vp)]
    L.gkr_ctx_destroy.argtypes = [vp]
    L.gkr_ctx_destroy.restype = None
    L.gkr_ctx_stream.argtypes = [vp]
    L.gkr_ctx_stream.restype = vp
    L.gkr_ctx_sync.argtypes = [vp]
    L.gkr_ctx_set_option.argtypes = [vp, C.c_char_p, i32]
    L.gkr_mimc7_multi_hash.argtypes = [vp, u32, vp, vp]
    L.gkr_mimc7_hash.argtypes = [vp, vp, vp]
    L.gkr_mimc7_round_constant.argtypes = [u32, vp]
    L.gkr_frontend_compile.argtypes = [vp, C.c_size_t, vp, C.c_size_t, C.POINTER(vp)]
    if hasattr(L, "gkr_frontend_compile_sym"):
        L.gkr_frontend_compile_sym.argtypes = [vp, C.c_size_t, vp, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(vp)]
        L.gkr_frontend_n_outputs.argtypes = [vp]
        L.gkr_frontend_n_outputs.restype = u32
        L.gkr_frontend_output.argtypes = [vp, u32, C.POINTER(u32), vp, C.POINTER(C.c_char_p)]
    L.gkr_frontend_n_circuits.argtypes = [vp]
    L.gkr_frontend_n_circuits.restype = u32
    L.gkr_frontend_n_public.argtypes = [vp]
    L.gkr_frontend_n_public.restype = u32
    L.gkr_frontend_circuit.argtypes = [vp, u32, C.POINTER(u32), C.POINTER(C.POINTER(LayerDesc)), C.POINTER(u32), C.POINTER(vp)]
    L.gkr_frontend_destroy.argtypes = [vp]
    L.gkr_frontend_destroy.restype = None
    L.gkr_circuit_create.argtypes = [vp, u32, C.POINTER(LayerDesc), C.POINTER(vp)]
    L.gkr_circuit_destroy.argtypes = [vp]
    L.gkr_circuit_destroy.restype = None
    L.gkr_witness_create.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(vp)]
    L.gkr_witness_eval.argtypes = [vp, vp, vp, C.POINTER(vp)]
    L.gkr_witness_layer.argtypes = [vp, vp, u32, vp]
    L.gkr_witness_destroy.argtypes = [vp]
    L.gkr_witness_destroy.restype = None
    L.gkr_prove.argtypes = [vp, vp, vp, C.POINTER(Transcript), C.POINTER(C.POINTER(ProofC))]
    L.gkr_verify.argtypes = [vp, vp, C.POINTER(ProofC), vp, C.POINTER(Transcript), C.POINTER(i32)]
    L.gkr_proof_free.argtypes = [C.POINTER(ProofC)]
    L.gkr_proof_free.restype = None
    L.gkr_sumcheck_prod.argtypes = [vp, u32, u32, C.POINTER(vp), i32, C.POINTER(Transcript), vp, vp, vp, vp]
    L.gkr_dev_table_synth.argtypes = [vp, u64, u64, u64, C.POINTER(vp)]
    L.gkr_dev_table_synth_strided.argtypes = [vp, u64, u64, u64, u64, u64, C.POINTER(vp)]
    L.gkr_comm_unique_id.argtypes = [vp]
    L.gkr_comm_init.argtypes = [vp, i32, i32, vp]
    L.gkr_comm_init_shared.argtypes = [vp, i32, i32, C.c_char_p]
    L.gkr_comm_destroy.argtypes = [vp]
    if hasattr(L, "gkr_comm_create"):
        L.gkr_comm_create.argtypes = [i32, C.POINTER(i32), C.POINTER(vp)]
    L.gkr_comm_destroy.restype = None
    L.gkr_sumcheck_prod_sharded.argtypes = [vp, u32, u32, C.POINTER(vp), C.POINTER(Transcript), vp, vp, vp, vp]
    L.gkr_dev_table_upload.argtypes = [vp, vp, u64, C.POINTER(vp)]
    L.gkr_dev_table_download.argtypes = [vp, vp, u64, vp]
    L.gkr_dev_table_free.argtypes = [vp, vp]
    if hasattr(L, "gkr_dev_table_eval"):
        L.gkr_dev_table_eval.argtypes = [vp, vp, u32, vp, vp]
    L.gkr_dev_table_free.restype = None
    L.gkr_fr_binop.argtypes = [vp, i32, vp, vp, vp, u64]
    L.gkr_eq_table.argtypes = [vp, vp, u32, vp]
    L.gkr_mobius.argtypes = [vp, vp, u32, vp, C.POINTER(u32), C.POINTER(u32)]
    L.gkr_line_restrict.argtypes = [vp, vp, u32, vp, vp, vp]
    L.gkr_ctx_stats.argtypes = [vp, C.POINTER(Stats), i32]
    L.gkr_ctx_profile.argtypes = [vp, i32, C.POINTER(Profile)]
    L.gkr_bench_field_mul.argtypes = [vp, i32, i32, i32, C.POINTER(C.c_double)]
    if hasattr(L, "gkr_selftest"):
        L.gkr_selftest.argtypes = [vp, u32, vp, vp]
    if hasattr(L, "gkr_batch_create"):
        L.gkr_batch_create.argtypes = [i32, i32, i32, C.POINTER(vp)]
        L.gkr_batch_load.argtypes = [vp, C.POINTER(Job), C.c_size_t]
        L.gkr_batch_prove.argtypes = [vp, C.POINTER(C.POINTER(ProofC)), C.POINTER(C.c_double)]
        L.gkr_batch_threads.argtypes = [vp]
        L.gkr_batch_lanes.argtypes = [vp]
        L.gkr_batch_set_option.argtypes = [vp, C.c_char_p, i32]
        L.gkr_batch_destroy.argtypes = [vp]
        L.gkr_batch_destroy.restype = None
        L.gkr_prove_many.argtypes = [i32, C.POINTER(Job), C.c_size_t, i32, i32, C.POINTER(C.POINTER(ProofC))]
        L.gkr_mimc7_multi_hash_many.argtypes = [vp, vp, u32, u32, vp]
    if hasattr(L, "gkr_fold_f64_constants"):       # absent from older builds loaded through GKR_B200_LIB for A/B runs
        L.gkr_fold_f64_constants.argtypes = [vp, vp]
    _LIB = L
    return L


def check(rc: int):
    if rc != 0:
        raise GkrError(rc, (lib().gkr_last_error() or b"").decode())
