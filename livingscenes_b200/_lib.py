"""ctypes binding of the C ABI declared in include/livingscenes_b200.h.

There is NO CPU fallback: if the shared library cannot be loaded every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ls_b200.so")

LS_MAX_LAYERS = 8
LS_KNN_K = 16
LS_HEAD_C = 16
LS_CODE_FLOATS = 1028

c_float_p = C.POINTER(C.c_float)
c_i64_p = C.POINTER(C.c_int64)
c_i32_p = C.POINTER(C.c_int32)


class EncLayerDesc(C.Structure):
    _fields_ = [
        ("c_in", C.c_int32), ("c_out", C.c_int32), ("down_factor", C.c_int32),
        ("attention", C.c_int32), ("global_conv", C.c_int32), ("_pad", C.c_int32),
        ("w0", C.c_void_p), ("w_src", C.c_void_p), ("w_dst", C.c_void_p),
        ("w_g1", C.c_void_p), ("w_g2", C.c_void_p),
        ("w_src_tc", C.c_void_p), ("w_dst_tc", C.c_void_p), ("w_g1_tc", C.c_void_p),
    ]


class EncoderDesc(C.Structure):
    _fields_ = [
        ("num_layers", C.c_int32), ("c_dim", C.c_int32), ("center_pred", C.c_int32),
        ("center_pred_scale", C.c_int32), ("scale_factor", C.c_float), ("neg_slope", C.c_float),
        ("layers", EncLayerDesc * LS_MAX_LAYERS),
        ("w_conv_c", C.c_void_p), ("w_inv_t", C.c_void_p), ("w_fc0_t", C.c_void_p),
        ("w_lin1", C.c_void_p), ("w_short", C.c_void_p), ("w_act2", C.c_float), ("_pad", C.c_int32),
        ("w_conv_c_tc", C.c_void_p),
    ]


class EncoderIO(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("B", C.c_int32), ("N", C.c_int32), ("normalize", C.c_int32), ("_pad", C.c_int32),
        ("center", C.c_void_p), ("scale", C.c_void_p), ("z_so3", C.c_void_p), ("z_inv", C.c_void_p),
        ("packed", C.c_void_p),
        ("knn_idx", C.c_void_p * LS_MAX_LAYERS), ("fps_idx", C.c_void_p * LS_MAX_LAYERS),
        ("feat", C.c_void_p * LS_MAX_LAYERS),
        ("scale0", C.c_void_p), ("x_norm", C.c_void_p),
        ("force_knn_idx", C.c_void_p * LS_MAX_LAYERS), ("force_fps_idx", C.c_void_p * LS_MAX_LAYERS),
    ]


class DecoderDesc(C.Structure):
    _fields_ = [
        ("latent", C.c_int32), ("hidden", C.c_int32), ("n_layers", C.c_int32), ("latent_in", C.c_int32),
        ("w", C.c_void_p * 12), ("b", C.c_void_p * 12),
        ("w0_zinv", C.c_void_p), ("w4_zinv", C.c_void_p),
        ("out_dims", C.c_int32 * 12), ("in_dims", C.c_int32 * 12),
        ("w_tc", C.c_void_p * 12),
        ("wt", C.c_void_p * 12), ("wt_tc", C.c_void_p * 12),
        ("wt4_h", C.c_void_p), ("wt4_u", C.c_void_p), ("wt4_h_tc", C.c_void_p), ("wt4_u_tc", C.c_void_p),
    ]


_PROTOS = {
    "ls_version": (C.c_int, []),
    "ls_last_error": (C.c_char_p, []),
    "ls_profile_enable": (C.c_int, [C.c_int32]),
    "ls_profile_read": (C.c_int, [c_i32_p, c_i32_p, c_float_p, C.c_int32, c_i32_p]),
    "ls_kernel_launches": (C.c_int64, []),
    "ls_encoder_workspace_bytes": (C.c_int, [C.POINTER(EncoderDesc), C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "ls_encoder_forward": (C.c_int, [C.POINTER(EncoderDesc), C.POINTER(EncoderIO), C.c_void_p, C.c_size_t, C.c_void_p]),
    "ls_tc_packed_floats": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "ls_tc_pack_weights": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "ls_vn_linear": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                               C.c_int32, C.c_int32, C.c_void_p]),
    "ls_set_tensor_cores": (C.c_int, [C.c_int32]),
    "ls_set_gemm_variant": (C.c_int, [C.c_int32]),
    "ls_set_wave_bytes": (C.c_int, [C.c_int64]),
    "ls_set_fps_fma": (C.c_int, [C.c_int32]),
    "ls_set_overlap": (C.c_int, [C.c_int32]),
    "ls_set_knn_tensor_cores": (C.c_int, [C.c_int32, C.c_float]),
    "ls_knn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ls_knn_tc_workspace_bytes": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "ls_knn_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ls_fps": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ls_fps_workspace_bytes": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "ls_fps_ex": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_size_t, C.c_void_p]),
    "ls_fps_masked": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ls_match_workspace_bytes": (C.c_int, [c_i32_p, c_i32_p, C.c_int32, C.POINTER(C.c_size_t)]),
    "ls_seq_match": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, c_i32_p, c_i32_p, C.c_int32, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_size_t, C.c_void_p]),
    "ls_mutual_nn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, c_i32_p, c_i32_p, C.c_int32, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_size_t, C.c_void_p]),
    "ls_seq_match_scored": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ls_sinkhorn_match": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ls_kabsch_batched": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ls_kabsch_from_codes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ls_icp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_float,
                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ls_sdf_workspace_bytes": (C.c_int, [C.POINTER(DecoderDesc), C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "ls_sdf_backward_workspace_bytes": (C.c_int, [C.POINTER(DecoderDesc), C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "ls_sdf_backward": (C.c_int, [C.POINTER(DecoderDesc)] + [C.c_void_p] * 5 + [C.c_int32, C.c_int32] + [C.c_void_p] * 7 +
                        [C.c_size_t, C.c_void_p]),
    "ls_mise_init": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ls_mise_collect": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "ls_mise_points": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]),
    "ls_mise_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_float, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ls_mise_to_dense": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "ls_mcubes_workspace_bytes": (C.c_int, [C.c_int32, C.POINTER(C.c_size_t)]),
    "ls_mcubes_count": (C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ls_mcubes_emit": (C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ls_sdf_decode": (C.c_int, [C.POINTER(DecoderDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)

_lib = None
_lock = threading.Lock()
launch_count = 0  # number of C-ABI compute calls issued (bench bookkeeping)


def lib():
    """The loaded shared library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"livingscenes_b200: CUDA extension {LIB_PATH} is missing. Build it with "
                        "`python -m livingscenes_b200._build` (needs nvcc); there is no CPU fallback.")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in _PROTOS.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().ls_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else 'unknown error'}")


def ptr(t):
    """Device pointer of a tensor as an int (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"livingscenes_b200: `{name}` must be a CUDA tensor; this package has no CPU path "
            "(the CPU restatement lives in oracle/ and is test infrastructure only).")


STAGE_NAMES = ("normalize", "fps", "gather", "gemm_tables", "knn_edgeconv", "global_conv", "head", "knn_filter", "knn_rerank")


def profile_enable(on: bool) -> None:
    check(lib().ls_profile_enable(int(on)), "ls_profile_enable")


def profile_read():
    """[(stage_name, layer, ms)] of the last encoder forward (synchronise the stream first)."""
    n = 8192  # wave scheduling issues a stage once per wave
    st, ly, ms, cnt = (C.c_int32 * n)(), (C.c_int32 * n)(), (C.c_float * n)(), C.c_int32(0)
    check(lib().ls_profile_read(st, ly, ms, n, C.byref(cnt)), "ls_profile_read")
    return [(STAGE_NAMES[st[i]], int(ly[i]), float(ms[i])) for i in range(cnt.value)]


def kernel_launches() -> int:
    return int(lib().ls_kernel_launches())


def set_tensor_cores(on: bool) -> None:
    """Route the packed-weight GEMMs through the tcgen05 3xTF32 kernel (default) or the FP32 SIMT kernel."""
    check(lib().ls_set_tensor_cores(int(on)), "ls_set_tensor_cores")


def set_gemm_variant(variant: int) -> None:
    """3 (default): activations in tensor memory (TS-form MMAs, 256-row weight tiles for K >= 128); 2: persistent
    warp-specialised kernel with both operands in shared memory; 1: the round-1 one-CTA-per-tile kernel."""
    check(lib().ls_set_gemm_variant(int(variant)), "ls_set_gemm_variant")


def set_wave_bytes(nbytes: int) -> None:
    """Table bytes per wave of the {table GEMM -> EdgeConv} schedule (0: whole batch per launch, as in round 1)."""
    check(lib().ls_set_wave_bytes(int(nbytes)), "ls_set_wave_bytes")


def set_fps_fma(on: bool) -> None:
    """FPS squared distance FMA-contracted (pytorch3d's CUDA kernel under nvcc's default) instead of separately rounded."""
    check(lib().ls_set_fps_fma(int(on)), "ls_set_fps_fma")


def set_overlap(on: bool) -> None:
    """Side-stream overlap of independent encoder stages (default on)."""
    check(lib().ls_set_overlap(int(on)), "ls_set_overlap")


def set_knn_tensor_cores(on: bool, kappa_scale: float = 1.0) -> None:
    """kNN graph through the tcgen05 candidate filter + exact re-rank (default) or the exact brute-force tiles."""
    check(lib().ls_set_knn_tensor_cores(int(on), float(kappa_scale)), "ls_set_knn_tensor_cores")


def tc_pack(weight_dev: torch.Tensor) -> torch.Tensor:
    """Pack a row-major [R, ldw] fp32 device matrix (true K = its column count) for the tcgen05 GEMM."""
    assert weight_dev.is_cuda and weight_dev.dtype == torch.float32 and weight_dev.dim() == 2
    w = weight_dev.contiguous()
    R, K = w.shape
    n = C.c_size_t(0)
    check(lib().ls_tc_packed_floats(R, K, C.byref(n)), "ls_tc_packed_floats")
    out = torch.empty(n.value, dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        check(lib().ls_tc_pack_weights(w.data_ptr(), R, K, K, out.data_ptr(), stream_ptr(w.device)), "ls_tc_pack_weights")
    out._ls_keep = w
    return out
