"""ctypes binding of libjpdvt_sm100.so (C ABI declared in include/jpdvt_b200.h).

There is NO fallback: if the shared library is missing and cannot be built, or the device is not sm_100, every entry
point raises.  Non-zero C status codes become RuntimeError carrying jpdvt_last_error_string(), which keeps the
reference callers' `try/except Exception` per-image behaviour (image_model/inference.py:257,367-370).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libjpdvt_sm100.so")

c_void_p, c_int, c_int64, c_double = C.c_void_p, C.c_int, C.c_int64, C.c_double


class JpdvtError(RuntimeError):
    pass


class Weights(C.Structure):
    _fields_ = [("depth", C.c_int32), ("tokens", C.c_int32), ("image_size", C.c_int32), ("reserved", C.c_int32)] + [
        (n, c_void_p) for n in (
            "w_patch", "b_embed", "w_in_t", "pos", "t_w0", "t_b0", "t_w2", "t_b2", "w_ada", "b_ada",
            "w_qkv", "b_qkv", "w_proj", "b_proj", "w_fc1", "b_fc1", "w_fc2", "b_fc2",
            "w_final", "b_final", "w_head1", "b_head1", "w_head2", "b_head2")]


class Workspace(C.Structure):
    _fields_ = [("rows", C.c_int64), ("cond_rows", C.c_int32), ("step_rows", C.c_int32)] + [
        (n, c_void_p) for n in ("x", "xn", "qkv", "attn", "hid", "y", "y32", "c", "silu_c", "silu_c_bf16", "mod",
                                "w_fold", "fold_u", "fold_v", "row_stats", "c_steps", "silu_c_steps", "mod_steps", "te_hid",
                                "x_embed")]


class Sampler(C.Structure):
    _fields_ = [("num_steps", C.c_int32), ("chain", C.c_int32)] + [
        (n, c_void_p) for n in ("step_ids", "timestep_map", "coef1", "coef2", "logvar", "step_noise")] + [
        ("step_noise_stride", C.c_int64)] + [(n, c_void_p) for n in ("x0", "sample", "traj_x0", "traj_sample", "noise_key")]


class WeightsT(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("w_qkv_t", "w_proj_t", "w_fc1_t", "w_fc2_t", "w_final_t", "w_head1_t", "w_ada_t", "t_w2_t")]


class Tape(C.Structure):
    _fields_ = [("rows", C.c_int64), ("batch", C.c_int32), ("reserved", C.c_int32)] + [
        (n, c_void_p) for n in ("cols", "x", "xn1", "qkv", "lse2", "att", "y1", "xn2", "hpre", "h", "y2", "xnf", "yfin", "yfin32",
                                "headpre", "feat", "tpre", "c", "silu_c", "silu_c_bf16", "mod", "thid")]


GRAD_FIELDS = ("w_patch", "b_patch", "w_in", "b_in", "t_w0", "t_b0", "t_w2", "t_b2", "w_ada", "b_ada", "w_qkv", "b_qkv",
               "w_proj", "b_proj", "w_fc1", "b_fc1", "w_fc2", "b_fc2", "w_final", "b_final", "w_head1", "b_head1", "w_head2",
               "b_head2")


class Grads(C.Structure):
    _fields_ = [(n, c_void_p) for n in GRAD_FIELDS]


class BwdScratch(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("dx", "dxn", "dy", "dh", "dqkv", "datt", "dpre", "dmod", "dmod_bf16", "small_f32",
                                        "small_bf16", "wgrad_scratch", "part", "zeros")]


MAX_PEERS = 8
MAX_F32_RANGES = 16


class PeerStep(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("shard_begin", C.c_int64), ("shard_end", C.c_int64),
                ("epoch", C.c_uint32), ("timeout_ms", C.c_uint32), ("grads", c_void_p * MAX_PEERS),
                ("weights_bf16", c_void_p * MAX_PEERS), ("params", c_void_p * MAX_PEERS), ("signals", c_void_p * MAX_PEERS),
                ("grads_mc", c_void_p), ("weights_mc", c_void_p), ("params_mc", c_void_p), ("n_f32_ranges", C.c_int32),
                ("reserved", C.c_int32), ("f32_ranges", C.c_int64 * (2 * MAX_F32_RANGES)), ("local_sync", c_void_p),
                ("status", c_void_p), ("epoch_dev", c_void_p)]


# JPDVT_ABI_VERSION in include/jpdvt_b200.h (2: LayerNorm-fold buffers, 3: per-step conditioning tables, 4: timestep-MLP
# scratch of its own, hoisted embedding, in-kernel Philox noise, loss kernels, tcgen05 attention backward, 5: peer-memory
# optimizer step, 6: device-side step count / barrier token for the graph-replayed training step - jpdvt_peer_step.epoch_dev,
# jpdvt_adamw_ema_dev, jpdvt_adamw_ema_peer_dev)
ABI_VERSION = 6

P = c_void_p
# name -> argument types (all return int status); kept in one table so tests can check it against the header
PROTOTYPES = {
    "jpdvt_device_check": [],
    "jpdvt_ln_modulate_fwd": [P, P, P, P, P, P, c_int64, P, c_int64, c_int, P],
    "jpdvt_gemm_bias": [P, P, P, P, P, c_int64, c_int, c_int, P],
    "jpdvt_gemm_bias_f32": [P, P, P, P, c_int64, c_int, c_int, P],
    "jpdvt_gemm_bias_gelu": [P, P, P, P, c_int64, c_int, c_int, P],
    "jpdvt_gemm_bias_gate": [P, P, P, P, c_int64, P, c_int64, c_int, c_int, c_int, P],
    "jpdvt_gemm_bias_gate_residual": [P, P, P, P, c_int64, P, c_int64, c_int, c_int, c_int, P],
    "jpdvt_gemm_bias_gate_residual_ln": [P, P, P, P, c_int64, P, P, P, c_int64, P, c_int64, c_int, c_int, c_int, P],
    "jpdvt_gemm_bias_gate_residual_copy": [P, P, P, P, c_int64, P, P, P, c_int64, c_int, c_int, c_int, P],
    "jpdvt_fold_ln_weights": [P, P, P, P, P, P, P, P, c_int, P],
    "jpdvt_gemm_ln_folded": [c_int, P, P, c_int, P, P, P, P, c_int64, c_int, c_int, P],
    "jpdvt_gemm_patch_embed": [P, P, P, P, P, P, P, c_int64, c_int, P],
    "jpdvt_final_head_fwd": [P, P, P, P, P, P, c_int64, P],
    "jpdvt_attention_fwd": [P, P, P, c_int, c_int, P],
    "jpdvt_patchify": [P, P, c_int, c_int, P],
    "jpdvt_unpatchify": [P, P, c_int, c_int, P],
    "jpdvt_timestep_embed": [P, c_int, P, P, P, P, P, P, P, P, P, P],
    "jpdvt_philox_normal": [P, P, c_int64, c_int, P, P],
    "jpdvt_posterior_step_philox": [P, P, P, c_int, P, P, P, P, P, P, c_int64, c_int64, P],
    "jpdvt_mse_loss_fwd": [P, P, c_int64, P, P, P, c_int, c_int, P, P, c_int, P],
    "jpdvt_mse_loss_bwd": [P, P, c_int64, P, P, P, c_int, c_int, P, P, P, c_int, P],
    "jpdvt_adaln_table": [P, c_int, P, P, P, c_int, P],
    "jpdvt_posterior_step": [P, P, P, P, P, P, P, P, P, P, c_int64, c_int64, P],
    "jpdvt_ddim_step": [P, P, P, P, P, P, P, P, P, P, P, c_int64, c_int64, P],
    "jpdvt_q_sample": [P, P, P, P, P, P, P, c_int64, c_int64, P],
    "jpdvt_assign_from_scores": [P, c_int, c_int, c_double, P, P, P],
    "jpdvt_assign_greedy_l1": [P, P, c_int, c_int, c_int, c_double, P, P, P, P],
    "jpdvt_gather_pieces": [P, P, P, P, c_int, c_int, c_int, c_int, P],
    "jpdvt_score_placements": [P, P, c_int, c_int, P, P, P, P],
    "jpdvt_crop_pieces": [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "jpdvt_gemm_wgrad": [P, P, P, P, c_int64, c_int, c_int, P],
    "jpdvt_gemm_dgelu": [P, P, P, P, c_int64, c_int, c_int, P],
    "jpdvt_gemm_dgrad": [P, P, P, P, P, P, c_int64, c_int, c_int, P],
    "jpdvt_attention_bwd": [P, P, P, P, P, P, c_int, c_int, P],
    "jpdvt_gate_bwd": [P, P, P, c_int64, P, P, c_int64, P, P, c_int, c_int, P],
    "jpdvt_ln_modulate_bwd": [P, P, P, c_int64, P, c_int, P, P, c_int64, P, P, c_int, c_int, P],
    "jpdvt_ln_gate_bwd": [P, P, P, c_int64, P, c_int, P, P, c_int64, P, P, P, c_int64, P, P, c_int64, P, c_int, c_int, P],
    "jpdvt_colsum_bf16": [P, c_int64, c_int, P, P],
    "jpdvt_colsum_f32": [P, c_int64, c_int, P, P],
    "jpdvt_train_forward": [C.POINTER(Weights), C.POINTER(Tape), P, P, P, P, P, c_int, P],
    "jpdvt_train_backward_head": [C.POINTER(Weights), C.POINTER(WeightsT), C.POINTER(Tape), C.POINTER(BwdScratch),
                                  C.POINTER(Grads), P, P, P],
    "jpdvt_train_backward_block": [C.POINTER(Weights), C.POINTER(WeightsT), C.POINTER(Tape), C.POINTER(BwdScratch),
                                   C.POINTER(Grads), c_int, P],
    "jpdvt_train_backward_embed": [C.POINTER(Weights), C.POINTER(WeightsT), C.POINTER(Tape), C.POINTER(BwdScratch),
                                   C.POINTER(Grads), P, P],
    "jpdvt_adamw_ema": [P, P, P, P, P, P, c_int64, c_int64] + [C.c_float] * 7 + [P],
    "jpdvt_adamw_ema_peer": [C.POINTER(PeerStep), P, P, P, P, c_int64] + [C.c_float] * 7 + [P],
    "jpdvt_adamw_ema_peer_dev": [C.POINTER(PeerStep), P, P, P, P, P] + [C.c_float] * 7 + [P],
    "jpdvt_adamw_ema_dev": [P, P, P, P, P, P, c_int64, P] + [C.c_float] * 7 + [P],
    "jpdvt_transpose_bf16": [P, P, c_int, c_int, c_int, P],
    "jpdvt_denoiser_forward": [C.POINTER(Weights), C.POINTER(Workspace), P, P, P, P, P, P, P, c_int, P],
    "jpdvt_sample_loop": [C.POINTER(Weights), C.POINTER(Workspace), C.POINTER(Sampler), P, P, c_int, c_int, c_int, P],
}
OTHER_SYMBOLS = ["jpdvt_abi_version", "jpdvt_last_error_string", "jpdvt_wgrad_scratch_floats", "jpdvt_train_wgrad_scratch_floats",
                 "jpdvt_bwd_part_floats", "jpdvt_mse_part_floats", "jpdvt_launch_count"]

_lock = threading.Lock()
_lib = None
_device_ok = False


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the library (building it in-tree first when sources changed and nvcc is present)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        override = os.environ.get("JPDVT_LIB_PATH")        # developer knob: A/B-time an alternative build of the library
        if override:
            if not os.path.exists(override):
                raise JpdvtError(f"JPDVT_LIB_PATH={override} does not exist")
            build_if_missing = False
        if build_if_missing:
            from . import build as _build
            try:
                if _build.needs_build():
                    _build.build()
            except Exception as e:  # no nvcc on this box: fall through to the prebuilt file, else fail loudly
                if not os.path.exists(LIB_PATH):
                    raise JpdvtError(f"libjpdvt_sm100.so is missing and could not be built: {e}") from e
        if not os.path.exists(LIB_PATH):
            raise JpdvtError(f"{LIB_PATH} not found - run `python -m jpdvt_mt_ntnu_b200.build` (no CPU fallback exists)")
        lib = C.CDLL(override or LIB_PATH)
        for name, args in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = c_int
        lib.jpdvt_wgrad_scratch_floats.argtypes = [c_int64, c_int, c_int]
        lib.jpdvt_wgrad_scratch_floats.restype = c_int64
        lib.jpdvt_train_wgrad_scratch_floats.argtypes = [c_int, c_int, c_int]
        lib.jpdvt_train_wgrad_scratch_floats.restype = c_int64
        lib.jpdvt_bwd_part_floats.argtypes = [c_int, c_int]
        lib.jpdvt_bwd_part_floats.restype = c_int64
        lib.jpdvt_mse_part_floats.argtypes = [c_int]
        lib.jpdvt_mse_part_floats.restype = c_int64
        lib.jpdvt_launch_count.argtypes = []
        lib.jpdvt_launch_count.restype = c_int64
        lib.jpdvt_abi_version.restype = c_int
        lib.jpdvt_last_error_string.restype = C.c_char_p
        got = lib.jpdvt_abi_version()
        if got != ABI_VERSION:       # a stale prebuilt library would be driven with mismatched ctypes structs
            raise JpdvtError(f"{override or LIB_PATH} has ABI version {got}, this package binds version {ABI_VERSION}: "
                             "rebuild it with `python -m jpdvt_mt_ntnu_b200.build --force`")
        _lib = lib
        return lib


def launch_count() -> int:
    """Kernels launched by libjpdvt_sm100.so in this process so far."""
    return int(load().jpdvt_launch_count())


def last_error() -> str:
    return load().jpdvt_last_error_string().decode("utf-8", "replace")


def check(status: int, what: str) -> None:
    if status != 0:
        raise JpdvtError(f"{what} failed (status {status}): {last_error()}")


def require_device() -> C.CDLL:
    """Library handle, after verifying a B200-class device is current.  Raises otherwise (no CPU path)."""
    global _device_ok
    lib = load()
    if not _device_ok:
        import torch
        if not torch.cuda.is_available():
            raise JpdvtError("jpdvt_mt_ntnu_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        check(lib.jpdvt_device_check(), "jpdvt_device_check")
        _device_ok = True
    return lib


def ptr(t) -> int:
    """Device pointer of a contiguous torch tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise JpdvtError("expected a CUDA tensor")
    if not t.is_contiguous():
        raise JpdvtError("expected a contiguous tensor")
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    """cudaStream_t of torch's current stream on `device` (default: the thread's current device).  Launch sites wrap
    themselves in `on_device(...)` so that the current device IS the tensors' device: a kernel launched while another
    device is current would run there, on a stream that does not order it against the torch ops of the tensors' device."""
    import torch
    return torch.cuda.current_stream(device).cuda_stream


class on_device:
    """`with on_device(dev):` - make `dev` the thread's current CUDA device for the launches inside (no-op when it already
    is; new threads, e.g. frontend.MicroBatcher's worker, start on device 0 whatever the model's device)."""

    def __init__(self, device):
        import torch
        self.idx = device.index if hasattr(device, "index") else int(device)
        if self.idx is None:
            self.idx = torch.cuda.current_device()
        self.prev = None

    def __enter__(self):
        import torch
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            import torch
            torch.cuda.set_device(self.prev)
        return False


def on_tensor_device(fn):
    """Decorator for the per-kernel wrappers: run `fn` with the first CUDA tensor argument's device current."""
    import functools

    @functools.wraps(fn)
    def inner(*args, **kw):
        import torch
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index == torch.cuda.current_device():
                    break
                with on_device(a.device):
                    return fn(*args, **kw)
        return fn(*args, **kw)
    return inner
