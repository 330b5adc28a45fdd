"""ctypes binding of libgeomae_b200.so (the C ABI declared in include/geomae_b200.h).

There is NO fallback: if the shared library is missing, or an entry point fails, a
RuntimeError is raised.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgeomae_b200.so")

_lib = None


class VoxelCfg(C.Structure):
    _fields_ = [("range_min", C.c_float * 3), ("range_max", C.c_float * 3),
                ("voxel_top", C.c_float * 3), ("voxel_med", C.c_float * 3), ("voxel_low", C.c_float * 3),
                ("ratio_med", C.c_int32 * 3), ("ratio_low", C.c_int32 * 3)]


class ScatterIO(C.Structure):
    _fields_ = [("points", C.c_void_p), ("frame_offsets", C.c_void_p),
                ("n_points", C.c_int64), ("stride", C.c_int32), ("n_frames", C.c_int32), ("cap", C.c_int64),
                ("bitmap", C.c_void_p), ("word_rank", C.c_void_p), ("scan_tmp", C.c_void_p),
                ("counts", C.c_void_p), ("pillar_coors", C.c_void_p), ("pillar_mean", C.c_void_p),
                ("point_pillar", C.c_void_p), ("med_mask", C.c_void_p), ("low_mask", C.c_void_p),
                ("med_ptr", C.c_void_p), ("low_ptr", C.c_void_p), ("med_mean", C.c_void_p),
                ("low_mean", C.c_void_p), ("coors_top", C.c_void_p), ("coors_med", C.c_void_p),
                ("coors_low", C.c_void_p)]


class WindowCfg(C.Structure):
    _fields_ = [("win_x", C.c_int32), ("win_y", C.c_int32), ("n_shifts", C.c_int32),
                ("shift_x", C.c_int32 * 2), ("shift_y", C.c_int32 * 2)]


class WindowIO(C.Structure):
    _fields_ = [("ptr_stride", C.c_int64), ("cand_count", C.c_void_p), ("cand_tok_off", C.c_void_p),
                ("cand_win_idx", C.c_void_p), ("n_windows", C.c_void_p), ("win_ptr", C.c_void_p),
                ("win_id", C.c_void_p), ("win_tok", C.c_void_p), ("tok_cell", C.c_void_p),
                ("tok_win", C.c_void_p), ("tok_pos", C.c_void_p)]


class LinearArgs(C.Structure):
    _fields_ = [("A", C.c_void_p), ("lda", C.c_int32), ("n_rows", C.c_int32), ("K", C.c_int32),
                ("pos_table", C.c_void_p), ("tok_cell", C.c_void_p), ("pos_slabs", C.c_int32), ("a_gelu", C.c_int32),
                ("W", C.c_void_p), ("ldw", C.c_int32), ("w_rows", C.c_int32), ("w_mn_major", C.c_int32),
                ("bias", C.c_void_p), ("N_total", C.c_int32),
                ("out", C.c_void_p), ("ldo", C.c_int32),
                ("add_src", C.c_void_p), ("ld_add", C.c_int32),
                ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("ln_eps", C.c_float), ("ln_in", C.c_void_p),
                ("ln_stats", C.c_void_p),
                ("gelu_u", C.c_void_p), ("ldu", C.c_int32),
                ("epilogue", C.c_int32), ("precision", C.c_int32), ("Wp_hi", C.c_void_p), ("Wp_lo", C.c_void_p),
                ("dot_src", C.c_void_p), ("ld_dot", C.c_int32), ("dot_out", C.c_void_p),
                ("ln_dgamma", C.c_void_p), ("ln_dbeta", C.c_void_p), ("ln_dcolsum", C.c_void_p),
                ("a_bf16", C.c_int32), ("out_bf16", C.c_int32)]


class WgradArgs(C.Structure):
    _fields_ = [("dY", C.c_void_p), ("ldy", C.c_int32), ("X", C.c_void_p), ("ldx", C.c_int32), ("n_rows", C.c_int32),
                ("pos_table", C.c_void_p), ("tok_cell", C.c_void_p), ("pos_slabs", C.c_int32), ("x_gelu", C.c_int32),
                ("dW", C.c_void_p), ("ldw", C.c_int32), ("db", C.c_void_p), ("M_total", C.c_int32),
                ("N_total", C.c_int32), ("precision", C.c_int32), ("dy_bf16", C.c_int32)]


class SRAWindows(C.Structure):
    _fields_ = [("win_ptr", C.c_void_p), ("win_tok", C.c_void_p), ("tok_win", C.c_void_p), ("tok_cell", C.c_void_p)]


class SRACtx(C.Structure):
    _fields_ = [("n_tokens", C.c_int64), ("d_model", C.c_int32), ("n_heads", C.c_int32), ("ffn", C.c_int32),
                ("precision", C.c_int32), ("pos_table", C.c_void_p), ("shift", SRAWindows * 2), ("pos16", C.c_void_p * 2)]


_LAYER_PARAMS = ["in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b",
                 "norm1_w", "norm1_b", "norm2_w", "norm2_b"]
_LAYER_GRADS = ["g_in_proj_w", "g_in_proj_b", "g_out_proj_w", "g_out_proj_b", "g_lin1_w", "g_lin1_b", "g_lin2_w",
                "g_lin2_b", "g_norm1_w", "g_norm1_b", "g_norm2_w", "g_norm2_b"]


class SRALayer(C.Structure):
    _fields_ = ([("shift", C.c_int32), ("ln_eps", C.c_float)] + [(k, C.c_void_p) for k in _LAYER_PARAMS] +
                [(k, C.c_void_p) for k in _LAYER_GRADS] +
                [(k, C.c_void_p * 2) for k in ("p_in_proj", "p_out_proj", "p_lin1", "p_lin2")])


class SRASaved(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("qkv", "attn", "lse", "s1", "st1", "y", "u", "s2", "st2", "z", "g", "xp", "xb")]


class ChainFwdArgs(C.Structure):
    _fields_ = ([("n_tokens", C.c_int64), ("mode", C.c_int32), ("x", C.c_void_p), ("attn", C.c_void_p)] +
                [(k, C.c_void_p) for k in ("p_out_proj", "p_lin1", "p_lin2", "p_in_proj_next", "out_proj_b", "lin1_b",
                                           "lin2_b", "in_proj_b_next", "norm1_w", "norm1_b", "norm2_w", "norm2_b")] +
                [("ln_eps", C.c_float), ("pos_table", C.c_void_p), ("tok_cell_next", C.c_void_p)] +
                [(k, C.c_void_p) for k in ("st1", "st2", "z", "xh1_16", "xh2_16", "u16", "g16", "xb16", "qkv16_next")])


class WgradLayerArgs(C.Structure):
    _fields_ = [("n_tokens", C.c_int64)] + [(k, C.c_void_p) for k in (
        "ds2_16", "g16", "du16", "xh1_16", "ds1_16", "attn16", "dqkv16", "xin16", "pos16", "norm1_w", "norm1_b", "in_scale",
        "in_shift", "g_lin2_w", "g_lin1_w", "g_lin1_b", "g_out_proj_w", "g_in_proj_w", "g_in_proj_b", "g_lin2_b",
        "g_out_proj_b")]


class ChainBwdArgs(C.Structure):
    _fields_ = ([("n_tokens", C.c_int64), ("mode", C.c_int32)] +
                [(k, C.c_void_p) for k in (
                    "dqkv16_up", "ds1_up", "p_in_proj_up", "dz_in", "xh2_16", "st2", "xh1_16", "st1", "u16", "attn16", "p_lin2",
                    "p_lin1", "p_out_proj", "norm2_w", "norm1_w", "ds2_16", "du16", "ds1_16", "dattn16", "ds1", "dd", "dx",
                    "g_norm2_w", "g_norm2_b", "g_norm1_w", "g_norm1_b")])


class LossArgs(C.Structure):
    _fields_ = [("rows", C.c_void_p), ("m", C.c_int64), ("reg_low", C.c_void_p), ("reg_med", C.c_void_p),
                ("reg_top", C.c_void_p), ("nor_top", C.c_void_p), ("cls_low", C.c_void_p), ("cls_med", C.c_void_p),
                ("normal", C.c_void_p), ("w_low", C.c_float), ("w_med", C.c_float), ("w_top", C.c_float),
                ("w_nor", C.c_float), ("w_cls_low", C.c_float), ("w_cls_med", C.c_float), ("ld", C.c_int32 * 6)]


def build_if_missing():
    if not os.path.exists(LIB_PATH):
        import subprocess
        subprocess.check_call(["bash", os.path.join(_HERE, "csrc", "build.sh")])


def lib():
    """The loaded library; raises if it cannot be loaded (no CPU/eager fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"geomae_b200: {LIB_PATH} is missing — build it with geomae_b200/csrc/build.sh "
                "(or __graft_entry__.build()); there is no fallback path")
        _lib = C.CDLL(LIB_PATH)
        _lib.geomae_last_error.restype = C.c_char_p
        for name in dir(_Sigs):
            if name.startswith("geomae_"):
                fn = getattr(_lib, name)
                fn.argtypes = getattr(_Sigs, name)
                fn.restype = C.c_int
        _lib.geomae_sra_scratch_floats.argtypes = [C.POINTER(SRACtx), C.c_int32]
        _lib.geomae_sra_scratch_floats.restype = C.c_int64
        _lib.geomae_peer_mailbox_doubles.argtypes = [C.c_int32]
        _lib.geomae_peer_mailbox_doubles.restype = C.c_int64
    return _lib


_p, _i64, _i32, _f3 = C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_float)


class PeerCtx(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("mailbox", C.c_void_p * 8), ("timeout_flag", C.c_void_p)]


class _Sigs:
    geomae_peer_enable_access = [_i32]
    geomae_peer_mailbox_create = [_i32, C.POINTER(C.c_void_p), _p]
    geomae_peer_buffer_create = [_i64, C.POINTER(C.c_void_p), _p]
    geomae_peer_reduce_shard = [C.POINTER(PeerCtx), C.POINTER(C.c_void_p), _i64, _i64, _p]
    geomae_peer_mailbox_open = [_p, C.POINTER(C.c_void_p)]
    geomae_peer_mailbox_close = [_p, _i32]
    geomae_peer_allreduce_f64 = [C.POINTER(PeerCtx), _p, _i32, C.c_double, C.c_double, C.c_uint64, _p]
    geomae_grid_size = [_f3, _f3, _f3, C.POINTER(C.c_int32)]
    geomae_dynamic_voxelize = [_p, _i64, _i32, _f3, _f3, _f3, _p, _p]
    geomae_voxel_scatter = [C.POINTER(VoxelCfg), C.POINTER(ScatterIO), _p]
    geomae_geom_targets = [C.POINTER(VoxelCfg), C.POINTER(ScatterIO), _i64, _p, _p, _p, _p, _p, _p]
    geomae_dense_targets = [C.POINTER(VoxelCfg), C.POINTER(ScatterIO), _p, _i64, _i32, _p, _p, _p, _p, _p, _p]
    geomae_window_candidates = [C.POINTER(VoxelCfg), C.POINTER(WindowCfg), _i32, C.POINTER(C.c_int32),
                                C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    geomae_coors_bitmap = [C.POINTER(VoxelCfg), _p, _i64, _i32, _p, _p, _p, _p, _p, _p]
    geomae_coors_rank = [C.POINTER(VoxelCfg), _p, _i64, _i32, _p, _p, _p, _p, _p, _p, _p]
    geomae_token_map = [_p, _i64, _p, _i64, _p]
    geomae_window_csr = [C.POINTER(VoxelCfg), C.POINTER(WindowCfg), C.POINTER(ScatterIO), _p, _i64,
                         C.POINTER(WindowIO), _p]
    geomae_window_drop = [C.POINTER(VoxelCfg), C.POINTER(WindowCfg), C.POINTER(ScatterIO), _p, _i64, _i32,
                          C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_uint64, _p, _p, _p]
    geomae_recover_bev = [_p, _p, _i64, _i32, _i32, _i32, _i32, _p, _p]
    geomae_recover_bev_bwd = [_p, _p, _i64, _i32, _i32, _i32, _p, _p]
    geomae_pos_table = [_i32, _i32, _i32, C.c_float, _p, _p]
    geomae_vfe_decorate = [_p, _i64, _i32, _p, _p, _p, _f3, _f3, _p, _p]
    geomae_scatter_reduce_fwd = [_p, _i64, _i32, _p, _p, _i64, _i32, _p, _p, _p]
    geomae_scatter_reduce_bwd = [_p, _i64, _i32, _p, _p, _p, _i32, _p, _p]
    geomae_sra_attention_fwd = [_p, _i64, _i32, _p, _p, _p, _p, _p, _p]
    geomae_sra_attention_bwd = [_p, _p, _p, _p, _i64, _i32, _p, _p, _p, _p, _p, _p]
    geomae_sra_attention_tc_fwd = [_p, _i64, _i32, _p, _p, _p, _p, _p, _i32, _p]
    geomae_sra_attention_tc_bwd = [_p, _p, _p, _p, _i64, _i32, _p, _p, _p, _p, _p, _i32, _p]
    geomae_tc_linear = [C.POINTER(LinearArgs), _p]
    geomae_tc_wgrad = [C.POINTER(WgradArgs), _p]
    geomae_sra_stack_forward = [C.POINTER(SRACtx), _i32, C.POINTER(SRALayer), C.POINTER(SRASaved), _p, _p]
    geomae_sra_stack_backward = [C.POINTER(SRACtx), _i32, C.POINTER(SRALayer), C.POINTER(SRASaved), _p, _p, _p, _p, _p]
    geomae_sra_stack2_forward = [C.POINTER(SRACtx), _i32, C.POINTER(SRALayer), C.POINTER(SRASaved), C.POINTER(SRALayer),
                                 C.POINTER(SRASaved), _p, _p]
    geomae_sra_stack2_backward = [C.POINTER(SRACtx), _i32, C.POINTER(SRALayer), C.POINTER(SRASaved),
                                  C.POINTER(SRALayer), C.POINTER(SRASaved), _p, _p, _p, _p, _p, _p, _p]
    geomae_sra_chain_fwd = [C.POINTER(ChainFwdArgs), _p]
    geomae_sra_chain_bwd = [C.POINTER(ChainBwdArgs), _p]
    geomae_pos_rows_bf16 = [_p, _p, _i64, _p, _p]
    geomae_sra_wgrad_layer = [C.POINTER(WgradLayerArgs), _p]
    geomae_layernorm_bwd = [_p, _p, _p, _p, _i64, _i32, _p, _p, _p, _p, _p]
    geomae_geom_loss_fwd = [C.POINTER(VoxelCfg), C.POINTER(ScatterIO), C.POINTER(LossArgs), _p, _p, _p, _p]
    geomae_geom_loss_bwd = [C.POINTER(VoxelCfg), C.POINTER(ScatterIO), C.POINTER(LossArgs), _p, _p, _p, _p, _p, _p, _p,
                            _p, _p]
    geomae_mask_split = [_p, _i32, C.c_double, C.c_uint64, _p, _p, _p]
    geomae_vfe0_forward = [_p, _i64, _i32, _p, _p, _p, _f3, _f3, _p, _p, _p, _p]
    geomae_colstats = [_p, _i64, _i32, _p, _p]
    geomae_vfe_bn_relu_max = [_p, _i64, _i32, _p, _p, _p, _p, C.c_float, _p, _p, C.c_float, C.c_float, _p, _i64, _p]
    geomae_vfe_cat = [_p, _i64, _p, _p, _p, _p, C.c_float, _p, _p, _p]
    geomae_vmax_decode = [_p, _i64, _p, _p]
    geomae_vfe1_backward = [_i32, _p, _i64, _p, _p, _p, _p, C.c_float, _p, _p, _p, _p, C.c_float, _p, _p]
    geomae_bn_backward_coeffs = [_i32, _p, _p, _p, C.c_float, _p, _p, _p, _p]
    geomae_vfe_gather_backward = [_p, _i64, _p, _p, _i64, _p]
    geomae_vfe0_backward = [_i32, _p, _i64, _i32, _p, _p, _p, _f3, _f3, _p, _p, _p, _p, C.c_float, _p, _p, _p, _p, _p,
                            C.c_float, _p, _p]
    geomae_pack_weights = [_i32, _p, _p, _p, _p, _p, _p]
    geomae_profile_enable = [_i32]
    geomae_profile_read = [_p, _p, _p, _p]
    geomae_augment_filter = [_p, _i64, _i32, _p, _i32, _p, _f3, _f3, _p, _p, _p, _i64, _p]
    geomae_decoder_tokens = [_p, _i64, _p, _i64, _i32, _p, _p]
    geomae_mask_token_grad = [_p, _i64, _i64, _i32, _p, _p]
    geomae_sweep_merge = [_p, _i64, _i32, _p, _i32, _p, _p, _p, _p, _i64, _p]
    geomae_adamw_step = [_p, _p, _p, _p, _i64, _i64, _p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                         C.c_float, C.c_float, _i64, _p, _p]


_timing = None      # list of (name, start_event, end_event) while bench.py measures per-call device time
_calls = {}         # name -> number of C-ABI calls (bench.py reports kernel launches from these)

# kernels launched per C-ABI call (upper bound for the optional ones), for the gpu_launches claim
LAUNCHES_PER_CALL = dict(dynamic_voxelize=1, voxel_scatter=9, augment_filter=3, sweep_merge=3, geom_targets=1, dense_targets=1, coors_bitmap=4, coors_rank=4,
                         token_map=1, window_csr=3, window_drop=2, pos_table=1, vfe_decorate=1, scatter_reduce_fwd=5,
                         scatter_reduce_bwd=1, sra_attention_fwd=1, sra_attention_bwd=2, sra_attention_tc_fwd=1,
                         sra_attention_tc_bwd=1, adamw_step=2,
                         tc_linear=1, tc_wgrad=1, layernorm_bwd=1, geom_loss_fwd=3, geom_loss_bwd=1)
# sra_stack_forward / _backward launch 5 / 11 kernels per layer: counted by the caller via add_launches()


def start_timing():
    global _timing
    _timing = []


def stop_timing():
    """-> {name: (total_ms, n_calls)}; synchronises the device."""
    global _timing
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in _timing or []:
        t, n = out.get(name, (0.0, 0))
        out[name] = (t + e0.elapsed_time(e1), n + 1)
    _timing = None
    return out


def reset_call_counts():
    _calls.clear()


def add_launches(n: int):
    """Kernel launches issued inside a multi-kernel C-ABI call (the SRA stack executors)."""
    _calls["__extra__"] = _calls.get("__extra__", 0) + n


def launch_count():
    return sum((1 if k == "__extra__" else LAUNCHES_PER_CALL.get(k, 1)) * v for k, v in _calls.items())


_fn_cache = {}


def run(what: str, *args):
    """Call geomae_<what>(*args); raise RuntimeError on a non-zero status."""
    fn = _fn_cache.get(what)
    if fn is None:
        fn = _fn_cache[what] = getattr(lib(), "geomae_" + what)
    _calls[what] = _calls.get(what, 0) + 1
    if _timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        _timing.append((what, e0, e1))
    else:
        rc = fn(*args)
    if rc != 0:
        raise RuntimeError(f"geomae_b200.{what} failed ({rc}): {lib().geomae_last_error().decode()}")


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"geomae_b200.{what} failed ({rc}): {lib().geomae_last_error().decode()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be contiguous."""
    if t is None:
        return None
    assert t.is_contiguous(), "geomae_b200 kernels need contiguous tensors"
    return t.data_ptr()


_stream_cache = None     # (device index, cuda stream handle) pinned for the duration of a training step


def pin_stream(device):
    """FlatTrainer pins the current stream of its device for one step: ~250 `torch.cuda.current_stream` look-ups per
    step (5 us each) otherwise.  Returns the previous pin; pass it back to `unpin_stream`."""
    global _stream_cache
    prev = _stream_cache
    _stream_cache = (torch.device(device).index, torch.cuda.current_stream(device).cuda_stream)
    return prev


def unpin_stream(prev=None):
    global _stream_cache
    _stream_cache = prev


class stream_override:
    """``with stream_override(torch_stream):`` — C-ABI calls made inside launch on that stream even while a step has its
    compute stream pinned (the input-stage work FlatTrainer runs on its side stream)."""

    def __init__(self, stream):
        self.stream = stream

    def __enter__(self):
        global _stream_cache
        self.prev = _stream_cache
        _stream_cache = (self.stream.device.index, self.stream.cuda_stream)
        return self

    def __exit__(self, *exc):
        global _stream_cache
        _stream_cache = self.prev
        return False


def stream_ptr(device=None):
    c = _stream_cache
    if c is not None and (device is None or getattr(device, "index", None) == c[0]):
        return c[1]
    return torch.cuda.current_stream(device).cuda_stream


def f3(vals):
    return (C.c_float * 3)(*[float(v) for v in vals])


def require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"geomae_b200: {name} must be a CUDA tensor (no CPU path exists)")
