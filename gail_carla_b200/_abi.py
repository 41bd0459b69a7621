"""ctypes binding of libgail_carla_b200.so (include/gail_carla_b200.h).

Every wrapper takes CUDA torch tensors, checks dtype / device / contiguity, and passes raw device pointers, sizes
and the current CUDA stream to the C ABI.  There is no CPU path: a missing library or a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgail_carla_b200.so")
_lib: Optional[C.CDLL] = None

EPI_STORE, EPI_BIAS_LRELU, EPI_BIAS, EPI_MASK = 0, 1, 2, 3
LDF = 25600 + 32          # feature-row pitch: 25600 conv features + 32 columns for metrics / action / zero pad
S2D_PER_SAMPLE = 96 * 96 * 16


class ConvGeom(C.Structure):
    _fields_ = [("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Hp", C.c_int), ("Wp", C.c_int), ("Cin", C.c_int),
                ("KH", C.c_int), ("KW", C.c_int), ("S", C.c_int), ("OH", C.c_int), ("OW", C.c_int), ("OHp", C.c_int),
                ("OWp", C.c_int), ("Cout", C.c_int), ("in_batch_stride", C.c_long), ("out_batch_stride", C.c_long)]


_P, _I, _L, _F, _D = C.c_void_p, C.c_int, C.c_long, C.c_float, C.c_double
_G = C.POINTER(ConvGeom)
_SIGNATURES = {
    "gc_gae_returns": [_P, _P, _P, _P, _P, _P, _I, _I, _F, _F, _P],
    "gc_adv_stats": [_P, _P, _P, _L, _P],
    "gc_adv_normalize": [_P, _P, _P, _P, _L, _P],
    "gc_ppo_loss_fwd_bwd": [_P] * 11 + [_I, _F, _F, _I, _F, _F, _F, _I, _I, _F, _P],
    "gc_policy_act": [_P, _P, _P, _P, _P, _I, _F, _F, _I, _P],
    "gc_welford_merge": [_P, _P, _L, _P, _P],
    "gc_gather_obs_s2d": [_P, _P, _P, _I, _P],
    "gc_gather_obs_u8_s2d": [_P, _P, _P, _I, _P],
    "gc_gather_pair_mix_u8_s2d": [_P, _P, _P, _P, _P, _P, _I, _P],
    "gc_gather_rows": [_P, _P, _P, _I, _I, _L, _P],
    "gc_mixup": [_P, _P, _P, _P, _I, _L, _P],
    "gc_metrics_features": [_P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P],
    "gc_metrics_features_bwd": [_P, _P, _P, _P, _L, _P, _I, _P],
    "gc_small_linear_fwd": [_P, _L, _P, _P, _P, _L, _I, _I, _I, _P],
    "gc_small_linear_bwd": [_P, _L, _P, _P, _L, _P, _L, _P, _P, _I, _I, _I, _I, _F, _P],
    "gc_disc_loss_seed": [_P, _P, _P, _I, _F, _P],
    "gc_grad_penalty": [_P, _P, _P, _I, _L, _F, _F, _F, _F, _F, _P],
    "gc_reward_epilogue": [_P, _P, _L, _P],
    "gc_colsum": [_P, _L, _L, _I, _P, _P],
    "gc_splitk_reduce": [_P, _I, _L, _I, _L, _P, _P, _L, _P, _L, _I, _F, _P],
    "gc_prep_conv_weight": [_P, _P, _P, _I, _I, _I, _P],
    "gc_unprep_conv_wgrad": [_P, _I, _P, _P, _I, _I, _I, _P],
    "gc_prep_fc1_weight": [_P, _P, _I, _I, _L, _P],
    "gc_unprep_fc1_wgrad": [_P, _I, _P, _I, _I, _L, _P],
    "gc_zero_block": [_P, _L, _L, _L, _P],
    "gc_grad_sumsq": [_P, _L, _F, _P, _P],
    "gc_clip_adam": [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _F, _F, _F, _F, _I, _P, _P],
    "gc_conv_fprop": [_G, _P, _P, _P, _P, _P, _P, _I, _F, _P],
    "gc_conv_dgrad": [_G, _P, _P, _P, _P, _P, _F, _P, _I, _P],
    "gc_conv_wgrad_splits": [_G],
    "gc_conv_wgrad": [_G, _P, _P, _P, _I, _P],
    "gc_linear_fwd": [_P, _L, _P, _L, _P, _P, _L, _I, _I, _I, _I, _F, _I, _P],
    "gc_linear_dgrad": [_P, _L, _P, _L, _P, _P, _L, _P, _L, _I, _I, _I, _F, _P, _I, _I, _P],
    "gc_linear_wgrad": [_P, _L, _P, _L, _P, _L, _I, _I, _I, _I, _P],
}
ABI_VERSION = 2            # == GC_ABI_VERSION of include/gail_carla_b200.h; bumped whenever a signature changes
EXPORTS = sorted(list(_SIGNATURES) + ["gc_last_error_string", "gc_abi_version", "gc_build_digest"])


def load_library() -> C.CDLL:
    """dlopen the in-tree library (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing - run `python -m gail_carla_b200.build` (no CPU fallback exists)")
    lib = C.CDLL(LIB_PATH)
    lib.gc_abi_version.restype = C.c_int
    lib.gc_abi_version.argtypes = []
    if lib.gc_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{LIB_PATH} has C-ABI version {lib.gc_abi_version()}, this package binds version {ABI_VERSION}: "
                           "rebuild with `python -m gail_carla_b200.build`")
    try:
        lib.gc_build_digest.restype = C.c_char_p
        lib.gc_build_digest.argtypes = []
        built = lib.gc_build_digest().decode()
    except AttributeError:
        built = ""
    from .build import source_digest
    want = source_digest()
    if want and built != want:      # a stale binary next to newer sources (e.g. after a pull): never run it
        raise RuntimeError(f"{LIB_PATH} was built from different sources (digest {built[:12]} != {want[:12]}): "
                           "rebuild with `python -m gail_carla_b200.build`")
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.gc_last_error_string.restype = C.c_char_p
    lib.gc_last_error_string.argtypes = []
    _lib = lib
    return lib


def _ptr(t: Optional[torch.Tensor], dtype=torch.float32):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("gail_carla_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if t.get_device() != torch.cuda.current_device():
        # kernels, tensor maps and the stream handle are issued on the CURRENT device: refuse pointers of another one
        raise RuntimeError(f"tensor lives on cuda:{t.get_device()} but the current device is cuda:{torch.cuda.current_device()}: "
                           "call torch.cuda.set_device(...) (or use `with torch.cuda.device(...)`) first")
    if t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


LAUNCHES = 0                       # kernels launched through the C ABI (graph replays add their captured count)
_LAUNCHES_PER_CALL = {"gc_welford_merge": 2, "gc_small_linear_bwd": 2, "gc_conv_wgrad_splits": 0}


def call(name: str, *args):
    """Invoke a C-ABI entry point; non-zero status -> RuntimeError(gc_last_error_string())."""
    global LAUNCHES
    LAUNCHES += _LAUNCHES_PER_CALL.get(name, 1)
    lib = load_library()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.gc_last_error_string().decode()}")


def _contig(*ts):
    for t in ts:
        if t is not None and not t.is_contiguous():
            raise ValueError("tensor must be contiguous")


# ------------------------------------------------------------------ rollout maths
def gae_returns(gail_rewards, value_preds, masks, returns, gamma, gae_lambda, adv_raw=None, stats=None):
    _contig(gail_rewards, value_preds, masks, returns, adv_raw)
    T, N = gail_rewards.shape[0], gail_rewards.shape[1]
    call("gc_gae_returns", _ptr(gail_rewards), _ptr(value_preds), _ptr(masks), _ptr(returns), _ptr(adv_raw),
         _ptr(stats, torch.float64), T, N, gamma, gae_lambda, _stream())


def adv_stats(returns, value_preds, stats, n):
    _contig(returns, value_preds)
    call("gc_adv_stats", _ptr(returns), _ptr(value_preds), _ptr(stats, torch.float64), n, _stream())


def adv_normalize(returns, value_preds, stats, out, n):
    _contig(returns, value_preds, out)
    call("gc_adv_normalize", _ptr(returns), _ptr(value_preds), _ptr(stats, torch.float64), _ptr(out), n, _stream())


def ppo_loss(head_out, actions, old_logp, value_old, returns, adv, adv_stats_, d_head, out_value, out_logp, loss_acc, B,
             logstd, activation, clip, value_coef, action_weight, mode, clipped_value=True, norm=None):
    """norm: number of samples the batch means run over (default B; a rank holding part of a global minibatch passes the
    global row count)."""
    _contig(head_out, actions, old_logp, value_old, returns, adv, d_head, out_value, out_logp)
    call("gc_ppo_loss_fwd_bwd", _ptr(head_out), _ptr(actions), _ptr(old_logp), _ptr(value_old), _ptr(returns), _ptr(adv),
         _ptr(adv_stats_, torch.float64), _ptr(d_head), _ptr(out_value), _ptr(out_logp), _ptr(loss_acc, torch.float64), B,
         float(logstd[0]), float(logstd[1]), int(bool(activation)), clip, value_coef, action_weight, mode,
         int(bool(clipped_value)), 0.0 if norm is None else 1.0 / float(norm), _stream())


def policy_act(head_out, noise, value, action, logp, B, logstd, activation):
    _contig(head_out, noise, value, action, logp)
    call("gc_policy_act", _ptr(head_out), _ptr(noise), _ptr(value), _ptr(action), _ptr(logp), B, float(logstd[0]),
         float(logstd[1]), int(bool(activation)), _stream())


def welford_merge(state, x, scratch2):
    _contig(x)
    call("gc_welford_merge", _ptr(state, torch.float64), _ptr(x), x.numel(), _ptr(scratch2, torch.float64), _stream())


# ------------------------------------------------------------------ data movement / small stages
def gather_obs_s2d(src, idx, out, B):
    """src: fp32 rows [*,3,192,192], or the uint8 table of a device-resident expert data set (decoded PNG bytes)."""
    _contig(src, idx, out)
    if src.dtype == torch.uint8:
        call("gc_gather_obs_u8_s2d", _ptr(src, torch.uint8), _ptr(idx, torch.int64), _ptr(out), B, _stream())
    else:
        call("gc_gather_obs_s2d", _ptr(src), _ptr(idx, torch.int64), _ptr(out), B, _stream())


def gather_pair_mix(src_e, idx_e, src_p, idx_p, alpha, out, B):
    """Expert rows, policy rows and their mix-up into out rows [0,B) | [B,2B) | [2B,3B) in one pass (uint8 sources)."""
    _contig(src_e, idx_e, src_p, idx_p, alpha, out)
    call("gc_gather_pair_mix_u8_s2d", _ptr(src_e, torch.uint8), _ptr(idx_e, torch.int64), _ptr(src_p, torch.uint8),
         _ptr(idx_p, torch.int64), _ptr(alpha), _ptr(out), B, _stream())


def gather_rows(src, idx, out, B, width, ldo):
    _contig(src, idx)
    call("gc_gather_rows", _ptr(src), _ptr(idx, torch.int64), _ptr(out), B, width, ldo, _stream())


def mixup(xe, xp, alpha, out, B, per_sample):
    call("gc_mixup", _ptr(xe), _ptr(xp), _ptr(alpha), _ptr(out), B, per_sample, _stream())


def metrics_features(metrics, emb, out, ldo, pad, B, action=None, metrics2=None, action2=None, alpha=None):
    _contig(metrics, metrics2, action, action2, alpha, emb)
    call("gc_metrics_features", _ptr(metrics), _ptr(metrics2), _ptr(action), _ptr(action2), _ptr(alpha), _ptr(emb),
         _ptr(out), ldo, pad, B, _stream())


def metrics_features_bwd(metrics, d_feat, ldf, d_emb, B, metrics2=None, alpha=None):
    call("gc_metrics_features_bwd", _ptr(metrics), _ptr(metrics2), _ptr(alpha), _ptr(d_feat), ldf, _ptr(d_emb), B, _stream())


def small_linear_fwd(x, ldx, w, bias, y, ldy, B, N, K):
    _contig(w, bias)
    call("gc_small_linear_fwd", _ptr(x), ldx, _ptr(w), _ptr(bias), _ptr(y), ldy, B, N, K, _stream())


def small_linear_bwd(x, ldx, w, dy, lddy, dx, lddx, dw, db, B, B_params, N, K, slope):
    call("gc_small_linear_bwd", _ptr(x), ldx, _ptr(w), _ptr(dy), lddy, _ptr(dx), lddx, _ptr(dw), _ptr(db), B, B_params, N, K,
         slope, _stream())


def disc_loss_seed(d, dd, acc, B, norm=None):
    call("gc_disc_loss_seed", _ptr(d), _ptr(dd), _ptr(acc, torch.float64), B, 0.0 if norm is None else 1.0 / float(norm), _stream())


def grad_penalty(g, u, acc, B, per_sample, lambda_, scales, norm=None):
    call("gc_grad_penalty", _ptr(g), _ptr(u), _ptr(acc, torch.float64), B, per_sample, lambda_, scales[0], scales[1], scales[2],
         0.0 if norm is None else 1.0 / float(norm), _stream())


def reward_epilogue(d, reward, n):
    call("gc_reward_epilogue", _ptr(d), _ptr(reward), n, _stream())


def colsum(x, ld, rows, Cc, out):
    call("gc_colsum", _ptr(x), ld, rows, Cc, _ptr(out), _stream())


def splitk_reduce(part, splits, M, N, ldp, bias, mask_src, ldm, out, ldo, epilogue, slope):
    call("gc_splitk_reduce", _ptr(part), splits, M, N, ldp, _ptr(bias), _ptr(mask_src), ldm, _ptr(out), ldo, epilogue, slope,
         _stream())


# ------------------------------------------------------------------ parameter layouts / optimiser
def prep_conv_weight(w, w_fprop, w_dgrad, Cout, Cin, layer1):
    _contig(w)
    call("gc_prep_conv_weight", _ptr(w), _ptr(w_fprop), _ptr(w_dgrad), Cout, Cin, int(layer1), _stream())


def unprep_conv_wgrad(part, splits, dw, Cout, Cin, layer1, dbias=None):
    call("gc_unprep_conv_wgrad", _ptr(part), splits, _ptr(dw), _ptr(dbias), Cout, Cin, int(layer1), _stream())


def prep_fc1_weight(w, w_gemm, out, tail, ld):
    _contig(w)
    call("gc_prep_fc1_weight", _ptr(w), _ptr(w_gemm), out, tail, ld, _stream())


def unprep_fc1_wgrad(part, splits, dw, out, tail, ld):
    call("gc_unprep_fc1_wgrad", _ptr(part), splits, _ptr(dw), out, tail, ld, _stream())


def zero_block(t: torch.Tensor, rows: int = 1, width: int = None, pitch: int = None):
    """Zero `rows` runs of `width` elements of `t` that start `pitch` elements apart (default: the whole contiguous tensor)."""
    if not t.is_cuda:
        raise RuntimeError("gail_carla_b200 kernels need CUDA tensors (there is no CPU fallback)")
    es = t.element_size()
    width = t.numel() if width is None else width
    pitch = width if pitch is None else pitch
    call("gc_zero_block", t.data_ptr(), pitch * es, rows, width * es, _stream())


def grad_sumsq(grad, n, sumsq, grad_scale=1.0):
    """sumsq[0] = sum (grad_scale * g)^2 (cleared first)."""
    call("gc_grad_sumsq", _ptr(grad), n, float(grad_scale), _ptr(sumsq, torch.float64), _stream())


def clip_adam(param, grad, exp_avg, exp_avg_sq, n, sumsq, max_norm, lr, beta1, beta2, eps, bc1, bc2, grad_scale=1.0,
              zero_grad=False, dev_hyper=None):
    call("gc_clip_adam", _ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), n, _ptr(sumsq, torch.float64),
         -1.0 if max_norm is None else float(max_norm), lr, beta1, beta2, eps, bc1, bc2, float(grad_scale), int(bool(zero_grad)),
         _ptr(dev_hyper), _stream())


# ------------------------------------------------------------------ tcgen05 contractions
def conv_fprop(geom: ConvGeom, x, w, bias, y, epilogue, slope=0.2, mask_src=None, mask_bits=None):
    call("gc_conv_fprop", C.byref(geom), _ptr(x), _ptr(w), _ptr(bias), _ptr(mask_src), _ptr(mask_bits, torch.int32), _ptr(y),
         epilogue, slope, _stream())


def conv_dgrad(geom: ConvGeom, dy, wd, dx, mask_src=None, slope=0.2, mask_bits=None, dbias_in=None, dbias_samples=0):
    """dbias_in (needs mask_bits): += per-channel sums of the masked dx over the first `dbias_samples` samples (0: all) - the
    bias gradient of the layer below, taken from the output tiles instead of a separate column-sum pass."""
    call("gc_conv_dgrad", C.byref(geom), _ptr(dy), _ptr(wd), _ptr(mask_src), _ptr(mask_bits, torch.int32), _ptr(dx), slope,
         _ptr(dbias_in), int(dbias_samples), _stream())


def conv_wgrad_splits(geom: ConvGeom) -> int:
    z = load_library().gc_conv_wgrad_splits(C.byref(geom))
    if z < 1:
        raise RuntimeError("gc_conv_wgrad_splits: " + load_library().gc_last_error_string().decode())
    return z


def conv_wgrad(geom: ConvGeom, dy, x, dw_partial, splits):
    call("gc_conv_wgrad", C.byref(geom), _ptr(dy), _ptr(x), _ptr(dw_partial), splits, _stream())


def linear_fwd(x, ldx, w, ldw, bias, y, ldy, M, N, K, epilogue, slope=0.2, splits=1):
    call("gc_linear_fwd", _ptr(x), ldx, _ptr(w), ldw, _ptr(bias), _ptr(y), ldy, M, N, K, epilogue, slope, splits, _stream())


def linear_dgrad(dy, lddy, w, ldw, dx, lddx, M, N, K, mask_src=None, ldm=0, slope=0.2, mask_bits=None, colsum=None, colsum_mod=0,
                 colsum_rows=0):
    """colsum (needs mask_bits): colsum[n % colsum_mod] += column sums of the masked dx over the first `colsum_rows` rows."""
    call("gc_linear_dgrad", _ptr(dy), lddy, _ptr(w), ldw, _ptr(mask_src), _ptr(mask_bits, torch.int32), ldm, _ptr(dx), lddx, M, N, K,
         slope, _ptr(colsum), int(colsum_mod), int(colsum_rows), _stream())


def linear_wgrad(dy, lddy, x, ldx, dw, lddw, M, N, K, splits=1):
    call("gc_linear_wgrad", _ptr(dy), lddy, _ptr(x), ldx, _ptr(dw), lddw, M, N, K, splits, _stream())
