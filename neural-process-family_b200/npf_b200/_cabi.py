"""ctypes binding of libnpf_b200.so (the C ABI declared in include/npf_b200.h).

The library is loaded lazily on the first kernel call so that module construction, ``state_dict``
handling and the other host-side logic can be exercised on a machine without a GPU.  There is NO
fallback: if the shared library is missing or a call fails, a ``RuntimeError`` is raised.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_long, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libnpf_b200.so")

# error codes / flags (mirror include/npf_b200.h)
NPF_OK, NPF_EINVAL, NPF_ECUDA, NPF_ENOTSUP = 0, -1, -2, -3
PREC_FP32, PREC_BF16, PREC_BF16X3 = 0, 1, 2
RELU_OUT, RELU_IN, ACCUM, ADD_DY, MASK_X = 1, 2, 4, 8, 16

P, I, L, F = c_void_p, c_int, c_long, c_float

# name -> argtypes; every entry returns int except the three bookkeeping calls
SIGNATURES = {
    "npf_linear_fwd": [P, I, P, I, P, P, I, I, I, I, I, P, P, I, I, P],
    "npf_linear_bwd_data": [P, I, P, I, P, I, I, I, I, P, I, I, I, P],
    "npf_linear_bwd_weight": [P, I, P, I, P, I, P, I, I, I, I, P, P, I, I, P],
    "npf_linear_bwd": [P, I, P, I, P, I, P, I, P, I, P, I, I, I, I, I, P],
    "npf_mlp_chain_fwd": [P, I, P, P, P, I, I, I, I, I, I, P],
    "npf_mlp_chain_bwd": [P, I, P, P, P, I, P, P, I, I, I, I, I, P],
    "npf_relu_bwd": [P, P, P, L, P],
    "npf_debug_set_trace": [P],
    "npf_p2p_alloc": [P, ctypes.c_size_t],
    "npf_p2p_free": [P],
    "npf_p2p_get_handle": [P, P],
    "npf_p2p_open": [P, P],
    "npf_p2p_close": [P],
    "npf_allreduce_mean_p2p": [P, P, P, P, I, I, L, P],
    "npf_allreduce_mean_p2p2": [P, P, P, P, I, I, L, P],
    "npf_setconv_fwd": [P, L, P, L, P, P, P, P, P, I, I, I, I, I, I, I, P],
    "npf_setconv_bwd": [P, L, P, L, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, P],
    "npf_dwconv_fwd": [P, P, P, P, P, I, I, I, I, I, I, I, P, P, P],
    "npf_dwconv_bwd": [P, P, P, P, P, P, I, I, I, I, I, I, I, P, P, P, P, P],
    "npf_resblock1d_fwd": [P, P, P, P, P, P, P, I, I, I, I, I, P],
    "npf_resblock1d_bwd": [P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, P],
    "npf_channel_stats": [P, P, P, P, L, I, P],
    "npf_channel_affine": [P, P, P, P, L, I, I, P],
    "npf_gridconv_in_fwd": [P, P, I, P, P, I, I, I, I, I, P],
    "npf_gridconv_in_bwd": [P, P, I, P, P, P, P, I, I, I, I, I, P],
    "npf_merge_relu_fwd": [P, P, P, I, I, I, I, I, P],
    "npf_merge_relu_bwd": [P, P, P, P, I, I, I, I, I, P],
    "npf_mean_pool_fwd": [P, P, I, I, I, P],
    "npf_mean_pool_bwd": [P, P, I, I, I, P],
    "npf_add_layernorm_fwd": [P, P, P, P, P, P, L, I, P],
    "npf_add_layernorm_bwd": [P, P, P, P, P, P, P, P, L, I, P],
    "npf_xattn_fwd": [P, P, P, P, P, I, I, I, I, I, I, F, I, P],
    "npf_xattn_bwd": [P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, F, I, P],
    "npf_gauss_head_fwd": [P, P, P, L, I, F, P],
    "npf_gauss_head_bwd": [P, P, P, P, L, I, F, P],
    "npf_gauss_nll_fwd": [P, P, P, P, I, I, L, P],
    "npf_gauss_nll_bwd": [P, P, P, P, P, P, I, I, L, P],
    "npf_latent_sample_fwd": [P, P, P, P, P, I, L, I, P],
    "npf_latent_sample_bwd": [P, P, P, P, P, P, I, L, I, P],
    "npf_global_latent_fwd": [P, P, I, I, I, P],
    "npf_global_latent_bwd": [P, P, I, I, I, P],
    "npf_adam_step": [P, P, P, P, L, I, F, F, F, F, F, F, P],
    "npf_sqnorm": [P, L, P, P],
    "npf_adam_step_clipped": [P, P, P, P, L, I, F, F, F, F, F, F, P, F, P],
    "npf_range_check": [P, L, F, F, P, P],
    "npf_random_subset": [P, I, I, I, c_ulonglong, P],
    "npf_random_mask": [P, I, I, I, c_ulonglong, P],
    "npf_select_points": [P, P, P, P, P, I, I, I, I, I, P],
    "npf_grid_select": [P, P, P, P, P, I, I, I, I, I, I, F, P],
    "npf_gp_sample": [P, P, P, P, P, I, I, I, I, F, F, F, F, P],
    "npf_gp_sample_hyp": [P, P, P, P, P, P, I, I, I, I, F, P],
}
BOOKKEEPING = {
    "npf_abi_version": (c_int, []),
    "npf_last_error": (c_char_p, []),
    "npf_launch_count": (c_ulonglong, []),
}

_lib = None


def load():
    """Load (once) and return the ctypes handle.  Raises RuntimeError if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"npf_b200: CUDA extension not built ({LIB_PATH} missing). Run `python __graft_entry__.py build` "
            "(or `make -C neural-process-family_b200/csrc`). There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = c_int
    for name, (restype, argtypes) in BOOKKEEPING.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    if lib.npf_abi_version() != 1:
        raise RuntimeError(f"npf_b200: ABI version mismatch ({lib.npf_abi_version()} != 1)")
    _lib = lib
    return lib


_timing = None  # name -> list of (start_event, end_event, algorithmic_bytes, flops) when enabled


def _lin(M, K, N):
    return 4 * (M * K + M * N + N * K), 2 * M * K * N


# Algorithmic HBM bytes / flops of one call, from its arguments (SURVEY.md section 8d conventions: every operand and
# result once, fp32).  Used only by the timing mode below.
ALGO = {
    "npf_linear_fwd": lambda a: _lin(a[7], a[8], a[9]),
    "npf_linear_bwd_data": lambda a: _lin(a[6], a[7], a[8]),
    "npf_linear_bwd_weight": lambda a: _lin(a[7], a[8], a[9]),
    # dY and X read once, dX written once: the same bytes as the data gradient alone; twice its flops
    "npf_linear_bwd": lambda a: (4 * a[11] * (a[13] + 2 * a[12]), 4 * a[11] * a[12] * a[13]),
    # X read once, every layer output written once (no intermediate read back)
    "npf_mlp_chain_fwd": lambda a: (4 * a[6] * a[7] * (a[5] + 1) + 4 * a[5] * a[7] * a[7], 2 * a[6] * a[7] * a[7] * a[5]),
    # keys/queries + values + feat (+ dens/stat): B*(K*C + Q*C)*4 dominates
    # dY read once, every layer input once, dX written once
    "npf_mlp_chain_bwd": lambda a: (4 * a[9] * a[10] * (a[8] + 1 + (1 if a[4] else 0)) + 4 * a[8] * a[10] * a[10], 4 * a[9] * a[10] * a[10] * a[8]),
    "npf_setconv_fwd": lambda a: (4 * a[9] * (a[10] * a[12] + a[11] * a[12] + 3 * a[11] + a[10]), 2 * a[9] * a[11] * a[10] * a[12]),
    "npf_setconv_bwd": lambda a: (4 * a[13] * (2 * a[14] * a[16] + 2 * a[15] * a[16] + 4 * a[15] + a[14]), 6 * a[13] * a[15] * a[14] * a[16]),
    # x read + y written (+ the residual only when it is a different tensor from x: in a one-conv ResConvBlock res IS x)
    "npf_dwconv_fwd": lambda a: (4 * a[5] * a[6] * a[7] * a[8] * (3 if (a[3] and a[3] != a[0]) else 2), 2 * a[5] * a[6] * a[7] * a[8] * a[9] * a[10]),
    # dy + x read, dx written
    "npf_dwconv_bwd": lambda a: (4 * a[6] * a[7] * a[8] * a[9] * 3, 4 * a[6] * a[7] * a[8] * a[9] * a[10] * a[11]),
    # X read, Y written (+ O written when saved): the intermediate never makes a round trip
    "npf_resblock1d_fwd": lambda a: (4 * a[7] * a[8] * a[9] * (3 if a[5] else 2) + 4 * a[9] * a[9], 2 * a[7] * a[8] * a[9] * (a[9] + a[10])),
    # dY + X read, dX written
    "npf_resblock1d_bwd": lambda a: (4 * a[10] * a[11] * a[12] * 3 + 4 * a[12] * a[12], 2 * a[10] * a[11] * a[12] * (3 * a[12] + 3 * a[13])),
    "npf_xattn_fwd": lambda a: (4 * a[5] * a[8] * (2 * a[6] * a[9] + a[7] * a[9] + a[7] * a[10] + a[6]), 2 * a[5] * a[8] * a[6] * a[7] * (a[9] + a[10])),
    "npf_xattn_bwd": lambda a: (4 * a[9] * a[12] * (4 * a[10] * a[13] + 2 * a[11] * a[13] + 2 * a[11] * a[14]), 5 * a[9] * a[12] * a[10] * a[11] * (a[13] + a[14])),
}


def enable_timing(on):
    """Per-entry-point CUDA-event timing on the launching stream (bench.py's kernel breakdown / roofline)."""
    global _timing
    _timing = {} if on else None


def collect_timing(by_shape=False):
    """name -> (total_ms, n_calls, algorithmic_bytes, flops); synchronises.  Clears the log.
    by_shape=True keys by (name, algorithmic bytes per call, flops per call) instead, i.e. one row per entry point AND
    problem size (the 128 x 128 x 75776 layer and the 2 x 128 head are different kernels behind the same entry point)."""
    import torch
    torch.cuda.synchronize()
    out = {}
    for name, evs in (_timing or {}).items():
        if by_shape:
            for a, b, nb, fl in evs:
                k = (name, nb, fl)
                ms, n = out.get(k, (0.0, 0))
                out[k] = (ms + a.elapsed_time(b), n + 1)
        else:
            out[name] = (sum(a.elapsed_time(b) for a, b, _, _ in evs), len(evs), sum(e[2] for e in evs), sum(e[3] for e in evs))
    if _timing is not None:
        _timing.clear()
    return out


def call(name, *args):
    """Invoke an entry point; raise on a non-zero return code (ValueError for NPF_EINVAL)."""
    lib = load()
    if _timing is not None:
        import torch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = getattr(lib, name)(*args)
        b.record()
        nbytes, flops = ALGO[name](args) if name in ALGO else (0, 0)
        _timing.setdefault(name, []).append((a, b, nbytes, flops))
    else:
        rc = getattr(lib, name)(*args)
    if rc != NPF_OK:
        msg = lib.npf_last_error().decode("utf-8", "replace")
        if rc == NPF_EINVAL:
            raise ValueError(f"{name}: {msg}")
        if rc == NPF_ENOTSUP:
            raise NotImplementedError(f"{name}: {msg}")
        raise RuntimeError(f"{name} failed (rc={rc}): {msg}")


def launch_count():
    return int(load().npf_launch_count())
