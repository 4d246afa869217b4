"""ctypes binding of libmbistft.so (the C ABI in include/mbistft.h).

The shared library is built in-tree (``make -C mb_istft_vits_b200/csrc`` or
``__graft_entry__.build()``).  There is deliberately no fallback: if the library is missing
or cannot be loaded this module raises, and every compute entry needs a B200.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmbistft.so")

MBV_ABI_VERSION = 1
MAX_UPS, MAX_KERNELS, MAX_DILATIONS = 4, 4, 3

VARIANTS = {"istft": 0, "mb": 1, "ms": 2}
PRECISIONS = {"fp32": 0, "tf32": 1, "bf16": 2, "fp16": 3}

FLAG_FORCE_SIMT = 4
FLAG_RESIDUAL_FP16 = 8
FLAG_FUSED_PAIR = 16
FLAG_CLUSTER_PAIRS = 32
FLAG_BRANCHES = 64
FLAG_SPLIT_TAIL = 128
FLAG_NO_CTA_PAIRS = 256
FLAG_NO_PW = 512
FLAG_NO_PAIR_SPLIT = 1024
FLAG_NO_PAIR_TM = 2048
FLAG_NO_CONV_TM = 4096

ERRORS = {0: "MBV_OK", -1: "MBV_ERR_INVALID", -2: "MBV_ERR_UNSUPPORTED", -3: "MBV_ERR_WEIGHTS",
          -4: "MBV_ERR_WORKSPACE", -5: "MBV_ERR_CUDA"}

# every symbol include/mbistft.h declares (tests check the .so exports all of them)
SYMBOLS = ["mbv_abi_version", "mbv_create", "mbv_destroy", "mbv_load_weights", "mbv_workspace_bytes",
           "mbv_flow_reverse", "mbv_decode", "mbv_flow_decode", "mbv_last_launch_count", "mbv_decode_flops",
           "mbv_flow_flops", "mbv_tail", "mbv_last_error", "mbv_set_profiling", "mbv_profile_read", "mbv_profile_read_launches", "mbv_pcm16", "mbv_expand_prior", "mbv_flow_forward",
           "mbv_posterior_workspace_bytes", "mbv_posterior_encode", "mbv_receptive_field", "mbv_stream_open",
           "mbv_stream_workspace_bytes", "mbv_stream_halo", "mbv_stream_push", "mbv_stream_close",
           "mbv_text_workspace_bytes", "mbv_text_encode", "mbv_tail_fused"]


class MbvConfig(C.Structure):
    _fields_ = [
        ("variant", C.c_int32), ("precision", C.c_int32), ("inter_channels", C.c_int32),
        ("hidden_channels", C.c_int32), ("upsample_initial_channel", C.c_int32), ("n_ups", C.c_int32),
        ("upsample_rates", C.c_int32 * MAX_UPS), ("upsample_kernel_sizes", C.c_int32 * MAX_UPS),
        ("resblock_type", C.c_int32), ("n_kernels", C.c_int32),
        ("resblock_kernel_sizes", C.c_int32 * MAX_KERNELS), ("n_dilations", C.c_int32),
        ("resblock_dilations", (C.c_int32 * MAX_DILATIONS) * MAX_KERNELS),
        ("n_fft", C.c_int32), ("hop", C.c_int32), ("subbands", C.c_int32), ("gin_channels", C.c_int32),
        ("flow_kernel", C.c_int32), ("flow_dilation_rate", C.c_int32), ("flow_layers", C.c_int32),
        ("flow_n", C.c_int32), ("device", C.c_int32), ("flags", C.c_int32),
    ]


class MbvTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.POINTER(C.c_float)), ("rank", C.c_int32),
                ("shape", C.c_int64 * 4)]


class MbvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libmbistft.so once; raise if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C mb_istft_vits_b200/csrc` (there is no CPU / PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, fp, i32 = C.c_void_p, C.c_void_p, C.c_int32
    lib.mbv_abi_version.restype = C.c_int
    lib.mbv_create.argtypes = [C.POINTER(MbvConfig), C.POINTER(vp)]
    lib.mbv_destroy.argtypes = [vp]
    lib.mbv_destroy.restype = None
    lib.mbv_load_weights.argtypes = [vp, C.POINTER(MbvTensor), i32]
    lib.mbv_workspace_bytes.argtypes = [vp, i32, i32, C.POINTER(C.c_size_t)]
    lib.mbv_flow_reverse.argtypes = [vp, fp, fp, fp, fp, i32, i32, vp, C.c_size_t, vp]
    lib.mbv_flow_forward.argtypes = [vp, fp, fp, fp, fp, i32, i32, vp, C.c_size_t, vp]
    lib.mbv_decode.argtypes = [vp, fp, fp, fp, fp, fp, fp, fp, i32, i32, vp, C.c_size_t, vp]
    lib.mbv_flow_decode.argtypes = [vp, fp, fp, fp, fp, fp, fp, fp, fp, i32, i32, vp, C.c_size_t, vp]
    lib.mbv_tail.argtypes = [vp, fp, fp, fp, fp, fp, i32, i32, vp]
    lib.mbv_tail_fused.argtypes = [vp, vp, fp, fp, fp, fp, i32, i32, vp]
    lib.mbv_last_launch_count.argtypes = [vp]
    lib.mbv_decode_flops.argtypes = [vp, i32, i32]
    lib.mbv_decode_flops.restype = C.c_double
    lib.mbv_flow_flops.argtypes = [vp, i32, i32]
    lib.mbv_flow_flops.restype = C.c_double
    lib.mbv_set_profiling.argtypes = [vp, i32]
    lib.mbv_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    lib.mbv_profile_read_launches.argtypes = [vp, C.POINTER(C.c_float), C.c_char_p, i32, i32, C.POINTER(i32)]
    lib.mbv_pcm16.argtypes = [vp, fp, vp, i32, i32, i32, vp, vp, vp]
    lib.mbv_expand_prior.argtypes = [vp, fp, fp, fp, fp, fp, C.c_float, i32, i32, i32, i32, fp, fp, fp, fp, fp, vp, vp]
    lib.mbv_posterior_workspace_bytes.argtypes = [vp, i32, i32, C.POINTER(C.c_size_t)]
    lib.mbv_posterior_encode.argtypes = [vp, fp, fp, fp, fp, fp, fp, i32, i32, vp, C.c_size_t, vp]
    lib.mbv_text_workspace_bytes.argtypes = [vp, i32, i32, C.POINTER(C.c_size_t)]
    lib.mbv_text_encode.argtypes = [vp, vp, fp, fp, fp, i32, i32, vp, C.c_size_t, vp]
    lib.mbv_receptive_field.argtypes = [vp]
    lib.mbv_stream_open.argtypes = [vp, i32, i32, C.POINTER(vp)]
    lib.mbv_stream_workspace_bytes.argtypes = [vp, C.POINTER(C.c_size_t)]
    lib.mbv_stream_halo.argtypes = [vp]
    lib.mbv_stream_push.argtypes = [vp, fp, i32, i32, fp, fp, i32, C.POINTER(C.c_int64), C.POINTER(i32), vp, C.c_size_t, vp]
    lib.mbv_stream_close.argtypes = [vp]
    lib.mbv_stream_close.restype = None
    lib.mbv_last_error.argtypes = [vp]
    lib.mbv_last_error.restype = C.c_char_p
    if lib.mbv_abi_version() != MBV_ABI_VERSION:
        raise ImportError("libmbistft.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def make_config(cfg, precision="bf16", device=0, flags=0) -> MbvConfig:
    """Geometry dict (mb_istft_vits_b200.configs) -> the C struct."""
    c = MbvConfig()
    c.variant = VARIANTS[cfg["variant"]]
    c.precision = PRECISIONS[precision]
    c.inter_channels = cfg["inter_channels"]
    c.hidden_channels = cfg["hidden_channels"]
    c.upsample_initial_channel = cfg["upsample_initial_channel"]
    ups = list(cfg["upsample_rates"])
    if len(ups) > MAX_UPS or len(cfg["resblock_kernel_sizes"]) > MAX_KERNELS:
        raise MbvError(-2, "too many upsample stages / resblock kernels")
    c.n_ups = len(ups)
    for i, (u, k) in enumerate(zip(ups, cfg["upsample_kernel_sizes"])):
        c.upsample_rates[i] = u
        c.upsample_kernel_sizes[i] = k
    c.resblock_type = int(cfg["resblock"])
    c.n_kernels = len(cfg["resblock_kernel_sizes"])
    nd = {len(d) for d in cfg["resblock_dilation_sizes"]}
    if len(nd) != 1 or max(nd) > MAX_DILATIONS:
        raise MbvError(-2, "resblock_dilation_sizes must all have the same length <= 3")
    c.n_dilations = nd.pop()
    for j, (k, dil) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
        c.resblock_kernel_sizes[j] = k
        for p, d in enumerate(dil):
            c.resblock_dilations[j][p] = d
    c.n_fft = cfg["gen_istft_n_fft"]
    c.hop = cfg["gen_istft_hop_size"]
    c.subbands = cfg["subbands"] if cfg["variant"] != "istft" else 1
    c.gin_channels = cfg.get("gin_channels", 0)
    c.flow_kernel, c.flow_dilation_rate, c.flow_layers, c.flow_n = 5, 1, 4, 4
    c.device = device
    c.flags = flags
    return c
