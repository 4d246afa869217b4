"""mb_istft_vits_b200 -- B200 (sm_100a) native flow-reverse + iSTFT waveform decoders of MB-iSTFT-VITS.

Host side (this package): geometry (`configs`), seeded synthetic weights (`synth`), the ctypes binding
(`lib`), the handle owner (`engine.Engine`) and the nn.Module shims (`modules`) that drop in behind
``SynthesizerTrn.infer()``.  Device side: ``libmbistft.so`` built from ``csrc/`` (hand-written CUDA).
"""
from . import configs, synth  # noqa: F401
from .configs import get_config  # noqa: F401


def __getattr__(name):
    # torch-dependent pieces are imported lazily so `import mb_istft_vits_b200` stays cheap
    if name in ("Engine", "fold_weight_norm"):
        from . import engine
        return getattr(engine, name)
    if name in ("NativeFlow", "NativeDecoder", "NativePosteriorEncoder", "NativeTextEncoder", "patch_synthesizer", "infer_native"):
        from . import modules
        return getattr(modules, name)
    if name == "StreamingDecoder":
        from . import streaming
        return streaming.StreamingDecoder
    if name == "HostStream":
        from . import pipeline
        return pipeline.HostStream
    raise AttributeError(name)
