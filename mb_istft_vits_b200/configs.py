"""Decoder / flow geometry of the reference's shipped configs.

Only the keys the hot path consumes are kept.  Values are
transcribed from the reference JSON ``model`` / ``data`` sections:

  ljs_mini_mb_istft_vits   configs/ljs_mini_mb_istft_vits.json:37-62
  ljs_mb_istft_vits        configs/ljs_mb_istft_vits.json:37-62
  ljs_ms_istft_vits        configs/ljs_ms_istft_vits.json:37-62
  ljs_istft_vits           configs/ljs_istft_vits.json:37-62
  ljs_mini_istft_vits      configs/ljs_mini_istft_vits.json:37-62
  uudb_spk8_istft_vits     configs/uudb_spk8_istft_vits.json:28-62   (MS decoder, 16 kHz, no g)
  uudb_ms_istft_vits_ms    configs/uudb_ms_istft_vits_ms.json:28-64  (MS decoder, 12 speakers, gin 256)

The flow hyper-parameters kernel 5 / dilation_rate 1 / n_layers 4 / n_flows 4 are hard-coded at
models.py:647 and therefore constants here too.
"""
from __future__ import annotations

import copy
import json
from typing import Any, Dict

FLOW_KERNEL = 5
FLOW_DILATION_RATE = 1
FLOW_LAYERS = 4
FLOW_N = 4

_BASE = dict(
    inter_channels=192, hidden_channels=192,
    resblock="1", resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    upsample_rates=[4, 4], upsample_initial_channel=512, upsample_kernel_sizes=[16, 16],
    gen_istft_n_fft=16, gen_istft_hop_size=4, subbands=4,
    gin_channels=0, n_speakers=0, sampling_rate=22050,
)


def _cfg(**kw) -> Dict[str, Any]:
    c = copy.deepcopy(_BASE)
    c.update(kw)
    return c


CONFIGS: Dict[str, Dict[str, Any]] = {
    "ljs_mini_mb_istft_vits": _cfg(variant="mb", hidden_channels=96, upsample_initial_channel=256),
    "ljs_mb_istft_vits": _cfg(variant="mb"),
    "ljs_ms_istft_vits": _cfg(variant="ms"),
    "ljs_istft_vits": _cfg(variant="istft", upsample_rates=[8, 8], subbands=1),
    "ljs_mini_istft_vits": _cfg(variant="istft", upsample_rates=[8, 8], subbands=1,
                                hidden_channels=96, upsample_initial_channel=256),
    "uudb_spk8_istft_vits": _cfg(variant="ms", sampling_rate=16000),
    "uudb_ms_istft_vits_ms": _cfg(variant="ms", sampling_rate=16000, gin_channels=256, n_speakers=12),
}


def get_config(name: str) -> Dict[str, Any]:
    if name not in CONFIGS:
        raise KeyError(f"unknown config {name!r}; known: {sorted(CONFIGS)}")
    return copy.deepcopy(CONFIGS[name])


def from_reference_json(path_or_dict) -> Dict[str, Any]:
    """Build the geometry dict from a reference ``configs/*.json`` (or its parsed dict), mapping the
    three decoder booleans of models.py:634-644 onto ``variant``."""
    d = path_or_dict
    if not isinstance(d, dict):
        with open(d) as f:
            d = json.load(f)
    m, data = d["model"], d.get("data", {})
    if m.get("mb_istft_vits"):
        variant = "mb"
    elif m.get("ms_istft_vits"):
        variant = "ms"
    elif m.get("istft_vits"):
        variant = "istft"
    else:
        raise ValueError("Decoder Error in json file")  # models.py:644
    c = _cfg(variant=variant)
    for k in ("inter_channels", "hidden_channels", "resblock", "resblock_kernel_sizes",
              "resblock_dilation_sizes", "upsample_rates", "upsample_initial_channel",
              "upsample_kernel_sizes", "gen_istft_n_fft", "gen_istft_hop_size"):
        c[k] = m[k]
    c["subbands"] = int(m.get("subbands") or 1)
    c["gin_channels"] = int(m.get("gin_channels", 0))
    c["n_speakers"] = int(data.get("n_speakers", 0))
    c["sampling_rate"] = int(data.get("sampling_rate", 22050))
    return c


def samples_per_frame(cfg) -> int:
    """Output samples per latent frame (256 for every shipped variant)."""
    up = 1
    for u in cfg["upsample_rates"]:
        up *= u
    return up * cfg["gen_istft_hop_size"] * (cfg["subbands"] if cfg["variant"] != "istft" else 1)


def receptive_field_frames(cfg) -> int:
    """Upper bound (latent frames, one side) of the decoder's receptive field: conv_pre (k 7), each polyphase upsampler
    (K / (2 S) input rows), each stage's widest ResBlock (sum over its convs of (k-1)/2 * d, plus (k-1)/2 per second conv
    of a ResBlock1), conv_post (k 7) with its 1-frame reflection pad, the 16/4 iSTFT overlap (4 frames) and the 63-tap
    synthesis FIR (8 sub-band samples = 2 hop blocks), each divided by the cumulative upsampling rate it runs at.
    ljs_mb: 24.95 -> 25 (SURVEY 3.3 measured +-24); single band [8,8]: 12.7 -> 13 (measured +-13)."""
    import math
    rf = 3.0
    rate = 1.0
    for u, k in zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"]):
        rf += math.ceil(k / (2 * u)) / rate
        rate *= u
        widest = 0
        for rk, dils in zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"]):
            r = sum((rk - 1) // 2 * d for d in dils)
            if str(cfg["resblock"]) == "1":
                r += (rk - 1) // 2 * len(dils)
            widest = max(widest, r)
        rf += widest / rate
    tail = 3 + 1 + 4 + (3 if cfg["variant"] != "istft" else 0)
    rf += tail / rate
    return int(math.ceil(rf))
