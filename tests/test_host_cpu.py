"""CPU-side checks: the C-ABI library loads and exports every declared symbol, geometry validation works
without a GPU, compute entries fail loudly without one, and the host-side weight folding matches the oracle."""
import ctypes as C
import os
import re

import pytest
import torch

import mbistft_oracle as orc
from mb_istft_vits_b200 import get_config, synth
from mb_istft_vits_b200 import lib as L
from mb_istft_vits_b200.configs import CONFIGS, from_reference_json, samples_per_frame
from mb_istft_vits_b200.engine import fold_weight_norm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    return os.path.exists(L.LIB_PATH)


@pytest.fixture(scope="module")
def lib():
    if not _built():
        import __graft_entry__ as g
        g.build()
    return L.load()


def test_library_exports_every_symbol_declared_in_header(lib):
    hdr = open(os.path.join(ROOT, "include", "mbistft.h")).read()
    declared = set(re.findall(r"\b(mbv_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"mbv_handle"}
    assert declared == set(L.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s)
    assert lib.mbv_abi_version() == 1


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_geometry_validation_and_workspace_without_gpu(lib, name):
    cfg = get_config(name)
    h = C.c_void_p()
    c = L.make_config(cfg, "bf16")
    assert lib.mbv_create(C.byref(c), C.byref(h)) == 0
    n = C.c_size_t()
    assert lib.mbv_workspace_bytes(h, 2, 50, C.byref(n)) == 0 and n.value > 0
    assert samples_per_frame(cfg) == 256
    lib.mbv_destroy(h)


def test_unsupported_geometry_is_rejected_not_approximated(lib):
    cfg = get_config("ljs_mb_istft_vits")
    for key, val in (("gen_istft_n_fft", 32), ("gen_istft_hop_size", 8), ("upsample_initial_channel", 500)):
        bad = dict(cfg)
        bad[key] = val
        h = C.c_void_p()
        c = L.make_config(bad, "bf16")
        rc = lib.mbv_create(C.byref(c), C.byref(h))
        assert rc == -2, (key, rc)
        assert lib.mbv_last_error(h)
        lib.mbv_destroy(h)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from mb_istft_vits_b200 import Engine
    cfg = get_config("ljs_mini_mb_istft_vits")
    with pytest.raises(RuntimeError):
        Engine(cfg, synth.make_state_dict(cfg), precision="fp32")
    # and at the C level: weights cannot be loaded (nothing computes on the host)
    h = C.c_void_p()
    c = L.make_config(cfg, "fp32")
    assert lib.mbv_create(C.byref(c), C.byref(h)) == 0
    t = (L.MbvTensor * 1)()
    buf = torch.zeros(4)
    t[0].name = b"x"
    t[0].data = C.cast(buf.data_ptr(), C.POINTER(C.c_float))
    t[0].rank = 1
    t[0].shape[0] = 4
    assert lib.mbv_load_weights(h, t, 1) == -5
    lib.mbv_destroy(h)


def test_weight_norm_fold_matches_oracle_and_torch():
    cfg = get_config("ljs_mini_mb_istft_vits")
    sd = synth.make_state_dict(cfg, seed=3, g_scale=1.4)
    eff = fold_weight_norm(sd)
    for k in ("dec.conv_pre", "dec.ups.0", "dec.ups.1", "dec.resblocks.4.convs2.1", "flow.flows.2.enc.in_layers.3"):
        w = eff[k + ".weight"]
        assert torch.allclose(w, orc.effective_weight(sd, k), atol=1e-7)
        assert torch.allclose(w, torch._weight_norm(sd[k + ".weight_v"], sd[k + ".weight_g"], 0), atol=1e-6)
    assert "dec.updown_filter" not in eff and not any(k.endswith("weight_g") for k in eff)
    assert eff["flow.flows.0.pre.weight"].shape == (96, 96, 1)


def test_from_reference_json_maps_the_three_decoder_booleans():
    d = {"model": {"mb_istft_vits": True, "ms_istft_vits": False, "istft_vits": False, "subbands": 4,
                   "inter_channels": 192, "hidden_channels": 192, "resblock": "1", "resblock_kernel_sizes": [3, 7, 11],
                   "resblock_dilation_sizes": [[1, 3, 5]] * 3, "upsample_rates": [4, 4],
                   "upsample_initial_channel": 512, "upsample_kernel_sizes": [16, 16], "gen_istft_n_fft": 16,
                   "gen_istft_hop_size": 4}, "data": {"sampling_rate": 22050, "n_speakers": 0}}
    assert from_reference_json(d) == get_config("ljs_mb_istft_vits")
    d["model"]["mb_istft_vits"] = False
    with pytest.raises(ValueError):
        from_reference_json(d)


@pytest.mark.parametrize("prec", ["fp32", "tf32", "bf16", "fp16"])
def test_every_precision_and_flag_combination_validates_without_gpu(lib, prec):
    """mbv_create accepts the four precisions (and the experimental flags on the 16-bit ones), rejects the fp16 residual
    flag on the fp32 / tf32 paths, and every new compute entry fails loudly without a device."""
    cfg = get_config("ljs_mb_istft_vits")
    for flags in (0, L.FLAG_FUSED_PAIR | L.FLAG_CLUSTER_PAIRS):
        h = C.c_void_p()
        c = L.make_config(cfg, prec, flags=flags)
        assert lib.mbv_create(C.byref(c), C.byref(h)) == 0
        lib.mbv_destroy(h)
    h = C.c_void_p()
    c = L.make_config(cfg, prec, flags=L.FLAG_RESIDUAL_FP16)
    rc = lib.mbv_create(C.byref(c), C.byref(h))
    assert (rc == 0) == (prec in ("bf16", "fp16"))
    if not torch.cuda.is_available():
        one = (C.c_float * 4)()
        p = C.cast(one, C.c_void_p)
        assert lib.mbv_expand_prior(h, p, p, p, None, p, 1.0, 1, 1, 1, 1, p, p, None, None, None, None, None) != 0
        assert lib.mbv_flow_forward(h, p, p, None, p, 1, 1, p, 1 << 20, None) != 0
    lib.mbv_destroy(h)


def test_receptive_field_library_and_host_formula_agree():
    """mbv_receptive_field (C, used by the streaming state) == configs.receptive_field_frames (host, decode_chunked) for every
    shipped geometry; 25 latent frames for the MB / MS decoders (SURVEY 3.3 measured +-24), 13 single-band (+-13)."""
    import ctypes as C
    from mb_istft_vits_b200 import configs, lib as L
    lib = L.load()
    want = {"ljs_mb_istft_vits": 25, "ljs_istft_vits": 13}
    for name in configs.CONFIGS:
        cfg = configs.get_config(name)
        h = C.c_void_p()
        ccfg = L.make_config(cfg, "bf16", 0, 0)
        assert lib.mbv_create(C.byref(ccfg), C.byref(h)) == 0
        rf = lib.mbv_receptive_field(h)
        lib.mbv_destroy(h)
        assert rf == configs.receptive_field_frames(cfg), name
        if name in want:
            assert rf == want[name]
