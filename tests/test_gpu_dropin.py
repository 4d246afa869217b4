"""The drop-in seam exercised on the REAL reference: `patch_synthesizer` on an unmodified `SynthesizerTrn`
(baseline/_ref, staged by baseline/stage_ref.py) against the same model's own `infer()` on the same GPU.

Reference entry points covered: SynthesizerTrn.infer (models.py:697-737), the direct decoder call of
synthesis_module.infer_z_only (synthesis_module.py:148-162: `model.dec(z * mask, g=g)`), voice_conversion
(models.py:790-798).  Tolerances: fp32 path <= 1e-3 of peak (held to 1e-4), tf32 <= 1e-3 of peak, bf16 >= 40 dB.
"""
import contextlib
import io

import pytest
import torch

import mbistft_oracle as orc
from mb_istft_vits_b200 import get_config, synth

pytestmark = pytest.mark.gpu


def _reference(cfg_name, seed=77):
    import ref_loader
    if not ref_loader.available():
        pytest.skip("reference not staged (python baseline/stage_ref.py)")
    cfg = get_config(cfg_name)
    sd = synth.make_state_dict(cfg, seed=1234)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        net = ref_loader.build_synthesizer(cfg, sd, device="cuda")
    return cfg, sd, net


def _gold_mode():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("cfg_name,prec", [("ljs_mini_mb_istft_vits", "fp32"), ("ljs_mini_mb_istft_vits", "tf32"),
                                           ("ljs_mini_mb_istft_vits", "bf16"), ("ljs_istft_vits", "bf16"),
                                           ("uudb_ms_istft_vits_ms", "fp32"), ("uudb_ms_istft_vits_ms", "bf16")])
def test_patched_synthesizer_infer_matches_reference_infer(cfg_name, prec):
    from mb_istft_vits_b200 import patch_synthesizer
    _gold_mode()
    cfg, sd, net = _reference(cfg_name)
    B, Tx = 2, 14
    x = torch.randint(1, 59, (B, Tx), device="cuda")
    x_len = torch.tensor([Tx, Tx - 4], device="cuda")
    sid = (torch.arange(B, device="cuda") % cfg["n_speakers"]) if cfg["n_speakers"] else None
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(5)
        o_r, omb_r, spec_r, phase_r, attn_r, ymask_r, (z_r, zp_r, _, _), t_r = net.infer(x, x_len, sid=sid, noise_scale=0.667)
        ref_flow, ref_dec = net.flow, net.dec
        eng = patch_synthesizer(net, cfg, precision=prec)
        assert type(net.flow).__name__ == "NativeFlow" and type(net.dec).__name__ == "NativeDecoder"
        torch.manual_seed(5)
        o, omb, spec, phase, attn, ymask, (z, zp, _, _), t = net.infer(x, x_len, sid=sid, noise_scale=0.667)
        torch.cuda.synchronize()
    assert torch.equal(zp, zp_r) and torch.equal(ymask, ymask_r) and torch.equal(attn, attn_r)
    assert set(t) == set(t_r) and {"flow", "waveform_decoder"} <= set(t)
    assert o.shape == o_r.shape and spec.shape == spec_r.shape and phase.shape == phase_r.shape
    if prec == "bf16":
        assert orc.snr_db(z.cpu(), z_r.cpu()) > 40.0
        assert orc.snr_db(o.cpu(), o_r.cpu()) > 40.0
    else:
        tol = 1e-4 if prec == "fp32" else 1e-3
        assert (z - z_r).abs().max() < tol * max(1.0, float(z_r.abs().max()))
        assert orc.max_abs_over_peak(o.cpu(), o_r.cpu()) < tol
        assert orc.max_abs_over_peak(spec.cpu(), spec_r.cpu()) < tol
    if omb_r is None:
        assert omb is None
    else:
        assert omb.shape == omb_r.shape
    # the second caller of the seam (synthesis_module.py:160): the decoder alone on z * mask
    g = net.emb_g(sid).unsqueeze(-1) if sid is not None else None
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        o2 = net.dec(z_r * ymask_r, g=g)[0]
        o2_r = ref_dec(z_r * ymask_r, g=g)[0]
    if prec == "bf16":
        assert orc.snr_db(o2.cpu(), o2_r.cpu()) > 40.0
    else:
        assert orc.max_abs_over_peak(o2.cpu(), o2_r.cpu()) < (1e-4 if prec == "fp32" else 1e-3)
    eng.close()


def test_infer_native_on_the_real_reference_modules():
    """infer_native: the reference's own enc_p / dp / emb_g modules + expand_prior + flow_decode on the library."""
    from mb_istft_vits_b200 import Engine, infer_native
    _gold_mode()
    cfg, sd, net = _reference("ljs_mini_mb_istft_vits")
    x = torch.randint(1, 59, (2, 17), device="cuda")
    x_len = torch.tensor([17, 11], device="cuda")
    eng = Engine(cfg, sd, precision="fp32")
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(9)
        o_r, _, spec_r, _, attn_r, ymask_r, (z_r, zp_r, m_r, logs_r), t_r = net.infer(x, x_len, noise_scale=0.5, length_scale=1.3)
        torch.manual_seed(9)
        o, _, spec, _, attn, ymask, (z, zp, m_e, logs_e), t = infer_native(net, eng, x, x_len, noise_scale=0.5, length_scale=1.3)
        torch.cuda.synchronize()
    assert set(t) == set(t_r)
    assert torch.equal(attn, attn_r) and torch.equal(ymask, ymask_r)
    assert torch.equal(m_e, m_r) and torch.equal(logs_e, logs_r)
    assert (zp - zp_r).abs().max() < 1e-5            # same generator consumption, expf vs torch.exp
    assert orc.max_abs_over_peak(o.cpu(), o_r.cpu()) < 1e-3
    eng.close()


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_patched_voice_conversion_matches_reference_voice_conversion(prec):
    """SynthesizerTrn.voice_conversion (models.py:790-798) with enc_q, flow and dec all swapped for the native modules
    (patch_synthesizer(posterior=True)) against the unmodified model's own voice_conversion on the same GPU."""
    import ref_loader
    from mb_istft_vits_b200 import patch_synthesizer
    if not ref_loader.available():
        pytest.skip("reference not staged (python baseline/stage_ref.py)")
    _gold_mode()
    cfg = get_config("uudb_ms_istft_vits_ms")
    sd = synth.make_state_dict(cfg, seed=1234, enc_q=True)
    with contextlib.redirect_stdout(io.StringIO()):
        net = ref_loader.build_synthesizer(cfg, sd, device="cuda")
    B, T = 2, 40
    y = torch.randn((B, 513, T), device="cuda").abs()
    y_len = torch.tensor([T, T - 9], device="cuda")
    sid_s, sid_t = torch.tensor([1, 4], device="cuda"), torch.tensor([7, 0], device="cuda")
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(21)
        o_r, omb_r, ymask_r, (z_r, zp_r, zhat_r) = net.voice_conversion(y, y_len, sid_s, sid_t)
        eng = patch_synthesizer(net, cfg, precision=prec, posterior=True)
        assert type(net.enc_q).__name__ == "NativePosteriorEncoder"
        torch.manual_seed(21)
        o, omb, ymask, (z, zp, zhat) = net.voice_conversion(y, y_len, sid_s, sid_t)
        torch.cuda.synchronize()
    assert torch.equal(ymask, ymask_r)
    if prec == "fp32":
        assert (z - z_r).abs().max() < 1e-4 * max(1.0, float(z_r.abs().max()))
        assert (zhat - zhat_r).abs().max() < 1e-4 * max(1.0, float(zhat_r.abs().max()))
        assert orc.max_abs_over_peak(o.cpu(), o_r.cpu()) < 1e-4
    else:
        assert orc.snr_db(z.cpu(), z_r.cpu()) > 40.0
        assert orc.snr_db(o.cpu(), o_r.cpu()) > 40.0
    eng.close()


def test_fully_native_infer_matches_reference_infer():
    """patch_synthesizer(text=True): text encoder, flow and decoder on the library (the duration predictor stays the
    reference module) against the unmodified model's infer() on the same GPU, fp32 path.  Durations come out of a ceil()
    (models.py:717-718), so the lengths are compared first."""
    import ref_loader
    from mb_istft_vits_b200 import patch_synthesizer
    if not ref_loader.available():
        pytest.skip("reference not staged (python baseline/stage_ref.py)")
    _gold_mode()
    cfg = get_config("ljs_mini_mb_istft_vits")
    sd = synth.make_state_dict(cfg, seed=1234, enc_p=True)
    torch.manual_seed(31)
    with contextlib.redirect_stdout(io.StringIO()):
        net = ref_loader.build_synthesizer(cfg, sd, device="cuda")
    x = torch.randint(1, 59, (2, 19), device="cuda")
    x_len = torch.tensor([19, 12], device="cuda")
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(6)
        o_r, _, _, _, attn_r, ymask_r, (z_r, zp_r, mp_r, _), _ = net.infer(x, x_len, noise_scale=0.5)
        eng = patch_synthesizer(net, cfg, precision="fp32", text=True)
        assert type(net.enc_p).__name__ == "NativeTextEncoder"
        torch.manual_seed(6)
        o, _, _, _, attn, ymask, (z, zp, mp, _), _ = net.infer(x, x_len, noise_scale=0.5)
        torch.cuda.synchronize()
    assert torch.equal(ymask, ymask_r) and torch.equal(attn, attn_r), "a duration moved across a ceil() boundary"
    assert (mp - mp_r).abs().max() < 1e-4 * max(1.0, float(mp_r.abs().max()))
    assert orc.max_abs_over_peak(o.cpu(), o_r.cpu()) < 1e-3
    eng.close()
