"""The oracle (oracle/mbistft_oracle.py) against every golden vector minted from the reference
itself (tools/make_golden.py).  CPU only.  Tolerance: 2e-5 of peak (fp32 reassociation only)."""
import pytest
import torch

import mbistft_oracle as orc
from helpers import GOLDEN_CASES, load_case
from mb_istft_vits_b200 import synth


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_oracle_matches_reference_golden(name):
    cfg, sd, t, meta = load_case(name)
    # the seeded inputs are reproducible from the generator alone
    if meta["zseed"] >= 0:
        z_p, mask, _ = synth.make_latents(cfg, meta["B"], meta["T"], seed=meta["zseed"], lengths=meta["lengths"])
        assert torch.equal(z_p, t["z_p"]) and torch.equal(mask, t["mask"])
    else:  # captured at the seam inside the reference's infer(): z_p is NOT masked on padded frames (models.py:729)
        z_p, mask = t["z_p"], t["mask"]
    g = t.get("g")
    if g is not None:
        assert torch.equal(sd["emb_g.weight"][meta["sid"]].unsqueeze(-1), g)
    z, (o, o_mb, spec, phase) = orc.flow_decode(sd, cfg, z_p, mask, g)
    assert (z - t["z"]).abs().max() < 1e-5
    assert orc.max_abs_over_peak(o, t["o"]) < 2e-5
    assert orc.max_abs_over_peak(spec, t["spec"]) < 2e-5
    assert (phase - t["phase"]).abs().max() < 2e-5
    if "o_mb" in t:
        assert orc.max_abs_over_peak(o_mb, t["o_mb"]) < 2e-5
    else:
        assert o_mb is None
    # padded frames of z are exactly zero after four masked couplings
    assert float((z * (1 - mask)).abs().max()) == 0.0
    if "z_fwd" in t:  # the forward direction (voice conversion), minted from the reference's flow(x, mask, g)
        assert (orc.flow_forward(sd, cfg, z_p, mask, g) - t["z_fwd"]).abs().max() < 1e-5


@pytest.mark.parametrize("name", ["mb", "ms", "istft"])
def test_closed_form_tail_matches_torch_istft(name):
    """The written-out inverse DFT / OLA / envelope that the CUDA tail kernel implements
    equals torch.istft as the reference calls it (stft.py:197-202)."""
    cfg, sd, t, meta = load_case(name)
    g = t.get("g")
    logits = orc.decoder_logits(sd, cfg, t["z"] * t["mask"], g)
    a = orc.decoder_tail(sd, cfg, logits, closed_form=False)[0]
    b = orc.decoder_tail(sd, cfg, logits, closed_form=True)[0]
    assert orc.max_abs_over_peak(b, a) < 1e-5


def test_flow_is_not_vacuous():
    """With the reference's zero-initialised post layers the flow reverse is a pure permutation
    (SURVEY 8c trap 1); the synthetic weights must exercise the WN stack."""
    cfg, sd, t, meta = load_case("mb")
    assert (t["z"] - t["z_p"]).abs().max() > 1e-2


@pytest.mark.parametrize("name", ["mini_mb", "ms_spk"])
def test_oracle_flow_forward_inverts_flow_reverse(name):
    """The two directions of the coupling block are exact inverses on the unmasked frames (mean-only couplings):
    forward(reverse(z_p)) == z_p * mask up to fp32 rounding."""
    cfg, sd, t, meta = load_case(name)
    g = t.get("g")
    back = orc.flow_forward(sd, cfg, orc.flow_reverse(sd, cfg, t["z_p"], t["mask"], g), t["mask"], g)
    assert (back - t["z_p"] * t["mask"]).abs().max() < 2e-5


@pytest.mark.parametrize("name", ["vc_ms_spk", "posterior_mini"])
def test_oracle_posterior_encoder_and_voice_conversion_match_reference(name):
    """PosteriorEncoder.forward (models.py:236-246) and, for the multi-speaker case, the whole voice_conversion path
    (models.py:790-798) against vectors minted from the reference's own enc_q / flow / dec (tools/make_golden.py vc)."""
    from helpers import load_vc_case
    cfg, sd, t = load_vc_case(name)
    g = t.get("g_src")
    z, m, logs, mask = orc.posterior_encoder(sd, cfg, t["y"], t["y_lengths"], g, t["noise"])
    assert torch.equal(mask, t["y_mask"])
    assert (m - t["m"]).abs().max() < 1e-5 and (logs - t["logs"]).abs().max() < 1e-5
    assert (z - t["z"]).abs().max() < 2e-5
    if "o" in t:
        o, _, _, (z2, z_p, z_hat) = orc.voice_conversion(sd, cfg, t["y"], t["y_lengths"], g, t["g_tgt"], t["noise"])
        assert (z_p - t["z_p"]).abs().max() < 2e-5 and (z_hat - t["z_hat"]).abs().max() < 2e-5
        assert orc.max_abs_over_peak(o, t["o"]) < 2e-5


@pytest.mark.parametrize("name", ["text_mb", "text_mini", "text_short"])
def test_oracle_text_encoder_matches_reference(name):
    """TextEncoder.forward (models.py:172-181) against vectors minted from the reference's own enc_p."""
    from helpers import load_text_case
    cfg, sd, t = load_text_case(name)
    x, m, logs, mask = orc.text_encoder(sd, t["tokens"], t["x_lengths"])
    assert torch.equal(mask, t["x_mask"])
    assert (x - t["x"]).abs().max() < 2e-5 and (m - t["m"]).abs().max() < 2e-5 and (logs - t["logs"]).abs().max() < 2e-5
