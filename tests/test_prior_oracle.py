"""Alignment expansion + prior sampling (models.py:717-729): the oracle restatement against vectors captured inside the
reference's own infer() (tools/make_golden.py prior).  CPU only."""
import os

import numpy as np
import pytest
import torch

import mbistft_oracle as orc
from helpers import GOLDEN_DIR


def load_prior(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: (torch.from_numpy(d[k]) if d[k].ndim else float(d[k])) for k in d.files}


@pytest.mark.parametrize("name", ["prior_mini", "prior_long"])
def test_oracle_expand_prior_equals_reference_infer(name):
    t = load_prior(name)
    z_p, y_mask, attn, m, logs, y_len = orc.expand_prior(t["m_p"], t["logs_p"], t["w_ceil"], t["noise"], t["noise_scale"], t["x_mask"])
    assert torch.equal(y_mask, t["y_mask"])
    assert torch.equal(attn, t["attn"])
    assert torch.equal(m, t["m_exp"]) and torch.equal(logs, t["logs_exp"])
    assert torch.equal(z_p, t["z_p"])
    assert y_len.tolist() == t["y_mask"].sum((1, 2)).long().tolist()
    # padded output frames are NOT zero in the reference: z_p = noise * noise_scale there (m = logs = 0)
    pad = (1 - t["y_mask"]).bool().expand_as(z_p)
    if pad.any():
        assert torch.equal(z_p[pad], (t["noise"] * 1.0 * t["noise_scale"])[pad])


def test_oracle_expand_prior_degenerate_durations():
    """All-zero durations give y_length 1 (clamp_min) with no token selected; a zero-duration token in the middle is
    skipped; durations of several frames repeat the token."""
    g = torch.Generator().manual_seed(5)
    m_p, logs_p = torch.randn((2, 4, 5), generator=g), torch.randn((2, 4, 5), generator=g) * 0.1
    w = torch.tensor([[[0., 0., 0., 0., 0.]], [[2., 0., 3., 1., 0.]]])
    noise = torch.randn((2, 4, 6), generator=g)
    z_p, y_mask, attn, m, logs, y_len = orc.expand_prior(m_p, logs_p, w, noise, 0.5)
    assert y_len.tolist() == [1, 6]
    assert float(attn[0].abs().sum()) == 0.0 and torch.equal(z_p[0], noise[0] * 0.5)
    assert attn[1, 0].argmax(-1).tolist() == [0, 0, 2, 2, 2, 3]
    assert torch.equal(m[1], m_p[1][:, [0, 0, 2, 2, 2, 3]])
