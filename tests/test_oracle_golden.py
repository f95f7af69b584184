"""The oracle restatement replays the fixtures produced by the UNMODIFIED reference
(oracle/make_goldens.py ran the reference itself in the build container)."""
import os

import pytest
import torch

from oracle import swin_ave_oracle as O
from oracle.make_goldens import probe_index

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["tiny3stage_fusion", "tiny3stage_multimodal", "swin_t_T2_fusion", "swin_t_T10_B2_fusion", "swin_b_T2_B1_fusion", "swin_l_T2_B1_fusion"])
def test_oracle_reproduces_reference(name):
    fx = torch.load(os.path.join(GOLD, name + ".pt"))
    cfg = O.SwinAVEConfig(**fx["cfg"])
    sd = O.make_state(cfg, seed=fx["seed"], dtype=torch.float64, **fx["state_kw"])
    assert list(sd.keys()) == fx["state_keys"]
    assert {k: tuple(t.shape) for k, t in sd.items()} == fx["state_shapes"]
    a, v, labels = O.make_inputs(cfg, fx["batch"], seed=fx["seed"], dtype=torch.float64)
    out, loss, g = O.train_step(sd, a, v, labels, cfg, fx["mode"])
    assert float((out - fx["logits"]).abs().max()) < 1e-10
    assert abs(float(loss - fx["loss"])) < 1e-12
    assert set(g) == set(fx["grad_norm"])
    for n, ref in fx["grads"].items():
        assert float((g[n] - ref).abs().max()) <= 1e-9 * max(1.0, float(ref.abs().max())), n
    for n, nrm in fx["grad_norm"].items():
        assert abs(float(g[n].norm()) - nrm) <= 1e-9 * max(nrm, 1e-30), n
        probe = g[n].reshape(-1)[probe_index(g[n].numel())]
        assert float((probe - fx["grad_probe"][n]).abs().max()) <= 1e-9 * max(1.0, float(probe.abs().max())), n


def test_index_buffers_and_mask():
    """Integer / mask work must be bit-exact (relative_position_index, t_relative_coords, shift mask)."""
    idx = O.relative_position_index(7)
    assert idx.shape == (49, 49) and int(idx.min()) == 0 and int(idx.max()) == 168
    assert int(idx[0, 0]) == 84 and int(idx[0, 48]) == 0 and int(idx[48, 0]) == 168
    rel = O.temporal_relative_coords(10)
    assert rel.shape == (100,) and int(rel[0]) == 9 and int(rel[9]) == 0 and int(rel[90]) == 18
    m = O.shifted_window_mask(14, 7, 3)
    assert m.shape == (4, 49, 49) and set(m.unique().tolist()) == {-100.0, 0.0}
    assert float(m[0].abs().sum()) == 0.0  # the first window never straddles the wrap-around
    tok = O.window_token_index(14, 7, 3)
    assert sorted(tok.reshape(-1).tolist()) == list(range(196))  # a permutation of the tokens
    assert int(tok[0, 0]) == 3 * 14 + 3


def test_fusion_with_zero_gates_is_multimodal():
    cfg = O.SwinAVEConfig(img_size=112, num_frames=2, embed_dim=32, depths=[2, 2, 2], num_heads=[1, 2, 4],
                          adapter_mlp_ratio=[0.25, 0.25, 0.125], label_dim=5)
    sd = O.make_state(cfg, seed=1, dtype=torch.float64, gate_lo=0.0, gate_hi=0.0)
    a, v, _ = O.make_inputs(cfg, 1, seed=1, dtype=torch.float64)
    assert torch.equal(O.forward(sd, a, v, cfg, "fusion"), O.forward(sd, a, v, cfg, "multimodal"))
