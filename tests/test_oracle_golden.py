"""Pin the CPU oracle (oracle/polydis_oracle.py) to fixtures produced by the unmodified reference.

The fixtures in tests/golden/*.npz were written by tests/golden/make_golden.py, which imports
/root/reference in the build container.  fp32 CPU results are deterministic up to thread-count
reassociation (~1e-7), hence the tight tolerances.
"""
import os
import random

import numpy as np
import pytest
import torch

from oracle import polydis_oracle as O
from polydis_b200.synth import synth_batch, pr_mat_to_grid
from polydis_b200.weights import make_state_dict, STATE_DICT_SPEC
from tests.golden.make_golden_probe import probe_indices


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_grid_builder_matches_reference_converter(golden_dir):
    g = _load(golden_dir, "grid.npz")
    x = pr_mat_to_grid(g["pr_mat"].astype(np.float32))
    assert np.array_equal(x, g["x"].astype(np.int64))
    x2, c2, pr2 = synth_batch(6, 11)
    assert np.array_equal(x2, g["x"]) and np.array_equal(pr2, g["pr_mat"]) and np.array_equal(c2, g["c"])


def test_grid_builder_rejects_overflow():
    pr = np.zeros((1, 32, 128), np.float32)
    pr[0, 0, 10:25] = 1
    with pytest.raises(ValueError):
        pr_mat_to_grid(pr)


def test_grid_builder_empty_and_full_steps():
    pr = np.zeros((2, 32, 128), np.float32)
    pr[1, 3, 20:34] = 32                      # 14 notes: the largest step that fits
    x = pr_mat_to_grid(pr)
    assert (x[0, :, 0, 0] == 128).all() and (x[0, :, 1, 0] == 129).all() and (x[0, :, 2:, 0] == 130).all()
    assert x[1, 3, 15, 0] == 129 and (x[1, 3, 1:15, 0] == np.arange(20, 34)).all()
    assert (x[1, 3, 1:15, 1:] == 1).all() and (x[1, 3, 15, 1:] == 2).all()


def test_state_dict_contract():
    sd = make_state_dict(0)
    assert len(sd) == 81 and sum(v.numel() for v in sd.values()) == 27310079
    assert [k for k, _, _ in STATE_DICT_SPEC] == list(sd.keys())


@pytest.mark.parametrize("tag", ["tf111", "tf000", "tf555"])
def test_oracle_training_matches_reference(golden_dir, tag):
    g = _load(golden_dir, f"train_{tag}.npz")
    B = int(g["B"])
    x, c, pr = (torch.from_numpy(a) for a in synth_batch(B, int(g["data_seed"])))
    sd = {k: v.requires_grad_(True) for k, v in
          make_state_dict(int(g["w_seed"]), gain=float(g["gain"]), eos_bias=float(g["eos_bias"])).items()}
    random.seed(int(g["rng_seed"]))
    plan = O.draw_plan(*[float(v) for v in g["tfr"]])
    e1, e2 = torch.from_numpy(g["eps_chd"]), torch.from_numpy(g["eps_rhy"])
    out = O.run(sd, x, c, pr, plan, e1, e2)
    losses = O.loss_function(x, c, *out, 0.1, (1, 0.5))
    np.testing.assert_allclose([float(v) for v in losses], g["losses"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(out[0].detach().numpy(), g["pitch"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out[1].detach().numpy(), g["dur"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out[2][0].detach().numpy(), g["mu_chd"], atol=2e-6)
    np.testing.assert_allclose(out[3][1].detach().numpy(), g["std_rhy"], atol=2e-6)
    np.testing.assert_allclose(out[4].detach().numpy(), g["root"], atol=2e-6)
    np.testing.assert_allclose(out[5].detach().numpy(), g["chroma"], atol=2e-6)
    np.testing.assert_allclose(out[6].detach().numpy(), g["bass"], atol=2e-6)
    losses[0].backward()
    for i, (name, _, _) in enumerate(STATE_DICT_SPEC):
        gr = sd[name].grad.reshape(-1).double()
        assert abs(float(gr.norm()) - g["grad_norm"][i]) <= 1e-4 * g["grad_norm"][i] + 1e-9, name
        pr_ = gr[torch.from_numpy(probe_indices(name, gr.numel()))].numpy()
        np.testing.assert_allclose(pr_, g["grad_probe"][i], rtol=1e-3,
                                   atol=1e-5 * g["grad_norm"][i] + 1e-10, err_msg=name)


@pytest.mark.parametrize("tag", ["w0", "w1"])
def test_oracle_greedy_tokens_match_reference(golden_dir, tag):
    g = _load(golden_dir, f"infer_{tag}.npz")
    x, c, pr = (torch.from_numpy(a) for a in synth_batch(int(g["B"]), int(g["data_seed"])))
    sd = make_state_dict(int(g["w_seed"]), gain=float(g["gain"]), eos_bias=float(g["eos_bias"]))
    est = O.inference(sd, pr, c)
    assert est.shape == (int(g["B"]), 32, 15, 6) and est.dtype == np.int64
    assert np.array_equal(est, g["est_x"].astype(np.int64))


def test_plan_consumes_random_like_the_reference():
    random.seed(3)
    O.draw_plan(0.5, 0.5, 0.5, training=True)
    a = random.random()
    random.seed(3)
    for _ in range(487):
        random.random()
    assert a == random.random()
    random.seed(3)
    O.draw_plan(0.0, 0.0, 0.0, training=False)
    b = random.random()
    random.seed(3)
    for _ in range(479):
        random.random()
    assert b == random.random()
