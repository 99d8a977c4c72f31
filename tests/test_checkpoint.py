"""Checkpoint interchange: ``itrs/itr_%04d.pkl`` files pickled from the UNMODIFIED reference policies by
tests/golden/make_golden_ckpt.py (exactly what garage/experiment/snapshotter.py:100-104 writes) load into this package's
policies without the reference installed, and reproduce the probabilities the reference computed from those weights."""
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CKPT = os.path.join(HERE, "golden", "ckpt")
KINDS = ["comm", "dec", "cent"]


def _expected():
    z = np.load(os.path.join(CKPT, "expected.npz"))
    return z, json.loads(str(z["meta"]))


@pytest.mark.parametrize("kind", KINDS)
def test_reference_checkpoint_loads_without_the_reference(kind):
    """CPU: the shim unpickler needs none of the reference's modules; the differentiable forward (torch ops) of the loaded
    policy equals the reference's probabilities"""
    from com_marl_b200 import checkpoint as ck
    assert not any(m.split(".")[0] in ("com_marl", "garage", "akro", "gym") for m in sys.modules if "com_marl_b200" not in m)
    z, meta = _expected()
    n, epoch = meta["n"], int(meta["files"][kind][4:8])
    pol = ck.load_policy(CKPT, epoch, device="cpu")            # exp_runners/testing.py:68-77 naming (zero-padded fallback)
    assert type(pol).__name__ == {"comm": "CommCategoricalMLPPolicy", "dec": "DecCategoricalMLPPolicy",
                                  "cent": "CentralizedCategoricalMLPPolicy"}[kind]
    assert pol._n_agents == n and pol._dec_obs_dim == meta["D"]
    obs, avail = torch.as_tensor(z["obs"]), torch.as_tensor(z["avail"])
    with torch.no_grad():
        if kind == "comm":
            d, _ = pol.forward(obs, avail, torch.as_tensor(z["adj"]).reshape(len(obs), -1),
                               torch.as_tensor(z["chan"]).reshape(len(obs), meta["L"] * n, n))
        else:
            d = pol.forward(obs, avail)
    assert np.abs(d.probs.numpy() - z[f"probs_{kind}"]).max() <= 1e-6
    with pytest.raises(FileNotFoundError):
        ck.load_policy(CKPT, 1234, device="cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_reference_checkpoint_runs_on_the_kernels(kind):
    """GPU: the loaded policy's rollout call (the fused kernels) reproduces the reference's probabilities"""
    from com_marl_b200 import checkpoint as ck
    z, meta = _expected()
    n, D, L = meta["n"], meta["D"], meta["L"]
    pol = ck.load_reference_checkpoint(os.path.join(CKPT, "itrs", meta["files"][kind]), device="cuda")
    if kind == "comm":
        _, infos = pol.get_actions(z["obs"], z["avail"], z["adj"], z["chan"], greedy=True)
    else:
        _, infos = pol.get_actions(z["obs"], z["avail"], greedy=True)
    probs = np.stack(infos["action_probs"])
    assert np.abs(probs - z[f"probs_{kind}"]).max() <= 1e-5
