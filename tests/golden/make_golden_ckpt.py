"""Checkpoint fixtures: the UNMODIFIED reference policies (Comm-DP, Obs-DP, CENT), imported through
tests/golden/ref_harness.py, are pickled exactly as the reference's snapshotter does it — ``pickle.dump(algo.policy)`` into
``itrs/itr_%04d.pkl`` (garage/experiment/snapshotter.py:100-104) — next to the probabilities the reference computes from
those weights on a fixed input.  It also checks the other direction once: a reference policy loads the plain state_dict
this package writes (com_marl_b200.checkpoint.save_state_dict) and reproduces the same probabilities.
Run in the build container (needs /root/reference):  python tests/golden/make_golden_ckpt.py
"""
import json
import os
import pickle
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_harness as H  # noqa: E402


def main():
    ns = H.load_reference()
    from com_marl.torch.policies.centralized_categorical_mlp_policy import CentralizedCategoricalMLPPolicy as Cent
    from com_marl.torch.policies.comm_categorical_mlp_policy import CommCategoricalMLPPolicy as Comm
    from com_marl.torch.policies.dec_categorical_mlp_policy import DecCategoricalMLPPolicy as Dec
    params = H.scenario_params("pp", 10, 1, 0.04, cap=2, loss=0.2)
    env = ns.PredatorPreyWrapper(centralized=True, other_agent_visible=True, params=params)
    genv = ns.GarageEnv(env)
    n, L = env.n_agents, params["n_gcn_layers"]
    D = genv.spec.observation_space.flat_dim // n
    out = os.path.join(HERE, "ckpt")
    os.makedirs(os.path.join(out, "itrs"), exist_ok=True)
    rng = np.random.default_rng(7)
    B = 5
    obs = rng.random((B, n * D)).astype(np.float32)
    avail = np.ones((B, n * 5), dtype=np.float32)
    adj = (rng.random((B, n, n)) < 0.7).astype(np.float32)
    chan = (rng.random((B, L, n, n)) < 0.8).astype(np.float32)
    for i in range(n):
        adj[:, i, i] = 1
        chan[:, :, i, i] = 1
    rec = dict(obs=obs, avail=avail, adj=adj, chan=chan)
    meta = dict(n=n, D=D, L=L, files={})
    for itr, (kind, build) in enumerate((("comm", lambda: Comm(genv.spec, n_agents=n, n_gcn_layers=L)),
                                         ("dec", lambda: Dec(genv.spec, n_agents=n, hidden_sizes=(128, 64, 32))),
                                         ("cent", lambda: Cent(genv.spec, n_agents=n, hidden_sizes=(128, 64, 32), hidden_nonlinearity=torch.tanh))), start=7):
        torch.manual_seed(100 + itr)
        pol = build()
        g = torch.Generator().manual_seed(itr)
        with torch.no_grad():
            for k, v in pol.state_dict().items():
                if k.endswith("bias"):
                    v.copy_((torch.rand(v.shape, generator=g) - 0.5) * 0.2)
        name = f"itr_{str(itr).zfill(4)}.pkl"
        with open(os.path.join(out, "itrs", name), "wb") as f:          # snapshotter.py:100-104
            pickle.dump(pol, f)
        with torch.no_grad():
            if kind == "comm":
                dist, _ = pol.forward(obs, avail, adj, chan, get_actions=True)
            else:
                dist = pol.forward(obs, avail, get_actions=True)
        rec[f"probs_{kind}"] = dist.probs.numpy()
        meta["files"][kind] = name
        # the other direction: the state_dict this package writes loads into a fresh reference policy
        from com_marl_b200.checkpoint import load_reference_checkpoint, save_state_dict
        mine = load_reference_checkpoint(os.path.join(out, "itrs", name), device="cpu")
        tmp = os.path.join(out, "_tmp_sd.pkl")
        save_state_dict(mine, tmp)
        torch.manual_seed(999)
        fresh = build()
        fresh.load_state_dict({k: torch.as_tensor(v) for k, v in pickle.load(open(tmp, "rb")).items()})
        os.remove(tmp)
        with torch.no_grad():
            d2 = fresh.forward(obs, avail, adj, chan, get_actions=True)[0] if kind == "comm" else fresh.forward(obs, avail, get_actions=True)
        assert np.array_equal(d2.probs.numpy(), rec[f"probs_{kind}"]), kind
        print(kind, name, os.path.getsize(os.path.join(out, "itrs", name)), "bytes; reference reloads our state_dict: ok")
    np.savez_compressed(os.path.join(out, "expected.npz"), meta=np.array(json.dumps(meta)), **rec)


if __name__ == "__main__":
    main()
