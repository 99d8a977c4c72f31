"""Golden vectors for the CENT policy (reference com_marl/torch/policies/centralized_categorical_mlp_policy.py:11-97):
the UNMODIFIED reference CentralizedCategoricalMLPPolicy is imported through tests/golden/ref_harness.py (stubs for the
absent third-party modules only), initialised under torch.manual_seed(1) with hidden_sizes=(128, 64, 32) and tanh as the
runners do (exp_runners/*/runner_*_cent.py:48-58, env_uitils.py:188-189), given non-zero biases, and evaluated on
concatenated observations taken from a reference rollout.  Run in the build container (needs /root/reference):
    python tests/golden/make_golden_cent.py
"""
import json
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402


def run_case(ns, Cent, name, scenario, params, B, seed, relu=False):
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    T = params["max_env_steps"]
    if scenario == "pp":
        env = ns.PredatorPreyWrapper(centralized=True, other_agent_visible=True, params=params)
    else:
        env = ns.CoverageWrapper(centralized=True, other_agent_visible=True, max_steps=T, params=params)
    genv = ns.GarageEnv(env)
    n = env.n_agents
    torch.manual_seed(1)
    pol = Cent(genv.spec, n_agents=n, hidden_sizes=(128, 64, 32),
               hidden_nonlinearity=torch.nn.functional.relu if relu else torch.tanh, name="centralized")
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, v in pol.state_dict().items():
            if k.endswith("linear.bias"):
                v.copy_((torch.rand(v.shape, generator=g) - 0.5) * 0.2)
    obs_l, av_l = [], []
    obs = env.reset()
    rng = np.random.default_rng(seed)
    for b in range(B):
        obs_l.append(np.asarray(obs, dtype=np.float32).reshape(-1))
        av = np.ones((n, 5), dtype=np.float32)
        if b % 3 == 2:
            av[rng.integers(0, n), rng.integers(0, 5)] = 0
        av_l.append(av.reshape(-1))
        for _ in range(3):
            obs, _, done, _ = env.step(rng.integers(0, 5, size=n))
    obs_b, av_b = np.stack(obs_l), np.stack(av_l)
    logits = {}
    hook = pol._output_layers[0].register_forward_hook(lambda m, i, o: logits.__setitem__("v", o.detach().numpy().copy()))
    with torch.no_grad():
        dist = pol.forward(obs_b, av_b, get_actions=True)
        ent = pol.entropy(torch.from_numpy(obs_b)[None], torch.from_numpy(av_b)[None]).numpy()
    hook.remove()
    sd = {f"w::{k}": v.detach().numpy() for k, v in pol.state_dict().items()}
    meta = dict(name=name, scenario=scenario, n=n, D=obs_b.shape[-1] // n, B=B, relu=int(relu))
    np.savez_compressed(os.path.join(HERE, f"centpol_{name}.npz"), meta=np.array(json.dumps(meta)), obs=obs_b, avail=av_b,
                        probs=dist.probs.numpy(), logits=logits["v"], entropy=ent, **sd)
    print(f"centpol_{name}: n={n} D={meta['D']} B={B} keys={sorted(pol.state_dict().keys())} probs[0,0]={dist.probs.numpy()[0, 0]}")


def main():
    ns = H.load_reference()
    from com_marl.torch.policies.centralized_categorical_mlp_policy import CentralizedCategoricalMLPPolicy as Cent
    P = H.scenario_params
    run_case(ns, Cent, "c1", "pp", P("pp", 10, 1, 0.04, cap=2, loss=0), B=6, seed=71)
    run_case(ns, Cent, "c2", "co", P("co", 10, 1, 0.03, loss=0), B=6, seed=72)
    run_case(ns, Cent, "c2relu", "co", P("co", 10, 1, 0.03, loss=0), B=5, seed=76, relu=True)
    run_case(ns, Cent, "c3", "pp", P("pp", 20, 2, 0.08, cap=4, loss=0.2), B=4, seed=73)
    run_case(ns, Cent, "c4", "co", P("co", 30, 2, 0.06, loss=0.1), B=3, seed=74)


if __name__ == "__main__":
    main()
