"""Batched evaluation driver (com_marl_b200.evaluate.eval_model, the device version of exp_runners/*/eval_*.py):
output structure of the reference, every episode replayed on the oracle with the recorded greedy actions, determinism and
independence of the env grouping."""
import os
import sys

import numpy as np
import pytest

from oracle import oracle as orc

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import ref_harness  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scen,kind", [("pp", "comm"), ("co", "comm"), ("pp", "dec")])
def test_eval_model_matches_oracle(scen, kind):
    from com_marl_b200.evaluate import VECTORS, eval_model
    from com_marl_b200.rollout import make_policy
    from com_marl_b200.scenario import ScenarioSpec
    T = 30
    params = ref_harness.scenario_params(scen, 10, 1, 0.08 if scen == "pp" else 0.03, cap=2, loss=0.2, max_env_steps=T)
    spec = ScenarioSpec.from_params(scen, params, seed=1)
    pol = make_policy(spec, kind=kind)
    n, B, steps = spec.n_agents, 37, 24                      # steps < T: most episodes are cut by max_env_steps
    episode_data, epi_success, epi_rewards, bound, traj, length = eval_model(
        spec, pol, n_eval_episodes=B, max_env_steps=steps, seed=5, groups=3, chunk=10, return_trajectory=True)
    assert len(episode_data) == B and len(epi_success) == B and set(epi_rewards) == set(VECTORS)
    # replay on the oracle with the recorded actions (no auto-reset: episodes stop being compared after their end)
    oenv = orc.OracleVecEnv(orc.spec_from_params(scen, params, seed=5), B)
    oenv.reset()
    alive = np.ones(B, dtype=bool)
    w = {k: v.cpu().numpy() for k, v in pol.state_dict().items()}
    for k in range(steps):
        if kind == "comm":
            _, pr, _ = orc.policy_forward(w, oenv.obs, None, oenv.adj, oenv.chan)
        else:
            _, pr = orc.policy_forward_dec(w, oenv.obs, None)
        a = traj["actions"][k]
        top = np.sort(pr, axis=-1)
        clear = (top[..., -1] - top[..., -2]) > 1e-4         # greedy == argmax wherever the oracle's argmax is not a near tie
        assert np.array_equal(a[alive][clear[alive]], np.argmax(pr, axis=-1)[alive][clear[alive]])
        oenv.step(a, auto_reset=False)
        assert np.array_equal(traj["reward"][k][alive], oenv.reward[alive])
        assert np.array_equal(traj["done"][k][alive], oenv.done[alive])
        alive &= ~(oenv.done.astype(bool))
        if not alive.any():
            break
    for b in range(B):
        step_success, step_data = episode_data[b]
        Tb = int(length[b])
        assert len(step_success) == Tb and all(len(step_data[v]) == Tb for v in VECTORS)
        assert step_data["reward"] == [float(x) for x in traj["reward"][:Tb, b]]
        assert epi_rewards["reward"][b] == np.sum(step_data["reward"]) and epi_rewards["step_cnt"][b] == Tb
        assert epi_success[b] == step_success[-1]
        assert Tb == steps or traj["done"][Tb - 1, b]
    # deterministic, and independent of the grouping / chunking
    again = eval_model(spec, pol, n_eval_episodes=B, max_env_steps=steps, seed=5, groups=1, chunk=24)
    assert again[1] == epi_success and again[2]["reward"] == epi_rewards["reward"] and again[2]["nodeDeg"] == epi_rewards["nodeDeg"]
