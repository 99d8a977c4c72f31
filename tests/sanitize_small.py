"""Smallest end-to-end exercise of every kernel, meant to run under compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tests/sanitize_small.py
    compute-sanitizer --tool racecheck python tests/sanitize_small.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from com_marl_b200.rollout import RolloutEngine, make_policy  # noqa: E402
from com_marl_b200.scenario import ScenarioSpec  # noqa: E402


def main():
    cases = [("pp", 10, 1, 0.04, 2, 0.3, 40, "auto", {"max_env_steps": 6}),        # lane groups of 4, tcgen05 policy, IID
             ("co", 10, 1, 0.03, 2, 0.0, 40, "fp32", {"max_env_steps": 6}),        # Coverage, FFMA policy
             ("pp", 20, 2, 0.08, 4, 0.2, 9, "auto", {"max_env_steps": 5}),         # n = 32: whole-warp groups
             ("pp", 30, 2, 0.08, 4, 0.0, 3, "auto", {"max_env_steps": 4})]         # n = 72: large-team FFMA kernel
    for scen, m, sen, den, cap, loss, B, math, over in cases:
        spec = ScenarioSpec.from_cli(scen, m, sen, den, cap=cap, loss=loss, seed=3, **over)
        pol = make_policy(spec, math=math)
        eng = RolloutEngine(spec, pol, B, ring=4, use_graph=False, record_attention=True)
        eng.reset()
        eng.run(12)
        torch.cuda.synchronize()
        eng.env.check_errors()
        pol.check_errors()
        assert np.isfinite(eng.traj["probs"].cpu().numpy()).all()
        print(scen, m, "n =", spec.n_agents, "episodes finished:", int(eng.local_stats()[0].item()))
    # Obs-DP policy (rows independent of the team) and env groups on separate streams inside a captured graph
    spec = ScenarioSpec.from_cli("co", 10, 1, 0.03, cap=2, loss=0.1, seed=3, max_env_steps=6)
    pol = make_policy(spec, kind="dec")
    eng = RolloutEngine(spec, pol, 50, ring=3, use_graph=True, groups=3)
    eng.reset()
    eng.run(9)
    torch.cuda.synchronize()
    eng.env.check_errors()
    pol.check_errors()
    print("obs-dp / env groups ok")
    # Gilbert-Elliot channel + comm-only entry point + mask converters
    spec = ScenarioSpec.from_cli("pp", 10, 1, 0.08, cap=2, loss=0.2, channel_type="GE", max_env_steps=5)
    from com_marl_b200.envs import BatchedEnv
    env = BatchedEnv(spec, 17)
    env.reset()
    for _ in range(7):
        env.step(np.random.default_rng(0).integers(0, 5, size=(17, spec.n_agents)).astype(np.int8))
    env.comm_update()
    env.dist_adj(); env.channels()
    torch.cuda.synchronize()
    env.check_errors()
    print("sanitize_small ok")


if __name__ == "__main__":
    main()
