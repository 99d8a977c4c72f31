"""Batched evaluation driver: the reference's eval loop with every episode running in parallel on the device.

Replaces ``eval_model`` of exp_runners/predatorprey/eval_pp.py:9-104 / exp_runners/coverage/eval_co.py (same loop): the
reference plays ``n_eval_episodes`` episodes one after another — reset, then up to ``max_env_steps`` greedy policy calls
and env steps, recording per step the reward details (``VECTORS`` of exp_runners/testing.py:209; ``nodeDeg`` is
``env.ave_deg`` after the step) and ``env.success`` — and returns ``(episode_data, epi_success, epi_rewards,
bound_return)`` for the CSV / MATLAB writers of testing.py:213-367.  Here the episodes are ``n_eval_episodes``
independent envs of one RolloutEngine (no auto-reset: an env that is done is simply ignored from then on), stepped by the
fused kernels with greedy action selection; only the small per-step outputs (reward, done, counts, ave_deg, success)
ever reach the host.
"""
import copy
from typing import Dict, List, Tuple

import numpy as np
import torch

from .rollout import RolloutEngine

VECTORS = ["reward", "capture_cnt", "step_cnt", "move_cnt", "penalty_cnt", "nodeDeg", "variable", "vars2"]


def _details(scenario: str, n: int, counts: np.ndarray, reward: float, ave_deg) -> Dict[str, float]:
    """one step's reward_details (predator_prey.py:440-448, coverage.py:308-315) + nodeDeg, from the integer counts"""
    c = counts.astype(np.float64)
    nf = float(n)
    if scenario == "pp":
        return dict(reward=reward, capture_cnt=int(c[0]), step_cnt=1, move_cnt=c[1] / nf, penalty_cnt=int(c[2]),
                    nodeDeg=ave_deg, variable=c[3] / nf, vars2=0)
    return dict(reward=reward, capture_cnt=c[0] / nf, step_cnt=1, move_cnt=c[1] / nf, penalty_cnt=c[2] / nf,
                nodeDeg=ave_deg, variable=c[3] / nf, vars2=c[4] / nf)


def eval_model(env, policy, itr=0, n_eval_episodes=100, max_env_steps=200, eval_greedy=True, render=False,
               inspect_steps=False, seed=1, flag=None, groups: int = 4, chunk: int = 50, return_trajectory: bool = False):
    """``env``: a com_marl_b200 gym wrapper (or anything with ``spec_b200``) or a ScenarioSpec.  Returns the reference's
    ``(episode_data, epi_success, epi_rewards, bound_return)``: ``episode_data[i] = (step_success, {vec: [per step]})``,
    ``epi_rewards[vec][i]`` = sum over the episode's steps (mean for ``nodeDeg``), ``epi_success[i]`` = ``env.success``
    when the episode ended."""
    if render:
        raise NotImplementedError("rendering is out of scope (DESIGN.md §8)")
    if flag is not None and flag[0]:
        return None, None, None, None
    spec = copy.copy(getattr(env, "spec_b200", env))
    spec.seed = int(seed)
    spec.max_path_length = 0
    n, B = spec.n_agents, int(n_eval_episodes)
    K = max(1, min(int(chunk), int(max_env_steps)))
    eng = RolloutEngine(spec, policy, B, device=policy.device, ring=K, greedy=bool(eval_greedy), use_graph=True,
                        groups=groups, auto_reset=False)
    eng.reset()
    length = np.zeros(B, dtype=np.int64)              # 0 = still running
    keep = {k: [] for k in ("reward", "done", "counts", "ave_deg", "success", "actions")}
    steps = 0
    while steps < max_env_steps and (length == 0).any():
        eng.run_chunk()
        t = eng.traj
        m = min(K, max_env_steps - steps)
        keep["reward"].append(t["reward"][:m].cpu().numpy())
        keep["done"].append(t["done"][:m].cpu().numpy())
        keep["counts"].append(t["counts"][:m].cpu().numpy())
        keep["ave_deg"].append(t["ave_deg"][1:m + 1].cpu().numpy())        # env.ave_deg AFTER each step
        keep["success"].append(t["success"][:m].cpu().numpy())
        if return_trajectory:
            keep["actions"].append(t["actions"][:m].cpu().numpy())
        d = keep["done"][-1]
        for b in np.nonzero(length == 0)[0]:
            hit = np.nonzero(d[:, b])[0]
            if hit.size:
                length[b] = steps + int(hit[0]) + 1
        steps += m
    eng.env.check_errors()
    policy.check_errors()
    length[length == 0] = steps                          # cut by max_env_steps (eval_pp.py:71)
    cat = {k: np.concatenate(v) for k, v in keep.items() if v}
    episode_data: List[Tuple[list, dict]] = []
    epi_success: List[int] = []
    epi_rewards: Dict[str, list] = {vec: [] for vec in VECTORS}
    fully = spec.rcom == 0
    for b in range(B):
        T = int(length[b])
        step_data = {vec: [] for vec in VECTORS}
        for k in range(T):
            deg = n if fully else np.float32(cat["ave_deg"][k, b])
            det = _details(spec.scenario, n, cat["counts"][k, b], float(cat["reward"][k, b]), deg)
            for vec in VECTORS:
                step_data[vec].append(det[vec])
        step_success = [int(x) for x in cat["success"][:T, b]]
        episode_data.append((step_success, step_data))
        epi_success.append(step_success[-1])
        for vec in VECTORS:
            epi_rewards[vec].append(np.mean(step_data[vec]) if vec == "nodeDeg" else np.sum(step_data[vec]))
    out = (episode_data, epi_success, epi_rewards, spec.bound_return)
    if return_trajectory:
        return out + ({k: cat[k] for k in ("actions", "reward", "done", "ave_deg", "success", "counts")}, length)
    return out
