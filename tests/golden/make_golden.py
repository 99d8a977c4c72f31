#!/usr/bin/env python
"""Generate golden input/output vectors by running the UNMODIFIED reference (cnuns/Com-MARL).

Run by hand in the build container (``/root/reference`` mounted):

    python tests/golden/make_golden.py            # (re)writes tests/golden/*.npz

The reference ships no tests and no golden vectors, so these fixtures — outputs of
the reference itself on seeded inputs — are what pins the oracle (``oracle/``) and, through it, the
CUDA path.  Random streams are *injected* at the points SURVEY.md §8c lists:

* prey moves  — ``numpy.random.choice`` is replaced while ``env.step`` runs by a feeder that returns
  ``cand[step][prey][trial]`` (predator_prey.py:396-407 consumes 1..5 draws per alive prey);
* packet loss — ``torch.rand`` is replaced while ``env.step``/``env.reset`` run by a feeder that
  returns planes of ``chan_u[update]`` (env_communication.py:212, gilbert_elliot_loss_model.py:139,143);
* spawn       — positions after every reference ``reset()`` are recorded (Python ``random`` is
  seeded but not replicated);
* actions     — supplied by this script.

The env is driven the way garage's VecEnvExecutor does (garage/sampler/vec_env_executor.py:19-45):
step, and when done, reset; the observation / comm state recorded for that step are the
post-reset ones.
"""
import json
import os
import random
import sys
from collections import deque

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import ref_harness as H  # noqa: E402
import streams  # noqa: E402


class RandFeeder:
    """Stands in for torch.rand while the reference env code runs."""

    def __init__(self):
        self.real = torch.rand
        self.planes = None
        self.k = 0
        self.max_k = 0

    def load(self, planes):
        self.planes = planes
        self.k = 0

    def __call__(self, *args, size=None, **kw):
        shape = tuple(size) if size is not None else tuple(args[0] if len(args) == 1 and not isinstance(args[0], int) else args)
        if self.planes is None:
            return self.real(size=shape)
        if len(shape) == 3:
            out = self.planes[self.k:self.k + shape[0]]
            self.k += shape[0]
        else:
            out = self.planes[self.k]
            self.k += 1
        self.max_k = max(self.max_k, self.k)
        out = torch.from_numpy(np.ascontiguousarray(out))
        assert tuple(out.shape) == shape, (out.shape, shape)
        return out


class ChoiceFeeder:
    """Stands in for numpy.random.choice while PredatorPrey.step runs."""

    def __init__(self):
        self.real = np.random.choice
        self.cand = None
        self.prey = None
        self.trial = 0
        self.consumed = 0

    def __call__(self, a, size=None, replace=True, p=None):
        if self.cand is None or self.prey is None:
            return self.real(a, size=size, replace=replace, p=p)
        v = int(self.cand[self.prey][self.trial])
        self.trial += 1
        self.consumed += 1
        return np.array([v])


def pack_rows(m):
    """(..., n, n) 0/1 -> packed little-endian bit rows uint8 (..., n, ceil(n/8))."""
    return np.packbits(np.asarray(m).astype(np.uint8), axis=-1, bitorder="little")


def guided_actions(env, n, rng):
    """BFS every agent toward its nearest unvisited free cell (generator-side helper; only used to
    reach the 'all cells covered' branch of coverage.py:378-382 within one episode)."""
    G = env._grid_shape[0]
    acts = []
    DR, DC = [1, 0, -1, 0], [0, -1, 0, 1]
    claimed = set()
    for i in range(n):
        start = tuple(env.agent_pos[i])
        prev = {start: None}
        q = deque([start])
        goal = None
        while q:
            cur = q.popleft()
            if cur != start and env._visited[cur[0]][cur[1]] == 0 and cur not in claimed:
                goal = cur
                break
            for a in range(4):
                nx = (cur[0] + DR[a], cur[1] + DC[a])
                if not (0 <= nx[0] < G and 0 <= nx[1] < G):
                    continue
                if env._base_grid[nx[0]][nx[1]] == 1 or nx in prev:
                    continue
                prev[nx] = (cur, a)
                q.append(nx)
        if goal is None:
            acts.append(int(rng.integers(0, 5)))
            continue
        claimed.add(goal)
        cur = goal
        a = 4
        while prev[cur] is not None:
            cur, a = prev[cur]
        # a little noise so collisions/penalties/noops still occur
        acts.append(a if rng.random() > 0.1 else int(rng.integers(0, 5)))
    return np.array(acts, dtype=np.int8)


def run_env_case(ns, name, scenario, params, steps, seed, bias_move=False, ge=None, guided=False,
                 max_path_length=None):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    T = params["max_env_steps"]
    if scenario == "pp":
        env = ns.PredatorPreyWrapper(centralized=True, other_agent_visible=True, params=params)
    else:
        env = ns.CoverageWrapper(centralized=True, other_agent_visible=True, max_steps=T, params=params)
    n = env.n_agents
    p = getattr(env, "n_preys", 0)
    L = params["n_gcn_layers"]
    if ge is not None:
        # SURVEY.md §8a "GE driver details": switch the unmodified update_communication_state to its GE branch
        env.channelType = "GE"
        env.Pgb, env.Pbg, env.GE_INIT = ge["Pgb"], ge["Pbg"], ge["GE_INIT"]
        env.loss_apply = ge["loss_apply"]
        env.Tmax = T
    planes = 2 * L + 1
    chan_u = streams.uniforms_f32(seed * 16 + 3, (steps + 1, planes, n, n))
    acts = streams.actions(seed * 16 + 1, (steps, n), bias_move=bias_move)
    cand = streams.prey_candidates(seed * 16 + 2, (steps, max(p, 1), 5))
    rng = np.random.default_rng(seed)

    rf, cf = RandFeeder(), ChoiceFeeder()
    torch.rand = rf
    np.random.choice = cf
    if scenario == "pp":
        orig_prm = env.prey_random_move

        def prm(prey_i):
            cf.prey, cf.trial = prey_i, 0
            try:
                return orig_prm(prey_i)
            finally:
                cf.prey = None

        env.prey_random_move = prm

    rec = {k: [] for k in ("obs", "reward", "details", "done", "success", "prey_alive", "agent_pos",
                           "prey_pos", "adj", "chan", "ave_deg", "visited", "total_capture")}
    spawn_a, spawn_p = [], []

    def snap_spawn():
        spawn_a.append(np.array([env.agent_pos[i] for i in range(n)], dtype=np.int8))
        if p:
            spawn_p.append(np.array([env.prey_pos[i] for i in range(p)], dtype=np.int8))

    def snap_state(obs):
        rec["obs"].append(np.asarray(obs, dtype=np.float32))
        rec["agent_pos"].append(np.array([env.agent_pos[i] for i in range(n)], dtype=np.int8))
        if p:
            rec["prey_pos"].append(np.array([env.prey_pos[i] for i in range(p)], dtype=np.int8))
        rec["adj"].append(pack_rows(env.dist_adj))
        rec["chan"].append(pack_rows(env.channels))
        rec["ave_deg"].append(float(env.ave_deg))
        if scenario == "co":
            rec["visited"].append(np.asarray(env._visited, dtype=np.uint8).copy())
            rec["total_capture"].append(int(env.total_capture_cnt))

    try:
        rf.load(chan_u[0])
        obs = env.reset()
        n_empty = int(getattr(env, "n_empty_cells", 0))
        snap_spawn()
        snap_state(obs)
        ts = 0
        for s in range(steps):
            if guided:
                acts[s] = guided_actions(env, n, rng)
            rf.load(chan_u[s + 1])
            cf.cand = cand[s]
            obs, (rew, det), done, info = env.step(np.array(acts[s], dtype=np.int64))
            ts += 1
            if max_path_length is not None and ts >= max_path_length:
                done = True
            rec["reward"].append(float(rew))
            rec["details"].append([float(det["capture_cnt"]), float(det["move_cnt"]), float(det["penalty_cnt"]),
                                   float(det["variable"]), float(det["vars2"])])
            rec["done"].append(bool(done))
            rec["success"].append(int(env.success))
            if p:
                rec["prey_alive"].append(np.asarray(info["prey_alive"], dtype=np.uint8).copy())
            if done:
                rf.load(chan_u[s + 1])
                obs = env.reset()
                ts = 0
                snap_spawn()
            snap_state(obs)
    finally:
        torch.rand = rf.real
        np.random.choice = cf.real

    meta = dict(name=name, scenario=scenario, params={k: v for k, v in params.items()}, steps=steps, seed=seed,
                n=n, p=p, L=L, T=T, ge=ge, planes=planes, n_empty_cells=n_empty,
                bound_return=float(env.bound_return), Rcom=int(env.Rcom), channelType=env.channelType,
                pl=float(env.pl), max_path_length=max_path_length, bias_move=bias_move, guided=guided,
                prey_draws_consumed=cf.consumed, chan_planes_used=rf.max_k,
                episodes=len(spawn_a))
    out = dict(meta=np.array(json.dumps(meta)), actions=acts.astype(np.int8),
               spawn_agent=np.stack(spawn_a),
               obs=np.stack(rec["obs"]), reward=np.array(rec["reward"], dtype=np.float64),
               details=np.array(rec["details"], dtype=np.float64), done=np.array(rec["done"], dtype=np.uint8),
               success=np.array(rec["success"], dtype=np.uint8), agent_pos=np.stack(rec["agent_pos"]),
               adj=np.stack(rec["adj"]), chan=np.stack(rec["chan"]),
               ave_deg=np.array(rec["ave_deg"], dtype=np.float64))
    if p:
        out.update(spawn_prey=np.stack(spawn_p), prey_alive=np.stack(rec["prey_alive"]),
                   prey_pos=np.stack(rec["prey_pos"]))
    if scenario == "co":
        out.update(visited=np.packbits(np.stack(rec["visited"]).reshape(len(rec["visited"]), -1), axis=-1, bitorder="little"),
                   total_capture=np.array(rec["total_capture"], dtype=np.int32))
    path = os.path.join(HERE, f"env_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name:18s} n={n:3d} steps={steps:4d} episodes={len(spawn_a):3d} "
          f"done={int(np.sum(rec['done'])):3d} success_max={max(rec['success'])} "
          f"reward_sum={np.sum(rec['reward']):9.2f} prey_draws={cf.consumed} -> {os.path.getsize(path) / 1024:.0f} KiB")
    return env


def run_ge_direct(ns):
    """gilbert_elliot_loss_model.get_init_state / get_next_state_matrix called directly."""
    out = {}
    rf = RandFeeder()
    torch.rand = rf
    try:
        for tag, n, seq, pgb, pbg in (("default", 8, 60, 0.0196, 0.282), ("busy", 33, 24, 0.3, 0.4)):
            u = streams.uniforms_f32(900 + n, (2 * seq + 1, n, n))
            rf.load(u)
            init = ns.ge.get_init_state(n, pgb, pbg)
            seq_states = ns.ge.get_next_state_matrix(seq, init.bool(), pgb, pbg, include_prev=True)
            out[f"{tag}_u"] = u
            out[f"{tag}_states"] = pack_rows(seq_states.numpy())
            out[f"{tag}_cfg"] = np.array([n, seq, pgb, pbg], dtype=np.float64)
    finally:
        torch.rand = rf.real
    np.savez_compressed(os.path.join(HERE, "ge_direct.npz"), **out)
    print("ge_direct written")


def run_policy_case(ns, name, scenario, params, B, seed, loss_for_masks, **policy_kwargs):
    """Reference CommCategoricalMLPPolicy forward on observations taken from a reference rollout."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    T = params["max_env_steps"]
    if scenario == "pp":
        env = ns.PredatorPreyWrapper(centralized=True, other_agent_visible=True, params=params)
    else:
        env = ns.CoverageWrapper(centralized=True, other_agent_visible=True, max_steps=T, params=params)
    genv = ns.GarageEnv(env)
    n = env.n_agents
    torch.manual_seed(1)  # policy weights = reference init under torch.manual_seed(1)
    pol = ns.CommCategoricalMLPPolicy(genv.spec, n_agents=n, **policy_kwargs)
    # non-zero biases so the bias paths are exercised (xavier init zeroes them)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, v in pol.state_dict().items():
            if k.endswith("linear.bias"):
                v.copy_((torch.rand(v.shape, generator=g) - 0.5) * 0.2)
    obs_l, adj_l, ch_l, av_l = [], [], [], []
    obs = env.reset()
    rng = np.random.default_rng(seed)
    for b in range(B):
        obs_l.append(np.asarray(obs, dtype=np.float32))
        adj_l.append(np.asarray(env.dist_adj, dtype=np.float32))
        ch_l.append(np.asarray(env.channels, dtype=np.float32))
        av = np.ones((n, 5), dtype=np.float32)
        if b % 3 == 2:  # a few masked actions (API supports them although these envs never mask)
            av[rng.integers(0, n), rng.integers(0, 5)] = 0
        av_l.append(av.reshape(-1))
        for _ in range(3):
            obs, _, done, _ = env.step(rng.integers(0, 5, size=n))
    obs_b, adj_b, ch_b, av_b = map(np.stack, (obs_l, adj_l, ch_l, av_l))
    if b >= 0 and loss_for_masks == "zero_row":
        ch_b[0, :, 1, :] = 0  # an agent whose every link (self included) is down: exercises the eps renorm
    logits = {}
    hook = pol.categorical_output_layer._output_layers[0].register_forward_hook(
        lambda m, i, o: logits.__setitem__("v", o.detach().numpy().copy()))
    with torch.no_grad():
        dist, attn = pol.forward(obs_b, av_b, adj_b, ch_b, get_actions=True)
    hook.remove()
    sd = {f"w::{k}": v.detach().numpy() for k, v in pol.state_dict().items()}
    meta = dict(name=name, scenario=scenario, n=n, D=obs_b.shape[-1] // n, B=B, L=params["n_gcn_layers"],
                policy_kwargs={k: (list(v) if isinstance(v, (tuple, list)) else v) for k, v in policy_kwargs.items()})
    np.savez_compressed(os.path.join(HERE, f"policy_{name}.npz"), meta=np.array(json.dumps(meta)),
                        obs=obs_b, avail=av_b, adj=pack_rows(adj_b), chan=pack_rows(ch_b),
                        probs=dist.probs.numpy(), attn=attn.numpy(), logits=logits["v"], **sd)
    print(f"policy_{name}: n={n} D={meta['D']} B={B} probs[0,0]={dist.probs.numpy()[0, 0]}")


def main():
    ns = H.load_reference()
    P = H.scenario_params
    # --- the five BASELINE.json configs (C1..C5) -------------------------------------
    run_env_case(ns, "pp_c1", "pp", P("pp", 10, 1, 0.04, cap=2, loss=0), steps=450, seed=11)
    run_env_case(ns, "co_c2", "co", P("co", 10, 1, 0.03, loss=0), steps=900, seed=12)
    run_env_case(ns, "pp_c3", "pp", P("pp", 20, 2, 0.08, cap=4, loss=0.2), steps=230, seed=13)
    run_env_case(ns, "co_c4", "co", P("co", 30, 2, 0.06, loss=0.1, max_env_steps=60), steps=130, seed=14)
    run_env_case(ns, "pp_c5", "pp", P("pp", 50, 2, 0.08, cap=4, loss=0, max_env_steps=30), steps=70, seed=15)
    # --- variants ----------------------------------------------------------------------------------
    run_env_case(ns, "pp_cap3", "pp", P("pp", 10, 2, 0.08, cap=3, loss=0.5), steps=260, seed=21, bias_move=True)
    run_env_case(ns, "pp_capture", "pp", P("pp", 6, 1, 0.08, cap=2, loss=0, n_agents=8, n_preys=3, max_env_steps=40,
                                           penalty=0.5, rm=0.25), steps=400, seed=22, bias_move=True)
    run_env_case(ns, "pp_rcom2", "pp", P("pp", 10, 1, 0.08, cap=4, loss=0.3, trRcom=2), steps=210, seed=23)
    run_env_case(ns, "pp_m30", "pp", P("pp", 30, 2, 0.08, cap=4, loss=0, max_env_steps=50), steps=110, seed=24)
    run_env_case(ns, "pp_mpl", "pp", P("pp", 10, 1, 0.04, cap=2, loss=1.0), steps=120, seed=25, max_path_length=25)
    run_env_case(ns, "co_m20", "co", P("co", 20, 2, 0.06, loss=0.3, max_env_steps=120), steps=260, seed=31, bias_move=True)
    run_env_case(ns, "co_full", "co", P("co", 10, 1, 0.03, loss=0, step_cost=0.05, rm=0.1), steps=700, seed=32, guided=True)
    run_env_case(ns, "co_hard", "co", P("co", 10, 2, 0.06, loss=0, obstComplex="Hard", trRcom=3), steps=450, seed=33, bias_move=True)
    # --- Gilbert-Elliot: unreachable through init_communication in the reference, driven directly ----
    for tag, init, la in (("good_l1", 1, 1), ("bad_l1", 0, 1), ("prop_l1", -1, 1), ("good_l0", 1, 0), ("bad_l0", 0, 0)):
        # (GE_INIT=-1, loss_apply=0) is broken in the reference itself: env_communication.py:121 adds a stray
        # leading axis (state becomes (1,L,n,n), then (L,L,n,n) after one step), so it has no defined
        # output to match and the engine rejects that combination.
        run_env_case(ns, f"pp_ge_{tag}", "pp", P("pp", 10, 1, 0.08, cap=2, loss=0.2, max_env_steps=40), steps=100, seed=41,
                     ge=dict(Pgb=0.2, Pbg=0.3, GE_INIT=init, loss_apply=la))
    run_env_case(ns, "co_ge_default", "co", P("co", 20, 1, 0.03, loss=0.2, max_env_steps=80), steps=170, seed=42,
                 ge=dict(Pgb=0.0196, Pbg=0.282, GE_INIT=1, loss_apply=1))
    run_ge_direct(ns)
    # --- policy forward ------------------------------------------------------------------------------
    run_policy_case(ns, "c1", "pp", P("pp", 10, 1, 0.04, cap=2, loss=0), B=6, seed=51, loss_for_masks=None)
    run_policy_case(ns, "c2", "co", P("co", 10, 1, 0.03, loss=0), B=6, seed=52, loss_for_masks=None)
    run_policy_case(ns, "c3", "pp", P("pp", 20, 2, 0.08, cap=4, loss=0.2), B=4, seed=53, loss_for_masks="zero_row")
    run_policy_case(ns, "c4", "co", P("co", 30, 2, 0.06, loss=0.1), B=3, seed=54, loss_for_masks=None)
    run_policy_case(ns, "c5", "pp", P("pp", 50, 2, 0.08, cap=4, loss=0), B=2, seed=55, loss_for_masks=None)
    run_policy_case(ns, "m30", "pp", P("pp", 30, 2, 0.08, cap=4, loss=0.3), B=2, seed=56, loss_for_masks="zero_row")


if __name__ == "__main__":
    main()
