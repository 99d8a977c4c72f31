"""garage sampler interface over the device rollout engine.

``DeviceRolloutSampler`` keeps the call contract of CentralizedMAOnPolicyVectorizedSampler
(com_marl/sampler/centralized_ma_on_policy_vectorized_sampler.py:20-245; base class
garage/sampler/base.py:5-47): ``__init__(algo, env, n_envs)``, ``start_worker()``,
``obtain_samples(itr, batch_size, whole_paths) -> list[dict]``, ``shutdown_worker()``; it reads
``algo.policy`` and ``algo.max_path_length`` and returns `paths` with the reference's keys and shapes
(SURVEY.md §8a a21).  Differences, on purpose:
  * n_envs real, independent env instances are stepped (the reference puts the SAME env object n_envs times
    in its list and keeps only env 0's reward — garage/sampler/vec_env_executor.py:27-28);
  * trajectories are recorded by the kernels into device buffers and split into per-episode dicts once,
    at the end of the call.
For training at scale skip the `paths` list and read ``engine.traj`` / ``obtain_batch`` directly.
"""
import time
from typing import List, Optional

import numpy as np
import torch

from . import _native as N
from .rollout import RolloutEngine


def _tabular():
    try:                                 # the reference records its timers in dowel's tabular
        from dowel import tabular
        return tabular
    except Exception:                    # dowel is optional here
        return None


def _unpack_bits(bits: np.ndarray, n: int) -> np.ndarray:
    b = np.ascontiguousarray(bits).view(np.uint32)
    cols = np.arange(n)
    return ((b[..., cols >> 5] >> (cols & 31).astype(np.uint32)) & 1).astype(np.float32)


class DeviceRolloutSampler:
    def __init__(self, algo, env, n_envs: Optional[int] = None, chunk: int = 25):
        self.algo = algo
        self.env = env
        base = getattr(env, "env", env)                      # GarageEnv-style wrapper or the wrapper itself
        self._spec = base.spec_b200
        self._n_envs = int(n_envs) if n_envs else 1
        self._n_agents = self._spec.n_agents
        self._chunk = int(chunk)
        self._engine: Optional[RolloutEngine] = None
        self.last_timers = {}

    def start_worker(self):
        spec = self._spec
        spec.max_path_length = int(self.algo.max_path_length or 0)
        policy = self.algo.policy
        # a policy without communication (Obs-DP) has no attention: the reference stores None per step then
        # (agent_info.get('attention_weights'), ...vectorized_sampler.py:186)
        self._comm = bool(getattr(policy, "comm", False))
        self._engine = RolloutEngine(spec, policy, self._n_envs, device=policy.device, ring=self._chunk,
                                     record_attention=self._comm, use_graph=False)

    def shutdown_worker(self):
        self._engine = None

    def obtain_samples(self, itr, batch_size=None, whole_paths=True) -> List[dict]:
        eng, spec = self._engine, self._spec
        n, p, L, B, D = self._n_agents, spec.n_preys, spec.n_layers, self._n_envs, spec.obs_dim
        if not batch_size:
            batch_size = self.algo.max_path_length * self._n_envs
        t0 = time.time()
        eng.reset()
        # per-env running buffers of host chunks
        keys = ("obs", "adj", "chan", "ave_deg", "actions", "probs", "attention", "reward", "done", "counts", "prey_alive")
        run = [{k: [] for k in keys} for _ in range(B)]
        paths, n_samples = [], 0
        details_fn = self._details
        bound = spec.bound_return
        while n_samples < batch_size:
            eng.run_chunk()
            t = {k: v.cpu().numpy() for k, v in eng.traj.items()}
            eng.env.check_errors()
            K = eng.K
            adj = _unpack_bits(t["adj_bits"][:K], n)                       # (K,B,n,n)
            chan = _unpack_bits(t["chan_bits"][:K], n)                     # (K,B,L,n,n)
            for b in range(B):
                done_idx = np.nonzero(t["done"][:, b])[0]
                bounds = [(int(dn) + 1, True) for dn in done_idx]
                if not bounds or bounds[-1][0] < K:
                    bounds.append((K, False))
                start = 0
                for stop, finished in bounds:
                    r = run[b]
                    m = stop - start
                    r["obs"].append(t["obs"][start:stop, b].reshape(m, n * D))
                    r["adj"].append(adj[start:stop, b].reshape(m, n * n))
                    r["chan"].append(chan[start:stop, b].reshape(m, L * n, n))
                    r["ave_deg"].append(t["ave_deg"][start:stop, b])
                    r["actions"].append(t["actions"][start:stop, b].astype(np.int64))
                    r["probs"].append(t["probs"][start:stop, b])
                    if self._comm:
                        r["attention"].append(t["attention"][start:stop, b])
                    r["reward"].append(t["reward"][start:stop, b])
                    r["done"].append(t["done"][start:stop, b].astype(bool))
                    r["counts"].append(t["counts"][start:stop, b])
                    r["prey_alive"].append(t["prey_alive_out"][start:stop, b, :max(p, 1)].astype(bool))
                    if finished:
                        c = {k: np.concatenate(v) for k, v in r.items() if v}
                        T_ = len(c["reward"])
                        agent_infos = {"action_probs": c["probs"]}
                        if self._comm:
                            agent_infos["attention_weights"] = c["attention"]
                        else:
                            c["attention"] = np.full(T_, None, dtype=object)
                        paths.append(dict(
                            observations=c["obs"], actions=c["actions"],
                            avail_actions=np.ones((T_, n * 5), dtype=np.int64),
                            rewards=c["reward"], rewards_details=np.asarray(details_fn(c["counts"], c["reward"])),
                            env_infos={"prey_alive": c["prey_alive"]} if p else {},
                            agent_infos=agent_infos,
                            dones=c["done"], dist_adjs=c["adj"],
                            ave_degs=(np.full(T_, n) if spec.rcom == 0 else c["ave_deg"]),
                            diameters=np.full(T_, n if spec.rcom == 0 else 0),
                            ave_trputs=np.full(T_, spec.ave_trput), attentions=c["attention"], channels=c["chan"],
                            # env.success read when the episode ended (…vectorized_sampler.py:194)
                            success=np.full(B, t["success"][stop - 1, b])))
                        n_samples += T_ * n
                        run[b] = {k: [] for k in keys}
                    start = stop
        self.last_timers = dict(TotalExecTime=time.time() - t0, BoundReturn=bound)
        tab = _tabular()
        if tab is not None:
            tab.record("PolicyExecTime", 0.0)
            tab.record("EnvExecTime", self.last_timers["TotalExecTime"])
            tab.record("ProcessExecTime", 0.0)
            tab.record("BoundReturn", bound)
        if whole_paths:
            return paths
        return _truncate_paths(paths, batch_size, n)

    def _details(self, counts, rewards):
        """per-step reward_details dicts rebuilt from the integer counts (predator_prey.py:440-448, coverage.py:308-315)"""
        n = float(self._n_agents)
        out = []
        pp = self._spec.scenario == "pp"
        for c, r in zip(counts.astype(np.float64), rewards):
            if pp:
                out.append(dict(reward=r, capture_cnt=int(c[0]), step_cnt=1, move_cnt=c[1] / n, penalty_cnt=int(c[2]),
                                variable=c[3] / n, vars2=0))
            else:
                out.append(dict(reward=r, capture_cnt=c[0] / n, step_cnt=1, move_cnt=c[1] / n, penalty_cnt=c[2] / n,
                                variable=c[3] / n, vars2=c[4] / n))
        return out


def _truncate_paths(paths, max_samples, n_agents):
    """garage.sampler.utils.truncate_paths semantics in agent-steps: drop whole paths from the end, then cut the
    last one."""
    paths = list(paths)
    total = sum(len(p["rewards"]) * n_agents for p in paths)
    while paths and total - len(paths[-1]["rewards"]) * n_agents >= max_samples:
        total -= len(paths.pop(-1)["rewards"]) * n_agents
    if paths and total > max_samples:
        last = paths.pop(-1)
        keep = len(last["rewards"]) - (total - max_samples) // n_agents
        cut = {}
        for k, v in last.items():
            if isinstance(v, dict):
                cut[k] = {kk: vv[:keep] for kk, vv in v.items()}
            elif k == "success":
                cut[k] = v
            else:
                cut[k] = v[:keep]
        paths.append(cut)
    return paths
