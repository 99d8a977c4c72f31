"""Device-resident rollout loop: policy forward + sampling -> env step -> comm update, no host round trip.

Replaces the hot loop of CentralizedMAOnPolicyVectorizedSampler.obtain_samples
(com_marl/sampler/centralized_ma_on_policy_vectorized_sampler.py:119-231): per iteration the reference
gathers dist_adj / channels from the envs (:123-127), calls policy.get_actions (:134), steps the
VecEnvExecutor (:143) and appends to Python lists (:158-191).  Here one iteration is two kernel launches
(cm_policy_forward, cm_env_step) on one stream; K consecutive iterations write into the K slots of a
trajectory ring in HBM (the kernels write the slots directly, there is no copy) and are captured once in
a CUDA graph that is replayed; episode accounting is accumulated by the env kernel (`stats`).
"""
from typing import Dict, Optional

import numpy as np
import torch

from .envs import BatchedEnv
from .policy import CommCategoricalMLPPolicy
from .scenario import ScenarioSpec
from .spaces import Box, Discrete, EnvSpec

STAT_KEYS = ("episodes", "return_sum", "length_sum", "success_sum", "c0_sum", "moved_sum", "c2_sum", "c3_sum", "c4_sum")


def make_policy(spec: ScenarioSpec, device="cuda", seed: Optional[int] = None, torch_seed: int = 1,
                math: str = "auto", kind: str = "comm"):
    """Policy with the reference initialisation under torch.manual_seed(torch_seed)."""
    n, D = spec.n_agents, spec.obs_dim
    env_spec = EnvSpec(Box(np.zeros(n * D, np.float32), np.ones(n * D, np.float32)), Discrete(5))
    torch.manual_seed(torch_seed)
    if kind == "dec":      # Obs-DP: no communication (dec_categorical_mlp_policy.py)
        from .policy import DecCategoricalMLPPolicy
        return DecCategoricalMLPPolicy(env_spec, n, device=device, seed=spec.seed if seed is None else seed)
    if kind == "cent":     # CENT: one MLP over the concatenated observation (centralized_categorical_mlp_policy.py)
        from .policy import CentralizedCategoricalMLPPolicy
        return CentralizedCategoricalMLPPolicy(env_spec, n, device=device, seed=spec.seed if seed is None else seed)
    return CommCategoricalMLPPolicy(env_spec, n, n_gcn_layers=spec.n_layers, device=device,
                                    seed=spec.seed if seed is None else seed, math=math)


class RolloutEngine:
    def __init__(self, spec: ScenarioSpec, policy: CommCategoricalMLPPolicy, n_envs: int, device="cuda", env_id0: int = 0,
                 ring: int = 8, record_attention: bool = False, greedy: bool = False, use_graph: bool = True,
                 groups: int = 1, auto_reset: bool = True):
        self.spec, self.policy, self.B = spec, policy, int(n_envs)
        self.device = torch.device(device)
        self.env = BatchedEnv(spec, n_envs, device=device, env_id0=env_id0, auto_reset=auto_reset)
        # Env groups: the envs are independent, so G contiguous groups form G independent policy -> step -> policy ...
        # chains.  Each chain runs on its own stream (forked / joined inside the captured graph): there is no device-wide
        # barrier between a step of one group and the next step of another, so partially filled waves of one launch
        # are covered by the launches of the other groups.  Results are identical for every G (global env ids key the
        # random streams).  Large teams (n > 64) share one policy scratch buffer and stay in a single group.
        G = max(1, min(int(groups), self.B)) if (spec.n_agents <= 64 or not getattr(policy, "comm", True)) else 1
        self.groups = G
        cuts = [self.B * g // G for g in range(G + 1)]
        self._ranges = [(cuts[g], cuts[g + 1]) for g in range(G) if cuts[g + 1] > cuts[g]]
        self._envs = [self.env] if len(self._ranges) == 1 else [self.env.slice(b0, b1) for b0, b1 in self._ranges]
        self._streams = None
        self.K, self.greedy, self.use_graph = int(ring), bool(greedy), bool(use_graph)
        e, K, B, dev = self.env, self.K, self.B, self.device
        n, L, W, D, p = e.n, e.L, e.W, e.D, max(e.p, 1)

        def z(shape, dt):
            return torch.zeros(shape, dtype=dt, device=dev)

        # trajectory ring: inputs of step k live in slot k, its outputs in slot k (+1 for obs / comm state)
        self.traj: Dict[str, torch.Tensor] = dict(
            obs=z((K + 1, B, n, D), torch.float32), adj_bits=z((K + 1, B, n, W), torch.int32),
            chan_bits=z((K + 1, B, L, n, W), torch.int32), ave_deg=z((K + 1, B), torch.float32),
            actions=z((K, B, n), torch.int8), probs=z((K, B, n, 5), torch.float32),
            reward=z((K, B), torch.float64), done=z((K, B), torch.uint8), counts=z((K, B, 6), torch.int32),
            prey_alive_out=z((K, B, p), torch.uint8), success=z((K, B), torch.uint8))
        if record_attention:
            self.traj["attention"] = z((K, B, n, n), torch.float32)
        self._graph = None
        self._warm = False
        self.steps_done = 0
        self.kernel_launches = 0

    # ---- single iteration --------------------------------------------------------------------------
    def _iteration(self, k: int, g: int = None):
        """policy forward + sampling, then env step, for ring slot k (group g of the envs, or all groups one after another)"""
        if g is None:
            for gi in range(len(self._ranges)):
                self._iteration(k, gi)
            return
        t, e = self.traj, self._envs[g]
        b0, b1 = self._ranges[g]
        self.policy.act_device(t["obs"][k, b0:b1], adj_bits=t["adj_bits"][k, b0:b1], chan_bits=t["chan_bits"][k, b0:b1], tick=e.tick, episode=e.episode,
                               greedy=self.greedy, probs=t["probs"][k, b0:b1], actions=t["actions"][k, b0:b1],
                               attention=t["attention"][k, b0:b1] if "attention" in t else None, env_id0=e.env_id0)
        e.step(t["actions"][k, b0:b1],
               out=dict(obs=t["obs"][k + 1, b0:b1], adj_bits=t["adj_bits"][k + 1, b0:b1], chan_bits=t["chan_bits"][k + 1, b0:b1],
                        ave_deg=t["ave_deg"][k + 1, b0:b1], reward=t["reward"][k, b0:b1], done=t["done"][k, b0:b1],
                        counts=t["counts"][k, b0:b1], prey_alive_out=t["prey_alive_out"][k, b0:b1],
                        success_out=t["success"][k, b0:b1]))
        self.kernel_launches += 2

    def _carry(self):
        """slot K (state after the last step of a chunk) becomes slot 0 of the next chunk"""
        for k in ("obs", "adj_bits", "chan_bits", "ave_deg"):
            self.traj[k][0].copy_(self.traj[k][self.K])

    def reset(self):
        t = self.traj
        self.env.stats.zero_()
        self.env.reset(out=dict(obs=t["obs"][0], adj_bits=t["adj_bits"][0], chan_bits=t["chan_bits"][0], ave_deg=t["ave_deg"][0]))
        self.steps_done = 0

    def _chunk_eager(self):
        if len(self._ranges) == 1:
            for k in range(self.K):
                self._iteration(k, 0)
            return
        # one chain per group, each on its own stream (fork / join on events: capturable)
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=self.device) for _ in self._ranges]
        main = torch.cuda.current_stream(self.device)
        fork = torch.cuda.Event()
        fork.record(main)
        for g, st in enumerate(self._streams):
            st.wait_event(fork)
            with torch.cuda.stream(st):
                for k in range(self.K):
                    self._iteration(k, g)
                join = torch.cuda.Event()
                join.record(st)
            main.wait_event(join)

    def _capture(self):
        g = torch.cuda.CUDAGraph()
        launches = self.kernel_launches
        with torch.cuda.graph(g):
            self._chunk_eager()          # recorded, not executed: the env state does not advance here
        self.kernel_launches = launches
        self._graph = g

    def run_chunk(self):
        """K rollout iterations (K * B * n agent-steps).  The first chunk after construction runs eagerly (it also
        warms the library's launch-geometry caches and the weight blob); later chunks replay one CUDA graph."""
        if self.steps_done:
            self._carry()
        if self.use_graph and self._warm:
            if self._graph is None:
                self._capture()
            self._graph.replay()
            self.kernel_launches += 2 * self.K * len(self._ranges)
        else:
            self._chunk_eager()
            self._warm = True
        self.steps_done += self.K

    def run(self, steps: int):
        assert steps % self.K == 0, "steps must be a multiple of the ring size"
        for _ in range(steps // self.K):
            self.run_chunk()

    # ---- episode statistics ------------------------------------------------------------------------
    def local_stats(self) -> torch.Tensor:
        """float64 [9] sums over this GPU's envs of the finished-episode accumulators (STAT_KEYS)."""
        return self.env.stats[:, 7:16].sum(dim=0)

    def agent_steps(self) -> int:
        return self.steps_done * self.B * self.env.n
