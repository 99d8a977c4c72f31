"""Rollout + PPO update loop on the device: what LocalRunnerWrapper.train drives in the reference
(custom_implement/local_runner_wrapper.py:41-43 obtain_samples -> CentralizedMAPPO.train_once,
centralized_ma_ppo.py:175-262), with the `paths` list skipped: the rollout engine records ``max_path_length`` steps of
every env into its device trajectory ring, ``DevicePPO.batch_from_trajectory`` pads the finished episodes, ``train_once``
updates policy and critic.  One process per GPU; every rank rolls out its own slice of the global env ids and the flat
gradient buckets are all-reduced inside the update (NCCL), so all ranks hold identical weights after every step."""
import time

import numpy as np
import torch
import torch.distributed as dist

from . import distributed as D
from .ppo import CommBaseCritic, DevicePPO
from .rollout import RolloutEngine, make_policy
from .spaces import Box, Discrete, EnvSpec


class DeviceTrainer:
    def __init__(self, spec, n_envs, device="cuda", env_id0=0, seed=1, **ppo_args):
        self.spec, self.device = spec, torch.device(device)
        n, Dobs = spec.n_agents, spec.obs_dim
        spec.max_path_length = spec.max_steps
        torch.manual_seed(seed)                              # identical initial weights on every rank
        self.policy = make_policy(spec, device=self.device)
        self.critic = CommBaseCritic(EnvSpec(Box(np.zeros(n * Dobs), np.ones(n * Dobs)), Discrete(5)), n,
                                     n_gcn_layers=spec.n_layers, device=self.device)
        self.algo = DevicePPO(self.policy, self.critic, **ppo_args)
        # one chunk = one episode horizon: every env finishes at least one episode per round (time limit)
        self.engine = RolloutEngine(spec, self.policy, n_envs, device=self.device, env_id0=env_id0, ring=spec.max_steps,
                                    use_graph=False)
        self.epoch = 0

    def train_epoch(self):
        """one round: rollout of max_steps steps for every env, then the PPO update.  Returns timings and statistics."""
        dev = self.device
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        np.random.seed(1000 + self.epoch)                    # the same path permutation on every rank is not required
        ev[0].record()
        self.engine.reset()
        self.engine.run_chunk()
        ev[1].record()
        batch = self.algo.batch_from_trajectory(self.engine.traj)
        ev[2].record()
        out = self.algo.train_once(batch=batch)
        ev[3].record()
        torch.cuda.synchronize(dev)
        self.engine.env.check_errors()
        stats = D.gather_stats(self.engine.local_stats())
        self.epoch += 1
        out.update(rollout_ms=ev[0].elapsed_time(ev[1]), batch_ms=ev[1].elapsed_time(ev[2]), update_ms=ev[2].elapsed_time(ev[3]),
                   agent_steps=self.engine.K * self.engine.B * self.spec.n_agents,
                   episode_stats=D.summarize_stats(stats, self.spec.scenario, self.spec.n_agents))
        return out

    def weights_checksum(self):
        """float64 sum of every policy + critic parameter: equal on all ranks when the gradient all-reduce works"""
        return float(sum(p.detach().double().sum() for m in (self.policy, self.critic) for p in m.parameters()))
