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
from .ppo import CommBaseCritic, DevicePPO, GaussianMLPBaseline
from .rollout import RolloutEngine, make_policy
from .spaces import Box, Discrete, EnvSpec


class DeviceTrainer:
    def __init__(self, spec, n_envs, device="cuda", env_id0=0, seed=1, kind="comm", groups=1, **ppo_args):
        """kind: 'comm' / 'dec' (runner_*_comm.py / runner_*_obsDP.py:61: Comm-DP / Obs-DP policy + CommBaseCritic), 'cent'
        (runner_*_cent.py:60-62: CENT policy + GaussianMLPBaseline(hidden_sizes=(64, 64, 64)))"""
        self.spec, self.device, self.seed = spec, torch.device(device), int(seed)
        n, Dobs = spec.n_agents, spec.obs_dim
        spec.max_path_length = spec.max_steps
        torch.manual_seed(seed)                              # identical initial weights on every rank
        self.policy = make_policy(spec, device=self.device, kind=kind)
        env_spec = EnvSpec(Box(np.zeros(n * Dobs), np.ones(n * Dobs)), Discrete(5))
        if kind != "cent":
            self.critic = CommBaseCritic(env_spec, n, n_gcn_layers=spec.n_layers, device=self.device)
        else:
            self.critic = GaussianMLPBaseline(env_spec, hidden_sizes=(64, 64, 64), device=self.device)
        self.algo = DevicePPO(self.policy, self.critic, **ppo_args)
        # one chunk = one episode horizon: every env finishes at least one episode per round (time limit)
        # the horizon is replayed from ONE captured CUDA graph: the kernel weight blobs are persistent buffers refreshed in
        # place after every update (RolloutEngine.run_chunk -> policy.refresh_weights), so the graph stays valid
        self.engine = RolloutEngine(spec, self.policy, n_envs, device=self.device, env_id0=env_id0, ring=spec.max_steps,
                                    use_graph=True, groups=groups)
        self.epoch = 0

    def train_epoch(self):
        """one round: rollout of max_steps steps for every env, then the PPO update.  Returns timings and statistics."""
        dev = self.device
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        self.engine.reset()
        self.engine.run_chunk()
        ev[1].record()
        batch = self.algo.batch_from_trajectory(self.engine.traj)
        ev[2].record()
        # the path permutation comes from a private generator keyed by (seed, epoch): the process-global numpy stream is the
        # user's (the same permutation on every rank is not required)
        perm = np.random.default_rng([self.seed, self.epoch]).permutation(int(batch["rewards"].shape[0]))
        out = self.algo.train_once(batch=batch, shuffled_ids=perm)
        ev[3].record()
        torch.cuda.synchronize(dev)
        self.engine.env.check_errors()
        stats = D.gather_stats(self.engine.local_stats())
        self.epoch += 1
        out.update(rollout_ms=ev[0].elapsed_time(ev[1]), batch_ms=ev[1].elapsed_time(ev[2]), update_ms=ev[2].elapsed_time(ev[3]),
                   agent_steps=self.engine.K * self.engine.B * self.spec.n_agents,
                   episode_stats=D.summarize_stats(stats, self.spec.scenario, self.spec.n_agents))
        out["tabular"] = self.tabular(batch, out)
        return out

    def tabular(self, batch, out):
        """The per-epoch columns CentralizedMAPPO.train_once records in dowel's tabular (centralized_ma_ppo.py:345-372 ->
        progress.csv), from the device batch of THIS rank and the whole-job episode statistics.  Episode-count columns
        (SuccessRate, Average*Count) come from the kernels' episode accounting over every episode that finished during the
        round; return columns from the padded batch like the reference (paths that finished inside the window).
        KLBefore (KL against the policy of the previous round) and GPUMemoryMax are not reproduced."""
        rewards, mask = batch["rewards"], batch["mask"]
        und = (rewards * mask).sum(1)
        es = out["episode_stats"]
        row = dict(Iteration=self.epoch - 1, NumTrajs=int(rewards.shape[0]) * self.spec.n_agents,
                   AverageDiscountedReturn=float(batch["returns"][:, 0].double().mean()), AverageReturn=float(und.mean()),
                   SuccessRate=es["SuccessRate"], AverageCaptureCount=es["AverageCaptureCount"],
                   AverageStepCount=es["AverageStepCount"], AverageMovingCount=es["AverageMovingCount"],
                   AveragePenaltyCount=es["AveragePenaltyCount"], AverageVariable=es["AverageVariable"], AverageVar2=es["AverageVar2"],
                   StdReturn=float(und.std(unbiased=False)), MaxReturn=float(und.max()), MinReturn=float(und.min()),
                   LossBefore=out["loss_before"], LossAfter=out["loss_after"], dLoss=out["loss_before"] - out["loss_after"],
                   KL=out["kl"], Entropy=out["entropy"], GradNorm=float(np.mean(out["grad_norms"])) if out["grad_norms"] else 0.0,
                   EpochTime=out["update_ms"] * 1e-3)
        # AveDegree: mean over paths of the per-path mean degree (:316-318); Diameter / AveTroughput are constants of the scenario
        t = self.engine.traj
        row["AveDegree"] = float(t["ave_deg"][:self.engine.K].double().mean()) if self.spec.rcom != 0 else float(self.spec.n_agents)
        row["Diameter"] = float(self.spec.n_agents if self.spec.rcom == 0 else 0)
        row["AveTroughput"] = float(self.spec.ave_trput)
        try:
            from dowel import tabular as dt
            for k, v in row.items():
                dt.record(k, v)
        except Exception:                      # dowel is optional here
            pass
        return row

    def weights_checksum(self):
        """float64 sum of every policy + critic parameter: equal on all ranks when the gradient all-reduce works"""
        return float(sum(p.detach().double().sum() for m in (self.policy, self.critic) for p in m.parameters()))
