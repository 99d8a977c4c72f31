"""Golden vectors for the PPO update, recorded from the UNMODIFIED reference:
CentralizedMAPPO.process_samples / _compute_loss / _compute_objective (centralized_ma_ppo.py:390-438, 540-589,
612-659), compute_advantages (garage/torch/algos/_utils.py:56-113), CommBaseCritic (comm_base_critic.py:11-120), the
reference's Adam (my_optimizer/adam.py:57-120), driven through the optimisation loop of train_once
(centralized_ma_ppo.py:207-262: shuffled path ids, minibatches, baseline loss, clip_grad_norm_, both optimizers).
train_once itself cannot run here (it queries torch.cuda memory counters and a LocalRunner), so this script calls the
reference's own methods in train_once's order.

Run by hand in the build container (needs /root/reference):  python tests/golden/make_golden_ppo.py
"""
import json
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402


def main():
    ns = H.load_reference()
    for name in ("tensorflow", "cma"):
        sys.modules.setdefault(name, types.ModuleType(name))
    from com_marl.torch.algos.centralized_ma_ppo import CentralizedMAPPO
    from com_marl.torch.baselines.comm_base_critic import CommBaseCritic
    from com_marl.sampler import CentralizedMAOnPolicyVectorizedSampler
    from garage.torch.algos import compute_advantages

    from com_marl.torch.baselines.gaussian_mlp_baseline import GaussianMLPBaseline
    from com_marl.torch.policies import CentralizedCategoricalMLPPolicy, DecCategoricalMLPPolicy
    only = set(sys.argv[1:])
    # kind: comm = runner_*_comm.py (Comm-DP + CommBaseCritic), dec = runner_*_obsDP.py:50-74 (Obs-DP + CommBaseCritic),
    #       cent = runner_*_cent.py:48-63 (CENT + GaussianMLPBaseline(64, 64, 64))
    for case, (scenario, m, sen, den, cap, loss, T, n_paths, kind) in dict(
            pp=("pp", 10, 1, 0.04, 2, 0.3, 14, 7, "comm"), co=("co", 10, 1, 0.03, 2, 0.0, 11, 6, "comm"),
            pp_dec=("pp", 10, 1, 0.04, 2, 0.3, 12, 6, "dec"), co_cent=("co", 10, 1, 0.03, 2, 0.0, 13, 7, "cent"),
            # BASELINE config 3's shape (n = 32, IID drops), a large team (n = 72 > one 64-row tile) and config 5's team (n = 200)
            pp_c3=("pp", 20, 2, 0.08, 4, 0.2, 10, 4, "comm"), pp_n72=("pp", 30, 2, 0.08, 4, 0.0, 6, 3, "comm"),
            pp_c5=("pp", 50, 2, 0.08, 4, 0.0, 4, 3, "comm")).items():
        if only and case not in only:
            continue
        comm = kind == "comm"
        seed = 5
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        params = H.scenario_params(scenario, m, sen, den, cap=cap, loss=loss, max_env_steps=T)
        if scenario == "pp":
            env = ns.PredatorPreyWrapper(centralized=True, other_agent_visible=True, params=params)
        else:
            env = ns.CoverageWrapper(centralized=True, other_agent_visible=True, max_steps=T, params=params)
        genv = ns.GarageEnv(env)
        n = env.n_agents
        torch.manual_seed(1)
        if kind == "comm":
            policy = ns.CommCategoricalMLPPolicy(genv.spec, n_agents=n)
        elif kind == "dec":
            policy = DecCategoricalMLPPolicy(genv.spec, n_agents=n, hidden_sizes=(128, 64, 32))
        else:
            policy = CentralizedCategoricalMLPPolicy(genv.spec, n_agents=n, hidden_nonlinearity=torch.tanh,
                                                     hidden_sizes=[128, 64, 32], name="centralized")
        critic = GaussianMLPBaseline(env_spec=genv.spec, hidden_sizes=(64, 64, 64)) if kind == "cent" \
            else CommBaseCritic(genv.spec, n_agents=n)
        critic_comm = kind != "cent"
        bl_loss = (lambda o, r, a, c: critic.compute_loss(o, r, a, c)) if critic_comm else (lambda o, r, a, c: critic.compute_loss(o, r))
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():       # non-zero biases (xavier init zeroes them)
            for mod in (policy, critic):
                for k, v in mod.state_dict().items():
                    if k.endswith("bias"):
                        v.copy_((torch.rand(v.shape, generator=g) - 0.5) * 0.2)
        algo = CentralizedMAPPO(env_spec=genv.spec, policy=policy, baseline=critic, max_path_length=T, discount=0.99,
                                center_adv=True, positive_adv=False, gae_lambda=0.97, policy_ent_coeff=0.1,
                                entropy_method="regularized", stop_entropy_gradient=False, clip_grad_norm=7,
                                optimization_n_minibatches=3, optimization_mini_epochs=2, device="cpu")
        sampler = CentralizedMAOnPolicyVectorizedSampler(algo, genv, n_envs=1)
        sampler.start_worker()
        paths = sampler.obtain_samples(0, batch_size=n_paths * T * n)
        # ragged lengths: cut some paths short (every array of a path has the time axis first)
        rng = np.random.default_rng(seed)
        cut = []
        for i, p in enumerate(paths[:n_paths]):
            keep = len(p["rewards"]) if i % 2 == 0 else int(rng.integers(3, len(p["rewards"])))
            q = {k: v[:keep] for k, v in p.items() if isinstance(v, np.ndarray) and v.shape[:1] == p["rewards"].shape[:1]}
            cut.append(q)
        paths = cut
        rec = dict()
        for i, p in enumerate(paths):
            for k in ("observations", "actions", "avail_actions", "rewards", "dist_adjs", "channels"):
                rec[f"path{i}::{k}"] = np.asarray(p[k])
        sd0 = {f"pol0::{k}": v.detach().numpy().copy() for k, v in policy.state_dict().items()}
        sd0.update({f"cri0::{k}": v.detach().numpy().copy() for k, v in critic.state_dict().items()})

        obs, avail, actions, rewards, valids, baselines, returns, dist_adjs, channels = algo.process_samples(0, [dict(p) for p in paths])
        captured = {}
        orig = algo._compute_objective

        def spy(advantages, *a, **k):
            captured["adv"] = advantages.detach().numpy().copy()
            return orig(advantages, *a, **k)

        algo._compute_objective = spy
        with torch.no_grad():
            loss_before = algo._compute_loss(0, obs, avail, actions, rewards, valids, baselines, dist_adjs, channels)
            ent = algo._compute_policy_entropy(obs, avail, dist_adjs, channels)
            if comm:
                ll = policy.log_likelihood(observations=obs, avail_actions=avail, dist_adj=dist_adjs, channels=channels, actions=actions)
            else:
                ll = policy.log_likelihood(obs, avail, actions)
            bl_loss0 = bl_loss(obs, returns, dist_adjs, channels)
        algo._compute_objective = orig
        raw_adv = compute_advantages(0.99, 0.97, algo.temp_max_path_length, baselines, rewards, "cpu")
        # ---- the optimisation loop of train_once (centralized_ma_ppo.py:207-262) ----
        algo._old_policy.load_state_dict(policy.state_dict())
        np.random.seed(11)
        step_size = int(np.ceil(len(rewards) / algo._optimization_n_minibatches))
        shuffled_ids = np.random.permutation(len(rewards))
        losses, bl_losses, gnorms = [], [], []
        for mini_epoch in range(algo._optimization_mini_epochs):
            for start in range(0, len(rewards), step_size):
                ids = shuffled_ids[start:min(start + step_size, len(rewards))]
                loss = algo._compute_loss(0, obs[ids], avail[ids], actions[ids], rewards[ids], valids[ids], baselines[ids],
                                          dist_adjs[ids], channels[ids])
                baseline_loss = bl_loss(obs[ids], returns[ids], dist_adjs[ids], channels[ids])
                algo._baseline_optimizer.zero_grad()
                baseline_loss.backward()
                algo._optimizer.zero_grad()
                loss.backward()
                torch.nn.utils.clip_grad_norm_(policy.parameters(), algo._clip_grad_norm)
                gnorms.append(float(policy.grad_norm()))
                algo._optimize(0, None, None, None, None, None, None, None)
                losses.append(float(loss)); bl_losses.append(float(baseline_loss))
        with torch.no_grad():
            loss_after = algo._compute_loss(0, obs, avail, actions, rewards, valids, baselines, dist_adjs, channels)
            kl = algo._compute_kl_constraint(obs, avail, dist_adjs, channels, actions)
        sd1 = {f"pol1::{k}": v.detach().numpy().copy() for k, v in policy.state_dict().items()}
        sd1.update({f"cri1::{k}": v.detach().numpy().copy() for k, v in critic.state_dict().items()})
        meta = dict(case=case, kind=kind, scenario=scenario, n=n, D=int(obs.shape[-1] // n), L=2, n_paths=len(paths), T=T,
                    Tmax=int(algo.temp_max_path_length), discount=0.99, gae_lambda=0.97, ent_coeff=0.1, clip=0.1, lr=3e-4,
                    adam_eps=1e-5, clip_grad_norm=7, n_minibatches=3, mini_epochs=2, map=m, sen=sen, den=den, cap=cap, loss=loss)
        np.savez_compressed(os.path.join(HERE, f"ppo_{case}.npz"), meta=np.array(json.dumps(meta, default=lambda o: float(o))),
                            valids=valids.numpy(), baselines=baselines.numpy(), returns=returns.numpy(), raw_adv=raw_adv.numpy(),
                            adv=captured["adv"], entropy=ent.numpy(), loglik=ll.numpy(), loss_before=float(loss_before),
                            baseline_loss0=float(bl_loss0), shuffled_ids=shuffled_ids, losses=np.array(losses),
                            baseline_losses=np.array(bl_losses), grad_norms=np.array(gnorms), loss_after=float(loss_after),
                            kl=float(kl), **rec, **sd0, **sd1)
        print(case, "paths", [len(p["rewards"]) for p in paths], "loss_before", float(loss_before), "losses", losses[:3],
              "gnorm", gnorms[:2], "loss_after", float(loss_after), "kl", float(kl))


if __name__ == "__main__":
    main()
