"""PPO update against golden vectors recorded from the UNMODIFIED reference
(tests/golden/make_golden_ppo.py: CentralizedMAPPO.process_samples / _compute_loss / the optimisation loop of train_once,
CommBaseCritic / GaussianMLPBaseline, the reference's Adam) for the three runner families: Comm-DP + CommBaseCritic (pp, co),
Obs-DP + CommBaseCritic (pp_dec), CENT + GaussianMLPBaseline (co_cent)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
# pp_c3: BASELINE config 3's team (n = 32, IID drops); pp_n72 / pp_c5: teams larger than one 64-row tile (n = 72, and config
# 5's n = 200, where the reference's per-step losses swing from -0.15 to +0.14: the joint ratio is a product over 200 agents)
CASES = ["pp", "co", "pp_dec", "co_cent", "pp_c3", "pp_n72", "pp_c5"]


class PPOCase:
    def __init__(self, name):
        z = np.load(os.path.join(HERE, "golden", f"ppo_{name}.npz"))
        self.z = z
        self.meta = json.loads(str(z["meta"]))
        m = self.meta
        self.paths = [{k: z[f"path{i}::{k}"] for k in ("observations", "actions", "avail_actions", "rewards", "dist_adjs", "channels")}
                      for i in range(int(m["n_paths"]))]
        self.sd = {tag: {k.split("::", 1)[1]: z[k] for k in z.files if k.startswith(tag + "::")} for tag in ("pol0", "cri0", "pol1", "cri1")}

    def padded_rewards(self):
        P, T = len(self.paths), max(len(p["rewards"]) for p in self.paths)
        r = np.zeros((P, T), dtype=np.float64)
        for i, p in enumerate(self.paths):
            r[i, :len(p["rewards"])] = p["rewards"]
        return r


@pytest.mark.parametrize("name", CASES)
def test_oracle_advantages_match_reference(name):
    """the numpy restatement of discount_cumsum / compute_advantages / center_adv reproduces the reference's tensors"""
    c = PPOCase(name)
    m = c.meta
    returns, raw, adv = orc.ppo_advantages(c.padded_rewards(), c.z["baselines"], c.z["valids"], m["discount"], m["gae_lambda"])
    assert np.abs(returns - c.z["returns"]).max() <= 1e-5 * max(1.0, np.abs(c.z["returns"]).max())
    assert np.abs(raw - c.z["raw_adv"]).max() <= 1e-5 * max(1.0, np.abs(c.z["raw_adv"]).max())
    assert np.abs(adv - c.z["adv"]).max() <= 2e-5 * max(1.0, np.abs(c.z["adv"]).max())


def _build(c, device="cuda", fused="auto"):
    import torch
    from com_marl_b200.policy import CentralizedCategoricalMLPPolicy, CommCategoricalMLPPolicy, DecCategoricalMLPPolicy
    from com_marl_b200.ppo import CommBaseCritic, DevicePPO, GaussianMLPBaseline
    from com_marl_b200.spaces import Box, Discrete, EnvSpec
    m = c.meta
    n, D, kind = int(m["n"]), int(m["D"]), m.get("kind", "comm")
    spec = EnvSpec(Box(np.zeros(n * D), np.ones(n * D)), Discrete(5))
    pol = {"comm": CommCategoricalMLPPolicy, "dec": DecCategoricalMLPPolicy, "cent": CentralizedCategoricalMLPPolicy}[kind](spec, n, device=device)
    cri = GaussianMLPBaseline(spec, hidden_sizes=(64, 64, 64), device=device) if kind == "cent" else CommBaseCritic(spec, n, device=device)
    assert set(cri.state_dict().keys()) == set(c.sd["cri0"].keys()) and set(pol.state_dict().keys()) == set(c.sd["pol0"].keys())
    pol.load_state_dict({k: torch.as_tensor(v) for k, v in c.sd["pol0"].items()})
    cri.load_state_dict({k: torch.as_tensor(v) for k, v in c.sd["cri0"].items()})       # reference checkpoints load as is
    algo = DevicePPO(pol, cri, discount=m["discount"], gae_lambda=m["gae_lambda"], policy_ent_coeff=m["ent_coeff"],
                     clip_grad_norm=m["clip_grad_norm"], optimization_n_minibatches=int(m["n_minibatches"]),
                     optimization_mini_epochs=int(m["mini_epochs"]), policy_lr=m["lr"], adam_eps=m["adam_eps"], fused=fused)
    return pol, cri, algo


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_process_samples_and_loss_match_reference(name):
    """critic baselines, returns, advantages (cm_ppo_advantages), entropy, log-likelihood and the PPO loss"""
    import torch
    c = PPOCase(name)
    pol, cri, algo = _build(c)
    b = algo.process_samples(c.paths)
    z = c.z
    close = lambda a, ref, tol=1e-5: np.abs(a.detach().cpu().numpy() - ref).max() <= tol * max(1.0, np.abs(ref).max())  # noqa: E731
    assert np.array_equal(b["valids"].cpu().numpy(), z["valids"])
    assert close(b["baselines"], z["baselines"])
    assert close(b["returns"], z["returns"])
    assert close(b["raw_adv"], z["raw_adv"], 2e-5)
    assert close(b["adv"], z["adv"], 5e-5)
    # the kernel against the oracle on the kernel's own baselines
    ret_o, raw_o, adv_o = orc.ppo_advantages(c.padded_rewards(), b["baselines"].cpu().numpy(), z["valids"], c.meta["discount"], c.meta["gae_lambda"])
    assert close(b["returns"], ret_o) and close(b["raw_adv"], raw_o) and close(b["adv"], adv_o, 2e-5)
    with torch.no_grad():
        d = algo._dist(b, None)
        assert close(d.entropy().mean(-1), z["entropy"])
        assert close(d.log_prob(b["actions"]).sum(-1), z["loglik"])
        assert abs(float(algo.compute_loss(b)) - float(z["loss_before"])) <= 2e-5
        bl0 = cri.compute_loss(b["obs"], b["returns"], b["dist_adjs"], b["channels"]) if algo._critic_comm \
            else cri.compute_loss(b["obs"], b["returns"])
        assert abs(float(bl0) - float(z["baseline_loss0"])) <= 1e-5 * max(1.0, abs(float(z["baseline_loss0"])))


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_train_once_matches_reference(name, fused):
    """the whole optimisation loop (2 mini-epochs x 3 minibatches, clip_grad_norm, both Adam steps through cm_adam_step):
    per-step losses, gradient norms and the final policy / critic weights equal the reference's — with the network forward /
    backward on torch autograd (fused=False) and on the hand-written kernels (fused=True: cm_ppo_net, the Comm-DP family)"""
    c = PPOCase(name)
    if fused and c.meta.get("kind", "comm") == "cent":
        pytest.skip("the hand-written update kernels cover the Comm-DP and Obs-DP families (CommBaseCritic); CENT keeps autograd")
    pol, cri, algo = _build(c, fused=fused)
    assert (algo._fused is not None) == fused
    z = c.z
    out = algo.train_once(paths=c.paths, shuffled_ids=z["shuffled_ids"])
    assert abs(out["loss_before"] - float(z["loss_before"])) <= 2e-5
    assert np.abs(np.array(out["losses"]) - z["losses"]).max() <= 1e-4
    assert np.abs(np.array(out["baseline_losses"]) - z["baseline_losses"]).max() <= 1e-4 * max(1.0, np.abs(z["baseline_losses"]).max())
    assert np.abs(np.array(out["grad_norms"]) - z["grad_norms"]).max() <= 1e-3 * max(1.0, np.abs(z["grad_norms"]).max())
    assert abs(out["loss_after"] - float(z["loss_after"])) <= 1e-4
    assert abs(out["kl"] - float(z["kl"])) <= 1e-5
    for tag, mod in (("pol1", pol), ("cri1", cri)):
        for k, v in mod.state_dict().items():
            ref, ref0 = c.sd[tag][k], c.sd[tag[:3] + "0"][k]
            step = np.abs(ref - ref0).max()                       # how far the reference moved this tensor
            assert np.abs(v.cpu().numpy() - ref).max() <= 0.02 * step + 1e-6, (tag, k)
    # the rollout kernels see the updated weights (the flat-bucket Adam bumps the parameter versions)
    import torch
    n, D = int(c.meta["n"]), int(c.meta["D"])
    obs = torch.rand((8, n, D), device="cuda")
    logits = torch.empty((8, n, 5), device="cuda")
    pol.act_device(obs, logits=logits, greedy=True)
    w = {k: v.cpu().numpy() for k, v in pol.state_dict().items()}
    kind = c.meta.get("kind", "comm")
    if kind == "comm":
        ref_logits = orc.policy_forward(w, obs.cpu().numpy(), None, np.ones((8, n, n), np.uint8), np.ones((8, 2, n, n), np.uint8))[0]
    else:
        ref_logits = (orc.policy_forward_dec if kind == "dec" else orc.policy_forward_cent)(w, obs.cpu().numpy(), None)[0]
    assert np.abs(logits.cpu().numpy() - ref_logits).max() <= 1e-5 * max(1.0, np.abs(ref_logits).max())


@pytest.mark.gpu
@pytest.mark.parametrize("scen,m,den,T", [("pp", 10, 0.04, 9), ("co", 10, 0.03, 13)])
def test_device_trajectory_batch_equals_paths(scen, m, den, T):
    """DevicePPO.batch_from_trajectory (no host `paths` list) builds exactly the padded batch process_samples builds from
    the per-episode dicts cut out of the same trajectory; one update round runs on it."""
    import sys
    import torch
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import ref_harness
    from com_marl_b200.ppo import CommBaseCritic, DevicePPO
    from com_marl_b200.rollout import RolloutEngine, make_policy
    from com_marl_b200.scenario import ScenarioSpec
    from com_marl_b200.spaces import Box, Discrete, EnvSpec
    params = ref_harness.scenario_params(scen, m, 1, den, cap=2, loss=0.2, max_env_steps=T)
    spec = ScenarioSpec.from_params(scen, params, seed=3)
    spec.max_path_length = T
    pol = make_policy(spec)
    n, D, L, B, K = spec.n_agents, spec.obs_dim, spec.n_layers, 37, 32
    cri = CommBaseCritic(EnvSpec(Box(np.zeros(n * D), np.ones(n * D)), Discrete(5)), n)
    algo = DevicePPO(pol, cri, optimization_mini_epochs=2, fused=False)
    eng = RolloutEngine(spec, pol, B, ring=K, use_graph=False)
    eng.reset()
    eng.run_chunk()
    b = algo.batch_from_trajectory(eng.traj)
    t = {k: v.cpu().numpy() for k, v in eng.traj.items()}
    unpack = lambda bits: ((bits.view(np.uint32)[..., np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(np.float32)  # noqa: E731
    paths = []
    for e in range(B):
        start = 0
        for k in np.nonzero(t["done"][:, e])[0]:
            sl = slice(start, int(k) + 1)
            T_ = int(k) + 1 - start
            paths.append(dict(observations=t["obs"][sl, e].reshape(T_, -1), actions=t["actions"][sl, e].astype(np.int64),
                              avail_actions=np.ones((T_, n * 5)), rewards=t["reward"][sl, e],
                              dist_adjs=unpack(t["adj_bits"][sl, e]).reshape(T_, n * n),
                              channels=unpack(t["chan_bits"][sl, e]).reshape(T_, L * n, n)))
            start = int(k) + 1
    assert len(paths) >= B * (K // T) - B
    ref = algo.process_samples(paths)
    for k in ("obs", "actions", "rewards", "valids", "dist_adjs", "channels", "avail", "mask"):
        assert torch.equal(b[k], ref[k]), k
    for k in ("baselines", "returns", "adv"):
        assert (b[k] - ref[k]).abs().max().item() <= 1e-5 * max(1.0, ref[k].abs().max().item()), k
    # the same trajectory through the hand-written update kernels (bit-row masks straight from the ring, no dense masks)
    import copy
    pol2, cri2 = copy.deepcopy(pol), copy.deepcopy(cri)
    algo2 = DevicePPO(pol2, cri2, optimization_mini_epochs=2, fused=True)
    b2 = algo2.batch_from_trajectory(eng.traj)
    assert "dist_adjs" not in b2 and torch.equal(b2["obs"], b["obs"]) and torch.equal(b2["valids"], b["valids"])
    for k in ("baselines", "returns", "adv"):
        assert (b2[k] - b[k]).abs().max().item() <= 2e-5 * max(1.0, b[k].abs().max().item()), k
    ids = np.random.RandomState(1).permutation(b["rewards"].shape[0])
    out = algo.train_once(batch=b, shuffled_ids=ids)
    assert np.isfinite(out["loss_after"]) and out["loss_after"] < out["loss_before"] and out["kl"] >= 0
    out2 = algo2.train_once(batch=b2, shuffled_ids=ids)
    for k in ("loss_before", "loss_after", "kl"):
        assert abs(out[k] - out2[k]) <= 1e-4 * max(1.0, abs(out[k])), k


@pytest.mark.gpu
def test_adam_step_kernel_matches_torch():
    import torch
    from com_marl_b200 import _native as N
    g = torch.Generator(device="cuda").manual_seed(0)
    n = 100003
    p = torch.randn(n, device="cuda", generator=g)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=3e-4, eps=1e-5)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        grad = torch.randn(n, device="cuda", generator=g)
        ref.grad = grad.clone() * 0.5
        opt.step()
        N.check("cm_adam_step", N.lib().cm_adam_step(N.ptr(p), N.ptr(grad), N.ptr(m), N.ptr(v), n, 3e-4, 0.9, 0.999, 1e-5, step, 0.5,
                                                     N.stream_ptr()))
        assert (p - ref.detach()).abs().max().item() <= 1e-6


@pytest.mark.gpu
def test_trainer_round_and_tabular_columns():
    """DeviceTrainer.train_epoch: rollout of one horizon + update; the tabular row carries the reference's progress.csv
    columns (centralized_ma_ppo.py:345-372) and its return columns agree with the kernels' episode accounting."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import ref_harness
    from com_marl_b200.scenario import ScenarioSpec
    from com_marl_b200.train import DeviceTrainer
    params = ref_harness.scenario_params("co", 10, 1, 0.03, cap=2, loss=0.0, max_env_steps=20)
    spec = ScenarioSpec.from_params("co", params, seed=2)
    tr = DeviceTrainer(spec, 64, optimization_mini_epochs=1)
    out = tr.train_epoch()
    for kind in ("dec", "cent"):            # the other two runner families run the same round
        o2 = DeviceTrainer(spec, 64, optimization_mini_epochs=1, kind=kind).train_epoch()
        assert np.isfinite(o2["loss_after"]) and o2["loss_after"] < o2["loss_before"] and o2["episode_stats"]["NumEpisodes"] == 64
    row = out["tabular"]
    for k in ("Iteration", "NumTrajs", "AverageDiscountedReturn", "AverageReturn", "SuccessRate", "AverageCaptureCount",
              "AverageStepCount", "AverageMovingCount", "AveragePenaltyCount", "AverageVariable", "AverageVar2", "StdReturn",
              "MaxReturn", "MinReturn", "LossBefore", "LossAfter", "dLoss", "KL", "Entropy", "GradNorm", "EpochTime", "AveDegree",
              "Diameter", "AveTroughput"):
        assert k in row and np.isfinite(row[k]), k
    # every env runs exactly one 20-step episode per round (Coverage cannot finish 86 cells in 20 steps): the batch's
    # mean return is the kernels' AverageReturn
    assert row["NumTrajs"] == 64 * spec.n_agents and out["episode_stats"]["NumEpisodes"] == 64
    assert abs(row["AverageReturn"] - out["episode_stats"]["AverageReturn"]) <= 1e-9 * max(1.0, abs(row["AverageReturn"]))
    assert row["MinReturn"] <= row["AverageReturn"] <= row["MaxReturn"]


@pytest.mark.gpu
def test_data_parallel_update_equals_union_update_on_gpu_kernels():
    """Two ranks (two processes sharing cuda:0, gloo) update on a 4 | 5 split of nine paths with the real kernels
    (cm_adam_step): the result equals the single-process update on the union batch, both ranks hold identical weights,
    unequal local slice counts do not deadlock (tests/ppo_dp_util.py)."""
    import ppo_dp_util
    single, ranks = ppo_dp_util.run("cuda")
    ppo_dp_util.check(single, ranks)
