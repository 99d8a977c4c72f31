"""GPU tests of the fused device rollout (policy forward + sampling -> env step -> comm update), its CUDA-graph
replay, the in-kernel episode accounting and the garage `paths` contract of DeviceRolloutSampler."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import oracle as orc

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import ref_harness  # noqa: E402  (only its params helper is used; it never touches /root/reference here)

pytestmark = pytest.mark.gpu

CASES = [("pp", 20, 2, 0.08, 4, 0.2, 256, {"max_env_steps": 30}), ("co", 10, 1, 0.03, 2, 0.0, 512, {"max_env_steps": 40}),
         ("pp", 30, 2, 0.08, 4, 0.0, 24, {"max_env_steps": 12})]


def _mk(scen, m, sen, den, cap, loss, over, seed=9):
    from com_marl_b200.scenario import ScenarioSpec
    params = ref_harness.scenario_params(scen, m, sen, den, cap=cap, loss=loss, **over)
    return params, ScenarioSpec.from_params(scen, params, seed=seed)


@pytest.mark.parametrize("cfg", CASES, ids=["pp20", "co10", "pp30_large_team"])
def test_fused_rollout_matches_oracle(cfg):
    """Every iteration of the device loop is replayed on the oracle with the kernel's own sampled actions:
    observations, rewards, dones and masks stay bit-identical across auto-resets; the sampled actions equal the
    oracle's inverse-CDF draw on the kernel's probabilities (Philox action stream)."""
    from com_marl_b200.rollout import RolloutEngine, make_policy
    scen, m, sen, den, cap, loss, B, over = cfg
    params, spec = _mk(scen, m, sen, den, cap, loss, over)
    pol = make_policy(spec)
    eng = RolloutEngine(spec, pol, B, ring=8, use_graph=True)
    oenv = orc.OracleVecEnv(orc.spec_from_params(scen, params, seed=9), B)
    eng.reset()
    oenv.reset()
    n = spec.n_agents
    ret = np.zeros(B); fin_ret = 0.0; fin_eps = 0; fin_len = 0; length = np.zeros(B)
    for chunk in range(10):
        eng.run_chunk()      # chunk 0 eager, then one captured graph replayed
        t = {k: v.cpu().numpy() for k, v in eng.traj.items()}
        for k in range(eng.K):
            assert np.array_equal(t["obs"][k], oenv.obs), f"obs differ at chunk {chunk} step {k}"
            want = oenv.sample_actions(t["probs"][k])
            assert np.array_equal(t["actions"][k], want), "sampled actions differ from the stream specification"
            oenv.step(t["actions"][k])
            assert np.array_equal(t["reward"][k], oenv.reward)
            assert np.array_equal(t["done"][k], oenv.done)
            assert np.array_equal(t["success"][k], oenv.success)
            ret += oenv.reward; length += 1
            d = oenv.done.astype(bool)
            fin_ret += ret[d].sum(); fin_len += length[d].sum(); fin_eps += int(d.sum())
            ret[d] = 0; length[d] = 0
        assert np.array_equal(t["obs"][eng.K], oenv.obs)
    eng.env.check_errors()
    st = eng.local_stats().cpu().numpy()
    assert fin_eps > 0 and st[0] == fin_eps and st[2] == fin_len
    assert st[1] == pytest.approx(fin_ret, rel=1e-12)
    assert eng.agent_steps() == 10 * eng.K * B * n
    assert eng.kernel_launches == 2 * 10 * eng.K


def test_graph_replay_equals_eager():
    from com_marl_b200.rollout import RolloutEngine, make_policy
    params, spec = _mk("pp", 10, 1, 0.08, 2, 0.3, {"max_env_steps": 20})
    pol = make_policy(spec)
    a = RolloutEngine(spec, pol, 300, ring=5, use_graph=True)
    b = RolloutEngine(spec, pol, 300, ring=5, use_graph=False)
    for e in (a, b):
        e.reset()
        e.run(40)
    for k in ("obs", "actions", "reward", "done", "chan_bits"):
        assert torch.equal(a.traj[k], b.traj[k]), k
    assert torch.equal(a.env.stats, b.env.stats)


@pytest.mark.parametrize("scen,groups", [("pp", 3), ("co", 4), ("pp", 7), ("pp_large", 2), ("pp_large", 3)])
def test_env_groups_equal_single_chain(scen, groups):
    """G independent env groups on G streams (forked / joined inside the captured graph) give exactly the trajectory of
    the single chain: the random streams are keyed by global env ids, the groups only remove device-wide barriers.
    pp_large: a team of 72 agents — the three-launch policy pipeline, whose groups hand their rows over through scratch
    buffers of their own (ws_slot)."""
    from com_marl_b200.rollout import RolloutEngine, make_policy
    B = 301
    if scen == "pp":
        params, spec = _mk("pp", 10, 1, 0.08, 2, 0.3, {"max_env_steps": 20})
    elif scen == "pp_large":
        params, spec = _mk("pp", 30, 2, 0.08, 4, 0.2, {"max_env_steps": 20})
        assert spec.n_agents > 64
        B = 37
    else:
        params, spec = _mk("co", 10, 1, 0.03, 2, 0.1, {"max_env_steps": 25})
    pol = make_policy(spec)
    a = RolloutEngine(spec, pol, B, ring=5, use_graph=True, groups=1)
    b = RolloutEngine(spec, pol, B, ring=5, use_graph=True, groups=groups)
    assert len(b._ranges) == groups
    for e in (a, b):
        e.reset()
        e.run(40)
    for k in a.traj:
        assert torch.equal(a.traj[k], b.traj[k]), k
    assert torch.equal(a.env.stats, b.env.stats)
    assert torch.equal(a.env.agent_pos, b.env.agent_pos) and torch.equal(a.env.tick, b.env.tick)
    assert b.kernel_launches == groups * a.kernel_launches
    b.env.check_errors()


def test_sampler_paths_contract():
    """obtain_samples returns the reference's path dicts (keys / shapes of SURVEY.md §8a a21)."""
    from types import SimpleNamespace
    from com_marl_b200.envs import PredatorPreyWrapper
    from com_marl_b200.rollout import make_policy
    from com_marl_b200.sampler import DeviceRolloutSampler
    params = ref_harness.scenario_params("pp", 10, 1, 0.08, cap=2, loss=0.2, max_env_steps=20)
    env = PredatorPreyWrapper(centralized=True, other_agent_visible=True, params=params)
    spec = env.spec_b200
    n, D, p, L = spec.n_agents, spec.obs_dim, spec.n_preys, spec.n_layers
    algo = SimpleNamespace(policy=make_policy(spec), max_path_length=20)
    sampler = DeviceRolloutSampler(algo, env, n_envs=16, chunk=10)
    sampler.start_worker()
    paths = sampler.obtain_samples(0, batch_size=16 * 20 * n)
    assert sum(len(pth["rewards"]) for pth in paths) * n >= 16 * 20 * n
    for pth in paths:
        T = len(pth["rewards"])
        assert 1 <= T <= 20
        assert pth["observations"].shape == (T, n * D) and pth["actions"].shape == (T, n) and pth["actions"].dtype == np.int64
        assert pth["avail_actions"].shape == (T, n * 5) and pth["dones"].shape == (T,) and pth["dones"][-1]
        assert not pth["dones"][:-1].any()
        assert pth["dist_adjs"].shape == (T, n * n) and pth["channels"].shape == (T, L * n, n)
        assert pth["attentions"].shape == (T, n, n) and pth["agent_infos"]["action_probs"].shape == (T, n, 5)
        assert pth["env_infos"]["prey_alive"].shape == (T, p) and pth["success"].shape == (16,)
        assert pth["rewards_details"].shape == (T,) and set(pth["rewards_details"][0]) >= {"reward", "capture_cnt", "move_cnt"}
        assert np.allclose(pth["agent_infos"]["action_probs"].sum(-1), 1.0, atol=1e-5)
        assert np.allclose(pth["attentions"].sum(-1), 1.0, atol=1e-5)
        # the time feature of the first observation is t/T = 0: the path starts at a reset observation
        assert pth["observations"][0].reshape(n, D)[0, -1] == 0.0
        # a path that ended early is a success (all preys captured), one that hit the limit is not
        assert bool(pth["success"][0]) == (not pth["env_infos"]["prey_alive"][-1].any())
    short = sampler.obtain_samples(1, batch_size=50 * n, whole_paths=False)
    assert sum(len(pth["rewards"]) for pth in short) * n <= 50 * n + n * 20
    sampler.shutdown_worker()


def test_host_buffer_api_equals_device_path():
    """BatchedEnv.step_host / policy.get_actions_host (the e2e path: pinned host buffers, H2D + D2H every step)
    produce the same trajectory as the device-resident calls."""
    from com_marl_b200.envs import BatchedEnv
    from com_marl_b200.rollout import make_policy
    params, spec = _mk("pp", 10, 1, 0.08, 2, 0.3, {"max_env_steps": 15})
    pol = make_policy(spec)
    B, n = 200, spec.n_agents
    host, devc = BatchedEnv(spec, B), BatchedEnv(spec, B)
    out = host.reset_host()
    devc.reset()
    probs = torch.empty((B, n, 5), device="cuda")
    actions = torch.empty((B, n), dtype=torch.int8, device="cuda")
    tick = torch.zeros((B,), dtype=torch.int32, device="cuda")
    h2d, d2h = host.host_step_bytes()
    assert h2d == B * n and d2h > B * n * spec.obs_dim * 4
    for s in range(40):
        acts_np, probs_np = pol.get_actions_host(out["obs"], out["adj_bits"], out["chan_bits"])
        tick.fill_(pol._host_calls - 1)
        pol.act_device(devc.obs, devc.adj_bits, devc.chan_bits, tick=tick, episode=torch.zeros_like(tick), probs=probs, actions=actions)
        assert np.array_equal(acts_np, actions.cpu().numpy()) and np.array_equal(probs_np, probs.cpu().numpy())
        out = host.step_host(acts_np if s % 2 else out["pinned"] and torch.from_numpy(acts_np.copy()).pin_memory())
        devc.step(actions)
        assert np.array_equal(out["obs"], devc.obs.cpu().numpy())
        assert np.array_equal(out["reward"], devc.reward.cpu().numpy()) and np.array_equal(out["done"], devc.done.cpu().numpy())
        assert np.array_equal(out["chan_bits"], devc.chan_bits.cpu().numpy())


@pytest.mark.parametrize("scen,kind,parts", [("pp", "comm", 3), ("co", "comm", 4), ("pp", "dec", 2), ("co", "cent", 2)])
def test_host_rollout_step_equals_device_rollout(scen, kind, parts):
    """HostRollout (cm_rollout_step_host: one C call = policy forward -> env step on device-resident state + D2H of what the
    sampler appends, per env part) delivers to pinned host memory exactly the trajectory RolloutEngine records on the device,
    with two steps in flight (submit before collect) and auto-resets inside the window."""
    from com_marl_b200.rollout import HostRollout, RolloutEngine, make_policy
    if scen == "pp":
        params, spec = _mk("pp", 10, 1, 0.08, 2, 0.3, {"max_env_steps": 12})
    else:
        params, spec = _mk("co", 10, 1, 0.03, 2, 0.1, {"max_env_steps": 14})
    pol = make_policy(spec, kind=kind)
    B, K, n = 203, 30, spec.n_agents
    eng = RolloutEngine(spec, pol, B, ring=K, use_graph=False, record_attention=(kind == "comm"))
    eng.reset()
    eng.run_chunk()
    t = {k: v.cpu().numpy() for k, v in eng.traj.items()}
    hr = HostRollout(spec, pol, B, parts=parts, host_slots=2, record_attention=(kind == "comm"))
    first = hr.reset()
    cat = lambda outs, k: np.concatenate([o[k] for o in outs], axis=0)  # noqa: E731
    assert np.array_equal(cat(first, "obs"), t["obs"][0]) and np.array_equal(cat(first, "adj_bits"), t["adj_bits"][0])
    h2d, d2h = hr.bytes_per_step()
    assert h2d == B * n and d2h >= B * n * (spec.obs_dim * 4 + 21)
    hr.submit()
    for k in range(K):
        if k + 1 < K:
            hr.submit()                      # step k + 1 is enqueued before step k's host buffers are read
        outs = hr.collect()
        assert [o["env_ids"] for o in outs] == [(b0, b1) for b0, b1 in hr.ranges]
        for key, ref in (("actions", t["actions"][k]), ("probs", t["probs"][k]), ("reward", t["reward"][k]), ("done", t["done"][k]),
                         ("counts", t["counts"][k]), ("success", t["success"][k]), ("obs", t["obs"][k + 1]),
                         ("adj_bits", t["adj_bits"][k + 1]), ("chan_bits", t["chan_bits"][k + 1]), ("ave_deg", t["ave_deg"][k + 1])):
            assert np.array_equal(cat(outs, key), ref), (key, k)
        if kind == "comm":
            assert np.array_equal(cat(outs, "attention"), t["attention"][k])
    assert t["done"].sum() > 0
    hr.check_errors()
    assert hr.kernel_launches == 2 * K * len(hr.ranges)


def test_graph_replay_sees_updated_weights():
    """The captured rollout graph holds raw pointers to the policy's kernel weight blobs; the blobs are persistent buffers
    refreshed in place, so a parameter update between replays (what DevicePPO / FlatAdam do) is picked up by the next replay
    — no stale or freed weight memory."""
    from com_marl_b200.rollout import RolloutEngine, make_policy
    params, spec = _mk("pp", 10, 1, 0.08, 2, 0.3, {"max_env_steps": 20})
    pol = make_policy(spec)
    a = RolloutEngine(spec, pol, 150, ring=4, use_graph=True)
    b = RolloutEngine(spec, pol, 150, ring=4, use_graph=False)
    for e in (a, b):
        e.reset()
        e.run(8)                 # a: eager chunk, then capture + replay
    assert a._graph is not None
    ptrs = (pol._blob.data_ptr(), pol._tc_blob.data_ptr())
    g = torch.Generator(device="cuda").manual_seed(3)
    with torch.no_grad():
        for p in pol.parameters():
            p.add_(0.3 * torch.randn(p.shape, device=p.device, generator=g))     # in-place update bumps the version counters
    junk = [torch.randn(1 << 20, device="cuda") for _ in range(8)]               # whatever was freed would be reused here
    for e in (a, b):
        e.run(8)
    assert (pol._blob.data_ptr(), pol._tc_blob.data_ptr()) == ptrs
    for k in ("obs", "actions", "probs", "reward", "done"):
        assert torch.equal(a.traj[k], b.traj[k]), k
    del junk


@pytest.mark.parametrize("case", ["pp", "co", "pp_c3"])
def test_sampler_paths_content_and_reference_layout(case):
    """DeviceRolloutSampler.obtain_samples against (1) the reference sampler's own paths stored in tests/golden/ppo_*.npz
    (recorded from CentralizedMAOnPolicyVectorizedSampler by make_golden_ppo.py): same keys, trailing shapes and dtypes
    (observations are float32 here — documented deviation), availability all ones; (2) an independent cut of the device
    trajectory a RolloutEngine with the same seed and env ids records (itself replayed on the oracle in
    test_fused_rollout_matches_oracle): every array of every path is exactly the episode's slice."""
    import json
    from types import SimpleNamespace
    from com_marl_b200.envs import CoverageWrapper, PredatorPreyWrapper
    from com_marl_b200.rollout import RolloutEngine, make_policy
    from com_marl_b200.sampler import DeviceRolloutSampler
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"ppo_{case}.npz"))
    m = json.loads(str(z["meta"]))
    scen, T = m["scenario"], int(m["T"])
    params = ref_harness.scenario_params(scen, int(m["map"]), int(m["sen"]), m["den"], cap=int(m["cap"]),
                                         loss=0.3 if case == "pp" else (0.2 if case == "pp_c3" else 0.0), max_env_steps=T)
    env = (PredatorPreyWrapper(centralized=True, other_agent_visible=True, params=params) if scen == "pp"
           else CoverageWrapper(centralized=True, other_agent_visible=True, max_steps=T, params=params))
    spec = env.spec_b200
    n, D, p, L, B, K = spec.n_agents, spec.obs_dim, spec.n_preys, spec.n_layers, 9, 2 * T + 3
    assert (n, D) == (m["n"], m["D"])
    pol = make_policy(spec)
    sampler = DeviceRolloutSampler(SimpleNamespace(policy=pol, max_path_length=T), env, n_envs=B, chunk=K)
    sampler.start_worker()
    paths = sampler.obtain_samples(0, batch_size=1)            # one chunk of K steps, whole paths
    # (1) layout of the reference sampler's paths
    for key in ("observations", "actions", "avail_actions", "rewards", "dist_adjs", "channels"):
        ref = z[f"path0::{key}"]
        for pth in paths:
            assert pth[key].shape[1:] == ref.shape[1:], key
            if key != "observations" and not (key == "dist_adjs" and spec.rcom == 0):
                assert pth[key].dtype == ref.dtype, (key, pth[key].dtype, ref.dtype)
        assert np.array_equal(np.unique(z["path0::avail_actions"]), [1]) and all((pth["avail_actions"] == 1).all() for pth in paths)
    # (2) content: an engine with the same seed / env ids records the same trajectory; cut it independently
    spec.max_path_length = T
    eng = RolloutEngine(spec, pol, B, ring=K, use_graph=False, record_attention=True)
    eng.reset()
    eng.run_chunk()
    t = {k: v.cpu().numpy() for k, v in eng.traj.items()}
    unpack = lambda bits: ((bits.view(np.uint32)[..., np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(np.float32)  # noqa: E731
    want = []
    for e in range(B):
        start = 0
        for k in np.nonzero(t["done"][:, e])[0]:
            want.append((e, start, int(k) + 1))
            start = int(k) + 1
    assert len(want) == len(paths) >= B
    for pth, (e, a, b) in zip(paths, want):
        sl, T_ = slice(a, b), b - a
        assert np.array_equal(pth["observations"], t["obs"][sl, e].reshape(T_, n * D))
        assert np.array_equal(pth["actions"], t["actions"][sl, e].astype(np.int64))
        assert np.array_equal(pth["rewards"], t["reward"][sl, e]) and pth["rewards"].dtype == np.float64
        assert np.array_equal(pth["dones"], t["done"][sl, e].astype(bool))
        assert np.array_equal(pth["dist_adjs"], unpack(t["adj_bits"][sl, e]).reshape(T_, n * n))
        assert np.array_equal(pth["channels"], unpack(t["chan_bits"][sl, e]).reshape(T_, L * n, n))
        assert np.array_equal(pth["agent_infos"]["action_probs"], t["probs"][sl, e])
        assert np.array_equal(pth["agent_infos"]["attention_weights"], t["attention"][sl, e])
        if p:
            assert np.array_equal(pth["env_infos"]["prey_alive"], t["prey_alive_out"][sl, e, :p].astype(bool))
        if spec.rcom != 0:
            assert np.array_equal(pth["ave_degs"], t["ave_deg"][sl, e])
        assert pth["success"][0] == t["success"][b - 1, e]
        for det, c, r in zip(pth["rewards_details"], t["counts"][sl, e], t["reward"][sl, e]):
            assert det["reward"] == r and det["step_cnt"] == 1
            assert det["capture_cnt"] == (int(c[0]) if scen == "pp" else c[0] / float(n)) and det["move_cnt"] == c[1] / float(n)
    sampler.shutdown_worker()


@pytest.mark.parametrize("scen,kind", [("pp", "comm"), ("co", "comm"), ("pp", "dec"), ("pp_large", "comm")])
def test_packed_observations_equal_fp32_rows(scen, kind):
    """cm_step_io.obs_bits (window bits + scalar columns, 24 bytes per agent) is exactly the fp32 observation row, and the
    tensor-core policy kernels fed with it (first operand built from the bits, low-order pass only for the K slice with the
    scalar columns) return bit-identical logits / probabilities / actions to the same kernels fed with the fp32 rows."""
    from com_marl_b200.envs import BatchedEnv
    from com_marl_b200.rollout import make_policy
    if scen == "pp":
        params, spec = _mk("pp", 20, 2, 0.08, 4, 0.2, {"max_env_steps": 30})
    elif scen == "pp_large":
        params, spec = _mk("pp", 30, 2, 0.08, 4, 0.0, {"max_env_steps": 12})
    else:
        params, spec = _mk("co", 30, 2, 0.06, 2, 0.1, {"max_env_steps": 40})
    B, n, D = 77, spec.n_agents, spec.obs_dim
    env = BatchedEnv(spec, B)
    bits = env.enable_obs_bits()
    env.reset()
    pol = make_policy(spec, kind=kind)
    acts = torch.randint(0, 5, (B, n), dtype=torch.int8, device="cuda")
    for _ in range(3):
        env.step(acts)
    nb = env.obs_nbits
    assert nb == D - (3 if spec.scenario == "pp" else 2)
    w = np.ascontiguousarray(bits.cpu().numpy()).view(np.uint32)                      # [B, n, 6]
    cols = np.arange(nb)
    unpacked = ((w[..., cols >> 5] >> (cols & 31).astype(np.uint32)) & 1).astype(np.float32)
    scal = np.ascontiguousarray(w[..., 3:3 + D - nb]).view(np.float32)
    assert np.array_equal(np.concatenate([unpacked, scal], axis=-1), env.obs.cpu().numpy())
    outs = []
    for packed in (False, True):
        probs = torch.empty((B, n, 5), device="cuda"); logits = torch.empty((B, n, 5), device="cuda")
        actions = torch.empty((B, n), dtype=torch.int8, device="cuda")
        kw = dict(obs_bits=bits, obs_nbits=nb) if packed else {}
        if kind == "comm":
            pol.act_device(env.obs, env.adj_bits, env.chan_bits, tick=env.tick, episode=env.episode, probs=probs, logits=logits, actions=actions, **kw)
        else:
            pol.act_device(env.obs, tick=env.tick, episode=env.episode, probs=probs, logits=logits, actions=actions, **kw)
        outs.append((logits.cpu(), probs.cpu(), actions.cpu()))
    pol.check_errors()
    for a, b in zip(*outs):
        assert torch.equal(a, b)
