"""GPU parity of the env / comm kernels, through the C ABI (com_marl_b200.envs.BatchedEnv -> cm_env_step).

1. against the golden fixtures recorded from the UNMODIFIED reference (injected prey-move / packet-loss
   streams, injected spawns, the reference's own action sequences): bit-exact observations, rewards,
   dones, counts, positions, adjacency / channel masks, ave_deg, success.
2. against the C oracle on thousands of envs in generated-stream mode (Philox spawn / prey walk / channel
   draws): bit-exact again, which also pins the production RNG stream specification.
"""
import numpy as np
import pytest
import torch

from golden_util import EnvCase, env_cases
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _spec_for(case, seed=0):
    from com_marl_b200.scenario import ScenarioSpec
    spec = ScenarioSpec.from_params(case.scenario, case.params, seed=seed, max_path_length=case.max_path_length,
                                    channel_type="GE" if case.ge else None)
    if case.ge:
        spec.pgb, spec.pbg, spec.ge_init, spec.loss_apply = case.ge["Pgb"], case.ge["Pbg"], case.ge["GE_INIT"], case.ge["loss_apply"]
    return spec


def _unpack_bits(bits, n):
    """int32 bit rows (..., W) -> uint8 (..., n)"""
    b = np.ascontiguousarray(bits).view(np.uint32)
    cols = np.arange(n)
    return ((b[..., cols >> 5] >> (cols & 31).astype(np.uint32)) & 1).astype(np.uint8)


@pytest.mark.parametrize("name", env_cases())
def test_env_kernel_matches_reference_golden(name):
    from com_marl_b200.envs import BatchedEnv
    case = EnvCase(name)
    z = case.z
    spec = _spec_for(case)
    B = 3                                   # the same episode in three env slots
    env = BatchedEnv(spec, B, auto_reset=True)
    n, p, L, S = case.n, case.p, case.L, case.steps
    rep = lambda a: np.broadcast_to(a[None], (B,) + a.shape)  # noqa: E731
    env.set_spawn_queue(rep(z["spawn_agent"]), rep(z["spawn_prey"]) if p else None)
    use_u = spec.channel in (2, 3)
    dev = env.device
    # trajectory buffers: the kernel writes slot s directly (zero-copy recording through `out=`)
    T = dict(obs=torch.zeros((S + 1, B, n, env.D), device=dev),
             adj_bits=torch.zeros((S + 1, B, n, env.W), dtype=torch.int32, device=dev),
             chan_bits=torch.zeros((S + 1, B, L, n, env.W), dtype=torch.int32, device=dev),
             ave_deg=torch.zeros((S + 1, B), device=dev),
             reward=torch.zeros((S, B), dtype=torch.float64, device=dev),
             done=torch.zeros((S, B), dtype=torch.uint8, device=dev),
             counts=torch.zeros((S, B, 6), dtype=torch.int32, device=dev),
             prey_alive_out=torch.zeros((S, B, max(p, 1)), dtype=torch.uint8, device=dev))
    apos = torch.zeros((S + 1, B, n), dtype=torch.int16, device=dev)
    ppos = torch.zeros((S + 1, B, max(p, 1)), dtype=torch.int16, device=dev)
    succ = torch.zeros((S, B), dtype=torch.uint8, device=dev)
    vis = torch.zeros((S + 1, B, env.G), dtype=torch.int64, device=dev)
    tot = torch.zeros((S + 1, B), dtype=torch.int32, device=dev)
    chan_u = torch.from_numpy(case.chan_u).to(dev) if use_u else None
    cand = torch.from_numpy(case.cand).to(dev)
    acts = torch.from_numpy(case.actions).to(dev)

    def slot(keys, s):
        return {k: T[k][s] for k in keys}

    env.reset(chan_u=chan_u[0][None].expand(B, -1, -1, -1) if use_u else None, out=slot(("obs", "adj_bits", "chan_bits", "ave_deg"), 0))
    apos[0], ppos[0], vis[0], tot[0] = env.agent_pos, env.prey_pos, env.visited, env.total_capture
    for s in range(S):
        out = slot(("obs", "adj_bits", "chan_bits", "ave_deg"), s + 1)
        out.update(slot(("reward", "done", "counts", "prey_alive_out"), s))
        env.step(acts[s][None].expand(B, -1), prey_cand=cand[s][None].expand(B, -1, -1) if p else None,
                 chan_u=chan_u[s + 1][None].expand(B, -1, -1, -1) if use_u else None, out=out)
        apos[s + 1], ppos[s + 1], succ[s], vis[s + 1], tot[s + 1] = env.agent_pos, env.prey_pos, env.success, env.visited, env.total_capture
    env.check_errors()
    H = {k: v.cpu().numpy() for k, v in T.items()}
    for b in range(B):
        assert np.array_equal(H["obs"][:, b].reshape(S + 1, -1), z["obs"]), "observations differ from the reference"
        assert np.array_equal(H["reward"][:, b], z["reward"]), "rewards differ"
        assert np.array_equal(H["done"][:, b].astype(bool), z["done"].astype(bool))
        assert np.array_equal(succ[:, b].cpu().numpy(), z["success"])
        # details rebuilt from integer counts exactly as the reference forms them
        c = H["counts"][:, b].astype(np.float64)
        nn = float(n)
        if case.scenario == "pp":
            det = np.stack([c[:, 0], c[:, 1] / nn, c[:, 2], c[:, 3] / nn, np.zeros(S)], axis=1)
        else:
            det = np.stack([c[:, k] / nn for k in range(5)], axis=1)
        assert np.array_equal(det, z["details"]), "reward_details differ"
        a_pos = np.ascontiguousarray(apos[:, b].cpu().numpy()).view(np.uint16)
        assert np.array_equal(np.stack([a_pos & 0xFF, a_pos >> 8], -1), z["agent_pos"].astype(np.uint16))
        if p:
            p_pos = np.ascontiguousarray(ppos[:, b].cpu().numpy()).view(np.uint16)
            assert np.array_equal(np.stack([p_pos & 0xFF, p_pos >> 8], -1), z["prey_pos"].astype(np.uint16))
            assert np.array_equal(H["prey_alive_out"][:, b], z["prey_alive"])
        adj = _unpack_bits(H["adj_bits"][:, b], n)
        ref_adj = np.unpackbits(z["adj"], axis=-1, count=n, bitorder="little")
        assert np.array_equal(adj, ref_adj), "adjacency differs"
        ch = _unpack_bits(H["chan_bits"][:, b], n)
        ref_ch = np.unpackbits(z["chan"], axis=-1, count=n, bitorder="little")
        assert np.array_equal(ch, ref_ch), "channel masks differ"
        assert np.array_equal(H["ave_deg"][:, b], z["ave_deg"].astype(np.float32))
        if case.scenario == "co":
            rows = np.ascontiguousarray(vis[:, b].cpu().numpy()).view(np.uint64)
            cols = np.arange(env.G, dtype=np.uint64)
            grid = ((rows[:, :, None] >> cols[None, None, :]) & np.uint64(1)).astype(np.uint8).reshape(S + 1, -1)
            ref_vis = np.unpackbits(z["visited"], axis=-1, count=env.G * env.G, bitorder="little")
            assert np.array_equal(grid, ref_vis), "visited map differs"
            assert np.array_equal(tot[:, b].cpu().numpy(), z["total_capture"])
    assert int(env.episode[0].item()) == case.meta["episodes"]


GENERATED = [
    # (tag, scenario, map, sen, den, cap, loss, B, steps, overrides, ge)
    ("c1", "pp", 10, 1, 0.04, 2, 0.0, 4096, 230, {}, None),
    ("c2", "co", 10, 1, 0.03, 2, 0.0, 4096, 430, {}, None),
    ("c3", "pp", 20, 2, 0.08, 4, 0.2, 1024, 210, {}, None),
    ("c4", "co", 30, 2, 0.06, 2, 0.1, 512, 70, {"max_env_steps": 60}, None),
    ("c5", "pp", 50, 2, 0.08, 4, 0.0, 96, 40, {"max_env_steps": 30}, None),
    ("cap3_rcom2", "pp", 10, 2, 0.08, 3, 0.5, 2048, 120, {"trRcom": 2, "max_env_steps": 50}, None),
    ("capture", "pp", 6, 1, 0.08, 2, 0.0, 2048, 150, {"n_agents": 8, "n_preys": 3, "max_env_steps": 40, "penalty": 0.5, "rm": 0.25}, None),
    ("ge_l1", "pp", 10, 1, 0.08, 2, 0.2, 1024, 90, {"max_env_steps": 40}, dict(Pgb=0.2, Pbg=0.3, GE_INIT=1, loss_apply=1)),
    ("ge_prop", "co", 20, 1, 0.03, 2, 0.2, 512, 100, {"max_env_steps": 45}, dict(Pgb=0.0196, Pbg=0.282, GE_INIT=-1, loss_apply=1)),
    ("ge_l0", "pp", 10, 1, 0.08, 2, 0.2, 512, 90, {"max_env_steps": 40}, dict(Pgb=0.2, Pbg=0.3, GE_INIT=0, loss_apply=0)),
    ("hard", "co", 10, 2, 0.06, 2, 1.0, 1024, 120, {"obstComplex": "Hard", "trRcom": 3, "max_env_steps": 50}, None),
    # maximum sizes: 256 agents + 256 preys on a 64 x 64 grid (full-width 64-bit rows); Coverage map 60 (grid 62), 216 agents
    ("max_pp", "pp", 64, 2, 0.08, 4, 0.1, 6, 26, {"n_agents": 256, "n_preys": 256, "max_env_steps": 12}, None),
    ("max_co", "co", 60, 2, 0.06, 2, 0.1, 6, 26, {"max_env_steps": 12}, None),
    ("tiny", "pp", 10, 0, 0.01, 2, 0.0, 300, 30, {"n_agents": 1, "n_preys": 1, "max_env_steps": 10}, None),
    # loss thresholds at the ends of (0, 1): the kernel compares the draws as integers against ceil(p 2^24) << 8, the oracle
    # as floats; teams that are not a multiple of 4 walk Philox blocks that straddle the row ends
    ("loss_hi", "pp", 10, 1, 0.08, 2, 0.999999, 512, 40, {"n_agents": 7, "n_preys": 3, "max_env_steps": 20}, None),
    ("loss_lo", "pp", 10, 1, 0.08, 2, 1e-7, 512, 40, {"n_agents": 13, "n_preys": 5, "max_env_steps": 20}, None),
    ("ge_ends", "pp", 10, 1, 0.08, 2, 0.2, 512, 60, {"n_agents": 9, "max_env_steps": 25}, dict(Pgb=0.999, Pbg=1e-6, GE_INIT=1, loss_apply=1)),
]


@pytest.mark.parametrize("cfg", GENERATED, ids=[g[0] for g in GENERATED])
def test_env_kernel_matches_oracle_generated_streams(cfg):
    """Thousands of envs, on-device Philox streams, auto-reset: every output and the final state equal the
    C oracle's.  Also covers env_id0 offsets (sharding invariance of the stream keys)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import ref_harness
    from com_marl_b200.envs import BatchedEnv
    from com_marl_b200.scenario import ScenarioSpec
    import streams
    tag, scen, m, sen, den, cap, loss, B, steps, over, ge = cfg
    params = ref_harness.scenario_params(scen, m, sen, den, cap=cap, loss=loss, **over)
    seed, env_id0 = 1234, 1000
    ospec = orc.spec_from_params(scen, params, seed=seed, ge=ge)
    oenv = orc.OracleVecEnv(ospec, B, env_id0=env_id0)
    spec = ScenarioSpec.from_params(scen, params, seed=seed, channel_type="GE" if ge else None)
    if ge:
        spec.pgb, spec.pbg, spec.ge_init, spec.loss_apply = ge["Pgb"], ge["Pbg"], ge["GE_INIT"], ge["loss_apply"]
    env = BatchedEnv(spec, B, env_id0=env_id0)
    n, p = spec.n_agents, spec.n_preys
    oenv.reset()
    env.reset()
    acts = streams.actions(77, (steps, B, n), bias_move=(tag in ("capture", "c2")))

    def compare(s):
        assert np.array_equal(env.obs.cpu().numpy(), oenv.obs), f"obs differ at step {s}"
        a, pp = env.positions()
        assert np.array_equal(a, oenv.apos), f"agent positions differ at step {s}"
        assert np.array_equal(_unpack_bits(env.adj_bits.cpu().numpy(), n), oenv.adj), f"adjacency differs at step {s}"
        assert np.array_equal(_unpack_bits(env.chan_bits.cpu().numpy(), n), oenv.chan), f"channels differ at step {s}"
        assert np.array_equal(env.ave_deg.cpu().numpy(), oenv.ave_deg)
        if p:
            assert np.array_equal(pp, oenv.ppos[:, :p])
            assert np.array_equal(env.prey_alive.cpu().numpy()[:, :p], oenv.alive[:, :p])
        else:
            assert np.array_equal(env.visited_grid().reshape(B, -1), oenv.visited)
            assert np.array_equal(env.total_capture.cpu().numpy(), oenv.total_capture)
        assert np.array_equal(env.step_count.cpu().numpy(), oenv.t)
        assert np.array_equal(env.episode.cpu().numpy().view(np.uint32), oenv.episode)

    compare(-1)
    ndone = 0
    for s in range(steps):
        env.step(acts[s])
        oenv.step(acts[s])
        assert np.array_equal(env.reward.cpu().numpy(), oenv.reward), f"reward differs at step {s}"
        assert np.array_equal(env.done.cpu().numpy(), oenv.done), f"done differs at step {s}"
        assert np.array_equal(env.counts.cpu().numpy(), oenv.counts), f"counts differ at step {s}"
        assert np.array_equal(env.success.cpu().numpy(), oenv.success)
        if p:
            assert np.array_equal(env.prey_alive_out.cpu().numpy()[:, :p], oenv.prey_alive_out[:, :p])
        ndone += int(oenv.done.sum())
        if s % 7 == 0 or s == steps - 1 or oenv.done.any():
            compare(s)
    env.check_errors()
    assert ndone > 0, "the case never exercised the auto-reset path"
    # episode accounting written by the kernel == sums of the oracle's per-step outputs
    st = env.stats.cpu().numpy()
    assert st[:, 7].sum() == ndone


def test_sharding_invariance():
    """Two shards with env_id0 offsets reproduce one big batch (stream keys use global env ids)."""
    from com_marl_b200.envs import BatchedEnv
    from com_marl_b200.scenario import ScenarioSpec
    import streams
    spec = ScenarioSpec.from_cli("pp", 10, 1, 0.08, cap=2, loss=0.3, seed=5)
    B, n = 512, spec.n_agents
    full = BatchedEnv(spec, B)
    lo, hi = BatchedEnv(spec, B // 2, env_id0=0), BatchedEnv(spec, B // 2, env_id0=B // 2)
    for e in (full, lo, hi):
        e.reset()
    acts = streams.actions(3, (60, B, n))
    for s in range(60):
        full.step(acts[s]); lo.step(acts[s][: B // 2]); hi.step(acts[s][B // 2:])
    both = torch.cat([lo.obs, hi.obs]).cpu().numpy()
    assert np.array_equal(full.obs.cpu().numpy(), both)
    assert np.array_equal(full.chan_bits.cpu().numpy(), torch.cat([lo.chan_bits, hi.chan_bits]).cpu().numpy())
    assert np.array_equal(full.reward.cpu().numpy(), torch.cat([lo.reward, hi.reward]).cpu().numpy())


def test_invalid_action_raises():
    from com_marl_b200.envs import BatchedEnv
    from com_marl_b200.scenario import ScenarioSpec
    spec = ScenarioSpec.from_cli("pp", 10, 1, 0.04)
    env = BatchedEnv(spec, 4)
    env.reset()
    a = np.zeros((4, spec.n_agents), dtype=np.int8)
    a[2, 1] = 7
    env.step(a)
    with pytest.raises(Exception, match="Action Not found"):
        env.check_errors()


def test_mask_pack_unpack_roundtrip():
    from com_marl_b200.policy import CommCategoricalMLPPolicy
    from com_marl_b200.envs import BatchedEnv
    from com_marl_b200.scenario import ScenarioSpec
    for n in (3, 32, 54, 200):
        dense = (torch.rand((7, n, n), device="cuda") > 0.4).float()
        bits = CommCategoricalMLPPolicy.pack_mask(dense, n)
        got = _unpack_bits(bits.cpu().numpy(), n)
        assert np.array_equal(got, dense.cpu().numpy().astype(np.uint8))
    spec = ScenarioSpec.from_cli("pp", 20, 2, 0.08, cap=4, loss=0.2)
    env = BatchedEnv(spec, 16)
    env.reset()
    d = env.channels().cpu().numpy()
    assert np.array_equal(d.astype(np.uint8), _unpack_bits(env.chan_bits.cpu().numpy(), spec.n_agents))


@pytest.mark.parametrize("name", ["pp_c1", "pp_c3", "pp_capture", "pp_rcom2", "co_c2", "co_hard", "co_c4"])
def test_gym_wrapper_contract_matches_reference(name):
    """The B=1 wrappers return what the reference wrappers return (shapes, tuple structure, details dict), and a WHOLE
    episode through the public reset()/step() API reproduces the reference's observations, rewards, details, done flags,
    prey_alive infos, success and comm attributes — PredatorPrey with the reference's prey-move candidates and packet-loss
    uniforms injected through the wrapper (inject_streams: the streams the fixture's generator fed the reference)."""
    from com_marl_b200.envs import PredatorPreyWrapper, CoverageWrapper
    case = EnvCase(name)
    z = case.z
    cls = PredatorPreyWrapper if case.scenario == "pp" else CoverageWrapper
    kw = dict(max_steps=case.T) if case.scenario == "co" else {}
    env = cls(centralized=True, other_agent_visible=True, params=case.params, **kw)
    env._vec.set_spawn_queue(z["spawn_agent"][None], z["spawn_prey"][None] if case.p else None)
    use_u = env.spec_b200.channel in (2, 3)
    env.inject_streams(prey_cand=case.cand if case.p else None, chan_u=case.chan_u if use_u else None)
    obs = env.reset()
    D = env.spec_b200.obs_dim
    assert obs.shape == (case.n * D,) and np.array_equal(obs, z["obs"][0])
    assert env.observation_space.flat_dim == obs.shape[0] and env.action_space.n == 5
    assert env.spec.observation_space is env.observation_space
    assert env.get_avail_actions().shape == (case.n * 5,)
    assert env.dist_adj.shape == (case.n, case.n) and env.channels.shape == (case.L, case.n, case.n)
    assert np.array_equal(env.channels.astype(np.uint8), case.unpack("chan", 0))
    assert env.bound_return == pytest.approx(case.meta["bound_return"])
    episodes = 0
    for s in range(case.steps):
        o, (r, det), done, info = env.step(case.actions[s])
        if s == 0:
            assert o.shape == obs.shape and isinstance(r, float) and isinstance(done, bool)
            assert set(det) == {"reward", "capture_cnt", "step_cnt", "move_cnt", "penalty_cnt", "variable", "vars2"}
        assert r == z["reward"][s] and done == bool(z["done"][s]), s
        assert [det["capture_cnt"], det["move_cnt"], det["penalty_cnt"], det["variable"], det["vars2"]] == list(z["details"][s])
        assert env.success == z["success"][s]
        if case.scenario == "pp":
            assert "prey_alive" in info and np.array_equal(info["prey_alive"], z["prey_alive"][s].astype(bool))
        if done:                                   # like the sampler's VecEnvExecutor: reset, the reset observation replaces o
            episodes += 1
            o = env.reset()
        assert np.array_equal(o, z["obs"][s + 1])
        assert np.array_equal(env.channels.astype(np.uint8), case.unpack("chan", s + 1))
        assert np.array_equal(np.asarray(env.dist_adj).astype(np.uint8), case.unpack("adj", s + 1))
        if env.spec_b200.rcom != 0:
            assert env.ave_deg == np.float32(z["ave_deg"][s + 1])
    assert episodes == int(z["done"].sum()) >= 1
    with pytest.raises(Exception, match="Action Not found"):
        env.step([9] * case.n)
