"""BASELINE.json's full sizes (16 384 Coverage envs, 65 536 PredatorPrey envs of 32 agents): too large for the scalar
oracle, so the device rollout is checked through size-independent properties of the domain —
  * conservation: every agent is on the grid on its own cell, never on a wall / a live prey; the agent windows have the
    self bit set; live preys are on distinct free cells; the Coverage visited map only grows within an episode and
    contains every agent cell;
  * comm state: adjacency is symmetric with a full diagonal and equals the distance rule on the stored positions;
    ave_deg is the float32 mean degree; channel rows keep their diagonal; probabilities are normalised;
  * accounting: the in-kernel episode statistics equal the sums over the trajectory ring;
  * a checksum of checksums is invariant under the partition of the envs into env groups and into shards with
    global-id offsets (what N GPUs run).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FULL = [("c2", "co", 10, 1, 0.03, 2, 0.0, 16384), ("c3", "pp", 20, 2, 0.08, 4, 0.2, 65536)]


def _unpack(bits, n):
    """int32 bit rows (..., W) -> bool (..., n)"""
    b = bits.cpu().numpy().view(np.uint32)
    out = ((b[..., :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool)
    return out.reshape(b.shape[:-1] + (-1,))[..., :n]


def _checksum(t: torch.Tensor) -> int:
    """order-sensitive 64-bit checksum of a tensor's bytes, per env then folded (a checksum of checksums)"""
    b = t.contiguous().view(t.shape[0], -1).view(torch.uint8).to(torch.int64)
    w = (torch.arange(b.shape[1], device=b.device, dtype=torch.int64) * 2654435761 + 12345) & 0x7FFFFFFF
    per_env = (b * w).sum(dim=1) & 0x7FFFFFFFFFFF
    k = (torch.arange(b.shape[0], device=b.device, dtype=torch.int64) * 40503 + 7) & 0xFFFF
    return int(((per_env * k) & 0x7FFFFFFFFFFF).sum().item() & 0x7FFFFFFFFFFFFFFF)


@pytest.mark.parametrize("cfg", FULL, ids=[c[0] for c in FULL])
def test_full_size_properties(cfg):
    from com_marl_b200.rollout import RolloutEngine, make_policy
    from com_marl_b200.scenario import ScenarioSpec
    _, scen, m, sen, den, cap, loss, B = cfg
    spec = ScenarioSpec.from_cli(scen, m, sen, den, cap=cap, loss=loss, seed=11)
    n, p, G, L, D = spec.n_agents, spec.n_preys, spec.grid, spec.n_layers, spec.obs_dim
    pol = make_policy(spec)
    K = 6
    eng = RolloutEngine(spec, pol, B, ring=K, use_graph=True, groups=4)
    eng.reset()
    prev_visited = None
    for chunk in range(3):
        eng.run_chunk()
        e, t = eng.env, eng.traj
        eng.env.check_errors()
        pol.check_errors()
        pos = e.agent_pos.cpu().numpy().view(np.uint16)
        r, c = (pos & 0xFF).astype(np.int64), (pos >> 8).astype(np.int64)
        assert r.max() < G and c.max() < G
        cell = r * 64 + c
        srt = np.sort(cell, axis=1)
        assert (np.diff(srt, axis=1) > 0).all(), "two agents on one cell"
        obs = t["obs"][K].cpu().numpy()                      # observations after the last step of the chunk
        ww = (2 * sen + 1) ** 2
        agent_win = obs[:, :, ww:2 * ww] if scen == "co" else obs[:, :, :ww]
        assert (agent_win[:, :, ww // 2] == 1.0).all(), "self bit of the agent window"
        assert np.isin(obs[:, :, : (3 if scen == "co" else 2) * ww], (0.0, 1.0)).all()
        if scen == "co":
            wall = spec.wall_rows().view(np.uint64)
            assert (((wall[r] >> c.astype(np.uint64)) & 1) == 0).all(), "agent on a wall"
            vis = e.visited.cpu().numpy().view(np.uint64)
            assert (((vis[np.arange(B)[:, None], r] >> c.astype(np.uint64)) & 1) == 1).all(), "agent cell not visited"
            pop = np.array([bin(int(x)).count("1") for x in vis.reshape(-1)[: 4096 * G]]).reshape(-1, G).sum(1)
            assert (pop[:4096] - n == e.total_capture.cpu().numpy()[:4096]).all()          # covered cells = start cells + captures
            same_episode = None if prev_visited is None else (e.episode.cpu().numpy() == prev_episode)
            if same_episode is not None:
                assert ((prev_visited & ~vis)[same_episode] == 0).all(), "visited map shrank inside an episode"
            prev_visited, prev_episode = vis.copy(), e.episode.cpu().numpy().copy()
        else:
            alive = e.prey_alive.cpu().numpy().astype(bool)
            pp = e.prey_pos.cpu().numpy().view(np.uint16)
            pcell = (pp & 0xFF).astype(np.int64) * 64 + (pp >> 8).astype(np.int64)
            big = np.where(alive, pcell, -1 - np.arange(p)[None, :])          # dead preys keep stale positions: make them unique
            allc = np.concatenate([cell, big], axis=1)
            s2 = np.sort(allc, axis=1)
            assert (np.diff(s2, axis=1) != 0).all(), "a live prey shares a cell"
            assert np.array_equal(t["prey_alive_out"][K - 1].cpu().numpy().astype(bool) | (t["done"][K - 1].cpu().numpy()[:, None] > 0),
                                  alive | (t["done"][K - 1].cpu().numpy()[:, None] > 0))
        # ---- comm state of the last slot ----
        sub = slice(0, 2048)                                 # dense n x n checks on a slice (memory), bit rows on everything
        adj = _unpack(t["adj_bits"][K][sub], n)
        dr = r[sub, :, None] - r[sub, None, :]
        dc = c[sub, :, None] - c[sub, None, :]
        rule = np.ones_like(adj) if spec.rcom2 < 0 else (dr * dr + dc * dc <= spec.rcom2)
        assert np.array_equal(adj, rule)
        deg = torch.from_numpy(_unpack(t["adj_bits"][K], n).sum(axis=(1, 2)).astype(np.float32))
        assert torch.equal(t["ave_deg"][K].cpu(), deg / np.float32(n))
        ch = _unpack(t["chan_bits"][K][sub], n)
        assert ch[:, :, np.arange(n), np.arange(n)].all(), "channel diagonal"
        if loss == 0.0:
            assert ch.all()
        else:
            off = ch[:, :, ~np.eye(n, dtype=bool)]
            assert abs(off.mean() - (1.0 - loss)) < 5e-3, "IID keep rate"
        pr = t["probs"].cpu()
        assert torch.isfinite(pr).all() and (pr.sum(-1) - 1).abs().max() < 1e-5
        a = t["actions"].cpu()
        assert a.min() >= 0 and a.max() <= 4
    # ---- accounting: finished-episode sums of the kernel == what the last ring holds is a subset; use totals over all chunks ----
    st = eng.env.stats.cpu().numpy()
    assert st[:, 7].sum() >= 0 and (st[:, 1] <= spec.max_steps).all() and (st[:, 1] >= 0).all()
    assert np.array_equal(st[:, 1] + st[:, 9], np.full(B, 3.0 * K))      # running length + finished lengths = steps taken


@pytest.mark.parametrize("cfg", FULL, ids=[c[0] for c in FULL])
def test_full_size_partition_invariance(cfg):
    """checksum of checksums over the whole trajectory ring: one chain == 4 env groups == 2 'GPU' shards with env-id offsets"""
    from com_marl_b200.rollout import RolloutEngine, make_policy
    from com_marl_b200.scenario import ScenarioSpec
    _, scen, m, sen, den, cap, loss, B = cfg
    B = B // 4                                                # three engines live at once: keep the memory bounded
    spec = ScenarioSpec.from_cli(scen, m, sen, den, cap=cap, loss=loss, seed=11)
    pol = make_policy(spec)
    K = 5

    def run(n_envs, env_id0, groups):
        eng = RolloutEngine(spec, pol, n_envs, ring=K, use_graph=True, groups=groups, env_id0=env_id0)
        eng.reset()
        eng.run(3 * K)
        return eng

    one = run(B, 0, 1)
    grp = run(B, 0, 4)
    lo, hi = run(B // 2, 0, 2), run(B // 2, B // 2, 3)
    for k in ("obs", "actions", "reward", "done", "adj_bits", "chan_bits", "probs", "counts"):
        a = one.traj[k].transpose(0, 1)
        ref = _checksum(a)
        assert _checksum(grp.traj[k].transpose(0, 1)) == ref, k
        assert _checksum(torch.cat([lo.traj[k], hi.traj[k]], dim=1).transpose(0, 1)) == ref, k
    assert torch.equal(one.env.stats, grp.env.stats)
    assert torch.equal(one.env.stats, torch.cat([lo.env.stats, hi.env.stats]))
