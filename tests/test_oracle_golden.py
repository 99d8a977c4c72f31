"""Pins the CPU oracle (oracle/) against outputs of the unmodified reference (tests/golden/*.npz).

Bit-exact for grid state, observations, adjacency / channel masks, rewards, dones, counts
(BASELINE.json north_star); policy logits within 1e-5 of the reference torch forward.
"""
import os

import numpy as np
import pytest

from golden_util import EnvCase, PolicyCase, env_cases, policy_cases
from oracle import oracle as orc


def run_oracle_case(case, B=1):
    spec = orc.spec_from_params(case.scenario, case.params, seed=0, max_path_length=case.max_path_length, ge=case.ge)
    env = orc.OracleVecEnv(spec, B)
    z = case.z
    sa = np.broadcast_to(z["spawn_agent"][None], (B,) + z["spawn_agent"].shape)
    sp = np.broadcast_to(z["spawn_prey"][None], (B,) + z["spawn_prey"].shape) if case.p else None
    env.set_spawn_queue(sa, sp)
    return spec, env


@pytest.mark.parametrize("name", env_cases())
def test_env_oracle_matches_reference(name):
    case = EnvCase(name)
    z = case.z
    spec, env = run_oracle_case(case)
    n = case.n
    if case.scenario == "co":
        assert spec["n_empty_cells"] == case.meta["n_empty_cells"]
    use_u = spec["chan_type"] in (orc.CH_IID, orc.CH_GE)

    def check(s):
        assert np.array_equal(env.obs[0].reshape(-1), z["obs"][s]), f"obs differs at update {s}"
        assert np.array_equal(env.apos[0], z["agent_pos"][s])
        if case.p:
            # the reference keeps stale positions for dead preys as well, so positions compare for all
            assert np.array_equal(env.ppos[0], z["prey_pos"][s])
        assert np.array_equal(env.adj[0], case.unpack("adj", s)), f"adjacency differs at update {s}"
        assert np.array_equal(env.chan[0], case.unpack("chan", s)), f"channels differ at update {s}"
        assert np.float32(z["ave_deg"][s]) == env.ave_deg[0]
        if case.scenario == "co":
            vis = np.unpackbits(z["visited"][s], count=env.G * env.G, bitorder="little")
            assert np.array_equal(env.visited[0], vis)
            assert env.total_capture[0] == z["total_capture"][s]

    env.reset(chan_u=case.chan_u[0][None] if use_u else None)
    check(0)
    for s in range(case.steps):
        env.step(case.actions[s][None], prey_cand=case.cand[s][None] if case.p else None,
                 chan_u=case.chan_u[s + 1][None] if use_u else None)
        assert env.reward[0] == z["reward"][s], f"reward differs at step {s}"
        assert bool(env.done[0]) == bool(z["done"][s]), f"done differs at step {s}"
        assert env.success[0] == z["success"][s]
        assert np.array_equal(env.details()[0], z["details"][s]), f"details differ at step {s}"
        if case.p:
            assert np.array_equal(env.prey_alive_out[0], z["prey_alive"][s])
        check(s + 1)
    assert env.episode[0] == case.meta["episodes"]


def test_ge_direct_matches_reference():
    """gilbert_elliot_loss_model.get_init_state / get_next_state_matrix called directly
    (not through an env): proportional init + a run of transitions."""
    import os
    from golden_util import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "ge_direct.npz"))
    for tag in ("default", "busy"):
        n, seq, pgb, pbg = z[f"{tag}_cfg"]
        n, seq = int(n), int(seq)
        u = z[f"{tag}_u"]
        ref = np.unpackbits(z[f"{tag}_states"], axis=-1, count=n, bitorder="little")
        params = dict(grid_size=30, n_agents=n, n_preys=n, Rsen=1, n_gcn_layers=seq + 1, max_env_steps=10, load=2,
                      capture_reward=10, step_cost=0.1, rm=0, penalty=0, mode="train", trpl=0.5, trRcom=9)
        spec = orc.spec_from_params("pp", params, ge=dict(Pgb=pgb, Pbg=pbg, GE_INIT=-1, loss_apply=1))
        env = orc.OracleVecEnv(spec, 1)
        # reset with L = seq+1 layers: init draw + seq transitions, include_prev=True == ref sequence
        env.reset(chan_u=u[None])
        assert np.array_equal(env.chan[0], ref)


@pytest.mark.parametrize("name", policy_cases())
def test_policy_oracle_matches_reference(name):
    c = PolicyCase(name)
    logits, probs, attn = orc.policy_forward(c.weights, c.obs, c.avail, c.adj, c.chan)
    scale = max(1.0, float(np.abs(c.logits).max()))
    assert np.abs(logits - c.logits).max() <= 1e-5 * scale
    assert np.abs(probs - c.probs).max() <= 1e-5
    assert np.abs(attn - c.attn).max() <= 1e-5
    # float64 evaluation agrees as well (the float32 reference is within rounding of the exact formula)
    logits64, _, _ = orc.policy_forward(c.weights, c.obs, c.avail, c.adj, c.chan, dtype=np.float64)
    assert np.abs(logits64 - c.logits).max() <= 1e-5 * scale


def _dec_cases():
    import glob
    return sorted(os.path.basename(p)[7:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "decpol_*.npz")))


@pytest.mark.parametrize("name", _dec_cases())
def test_dec_policy_oracle_matches_reference(name):
    """Obs-DP forward restatement (oracle.policy_forward_dec) against the vectors recorded from the unmodified reference
    DecCategoricalMLPPolicy (tests/golden/make_golden_dec.py)."""
    import json
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"decpol_{name}.npz"))
    meta = json.loads(str(z["meta"]))
    n, D, B = meta["n"], meta["D"], meta["B"]
    w = {k[3:]: z[k] for k in z.files if k.startswith("w::")}
    logits, probs = orc.policy_forward_dec(w, z["obs"].reshape(B, n, D), z["avail"].reshape(B, n, 5))
    assert np.abs(logits - z["logits"].reshape(B, n, 5)).max() <= 1e-5
    assert np.abs(probs - z["probs"].reshape(B, n, 5)).max() <= 1e-5


def _cent_cases():
    import glob
    return sorted(os.path.basename(p)[8:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "centpol_*.npz")))


@pytest.mark.parametrize("name", _cent_cases())
def test_cent_policy_oracle_matches_reference(name):
    """CENT forward restatement (oracle.policy_forward_cent) against the vectors recorded from the unmodified reference
    CentralizedCategoricalMLPPolicy (tests/golden/make_golden_cent.py)."""
    import json
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"centpol_{name}.npz"))
    meta = json.loads(str(z["meta"]))
    n, B = meta["n"], meta["B"]
    w = {k[3:]: z[k] for k in z.files if k.startswith("w::")}
    logits, probs = orc.policy_forward_cent(w, z["obs"], z["avail"].reshape(B, n, 5), relu=bool(meta["relu"]))
    assert np.abs(logits - z["logits"].reshape(B, n, 5)).max() <= 1e-5 * max(1.0, np.abs(z["logits"]).max())
    assert np.abs(probs - z["probs"].reshape(B, n, 5)).max() <= 1e-5
