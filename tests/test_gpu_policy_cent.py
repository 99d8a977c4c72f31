"""CENT policy (CentralizedCategoricalMLPPolicy) on policy_cent_kernel: golden vectors recorded from the unmodified
reference (tests/golden/make_golden_cent.py), the float64 oracle on large ragged batches (team sizes 1..256, K = n*D up to
13 568), the sampling specification, the reference call surface, and the device rollout against the oracle."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[8:-4] for p in glob.glob(os.path.join(GOLDEN, "centpol_*.npz")))


def _policy(n, D, weights=None, seed=3, relu=False, math="auto"):
    from com_marl_b200.policy import CentralizedCategoricalMLPPolicy
    from com_marl_b200.spaces import Box, Discrete, EnvSpec
    pol = CentralizedCategoricalMLPPolicy(EnvSpec(Box(np.zeros(n * D), np.ones(n * D)), Discrete(5)), n, seed=seed,
                                          hidden_nonlinearity=torch.relu if relu else torch.tanh, math=math)
    if weights is not None:
        pol.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()})
    return pol


def _bits(avail):
    return torch.from_numpy(((avail != 0) * np.array([1, 2, 4, 8, 16])).sum(-1).astype(np.uint8)).cuda()


@pytest.mark.parametrize("math", ["fp32", "tc"])
@pytest.mark.parametrize("name", CASES)
def test_cent_policy_kernel_matches_reference_golden(name, math):
    z = np.load(os.path.join(GOLDEN, f"centpol_{name}.npz"))
    meta = json.loads(str(z["meta"]))
    n, D, B = meta["n"], meta["D"], meta["B"]
    w = {k[3:]: z[k] for k in z.files if k.startswith("w::")}
    pol = _policy(n, D, w, relu=bool(meta["relu"]), math=math)
    assert set(pol.state_dict().keys()) == set(w.keys())            # reference checkpoints load as they are
    obs = torch.from_numpy(z["obs"]).cuda()
    probs = torch.empty((B, n, 5), device="cuda"); logits = torch.empty((B, n, 5), device="cuda")
    pol.act_device(obs, avail_bits=_bits(z["avail"].reshape(B, n, 5)), greedy=True, probs=probs, logits=logits,
                   actions=torch.empty((B, n), dtype=torch.int8, device="cuda"))
    pol.check_errors()
    ref_logits = z["logits"].reshape(B, n, 5)
    assert np.abs(logits.cpu().numpy() - ref_logits).max() <= 1e-5 * max(1.0, np.abs(ref_logits).max())
    assert np.abs(probs.cpu().numpy() - z["probs"].reshape(B, n, 5)).max() <= 1e-5
    # the torch training path (forward / entropy) is the same formula
    with torch.no_grad():
        dist = pol.forward(torch.from_numpy(z["obs"]).cuda(), torch.from_numpy(z["avail"]).cuda())
        ent = pol.entropy(torch.from_numpy(z["obs"]).cuda()[None], torch.from_numpy(z["avail"]).cuda()[None])
    assert np.abs(dist.probs.cpu().numpy() - z["probs"].reshape(B, n, 5)).max() <= 1e-5
    assert np.abs(ent.cpu().numpy() - z["entropy"]).max() <= 1e-5


@pytest.mark.parametrize("n,D,B,relu", [(3, 29, 16384, False), (4, 21, 5001, False), (32, 53, 2048, False), (54, 77, 777, True),
                                        (200, 53, 130, False), (256, 53, 65, False), (1, 5, 100, False), (17, 7, 63, True)])
@pytest.mark.parametrize("math", ["fp32", "tc"])
def test_cent_policy_kernel_matches_oracle_batched(n, D, B, relu, math):
    rng = np.random.default_rng(n * 1000 + D)
    pol = _policy(n, D, relu=relu, math=math)
    with torch.no_grad():
        for k, v in pol.state_dict().items():
            if k.endswith("bias"):
                v.copy_(torch.from_numpy(rng.uniform(-0.2, 0.2, size=tuple(v.shape)).astype(np.float32)))
    obs = rng.random((B, n * D), dtype=np.float32)
    avail = (rng.random((B, n, 5)) < 0.9).astype(np.float32)
    avail[..., 4] = 1.0
    u = rng.random((B, n), dtype=np.float32)
    probs = torch.empty((B, n, 5), device="cuda"); logits = torch.empty((B, n, 5), device="cuda")
    actions = torch.empty((B, n), dtype=torch.int8, device="cuda")
    pol.act_device(torch.from_numpy(obs).cuda(), avail_bits=_bits(avail), sample_u=torch.from_numpy(u).cuda(), probs=probs,
                   logits=logits, actions=actions)
    pol.check_errors()
    w = {k: v.cpu().numpy() for k, v in pol.state_dict().items()}
    ref_logits, ref_probs = orc.policy_forward_cent(w, obs, avail, relu=relu, dtype=np.float64)
    assert np.abs(logits.cpu().numpy() - ref_logits).max() <= 1e-5 * max(1.0, np.abs(ref_logits).max())
    assert np.abs(probs.cpu().numpy() - ref_probs).max() <= 1e-5
    # sampling == inverse CDF of the kernel's own probabilities (sequential fp32 cumulative sum)
    pr = probs.cpu().numpy()
    cdf = np.zeros_like(pr)
    acc = np.zeros(pr.shape[:-1], dtype=np.float32)
    for a in range(5):
        acc = (acc + pr[..., a]).astype(np.float32)
        cdf[..., a] = acc
    exp = (u[..., None] >= cdf).sum(-1)
    last = 4 - np.argmax((pr > 0)[..., ::-1], axis=-1)
    exp = np.where(exp > 4, last, exp)
    assert np.array_equal(actions.cpu().numpy(), exp)


def test_cent_get_actions_contract():
    n, D, B = 4, 21, 7
    pol = _policy(n, D)
    rng = np.random.default_rng(0)
    obs = rng.random((B, n * D), dtype=np.float32)
    acts, infos = pol.get_actions(obs, np.ones((B, n * 5), dtype=np.float32))
    assert acts.shape == (B, n) and acts.dtype == np.int64 and len(infos["action_probs"]) == B
    g, ginf = pol.get_actions(obs, np.ones((B, n * 5), dtype=np.float32), greedy=True)
    assert np.array_equal(g, np.argmax(np.stack(ginf["action_probs"]), axis=-1))
    a1, _ = pol.get_actions(obs[0], np.ones(n * 5, dtype=np.float32))
    assert a1.shape == (n,)
    assert not hasattr(pol, "comm") and pol.centralized and pol.vectorized and not pol.recurrent     # the sampler's switch (:133)
    ll = pol.log_likelihood(torch.from_numpy(obs).cuda()[None], torch.ones((1, B, n * 5), device="cuda"),
                            torch.from_numpy(acts).cuda()[None])
    assert ll.shape == (1, B)
    with pytest.raises(NotImplementedError):
        from com_marl_b200.policy import CentralizedCategoricalMLPPolicy
        from com_marl_b200.spaces import Box, Discrete, EnvSpec
        CentralizedCategoricalMLPPolicy(EnvSpec(Box(np.zeros(8), np.ones(8)), Discrete(5)), 2, hidden_sizes=(32, 32))


@pytest.mark.parametrize("scen", ["pp", "co"])
def test_cent_rollout_matches_oracle(scen):
    """device rollout with the CENT policy: every step replayed on the oracle with the kernel's own actions"""
    import sys
    sys.path.insert(0, GOLDEN)
    import ref_harness
    from com_marl_b200.rollout import RolloutEngine, make_policy
    from com_marl_b200.scenario import ScenarioSpec
    over = {"max_env_steps": 25}
    params = ref_harness.scenario_params(scen, 10, 1, 0.08 if scen == "pp" else 0.03, cap=2, loss=0.2, **over)
    spec = ScenarioSpec.from_params(scen, params, seed=9)
    B = 300
    pol = make_policy(spec, kind="cent")
    if scen == "pp":
        pol.math = "tc"            # env groups on parallel streams, each with its own first-layer scratch
    eng = RolloutEngine(spec, pol, B, ring=6, use_graph=True, groups=3)
    oenv = orc.OracleVecEnv(orc.spec_from_params(scen, params, seed=9), B)
    eng.reset(); oenv.reset()
    w = {k: v.cpu().numpy() for k, v in pol.state_dict().items()}
    for chunk in range(6):
        eng.run_chunk()
        t = {k: v.cpu().numpy() for k, v in eng.traj.items()}
        for k in range(eng.K):
            assert np.array_equal(t["obs"][k], oenv.obs)
            _, ref_probs = orc.policy_forward_cent(w, oenv.obs, None)
            assert np.abs(t["probs"][k] - ref_probs).max() <= 1e-5
            assert np.array_equal(t["actions"][k], oenv.sample_actions(t["probs"][k]))
            oenv.step(t["actions"][k])
            assert np.array_equal(t["reward"][k], oenv.reward) and np.array_equal(t["done"][k], oenv.done)
    eng.env.check_errors(); pol.check_errors()


def test_cent_error_codes():
    """C-ABI error behaviour of kind = CM_POLICY_CENT: math 0 / 1 only, math = 1 needs its operands, no attention output"""
    import ctypes as C
    from com_marl_b200 import _native as N
    n, D, B = 3, 29, 8
    pol = _policy(n, D)
    obs = torch.zeros((B, n * D), device="cuda")
    probs = torch.empty((B, n, 5), device="cuda")
    io = N.PolicyIO()
    io.n_envs, io.weights, io.obs, io.probs = B, N.ptr(pol.weight_blob()), N.ptr(obs), N.ptr(probs)
    for math, attn, expect in ((2, False, N.CM_EUNSUPPORTED), (1, False, N.CM_EINVAL), (0, True, N.CM_EINVAL), (0, False, N.CM_OK)):   # math = 1 without tc_weights / workspace
        desc = N.PolicyDesc(n, D, 1, 0, 1, math, 1, 0, N.POLICY_CENT, 0)
        io.attention = N.ptr(torch.empty((B, n, n), device="cuda")) if attn else None
        assert N.lib().cm_policy_forward(C.byref(desc), C.byref(io), N.stream_ptr()) == expect, (math, attn)
    torch.cuda.synchronize()
    assert torch.isfinite(probs).all()
    assert N.lib().cm_policy_cent_blob_floats(n, D) == n * D * 128 + 128 + 128 * 64 + 64 + 64 * 32 + 32 + 32 * 5 * n + 5 * n
