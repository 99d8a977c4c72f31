"""The hand-written forward / backward of the PPO update (cm_ppo_net, csrc/ppo_net_kernels.cu) against the torch autograd
graph of the same formula (com_marl_b200/policy.py::forward, ppo.py::CommBaseCritic — themselves held against the unmodified
reference by tests/test_ppo.py): losses, per-step outputs and EVERY parameter gradient, on random ragged batches with random
adjacency / channel masks and availability, for team sizes on both sides of every tiling boundary of the kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make(n, D, L, P, T, seed, residual=True, attention_type="general", sizes=None, ent_coeff=0.1):
    import torch
    from com_marl_b200.policy import CommCategoricalMLPPolicy
    from com_marl_b200.ppo import CommBaseCritic, DevicePPO
    from com_marl_b200.spaces import Box, Discrete, EnvSpec
    torch.manual_seed(seed)
    spec = EnvSpec(Box(np.zeros(n * D), np.ones(n * D)), Discrete(5))
    kw = {} if sizes is None else dict(encoder_hidden_sizes=sizes[0], embedding_dim=sizes[1], categorical_mlp_hidden_sizes=sizes[2])
    ckw = {} if sizes is None else dict(encoder_hidden_sizes=sizes[0], embedding_dim=sizes[1], decoder_hidden_sizes=sizes[3])
    mk = lambda: (CommCategoricalMLPPolicy(spec, n, n_gcn_layers=L, residual=residual, attention_type=attention_type, **kw),  # noqa: E731
                  CommBaseCritic(spec, n, n_gcn_layers=L, residual=residual, attention_type=attention_type, **ckw))
    pol_a, cri_a = mk()
    pol_f, cri_f = mk()
    pol_f.load_state_dict(pol_a.state_dict())
    cri_f.load_state_dict(cri_a.state_dict())
    with torch.no_grad():                       # biases away from their zero initialisation, log-std away from 0
        for m in (pol_a, cri_a):
            for k, p in m.named_parameters():
                if k.endswith("bias") or k.endswith("_init_std"):
                    p.add_(0.1 * torch.randn_like(p))
        pol_f.load_state_dict(pol_a.state_dict())
        cri_f.load_state_dict(cri_a.state_dict())
    auto = DevicePPO(pol_a, cri_a, fused=False, policy_ent_coeff=ent_coeff)
    fused = DevicePPO(pol_f, cri_f, fused=True, policy_ent_coeff=ent_coeff)
    g = torch.Generator(device="cuda").manual_seed(seed)
    rnd = lambda *s: torch.rand(*s, device="cuda", generator=g)  # noqa: E731
    valids = torch.randint(1, T + 1, (P,), device="cuda", generator=g).to(torch.int32)
    valids[0] = T
    adj = (rnd(P, T, n, n) < 0.6).float()
    adj = torch.maximum(adj, torch.eye(n, device="cuda"))
    chan = (rnd(P, T, L, n, n) < 0.7).float()
    chan[:, :, :, torch.arange(n), torch.arange(n)] = 1.0
    if n > 1:
        adj[0, 0, 0, :] = 0.0                    # a row with no neighbour at all (the + 1e-12 of the renormalisation)
    avail = (rnd(P, T, n, 5) < 0.8).float()
    avail[..., 4] = 1.0
    actions = torch.multinomial(avail.reshape(-1, 5), 1).reshape(P, T, n)
    b = dict(obs=rnd(P, T, n * D), avail=avail.reshape(P, T, n * 5), actions=actions, rewards=rnd(P, T).double(),
             dist_adjs=adj.reshape(P, T, n * n), channels=chan.reshape(P, T, L * n, n), valids=valids)
    t = torch.arange(T, device="cuda")[None, :] < valids[:, None]
    for k in ("obs", "rewards"):
        b[k] = b[k] * t.reshape(P, T, 1).to(b[k].dtype) if b[k].dim() == 3 else b[k] * t.to(b[k].dtype)
    return auto, fused, b


def _rel(a, ref):
    return float((a - ref).abs().max() / max(1e-12, float(ref.abs().max())))


@pytest.mark.parametrize("n,D,L,P,T", [(3, 29, 2, 7, 9), (4, 21, 2, 40, 11), (8, 21, 1, 5, 6), (9, 53, 3, 4, 5), (32, 53, 2, 5, 7),
                                       (33, 53, 2, 3, 4), (54, 77, 2, 3, 5), (64, 53, 2, 2, 3), (65, 53, 2, 2, 3), (130, 53, 2, 2, 2),
                                       (200, 53, 2, 2, 3), (225, 21, 2, 1, 2), (256, 21, 4, 1, 2)])
def test_fused_losses_and_gradients_equal_autograd(n, D, L, P, T):
    import torch
    auto, fused, b0 = _make(n, D, L, P, T, seed=n)
    ba = auto.finish_batch(dict(b0))
    bf = fused.finish_batch(dict(b0))
    assert _rel(bf["baselines"], ba["baselines"]) <= 2e-5
    F = fused._fused
    f = bf["_flat"]
    ids = np.arange(P)[::-1].copy()[: max(1, P - 1)]
    tid = torch.as_tensor(ids, device="cuda")
    # ---- policy: forward outputs, loss and gradient on a minibatch, with a perturbed "old" log-likelihood so that
    # ratios fall on both sides of the clip range
    with torch.no_grad():
        d = auto._dist(ba, None)
        ll_ref, ent_ref = d.log_prob(ba["actions"]).sum(-1), d.entropy().mean(-1)
        F.pol_map.refresh()
        out = F.policy_call(f, None, want_probs=True)
    assert _rel(out["ll"].reshape(P, T), ll_ref) <= 2e-5
    assert _rel(out["entropy"].reshape(P, T), ent_ref) <= 2e-5
    assert _rel(out["probs"].reshape(P, T, n, 5), d.probs) <= 2e-5
    old = ll_ref + 0.15 * torch.randn_like(ll_ref)
    auto.opt.zero_grad()
    loss_ref = auto.compute_loss(ba, tid, old)
    loss_ref.backward()
    idx = F.step_index(f, ids, valid_only=True)
    res = F.policy_call(f, idx, adv=ba["adv"].reshape(-1).index_select(0, idx), old_ll=old.reshape(-1).index_select(0, idx),
                        backward=True, inv_count=1.0 / idx.numel())
    g = torch.empty_like(fused.opt.grad)
    F.pol_map.scatter_grad(g)
    # (the joint log-likelihood is a sum over n agents: at n = 200 it is ~ -300, one fp32 ulp of it moves the ratio by 3e-5)
    tol = max(1.0, n / 32.0)
    assert abs(float(res["loss"]) - float(loss_ref.detach())) <= 2e-5 * tol * max(1.0, abs(float(loss_ref.detach())))
    gref = auto.opt.grad
    assert float(gref.abs().max()) > 0
    o = 0
    for k, p in auto.policy.named_parameters():
        e = _rel(g[o:o + p.numel()], gref[o:o + p.numel()])
        assert e <= 2e-4 * tol, ("policy", k, e)
        o += p.numel()
    assert _rel(g, gref) <= 1e-4 * tol
    # ---- critic: loss and gradient on the padded rows of the minibatch
    auto.baseline_opt.zero_grad()
    bl_ref = auto.baseline.compute_loss(ba["obs"][tid], ba["returns"][tid], ba["dist_adjs"][tid], ba["channels"][tid])
    bl_ref.backward()
    F.cri_map.refresh()
    idx_all = F.step_index(f, ids, valid_only=False)
    res = F.critic_call(f, idx_all, returns=ba["returns"].reshape(-1).index_select(0, idx_all), backward=True)
    g = torch.empty_like(fused.baseline_opt.grad)
    F.cri_map.scatter_grad(g)
    assert abs(float(res["loss"]) - float(bl_ref.detach())) <= 2e-5 * max(1.0, abs(float(bl_ref.detach())))
    gref = auto.baseline_opt.grad
    o = 0
    for k, p in auto.baseline.named_parameters():
        e = _rel(g[o:o + p.numel()], gref[o:o + p.numel()])
        assert e <= 2e-4, ("critic", k, e)
        o += p.numel()


@pytest.mark.parametrize("kw", [dict(residual=False), dict(attention_type="dot"), dict(sizes=((96,), 48, (100, 40, 24), (56,))),
                                dict(ent_coeff=0.0)])
def test_fused_variants(kw):
    """no residual connection, 'dot' attention (no W_a parameter), narrower layers (zero-padded into the kernel widths), no entropy"""
    import torch
    auto, fused, b0 = _make(5, 21, 2, 6, 8, seed=11, **kw)
    ba, bf = auto.finish_batch(dict(b0)), fused.finish_batch(dict(b0))
    assert _rel(bf["baselines"], ba["baselines"]) <= 2e-5
    F, f = fused._fused, bf["_flat"]
    with torch.no_grad():
        d = auto._dist(ba, None)
        old = d.log_prob(ba["actions"]).sum(-1) + 0.1 * torch.randn((6, 8), device="cuda")
    auto.opt.zero_grad()
    auto.compute_loss(ba, None, old).backward()
    F.pol_map.refresh()
    idx = F.step_index(f, np.arange(6), valid_only=True)
    F.policy_call(f, idx, adv=ba["adv"].reshape(-1).index_select(0, idx), old_ll=old.reshape(-1).index_select(0, idx), backward=True,
                  inv_count=1.0 / idx.numel())
    g = torch.empty_like(fused.opt.grad)
    F.pol_map.scatter_grad(g)
    assert _rel(g, auto.opt.grad) <= 1e-4
    auto.baseline_opt.zero_grad()
    auto.baseline.compute_loss(ba["obs"], ba["returns"], ba["dist_adjs"], ba["channels"]).backward()
    F.cri_map.refresh()
    F.critic_call(f, None, returns=ba["returns"].reshape(-1), backward=True)
    g = torch.empty_like(fused.baseline_opt.grad)
    F.cri_map.scatter_grad(g)
    assert _rel(g, auto.baseline_opt.grad) <= 1e-4


def test_fused_train_once_equals_autograd_train_once():
    """the whole optimisation loop on both paths from the same start: per-step losses, gradient norms, final weights"""
    import torch
    auto, fused, b0 = _make(6, 21, 2, 12, 10, seed=5)
    ids = np.random.RandomState(0).permutation(12)
    oa = auto.train_once(batch=auto.finish_batch(dict(b0)), shuffled_ids=ids)
    of = fused.train_once(batch=fused.finish_batch(dict(b0)), shuffled_ids=ids)
    for k in ("loss_before", "loss_after", "kl", "entropy"):
        assert abs(oa[k] - of[k]) <= 1e-4 * max(1.0, abs(oa[k])), k
    for k in ("losses", "baseline_losses", "grad_norms"):
        assert np.abs(np.array(oa[k]) - np.array(of[k])).max() <= 2e-4 * max(1.0, np.abs(np.array(oa[k])).max()), k
    for (k, pa), (_, pf) in zip(auto.policy.state_dict().items(), fused.policy.state_dict().items()):
        assert (pa - pf).abs().max().item() <= 2e-5, k
    for (k, pa), (_, pf) in zip(auto.baseline.state_dict().items(), fused.baseline.state_dict().items()):
        assert (pa - pf).abs().max().item() <= 2e-5, k


def test_chunked_walk_equals_single_chunk():
    """a workspace that holds only a few steps at a time gives the same gradient (chunks add up)"""
    import torch
    _, fused, b0 = _make(7, 21, 2, 5, 9, seed=3)
    bf = fused.finish_batch(dict(b0))
    F, f = fused._fused, bf["_flat"]
    F.pol_map.refresh()
    idx = F.step_index(f, np.arange(5), valid_only=True)
    adv = bf["adv"].reshape(-1).index_select(0, idx)
    F.policy_call(f, idx, adv=adv, backward=True, inv_count=1.0 / idx.numel())
    g1 = F.pol_map.grad.clone()
    F.chunk_steps, F._ws = 5, None
    F.policy_call(f, idx, adv=adv, backward=True, inv_count=1.0 / idx.numel())
    assert _rel(F.pol_map.grad, g1) <= 1e-5


@pytest.mark.parametrize("n,D", [(4, 21), (32, 53), (200, 53)])
def test_fused_obs_dp_policy_equals_autograd(n, D):
    """the Obs-DP runner family (DecCategoricalMLPPolicy + CommBaseCritic, runner_*_obsDP.py): cm_ppo_net kind = CM_NET_POLICY_DEC"""
    import torch
    from com_marl_b200.policy import DecCategoricalMLPPolicy
    from com_marl_b200.ppo import CommBaseCritic, DevicePPO
    from com_marl_b200.spaces import Box, Discrete, EnvSpec
    _, _, b0 = _make(n, D, 2, 3, 4, seed=n + 1)
    torch.manual_seed(n)
    spec = EnvSpec(Box(np.zeros(n * D), np.ones(n * D)), Discrete(5))
    pa, pf = DecCategoricalMLPPolicy(spec, n), DecCategoricalMLPPolicy(spec, n)
    ca, cf = CommBaseCritic(spec, n), CommBaseCritic(spec, n)
    with torch.no_grad():
        for p in pa.parameters():
            p.add_(0.05 * torch.randn_like(p))
    pf.load_state_dict(pa.state_dict())
    cf.load_state_dict(ca.state_dict())
    auto, fused = DevicePPO(pa, ca, fused=False), DevicePPO(pf, cf, fused=True)
    assert fused._fused is not None and fused._fused.dec
    ba, bf = auto.finish_batch(dict(b0)), fused.finish_batch(dict(b0))
    assert _rel(bf["baselines"], ba["baselines"]) <= 2e-5
    F, f = fused._fused, bf["_flat"]
    with torch.no_grad():
        d = auto._dist(ba, None)
        ll = d.log_prob(ba["actions"]).sum(-1)
        F.pol_map.refresh()
        out = F.policy_call(f, None, want_probs=True)
    assert _rel(out["ll"].reshape(3, 4), ll) <= 2e-5 and _rel(out["probs"].reshape(3, 4, n, 5), d.probs) <= 2e-5
    old = ll + 0.1 * torch.randn_like(ll)
    auto.opt.zero_grad()
    loss_ref = auto.compute_loss(ba, None, old)
    loss_ref.backward()
    idx = F.step_index(f, np.arange(3), valid_only=True)
    res = F.policy_call(f, idx, adv=ba["adv"].reshape(-1).index_select(0, idx), old_ll=old.reshape(-1).index_select(0, idx), backward=True,
                        inv_count=1.0 / idx.numel())
    g = torch.empty_like(fused.opt.grad)
    F.pol_map.scatter_grad(g)
    tol = max(1.0, n / 32.0)
    assert abs(float(res["loss"]) - float(loss_ref.detach())) <= 2e-5 * tol * max(1.0, abs(float(loss_ref.detach())))
    assert _rel(g, auto.opt.grad) <= 1e-4 * tol
    oa = auto.train_once(batch=ba, shuffled_ids=np.arange(3))
    of = fused.train_once(batch=bf, shuffled_ids=np.arange(3))
    for k in ("loss_before", "loss_after", "kl"):
        assert abs(oa[k] - of[k]) <= 2e-4 * tol * max(1.0, abs(oa[k])), k
