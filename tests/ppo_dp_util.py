"""World-size-2 harness for the data-parallel PPO update (com_marl_b200.ppo.DevicePPO.train_once): rank r of `world`
ranks updates on its slice of a synthetic padded batch; the result must equal the single-process update on the union
batch (one minibatch per epoch: the union's minibatch is then the union of the ranks' minibatches), the ranks must
end with identical weights, and unequal path counts (different numbers of local slices under 3 minibatches) must not
deadlock.  Used on CPU with gloo (tests/test_host_cpu.py; the Adam kernel is replaced by the same formula in torch
there) and on the GPU with the real kernels (tests/test_ppo.py, two processes sharing cuda:0 over gloo)."""
import math
import os

import numpy as np
import torch


def torch_adam_step(self, grad_scale=1.0):
    """the formula of cm_adam_step (csrc/ppo_kernels.cu; my_optimizer/adam.py:57-120) in torch ops, for CPU runs"""
    self.steps += 1
    g = self.grad * float(grad_scale)
    b1, b2 = self.betas
    self.exp_avg.mul_(b1).add_(g, alpha=1 - b1)
    self.exp_avg_sq.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** self.steps, 1 - b2 ** self.steps
    denom = (self.exp_avg_sq.sqrt() / math.sqrt(bc2)).add_(self.eps)
    self.flat.addcdiv_(self.exp_avg, denom, value=-self.lr / bc1)


def synthetic_batch(P, T, n, D, L, device, seed):
    """padded PPO batch [P, T, ...] with ragged valid lengths, random masks / advantages / returns"""
    g = torch.Generator().manual_seed(seed)
    valids = torch.randint(max(1, T // 2), T + 1, (P,), generator=g).to(torch.int32)
    valids[0] = T
    mask = torch.arange(T)[None, :] < valids[:, None]
    adj = (torch.rand((P, T, n, n), generator=g) < 0.7).float()
    adj = torch.maximum(adj, torch.eye(n)[None, None])
    chan = (torch.rand((P, T, L, n, n), generator=g) < 0.8).float()
    chan = torch.maximum(chan, torch.eye(n)[None, None, None])
    b = dict(obs=torch.rand((P, T, n * D), generator=g), avail=torch.ones((P, T, n * 5)),
             actions=torch.randint(0, 5, (P, T, n), generator=g), rewards=torch.randn((P, T), generator=g).double(),
             dist_adjs=adj.reshape(P, T, n * n), channels=chan.reshape(P, T, L * n, n), valids=valids,
             returns=torch.randn((P, T), generator=g), adv=torch.randn((P, T), generator=g), mask=mask)
    for k in ("obs", "dist_adjs", "channels", "rewards", "returns", "adv"):     # zero padding like process_samples
        pad = ~mask
        b[k][pad] = 1.0 if k in ("dist_adjs", "channels") else 0.0
    return {k: v.to(device) for k, v in b.items()}


def take(b, lo, hi):
    return {k: v[lo:hi] for k, v in b.items()}


def build(n, D, L, device, n_minibatches, mini_epochs=2):
    from com_marl_b200.policy import CommCategoricalMLPPolicy
    from com_marl_b200.ppo import CommBaseCritic, DevicePPO, FlatAdam
    from com_marl_b200.spaces import Box, Discrete, EnvSpec
    if torch.device(device).type == "cpu":
        FlatAdam.step = torch_adam_step
    torch.manual_seed(11)
    spec = EnvSpec(Box(np.zeros(n * D), np.ones(n * D)), Discrete(5))
    pol = CommCategoricalMLPPolicy(spec, n, n_gcn_layers=L, device=device)
    cri = CommBaseCritic(spec, n, n_gcn_layers=L, device=device)
    algo = DevicePPO(pol, cri, optimization_n_minibatches=n_minibatches, optimization_mini_epochs=mini_epochs)
    return pol, cri, algo


def flat_weights(pol, cri):
    return torch.cat([p.detach().reshape(-1).double().cpu() for m in (pol, cri) for p in m.parameters()]).numpy()


def worker(rank, world, port, q, device, cuts):
    """cuts = path index boundaries of the ranks inside the union batch, e.g. (0, 4, 9)"""
    os.environ.update(RANK=str(rank), LOCAL_RANK="0", WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    torch.set_num_threads(1)
    if world > 1:
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    n, D, L, T = 3, 7, 2, 6
    union = synthetic_batch(cuts[-1], T, n, D, L, device, seed=5)
    out = {}
    # (1) one minibatch per epoch: the ranks' weighted all-reduce reproduces the union update
    pol, cri, algo = build(n, D, L, device, n_minibatches=1)
    mine = take(union, cuts[rank], cuts[rank + 1]) if world > 1 else union
    r = algo.train_once(batch=mine, shuffled_ids=np.arange(mine["rewards"].shape[0]))
    out["union_weights"] = flat_weights(pol, cri)
    out["union_gnorms"] = r["grad_norms"]
    # (2) three minibatches with unequal path counts (4 paths -> 2 local slices, 5 paths -> 3): no deadlock, equal weights
    pol, cri, algo = build(n, D, L, device, n_minibatches=3)
    r = algo.train_once(batch=mine, shuffled_ids=np.arange(mine["rewards"].shape[0]))
    out["mb3_weights"] = flat_weights(pol, cri)
    out["mb3_steps"] = len(r["losses"])
    q.put((rank, out))
    if world > 1:
        dist.destroy_process_group()


def run(device, timeout=240):
    """-> (single-process result, {rank: result}) for the 4 | 5 split of 9 paths"""
    import socket
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    cuts = (0, 4, 9)

    def launch(world):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        q = ctx.Queue()
        procs = [ctx.Process(target=worker, args=(r, world, port, q, device, cuts)) for r in range(world)]
        for p in procs:
            p.start()
        res = dict(q.get(timeout=timeout) for _ in range(world))
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        return res

    return launch(1)[0], launch(2)


def check(single, ranks):
    a, b = ranks[0], ranks[1]
    # identical weights on both ranks, bit for bit (same all-reduced gradients, same Adam)
    assert np.array_equal(a["union_weights"], b["union_weights"]) and np.array_equal(a["mb3_weights"], b["mb3_weights"])
    # 2-rank update on the 4 | 5 split == 1-rank update on the union of 9 paths
    ref = single["union_weights"]
    moved = np.abs(ref - a["union_weights"]).max()
    assert moved <= 2e-6, moved
    assert np.allclose(single["union_gnorms"], a["union_gnorms"], rtol=1e-4, atol=1e-6)
    # unequal slice counts: both ranks ran max(2, 3) = 3 optimizer steps per epoch
    assert a["mb3_steps"] == b["mb3_steps"] == 2 * 3 and single["mb3_steps"] == 2 * 3
