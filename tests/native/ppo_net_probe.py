"""one forward + backward of the policy on cm_ppo_net at a given shape (for ncu / timing):  python tests/native/ppo_net_probe.py n D steps [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from com_marl_b200.policy import CommCategoricalMLPPolicy  # noqa: E402
from com_marl_b200.ppo import CommBaseCritic, DevicePPO  # noqa: E402
from com_marl_b200.spaces import Box, Discrete, EnvSpec  # noqa: E402

n, D, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
spec = EnvSpec(Box(np.zeros(n * D), np.ones(n * D)), Discrete(5))
pol, cri = CommCategoricalMLPPolicy(spec, n), CommBaseCritic(spec, n)
algo = DevicePPO(pol, cri, fused=True)
F = algo._fused
W = (n + 31) // 32
g = torch.Generator(device="cuda").manual_seed(0)
f = dict(P=1, T=S, obs=torch.rand((S, n, D), device="cuda", generator=g), actions=torch.randint(0, 5, (S, n), device="cuda", generator=g),
         adj=torch.randint(-2**31, 2**31 - 1, (S, n, W), device="cuda", generator=g, dtype=torch.int64).to(torch.int32),
         chan=torch.randint(-2**31, 2**31 - 1, (S, 2, n, W), device="cuda", generator=g, dtype=torch.int64).to(torch.int32),
         avail=torch.full((S, n), 31, dtype=torch.uint8, device="cuda"), valid=torch.ones(S, dtype=torch.uint8, device="cuda"))
adv = torch.randn(S, device="cuda", generator=g)
F.pol_map.refresh()
F.cri_map.refresh()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for r in range(reps):
    ev[0].record()
    F.policy_call(f, None, adv=adv, backward=True, inv_count=1.0 / S)
    ev[1].record()
    F.critic_call(f, None, returns=adv, backward=True)
    ev[2].record()
    F.policy_call(f, None)
    ev[3].record()
    torch.cuda.synchronize()
    print(f"n {n} D {D} steps {S} rows {S * n}: policy fwd+bwd {ev[0].elapsed_time(ev[1]):.3f} ms, critic fwd+bwd {ev[1].elapsed_time(ev[2]):.3f} ms, "
          f"policy fwd {ev[2].elapsed_time(ev[3]):.3f} ms")
