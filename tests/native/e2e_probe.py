"""Where does the e2e (host-buffer) step go?  copy bandwidths by size, host enqueue time vs device time per phase."""
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from com_marl_b200.scenario import ScenarioSpec
from com_marl_b200.rollout import make_policy
from com_marl_b200.envs import BatchedEnv
dev = torch.device('cuda', 0)
for mb in (0.05, 0.4, 1.0, 5.7, 23.0):
    nb = int(mb * 1e6)
    h = torch.empty(nb, dtype=torch.uint8, pin_memory=True); d = torch.empty(nb, dtype=torch.uint8, device=dev)
    for name, (dst, src) in (("h2d", (d, h)), ("d2h", (h, d))):
        for _ in range(3): dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): dst.copy_(src, non_blocking=True)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        t0 = time.perf_counter()
        for _ in range(20): dst.copy_(src, non_blocking=True)
        t1 = time.perf_counter(); torch.cuda.synchronize()
        print(f"{name} {mb:5.2f} MB: {ms*1e3:7.1f} us  {nb/ms/1e6:6.1f} GB/s   host enqueue {1e6*(t1-t0)/20:5.1f} us")
spec = ScenarioSpec.from_cli('co', 10, 1, 0.03)
B = 16384
env = BatchedEnv(spec, B); pol = make_policy(spec)
out = env.reset_host()
def it(out, prof=None):
    pin = out["pinned"]
    t0 = time.perf_counter()
    a, p, ev = pol.get_actions_host(pin["obs"], pin["adj_bits"], pin["chan_bits"], inputs_arena=True, sync=False)
    t1 = time.perf_counter(); ev.synchronize(); t2 = time.perf_counter()
    o, ev = env.step_host(a, sync=False)
    t3 = time.perf_counter(); ev.synchronize(); t4 = time.perf_counter()
    if prof is not None: prof.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
    return o
for _ in range(5): out = it(out)
prof = []
for _ in range(100): out = it(out, prof)
m = np.array(prof).mean(0) * 1e6
print("policy phase: enqueue %.0f us, wait %.0f us | env phase: enqueue %.0f us, wait %.0f us | total %.0f us" % (*m, m.sum()))
