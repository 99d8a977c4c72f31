"""Timeline of the two-half e2e pipeline: device start/end of every phase (CUDA events) and host enqueue times."""
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from com_marl_b200.scenario import ScenarioSpec
from com_marl_b200.rollout import make_policy
from com_marl_b200.envs import BatchedEnv
dev = torch.device('cuda', 0)
spec = ScenarioSpec.from_cli('co', 10, 1, 0.03)
B, NB = 16384, 2
Bh = B // NB
envs = [BatchedEnv(spec, Bh, env_id0=h * Bh) for h in range(NB)]
pol = make_policy(spec)
outs = [e.reset_host() for e in envs]
streams = [torch.cuda.Stream(dev) for _ in range(NB)]
torch.cuda.synchronize()
E = lambda: torch.cuda.Event(enable_timing=True)
def pol_phase(h, rec=None):
    pin = outs[h]["pinned"]
    if rec is not None: s = E(); s.record(streams[h])
    a, _, ev = pol.get_actions_host(pin["obs"], pin["adj_bits"], pin["chan_bits"], inputs_arena=True, slot=h, sync=False, stream=streams[h])
    if rec is not None: e = E(); e.record(streams[h]); rec.append(("pol%d" % h, s, e, time.perf_counter()))
    return a, ev
def env_phase(h, acts, rec=None):
    if rec is not None: s = E(); s.record(streams[h])
    o, ev = envs[h].step_host(acts, sync=False, stream=streams[h])
    if rec is not None: e = E(); e.record(streams[h]); rec.append(("env%d" % h, s, e, time.perf_counter()))
    return o, ev
a1, ev = pol_phase(1); ev.synchronize()
def iteration(rec=None):
    global a1
    a0, e0 = pol_phase(0, rec); o1, e1 = env_phase(1, a1, rec)
    e0.synchronize(); e1.synchronize(); outs[1] = o1
    if rec is not None: rec.append(("sync", None, None, time.perf_counter()))
    o0, e0 = env_phase(0, a0, rec); a1, e1 = pol_phase(1, rec)
    e0.synchronize(); e1.synchronize(); outs[0] = o0
    if rec is not None: rec.append(("sync", None, None, time.perf_counter()))
for _ in range(10): iteration()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100): iteration()
torch.cuda.synchronize()
print("iteration: %.0f us" % ((time.perf_counter() - t0) / 100 * 1e6))
base = E(); base.record(); torch.cuda.synchronize()
rec = []; th = time.perf_counter()
base.record(streams[0]); 
for _ in range(3): iteration(rec)
torch.cuda.synchronize()
for name, s, e, t in rec[-6:] if False else rec:
    if s is None: print("  host %7.0f us  sync done" % ((t - th) * 1e6))
    else: print("  host %7.0f us  %s  device %7.0f -> %7.0f us" % ((t - th) * 1e6, name, base.elapsed_time(s) * 1e3, base.elapsed_time(e) * 1e3))
