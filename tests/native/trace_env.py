import sys, torch, numpy as np
sys.path.insert(0, '.')
from com_marl_b200.scenario import ScenarioSpec
from com_marl_b200.envs import BatchedEnv
cfg = sys.argv[1] if len(sys.argv) > 1 else 'c2'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
spec = {'c2': ScenarioSpec.from_cli('co',10,1,0.03), 'c1': ScenarioSpec.from_cli('pp',10,1,0.04,cap=2), 'c3': ScenarioSpec.from_cli('pp',20,2,0.08,cap=4,loss=0.2)}[cfg]
env = BatchedEnv(spec, B); env.reset()
acts = torch.randint(0, 5, (B, spec.n_agents), dtype=torch.int8, device='cuda')
for _ in range(5): env.step(acts)
torch.cuda.synchronize()
t = env.stats.cpu().numpy().view(np.int64)[0, :11]
names = ['start','wall sync','scalars loaded','staged+bitmaps','actions','step logic','outputs+stats','(reset)','state written','obs written','comm done']
for k in range(1, 11): print(f"{names[k]:18s} +{t[k]-t[k-1]:7d}   total {t[k]-t[0]:7d}")
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): env.step(acts)
e1.record(); torch.cuda.synchronize()
print(cfg, B, 'env ms', e0.elapsed_time(e1)/50)
