import sys, torch, numpy as np
sys.path.insert(0, '.')
from com_marl_b200.scenario import ScenarioSpec
from com_marl_b200.rollout import make_policy
from com_marl_b200.envs import BatchedEnv
cfg = sys.argv[1] if len(sys.argv) > 1 else 'c2'
spec = {'c2': ScenarioSpec.from_cli('co',10,1,0.03), 'c3': ScenarioSpec.from_cli('pp',20,2,0.08,cap=4,loss=0.2),
        'c5': ScenarioSpec.from_cli('pp',50,2,0.08,cap=4)}[cfg]
B = {'c2':16384,'c3':16384,'c5':2048}[cfg]
env = BatchedEnv(spec, B); env.reset()
pol = make_policy(spec)
n=spec.n_agents
probs=torch.empty((B,n,5),device='cuda'); acts=torch.empty((B,n),dtype=torch.int8,device='cuda')
for _ in range(5): pol.act_device(env.obs, env.adj_bits, env.chan_bits, tick=env.tick, episode=env.episode, probs=probs, actions=acts)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): pol.act_device(env.obs, env.adj_bits, env.chan_bits, tick=env.tick, episode=env.episode, probs=probs, actions=acts)
e1.record(); torch.cuda.synchronize()
print(cfg, 'policy ms', e0.elapsed_time(e1)/50)
