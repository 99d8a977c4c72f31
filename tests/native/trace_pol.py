import sys, torch, numpy as np, ctypes as C
sys.path.insert(0, '.')
from com_marl_b200.scenario import ScenarioSpec
from com_marl_b200.rollout import make_policy
from com_marl_b200.envs import BatchedEnv
from com_marl_b200 import _native as N
cfg = sys.argv[1] if len(sys.argv) > 1 else 'c2'
spec = {'c2': ScenarioSpec.from_cli('co',10,1,0.03), 'c3': ScenarioSpec.from_cli('pp',20,2,0.08,cap=4,loss=0.2)}[cfg]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
env = BatchedEnv(spec, B); env.reset()
pol = make_policy(spec)
n=spec.n_agents
probs=torch.empty((B,n,5),device='cuda'); acts=torch.empty((B,n),dtype=torch.int8,device='cuda')
pol._workspace = torch.zeros(8192, dtype=torch.float32, device='cuda')
orig = pol.act_device
def run():
    n_, D, L = pol._n_agents, pol._dec_obs_dim, pol.n_gcn_layers
    desc = N.PolicyDesc(n_, D, L, 1, 0, 1, pol.seed, 0)
    io = N.PolicyIO(); io.n_envs = B; io.weights = N.ptr(pol.weight_blob()); io.tc_weights = N.ptr(pol.tc_weight_blob())
    io.obs=N.ptr(env.obs); io.adj_bits=N.ptr(env.adj_bits); io.chan_bits=N.ptr(env.chan_bits); io.tick=N.ptr(env.tick); io.episode=N.ptr(env.episode)
    io.probs=N.ptr(probs); io.actions=N.ptr(acts); io.workspace=N.ptr(pol._workspace); io.workspace_bytes=32768; io.error_flag=N.ptr(pol._tc_error)
    N.check('f', N.lib().cm_policy_forward(C.byref(desc), C.byref(io), N.stream_ptr()))
for _ in range(3): run()
torch.cuda.synchronize()
t = pol._workspace.cpu().numpy().view(np.int64)[:512].reshape(-1,8)[:48]
prev3 = None
print("stage  epi  sync  wwait0 issue0 wwait1 issue1+commit  mmawait  total")
for i,(a,b,c,d,w0,w1,_,_) in enumerate(t):
    if a == 0: continue
    epi = (a - prev3) if prev3 is not None else 0
    if w1: print(f"{i:3d} {epi:6d} {b-a:6d} {w0-b:6d} {'':6s} {w1-w0:6d}(wait1+issue0) {c-w1:6d} {d-c:6d} {d-(prev3 if prev3 else a):7d}")
    else: print(f"{i:3d} {epi:6d} {b-a:6d} {w0-b:6d} {c-w0:6d} {'':6s} {'':6s} {d-c:6d} {d-(prev3 if prev3 else a):7d}")
    prev3 = d
x = pol._workspace.cpu().numpy().view(np.int64)[600:624]
names = {20:'kernel entry',21:'kernel exit',0:'scores start',1:'softmax done',2:'HW0->KV + barrier',3:'aggr l0',4:'epi l0',5:'mma HW1',6:'HW1->KV',7:'aggr l1',8:'epi l1',16:'head done',17:'final done',18:'tile start',19:'obs staged'}
base = x[18]
for k in sorted(names):
    if x[k]: print(f"{names[k]:20s} {x[k]-base:8d}")

e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(32): run()
g.replay(); torch.cuda.synchronize()
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print('graph of 32 launches: us per launch', e0.elapsed_time(e1)/32*1e3)

gt = pol._workspace.cpu().numpy().view(np.int64)[1024:1024+2*296].reshape(-1,2)
gt = gt[gt[:,0] > 0]
t0 = gt[:,0].min()
st, en = (gt[:,0]-t0)/1e3, (gt[:,1]-t0)/1e3
print('CTAs', len(gt), 'start us: min %.1f med %.1f max %.1f' % (st.min(), np.median(st), st.max()), ' end us: min %.1f med %.1f max %.1f' % (en.min(), np.median(en), en.max()), ' dur med %.1f max %.1f' % (np.median(en-st), (en-st).max()))
print('start by block idx (every 37th):', [(int(i), round(float(st[i]),1), round(float(en[i]),1)) for i in range(0, len(gt), 37)])
