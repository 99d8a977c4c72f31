"""Kernel time and device-rollout throughput of the three policy architectures of the reference README (Comm-DP, Obs-DP,
CENT) on one config:   python tools/policy_kinds_bench.py c2 [--envs N] [--kinds comm,dec,cent]
Prints one JSON line per kind: policy ms per launch over the whole batch (CUDA events around 50 launches on the launching
stream, after warm-up) and agent-steps/s of the CUDA-graph rollout loop (RolloutEngine, like bench.py's `value`)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from com_marl_b200.rollout import RolloutEngine, make_policy  # noqa: E402
from com_marl_b200.scenario import ScenarioSpec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default="c2")
    ap.add_argument("--envs", type=int, default=0)
    ap.add_argument("--kinds", default="comm,dec,cent")
    ap.add_argument("--steps", type=int, default=256)
    a = ap.parse_args()
    scen, params = bench.params_for(a.config)
    spec = ScenarioSpec.from_params(scen, params, seed=1)
    B = a.envs or bench.CONFIGS[a.config][6]
    n, D = spec.n_agents, spec.obs_dim
    for kind in a.kinds.split(","):
        pol = make_policy(spec, kind=kind)
        groups = 4 if (n <= 64 or kind != "comm") else 1
        eng = RolloutEngine(spec, pol, B, ring=32, use_graph=True, groups=groups)
        eng.reset()
        eng.run(96)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run(a.steps); e1.record(); torch.cuda.synchronize()
        ms_step = e0.elapsed_time(e1) / a.steps
        env = eng.env
        probs = torch.empty((B, n, 5), device="cuda"); acts = torch.empty((B, n), dtype=torch.int8, device="cuda")
        call = lambda: pol.act_device(env.obs, adj_bits=env.adj_bits, chan_bits=env.chan_bits, tick=env.tick,  # noqa: E731
                                      episode=env.episode, probs=probs, actions=acts)
        for _ in range(5):
            call()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            call()
        e1.record(); torch.cuda.synchronize()
        ms_pol = e0.elapsed_time(e1) / 50
        pol.check_errors(); eng.env.check_errors()
        flop = 2 * (n * D * 128 + 128 * 64 + 64 * 32 + 32 * 5 * n) / n if kind == "cent" else None
        print(json.dumps({"config": a.config, "kind": kind, "envs": B, "n_agents": n, "obs_dim": D,
                          "policy_ms_per_launch_whole_batch": round(ms_pol, 5), "rollout_ms_per_step": round(ms_step, 5),
                          "agent_steps_per_s": B * n / ms_step * 1e3,
                          "math": getattr(pol, "math", None), "cent_flop_per_agent": flop,
                          "cent_tflops": None if flop is None else flop * B * n / ms_pol / 1e9,
                          "obs_read_GBps": B * n * D * 4 / ms_pol / 1e6}), flush=True)
        del eng, pol
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
