"""Where the PPO update's GPU time goes: one DeviceTrainer round under torch.profiler, kernels grouped by name.
    python tools/ppo_profile.py --config c5 --envs 32 --mini-epochs 2"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c1")
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--mini-epochs", type=int, default=2)
    ap.add_argument("--fused", default="auto", choices=["auto", "0", "1"])
    args = ap.parse_args()
    import bench
    from com_marl_b200.scenario import ScenarioSpec
    from com_marl_b200.train import DeviceTrainer
    scen, params = bench.params_for(args.config)
    spec = ScenarioSpec.from_params(scen, params, seed=1)
    tr = DeviceTrainer(spec, args.envs, optimization_mini_epochs=args.mini_epochs, fused={"auto": "auto", "0": False, "1": True}[args.fused])
    tr.train_epoch()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        out = tr.train_epoch()
        torch.cuda.synchronize()
    print(f"{args.config} envs {args.envs}: rollout {out['rollout_ms']:.1f} ms, batch {out['batch_ms']:.1f} ms, update {out['update_ms']:.1f} ms "
          f"({len(out['losses'])} optimizer steps, {out['n_paths']} paths)")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))


if __name__ == "__main__":
    main()
