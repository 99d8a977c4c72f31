"""Device -> host copy bandwidth per GPU, alone and with every rank copying at once (torchrun): what bounds the e2e number
when several ranks stream their step results into host memory.  Prints one JSON line (rank 0).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/pcie_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from com_marl_b200 import distributed as D
    rank, local_rank, world = D.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    nbytes = 512 << 20
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    hsrc = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    ddst = torch.empty(nbytes, dtype=torch.uint8, device=dev)

    def run(fn, reps=8):
        fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9

    def barrier():
        if world > 1:
            dist.barrier()

    out = {}
    for name, fn in (("d2h", lambda: dst.copy_(src, non_blocking=True)), ("h2d", lambda: ddst.copy_(hsrc, non_blocking=True))):
        alone = []
        for r in range(world):                    # one rank at a time
            barrier()
            if r == rank:
                alone.append(run(fn))
            barrier()
        barrier()
        together = run(fn, reps=16)               # every rank at once
        barrier()
        vals = torch.tensor([alone[0], together], dtype=torch.float64, device=dev)
        if world > 1:
            lst = [torch.zeros_like(vals) for _ in range(world)]
            dist.all_gather(lst, vals)
        else:
            lst = [vals]
        out[name] = {"alone_GBps_per_rank": [round(float(v[0]), 1) for v in lst], "all_ranks_at_once_GBps_per_rank": [round(float(v[1]), 1) for v in lst],
                     "sum_at_once_GBps": round(float(sum(v[1] for v in lst)), 1)}
    # both directions of ONE GPU at once, each on a copy stream of its own (rank 0; the others wait): does a step that uploads
    # while the previous one downloads get the sum of the two directions?
    barrier()
    if rank == 0:
        s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        reps = 8

        def both():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            e0.record()
            s_up.wait_event(e0); s_down.wait_event(e0)
            for _ in range(reps):
                with torch.cuda.stream(s_up):
                    ddst.copy_(hsrc, non_blocking=True)
                with torch.cuda.stream(s_down):
                    dst.copy_(src, non_blocking=True)
            for st in (s_up, s_down):
                j = torch.cuda.Event(); j.record(st); torch.cuda.current_stream(dev).wait_event(j)
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) * 1e-3

        both()
        t = both()
        out["h2d_and_d2h_one_gpu_two_streams"] = {"GBps_each_direction": round(reps * nbytes / t / 1e9, 1), "GBps_sum": round(2 * reps * nbytes / t / 1e9, 1)}
    barrier()
    if rank == 0:
        print(json.dumps({"n_gpus": world, "transfer_bytes": nbytes, **out}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
