#!/usr/bin/env python3
"""One search sharded over the GPUs of this box (run under torchrun, one process per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/dist_check.py <golden-case-name | --args ...>

Every rank joins the NCCL communicator through the C ABI (ries_dist_init), calls ries_main with the same
arguments, and checks its report against the committed golden text (byte-identical, `Time:` normalised).
Rank 0 prints one summary line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    os.environ["RIES_DEVICE"] = str(local)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import ries_b200
    from conftest import golden_cases, golden_text, normalise
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(ries_b200.dist_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ries_b200.dist_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
    names = sys.argv[1:] or ["l5"]
    ok = True
    report = []
    for name in names:
        case = [c for c in golden_cases() if c["name"] == name][0]
        args = case["args"] + ["--no-slow-messages"]
        ries_b200.call_main(args)  # warm-up (graphs, arenas)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        code, text = ries_b200.call_main(args)
        dt = time.perf_counter() - t0
        st = ries_b200.last_stats()
        shown = code & 0xff if code != 0 else 0
        same = normalise(text, "ries") + "exit=%d\n" % shown == golden_text(name)
        flag = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = ok and bool(flag.item())
        report.append({"case": name, "identical_on_every_rank": bool(flag.item()), "seconds": round(dt, 4),
                       "sharded_passes": st["sharded_passes"], "gather_MB": round(st["gather_bytes"] / 1e6, 1), "gather_ms": round(st["gather_ms"], 2),
                       "stage_ms": [round(x, 2) for x in st["stage_ms"]]})
    ries_b200.dist_finalize()
    if rank == 0:
        print(json.dumps({"world": world, "ok": ok, "cases": report}))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
