"""Multi-GPU inference check (run under torchrun, NCCL): `sharding.separate_sharded` — contiguous batch shards, no
data-path collective, NCCL all-gather of the separated waveforms — must reproduce the single-GPU forward of the whole
batch bit for bit, including an uneven split (B = 5 over the ranks) and ragged-length balancing.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tools/sharded_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse_b200  # noqa: E402,F401
from cse_b200 import sharding, synth  # noqa: E402
from cse_b200.models.ContSep import Sepformer  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = f"cuda:{local}"
    m = Sepformer(2, add_mt=True)
    m.add_mt_pipeline()
    m.load_state_dict(synth.make_state_dict("contsep", 2, seed=3))
    m = m.to(dev).eval()
    for B in (5, 2 * world):
        mix, _ = synth.make_mixture(B, 8000, 2, seed=40 + B)
        ctx = synth.make_context(B, 1, seed=40 + B)
        mix, ctx = mix.to(dev), ctx.to(dev)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            whole = m(mix, ctx)[0]
            got = sharding.separate_sharded(lambda a, c: m(a, c)[0], mix, ctx)
        assert got.shape == whole.shape, (got.shape, whole.shape)
        assert torch.equal(got, whole), f"rank {rank}: sharded result differs (B={B}): {(got - whole).abs().max().item()}"
    shards = sharding.balanced_shards([64000, 12000, 30000, 50000, 20000, 40000, 8000], world)
    assert sorted(sum(shards, [])) == list(range(7))
    dist.barrier()
    if rank == 0:
        print(f"sharded_check: OK on {world} GPUs (NCCL all-gather of the separated waveforms == single-GPU forward)")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
