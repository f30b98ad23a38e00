"""Full-size runs of BASELINE.json configs[2..4] (C3, C4, C5), one process per GPU (plain python for N = 1,
torchrun for N > 1); the work itself lives in gapit_b200/fullsize.py (bench.py runs the same legs).

    python tools/run_config.py --config c3            # 20,000 taxa x 1,000,000 SNPs: kinship + eigh + REML + MLM scan
    python tools/run_config.py --config c4            # 50,000 taxa x 1,000,000 SNPs: VanRaden kinship only
    python tools/run_config.py --config c5            # 10,000 taxa x 2,000,000 SNPs x 100 traits, one eigendecomposition
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_config.py --config c3
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import gapit_b200 as gb
from gapit_b200 import fullsize


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=sorted(fullsize.CONFIGS))
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--m", type=int, default=0)
    ap.add_argument("--traits", type=int, default=-1)
    ap.add_argument("--blocks", type=int, default=8, help="marker blocks (the 8-GPU layout's shards)")
    ap.add_argument("--slices", type=int, default=0)
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = gb.Context(local, slices=args.slices or None)
    stream = torch.cuda.Stream(device=dev)      # one non-default stream for torch, NCCL and libgapitcuda alike
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    out = fullsize.run_config(args.config, ctx, dev, rank, world, n=args.n or None, m=args.m or None,
                              T=None if args.traits < 0 else args.traits, blocks=args.blocks, slices=ctx.slices,
                              exchange=args.exchange)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
