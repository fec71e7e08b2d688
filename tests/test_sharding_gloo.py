"""CPU, world_size 2, gloo: the multi-GPU path is pure sharding (tiles / frames dealt to ranks, no data-path
collective); what needs a rendezvous is only the timing protocol of bench.py (barrier + max over ranks) and the
host-side gather of per-rank results.  This test runs both with the oracle standing in for the device."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from grokimagecompression_b200 import params as P
    from grokimagecompression_b200.synth import synthetic_planes
    import oracle_pipeline as OP
    width, height, tile = 256, 128, (64, 64)
    img = synthetic_planes(width, height, 3, 8, seed=11)
    tiles = P.image_tiles(width, height, 3, 8, True, tile, 4)
    planes = P.split_planes(img, width, height, tile)
    mine = [t for t in range(len(tiles)) if t % world == rank]          # tile t -> rank t mod G
    blocks, _ = OP.encode_tiles([tiles[t] for t in mine], [planes[3 * t + c] for t in mine for c in range(3)])
    nbytes = sum(len(b["data"]) for b in blocks)
    # timing protocol: max over ranks of a per-rank duration
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # host gathers the per-rank code-block payloads in tile order
    gathered = [None] * world
    dist.all_gather_object(gathered, [(mine[b["tileno"]], b["compno"], b["resno"], b["orient"], b["x0"], b["y0"], b["data"]) for b in blocks])
    if rank == 0:
        q.put((float(t.item()), nbytes, sorted(x for g in gathered for x in g)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from grokimagecompression_b200 import params as P
    from grokimagecompression_b200.synth import synthetic_planes
    import oracle_pipeline as OP
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tmax, nbytes0, merged = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    width, height, tile = 256, 128, (64, 64)
    img = synthetic_planes(width, height, 3, 8, seed=11)
    tiles = P.image_tiles(width, height, 3, 8, True, tile, 4)
    blocks, _ = OP.encode_tiles(tiles, P.split_planes(img, width, height, tile))
    single = sorted((b["tileno"], b["compno"], b["resno"], b["orient"], b["x0"], b["y0"], b["data"]) for b in blocks)
    assert merged == single
