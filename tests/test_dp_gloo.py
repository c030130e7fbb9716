"""World-size-2 CPU (gloo) test of the data-parallel host logic: sample sharding, flat gradient buckets, the
asynchronous all-reduce and the 1/N scaling that the fused Adam kernel applies.  No CUDA kernels run (DRY_RUN); the
arithmetic of the exchange is checked with torch CPU ops on the buckets themselves."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from objgan_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from objgan_b200 import lib, model, trainer
    lib.DRY_RUN = True
    model.FAST_INIT = True                             # values are never looked at in DRY_RUN
    torch.manual_seed(100 + rank)                      # different init per rank on purpose
    t = trainer.StepATrainer(device="cpu")
    assert t.world == world
    t.broadcast_parameters()                           # now identical to rank 0
    flat0 = t.bG.flat.clone()
    gathered = [torch.zeros_like(flat0) for _ in range(world)]
    dist.all_gather(gathered, flat0)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    # gradient exchange: bucket.grad = rank + 1 everywhere -> SUM = 3, Adam sees SUM / world through gscale
    for b in [t.bG, *t.bD]:
        b.grad.fill_(float(rank + 1))
    works = [t._allreduce(b) for b in [t.bG, *t.bD]]
    for w in works:
        w.wait()
    summed = all(bool((b.grad == 3.0).all()) for b in [t.bG, *t.bD])
    # sharding of a global batch
    inp = synth.make_inputs(4, seed=3, parity=True)
    sh = synth.shard(inp, rank, world)
    ok_shard = sh["z"].shape[0] == 2 and torch.equal(sh["z"], inp["z"][rank * 2:(rank + 1) * 2]) \
        and sh["hmaps"][2].shape[0] == 2 and sh["slabels_feat"].shape[2] == int(sh["num_rois"].max())
    # one dry-run step exercises the step's control flow with world > 1 (all-reduce calls included)
    t.step(sh)
    out[rank] = (same, summed, ok_shard, t.bG.step)
    dist.destroy_process_group()


def test_dp_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        same, summed, ok_shard, step = out[r]
        assert same and summed and ok_shard and step == 1


def _worker_b(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import random
    from objgan_b200 import lib, model, trainer
    lib.DRY_RUN = True
    model.FAST_INIT = True                # values are never looked at in DRY_RUN
    torch.manual_seed(200 + rank)
    t = trainer.StepBTrainer(device="cpu")
    t.broadcast_parameters()
    flat = t.bObj[1].flat.clone()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    inp = synth.make_inputs(4, seed=3, parity=True)
    if rank == 1:                       # this rank sees no roi at all: its object-D losses are the int 0
        inp["num_rois"] = inp["num_rois"].clone()
    sh = synth.shard(inp, rank, world)
    if rank == 1:
        sh["fm_rois"] = sh["fm_rois"].clone()
        sh["fm_rois"][..., 2:4] = 0.5   # every box below the 1.25-cell filter -> feat_select keeps nothing
        sh["rois"] = [r.clone() for r in sh["rois"]]
        sh["rois"][0][..., 2:4] = 0.5
    random.seed(7 + rank)
    res = t.step(sh)
    out[rank] = (same, t.bG.step, [b.step for b in t.bShp], [b.step for b in t.bObj],
                 res["errObjSSD"] is None, res["errObjLSD"] is None)
    dist.destroy_process_group()


def test_dp_world2_gloo_step_b():
    """The complete step under data parallelism: a rank whose shard has no usable roi still takes part in the object
    discriminators' exchange (zero gradient) and all ranks take the same optimiser steps."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_b, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0][0] and out[1][0]
    assert out[0][1] == out[1][1] == 1 and out[0][2] == out[1][2] == [1, 1, 1]
    assert out[0][3] == out[1][3]                       # identical optimiser steps on both ranks
    assert out[1][4] and out[1][5]                      # rank 1 had no object loss of its own
