"""Two-GPU data-parallel parity: one Step-A step on 2 ranks (one process per GPU, NCCL all-reduce of
each network's gradient bucket) against oracle.step_a_dp -- the mean over the two shards' reference gradients, one Adam
step from it, the generator update through the discriminators after THEIR averaged step.

Needs 2 GPUs: run with  gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp2.py -m gpu -q  (skipped on one GPU)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, engine, out_dir):
    import torch.distributed as dist
    from objgan_b200 import ops, synth, trainer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ops.CONV_ENGINE = engine
    t = trainer.StepATrainer(device=f"cuda:{rank}", seed=100 + rank)          # different seeds: broadcast must fix that
    t.broadcast_parameters()
    if rank == 0:
        torch.save({"g": {k: v.cpu() for k, v in t.netG.state_dict().items()},
                    "d": [{k: v.cpu() for k, v in d.state_dict().items()} for d in t.netsPatD]},
                   os.path.join(out_dir, "init.pt"))
    full = synth.make_inputs(2 * world, seed=61, parity=True)
    mine = synth.shard(full, rank, world)
    out = t.step(t.to_device(mine))
    torch.cuda.synchronize()
    res = {"losses": {k: float(out[k]) for k in ("errPatD0", "errPatD1", "errPatD2", "errG", "kl")},
           "g_grad": {k: (p.grad / world).cpu() for k, p in t.netG.named_parameters()},
           "d_grad": [{k: (p.grad / world).cpu() for k, p in d.named_parameters()} for d in t.netsPatD],
           "g": {k: p.detach().cpu() for k, p in t.netG.named_parameters()},
           "flat_g": t.bG.flat.detach().cpu()}
    torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("engine", ["simt", "f16x3"])
def test_two_rank_step_matches_mean_of_shard_oracles(engine, tmp_path):
    import torch.multiprocessing as mp
    from objgan_b200 import synth
    from oracle import objgan_oracle as O
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), engine, str(tmp_path)), nprocs=world, join=True)
    init = torch.load(tmp_path / "init.pt")
    r = [torch.load(tmp_path / f"rank{i}.pt") for i in range(world)]
    # replicas stay identical: same parameters after the step on both ranks
    assert torch.equal(r[0]["flat_g"], r[1]["flat_g"])
    full = synth.make_inputs(2 * world, seed=61, parity=True)
    shards = [synth.shard(full, i, world) for i in range(world)]
    state = O.StepAState(init["g"], init["d"])
    keep = {}
    want = O.step_a_dp(state, shards, keep=keep)
    for i in range(world):
        for k, v in want[i].items():
            assert abs(r[i]["losses"][k] - v) <= 2e-3 * max(1.0, abs(v)), (i, k, r[i]["losses"][k], v)
    l2 = 5e-3 if engine == "simt" else 3e-2
    for j in range(3):                       # discriminator gradients: mean over the shards
        for k, gref in keep["d_grads"][j].items():
            got = r[0]["d_grad"][j][k]
            rel = ((got.double() - gref.double()).norm() / gref.double().norm().clamp(min=1e-30)).item()
            assert rel <= l2, ("PatD%d" % j, k, rel)
    # generator gradient (through the discriminators after their averaged Adam step: sign descent amplifies noise-level
    # entries, test_step_a_parity): direction + per-tensor L2 as in the single-GPU step test
    keys = [k for k in state.g_keys if not k.endswith("conv3x3.1.bias")]
    dot = sum((r[0]["g_grad"][k].double() * keep["g_grads"][k].double()).sum() for k in keys)
    na = sum((r[0]["g_grad"][k].double() ** 2).sum() for k in keys).sqrt()
    nb = sum((keep["g_grads"][k].double() ** 2).sum() for k in keys).sqrt()
    assert float(dot / (na * nb)) > (0.9999 if engine == "simt" else 0.999), float(dot / (na * nb))
    for k in keys:                           # parameters after Adam where the gradient sign is unambiguous
        gr = keep["g_grads"][k]
        sel = gr.abs() > 0.2 * gr.abs().max()
        if sel.any():
            assert (r[0]["g"][k] - state.g[k])[sel].abs().max().item() < 5e-5, k
