"""GPU smoke test of the complete training step (StepBTrainer: all nine optimisers).  Every loss and network it
composes has its own parity test in test_gpu_parity.py; this one checks that the composition runs on the device,
stays finite, and that every optimiser that had work to do moved its parameters.  (File name sorts last on purpose.)"""
import random

import pytest
import torch

from objgan_b200 import synth, trainer

pytestmark = pytest.mark.gpu


def test_step_b_runs():
    t = trainer.StepBTrainer(device="cuda", seed=3)
    inp = synth.make_inputs(2, seed=4, parity=True)
    inp.pop("eps")
    dev = t.to_device(inp)
    before = {id(b): b.flat.clone() for b in [t.bG, *t._d_buckets()]}
    random.seed(5)
    out = t.step(dev)
    torch.cuda.synchronize()
    for k in ("errG", "kl", "errPatD0", "errPatD1", "errPatD2", "errShpD0", "errShpD1", "errShpD2"):
        assert torch.isfinite(out[k]).all(), k
    for k in ("errObjSSD", "errObjLSD"):
        assert out[k] is None or torch.isfinite(out[k]).all(), k
    assert all(torch.isfinite(v).all() for v in out["logs"].values())
    for b in [t.bG, *t.bD, *t.bShp]:
        assert b.step == 1 and torch.isfinite(b.flat).all() and not torch.equal(b.flat, before[id(b)])
    for b, k in zip(t.bObj, ("errObjSSD", "errObjLSD")):
        assert b.step == (0 if out[k] is None else 1)
        assert torch.isfinite(b.flat).all()
    assert torch.isfinite(t.bG.avg).all()
