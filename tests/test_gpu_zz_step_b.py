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


@pytest.mark.skipif(__import__("os").environ.get("OBJGAN_RUN_UNVALIDATED") != "1",
                    reason="written after the round's GPU budget was spent; enable with OBJGAN_RUN_UNVALIDATED=1")
def test_step_b_parity(monkeypatch):
    """One complete step against oracle.step_b (exact-fp32 engine): every discriminator loss, the generator loss and
    KL, and the fake images.  Both sides draw permute_seg's shuffles from Python's ``random`` in the same order."""
    from objgan_b200 import ops
    from oracle import objgan_oracle as O
    monkeypatch.setattr(ops, "CONV_ENGINE", "simt")
    t = trainer.StepBTrainer(device="cuda", seed=3)
    cpu = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    st = O.StepBState(cpu(t.netG), [cpu(d) for d in t.netsPatD], [cpu(d) for d in t.netsShpD], cpu(t.netObjSSD),
                      cpu(t.netObjLSD))
    inp = synth.make_inputs(4, seed=4, parity=True)
    random.seed(5)
    want = O.step_b(st, inp)
    dev = t.to_device(inp)
    random.seed(5)
    got = t.step(dev)
    for k in ("errPatD0", "errPatD1", "errPatD2", "errShpD0", "errShpD1", "errShpD2", "errObjSSD", "errObjLSD",
              "errG", "kl"):
        if want[k] is None:
            assert got[k] is None, k
        else:
            assert abs(float(got[k]) - want[k]) <= 2e-3 * max(1.0, abs(want[k])), (k, float(got[k]), want[k])
