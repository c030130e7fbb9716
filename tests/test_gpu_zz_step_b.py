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


def _cpu_sd(m):
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


def _check_weights(net, sd_oracle, grads_clear=None, tol=5e-5, what=""):
    """Parameters after one Adam step.  The first Adam step is sign descent (every entry moves by ~lr), so an entry
    whose gradient is rounding noise may legitimately move the other way: compare the fraction of entries that agree
    to 0.1 * lr and the worst deviation (bounded by 2 * lr + tol by construction), like test_step_a_parity."""
    lr = 2e-4
    for k, p in net.named_parameters():
        a, b = p.detach().cpu(), sd_oracle[k]
        d = (a - b).abs()
        assert torch.isfinite(a).all(), (what, k)
        assert d.max().item() <= 2 * lr + tol, (what, k, d.max().item())
        agree = (d <= 0.1 * lr).float().mean().item()
        if k.endswith(("shp_code.1.bias", "conv3x3.1.bias")):
            continue        # bias ahead of InstanceNorm: exact gradient is zero, the step direction is noise
        assert agree >= 0.97, (what, k, agree)


@pytest.mark.parametrize("engine", ["simt", "f16x3"])
def test_step_b_parity(engine, monkeypatch):
    """One complete step (ref: trainer.py:385-462) against oracle.step_b: every discriminator loss, the generator loss
    and KL, the fake images, and the weights of all nine networks + the generator's EMA after their Adam steps.  Both
    sides draw permute_seg's shuffles from Python's ``random`` in the same order.  oracle.step_b is itself pinned
    against the reference's own loop body in tests/test_oracle_vs_reference.py::test_step_b_reference_loop."""
    from objgan_b200 import ops
    from oracle import objgan_oracle as O
    monkeypatch.setattr(ops, "CONV_ENGINE", engine)
    t = trainer.StepBTrainer(device="cuda", seed=3)
    st = O.StepBState(_cpu_sd(t.netG), [_cpu_sd(d) for d in t.netsPatD], [_cpu_sd(d) for d in t.netsShpD],
                      _cpu_sd(t.netObjSSD), _cpu_sd(t.netObjLSD))
    inp = synth.make_inputs(3, seed=4, parity=True)
    random.seed(5)
    want = O.step_b(st, inp)
    dev = t.to_device(inp)
    random.seed(5)
    got = t.step(dev)
    ltol = 2e-3 if engine == "simt" else 5e-3
    for k in ("errPatD0", "errPatD1", "errPatD2", "errShpD0", "errShpD1", "errShpD2", "errObjSSD", "errObjLSD",
              "errG", "kl"):
        if want[k] is None:
            assert got[k] is None, k
        else:
            assert abs(float(got[k]) - want[k]) <= ltol * max(1.0, abs(want[k])), (k, float(got[k]), want[k])
    for k, v in want["terms"].items():
        assert abs(float(got["logs"][k]) - v) <= ltol * max(1.0, abs(v)), (k, float(got["logs"][k]), v)
    for i in range(3):
        a, b = got["fake_imgs"][i].cpu(), want["fake"][i]
        assert (a - b).abs().max().item() <= 1e-3 * b.abs().max().item(), i
    for i, d in enumerate(t.netsPatD):
        _check_weights(d, st.ds[i], what=f"PatD{i}")
    for i, d in enumerate(t.netsShpD):
        _check_weights(d, st.extra[i], what=f"ShpD{i}")
    for j, (d, k) in enumerate(((t.netObjSSD, "errObjSSD"), (t.netObjLSD, "errObjLSD"))):
        _check_weights(d, st.extra[3 + j], what=k)
        assert t.bObj[j].step == st.x_step[3 + j]
    _check_weights(t.netG, st.g, what="G")
    ema = t.bG.ema_state_dict()
    for k in st.g_keys:
        assert (ema[k].cpu() - st.g_avg[k]).abs().max().item() <= 1e-3 * (2 * 2e-4 + 5e-5) + 1e-7, k
