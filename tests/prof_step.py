"""GPU-time breakdown of one training step by kernel name (torch.profiler / CUPTI; guidance only, not a bench value).

    python tests/prof_step.py b        # Step-B (eager)          -> gpurun_out/prof_step_b.txt
    python tests/prof_step.py a        # Step-A (eager launches) -> gpurun_out/prof_step_a.txt
Also prints how long the host needs to ENQUEUE a step (no synchronisation) next to the GPU time of the step: if the
two are equal the step is bound by Python / launch overhead, not by the kernels."""
import collections
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from objgan_b200 import synth, trainer

which = sys.argv[1] if len(sys.argv) > 1 else "b"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
random.seed(1234)
tr = (trainer.StepBTrainer if which == "b" else trainer.StepATrainer)(device="cuda", seed=1234)
host = trainer.pin(synth.make_inputs(B, seed=1234, parity=False))
host.pop("eps")
dev = tr.to_device(host)
for _ in range(2):
    tr.step(dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    tr.step(dev)
e1.record()
t_enq = (time.perf_counter() - t0) / 3
torch.cuda.synchronize()
t_gpu = e0.elapsed_time(e1) / 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(dev)
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
busy = 0.0
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        k = ev.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:90]
        tot[k][0] += 1
        tot[k][1] += ev.device_time
        busy += ev.device_time
os.makedirs("gpurun_out", exist_ok=True)
with open(f"gpurun_out/prof_step_{which}.txt", "w") as f:
    print(f"step {which} B={B}: host enqueue {t_enq * 1e3:.1f} ms/step, GPU (events) {t_gpu:.1f} ms/step, "
          f"sum of kernel times {busy / 1e3:.1f} ms, kernels {sum(v[0] for v in tot.values())}", file=f)
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:60]:
        print(f"{us / 1e3:9.3f} ms {100 * us / busy:5.1f}% {n:5d}  {k}", file=f)
print(open(f"gpurun_out/prof_step_{which}.txt").read())
