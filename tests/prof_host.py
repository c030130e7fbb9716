"""Host-side (Python) cost of one eager training step on the GPU box: cProfile of tr.step without synchronisation.
    python tests/prof_host.py b  -> gpurun_out/prof_host_b.txt"""
import cProfile
import io
import os
import pstats
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from objgan_b200 import synth, trainer

which = sys.argv[1] if len(sys.argv) > 1 else "b"
random.seed(1)
tr = (trainer.StepBTrainer if which == "b" else trainer.StepATrainer)(device="cuda", seed=1234)
full = synth.make_inputs(16, seed=1234, parity=False)
full.pop("eps")
dev = tr.to_device(trainer.pin(synth.compact(full)))
for _ in range(2):
    tr.step(dev)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
tr.step(dev)
pr.disable()
torch.cuda.synchronize()
out = io.StringIO()
st = pstats.Stats(pr, stream=out)
st.sort_stats("tottime").print_stats(45)
st.sort_stats("cumulative").print_stats(35)
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/prof_host_{which}.txt", "w").write(out.getvalue())
