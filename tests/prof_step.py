"""Kernel time table of one Step-A step via torch.profiler (CUPTI); guidance only, not a bench value."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from objgan_b200 import synth, trainer
tr = trainer.StepATrainer(device="cuda", seed=1234)
inp = synth.make_inputs(16, seed=1234, parity=False)
inp.pop("eps")
dev = tr.to_device(inp)
for _ in range(2):
    tr.step(dev)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step(dev)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
