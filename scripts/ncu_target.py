"""Small deterministic workload for ncu: python scripts/ncu_target.py L chi layers B"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch
from qsp_b200.engine import Engine
L, chi, layers, B = (int(x) for x in sys.argv[1:5])
dev = torch.device("cuda:0")
eng = Engine()
raw, b0 = synth_batch(torch, B, L, chi, dev, 1)
st = eng.prepare_target(raw, b0)
res = eng.seq_prepare(raw, b0, layers, max_bond=chi, encoded=False)
torch.cuda.synchronize()
print("status", res.status.unique().tolist(), "dis", res.dis_overlaps[layers - 1].abs().mean().item())
