"""Determinism check of the chunked / threaded large-bond path: repeated calls must agree bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch
from qsp_b200.engine import Engine
eng = Engine()
B, L, chi, layers = 6, 14, 70, 2
raw, bonds = synth_batch(torch, B, L, chi, eng.device, 9)
assert int(eng.prepare_target(raw, bonds).max()) == 0
ones = [eng.seq_prepare(raw, bonds, layers, max_bond=chi) for _ in range(3)]
manys = [eng.seq_prepare(raw, bonds, layers, max_bond=chi, chunk=2) for _ in range(6)]
seq = [eng.seq_prepare(raw, bonds, layers, max_bond=chi, chunk=2, overlap_chunks=False) for _ in range(2)]
torch.cuda.synchronize()
def d(a, b): return float((a.dis_overlaps - b.dis_overlaps).abs().max()), float((torch.view_as_real(a.G) - torch.view_as_real(b.G)).abs().max())
print("one vs one", [d(ones[0], x) for x in ones[1:]])
print("many vs many", [d(manys[0], x) for x in manys[1:]])
print("seq vs seq", [d(seq[0], x) for x in seq[1:]])
print("seq vs many", [d(seq[0], x) for x in manys])
print("one vs seq", d(ones[0], seq[0]))
print("dis", ones[0].dis_overlaps[:, 0].tolist())
