"""Repeat the scenario of tests/test_gpu_exports.py::test_overlapped_chunks_large_bonds and print the differences
between one resident call and the chunked, two-thread / two-stream call.  python scripts/stress_chunks.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch
from qsp_b200.engine import Engine
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B, L, chi, layers = 6, 14, 70, 2
ref = None
for it in range(iters):
    eng = Engine()
    raw, bonds = synth_batch(torch, B, L, chi, eng.device, 9)
    assert int(eng.prepare_target(raw, bonds).max()) == 0
    one = eng.seq_prepare(raw, bonds, layers, max_bond=chi)
    many = eng.seq_prepare(raw, bonds, layers, max_bond=chi, chunk=2)
    torch.cuda.synchronize()
    d = lambda a, b: float((a - b).abs().max())
    if ref is None:
        ref = one
    print(it, "one-many dis %.2e ov %.2e sv %.2e G %.2e | one-ref dis %.2e | many-ref dis %.2e per-instance %s" % (
        d(one.dis_overlaps, many.dis_overlaps), d(one.overlaps[-1], many.overlaps[-1]), d(one.d2_svals, many.d2_svals),
        d(torch.view_as_real(one.G), torch.view_as_real(many.G)), d(one.dis_overlaps, ref.dis_overlaps),
        d(many.dis_overlaps, ref.dis_overlaps), (one.dis_overlaps - many.dis_overlaps).abs().max(0).values.tolist()), flush=True)
