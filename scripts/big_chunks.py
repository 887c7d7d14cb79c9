"""cfg-5 shape: one layer with different chunkings (L2 residency of the Jacobi working set vs parallelism)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch
from qsp_b200.engine import Engine
L, chi, B = (int(x) for x in sys.argv[1:4])
eng = Engine()
raw, b0 = synth_batch(torch, B, L, chi, eng.device, 1)
st = eng.prepare_target(raw, b0); torch.cuda.synchronize()
ref = None
for chunk, ov in [(16, True), (16, False), (8, False), (32, False), (8, True)]:
    torch.cuda.synchronize(); t0 = time.time()
    res = eng.seq_prepare(raw, b0, 1, max_bond=chi, encoded=False, chunk=chunk, overlap_chunks=ov)
    torch.cuda.synchronize(); dt = time.time() - t0
    if ref is None: ref = res.dis_overlaps.clone()
    print(f"chunk {chunk} overlap {ov}: {dt:.2f} s/layer  status {res.status.unique().tolist()}  max diff vs first {float((res.dis_overlaps - ref).abs().max()):.1e}", flush=True)
