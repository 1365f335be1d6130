"""Same-process A/B of the two edge-kernel implementations (arreau_debug_set_edge_variant) at C2 / C3 sizes:
bitwise comparison of the kernel slabs and CUDA-event timing, alternating."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from arreau_b200 import _lib  # noqa: E402
from arreau_b200.engine import DenoiseEngine  # noqa: E402
from arreau_b200.tables import build_tables  # noqa: E402
from arreau_b200.weights import PonitaWeights  # noqa: E402

G, n, radius, cap = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1024, 40, 5.0, 8)
dev = torch.device("cuda:0")
lib = _lib.load()
lib.arreau_debug_set_edge_variant.argtypes = [C.c_int]
sd, ori, fw = bench.load_weights(n)
eng = DenoiseEngine(PonitaWeights(sd, ori, device=dev), build_tables(1000, 90), fw, [n] * G, radius, cap, precision="fp16", device=dev)
eng.set_state(*bench.teacher_state(G, n, 5, 400))
eng.prepare_inputs(400)
eng._ensure_capacity()
eng.build_graph()
torch.cuda.synchronize()
E = eng.num_edges()
w = eng.w.t


def run():
    _lib.call("arreau_edge_kernels_f16", eng.dir.data_ptr(), eng.dist.data_ptr(), eng.lattice.data_ptr(),
              eng.crystal_of_atom.data_ptr(), eng.src.data_ptr(), eng.row_ptr.data_ptr() + 4 * eng.N, eng.edge_capacity,
              w["ori"].data_ptr(), w["edge_w1_img"].data_ptr(), w["edge_w_img"].data_ptr(), w["b2"].data_ptr(), eng.radius,
              eng.kernels.data_ptr(), eng.stream)


outs = {}
for v in (2, 3):
    assert lib.arreau_debug_set_edge_variant(v) == 0
    eng.kernels.zero_()
    run()
    torch.cuda.synchronize()
    outs[v] = eng.kernels[:, :E].clone() if E * 5 * 4096 < 20e9 else eng.kernels[:, : E // 8].clone()
for v in (2, 3):
    same = torch.equal(outs[2].view(torch.int16), outs[v].view(torch.int16))
    diff = float((outs[2].float() - outs[v].float()).abs().max())
    print(f"G={G} n={n} cap={cap} E={E}: v2 vs v{v} bitwise equal: {same}, max abs diff {diff:.3e}, max |v2| {float(outs[2].float().abs().max()):.3e}")
del outs
times = {2: [], 3: []}
for rep in range(6):
    for v in (2, 3):
        lib.arreau_debug_set_edge_variant(v)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            run()
        b.record()
        torch.cuda.synchronize()
        if rep:
            times[v].append(a.elapsed_time(b) / 5)
flops = 2.0 * E * 16 * (83 * 128 + 128 * 256 + 5 * 256 * 128)
for v in (2, 3):
    t = float(np.median(times[v]))
    print(f"variant {v}: {t:.4f} ms per launch (median of {len(times[v])} x 5), {flops / t / 1e9:.1f} TFLOP/s; all {np.round(times[v], 4).tolist()}")
lib.arreau_debug_set_edge_variant(3)
