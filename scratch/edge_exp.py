"""Timing decomposition of the edge tc kernel with debug flags (results are wrong when a flag is set)."""
import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, '.')
import bench
from arreau_b200 import _lib
from arreau_b200.engine import DenoiseEngine
from arreau_b200.tables import build_tables
from arreau_b200.weights import PonitaWeights
dev = torch.device('cuda')
G, n = 1024, 40
sd, ori, fw = bench.load_weights(n)
eng = DenoiseEngine(PonitaWeights(sd, ori, device=dev), build_tables(1000, 90), fw, [n] * G, 5.0, 8, precision='fp16', device=dev)
eng.set_state(*bench.teacher_state(G, n, 0, 500))
eng.predict_scores(500)
torch.cuda.synchronize()
lib = _lib.load()
w = eng.w.t
nep = eng.row_ptr.data_ptr() + 4 * eng.N
def run():
    _lib.call("arreau_edge_kernels_f16", eng.dir.data_ptr(), eng.dist.data_ptr(), eng.lattice.data_ptr(),
              eng.crystal_of_atom.data_ptr(), eng.src.data_ptr(), nep, eng.edge_capacity, w["ori"].data_ptr(),
              w["edge_w1_img"].data_ptr(), w["edge_w_img"].data_ptr(), w["b2"].data_ptr(), eng.radius,
              eng.kernels.data_ptr(), eng.stream)
for _ in range(2): run()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): run()
b.record(); torch.cuda.synchronize()
print('edge kernel ms', a.elapsed_time(b) / 5)
