"""clock64 timeline of CTA 0 of the v2 edge kernel (tiles 2..9 of that CTA): MMA warp and epilogue warp 4, per job."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from arreau_b200 import _lib  # noqa: E402
from arreau_b200.engine import DenoiseEngine  # noqa: E402
from arreau_b200.tables import build_tables  # noqa: E402
from arreau_b200.weights import PonitaWeights  # noqa: E402

dev = torch.device("cuda")
G, n = 1024, 40
sd, ori, fw = bench.load_weights(n)
eng = DenoiseEngine(PonitaWeights(sd, ori, device=dev), build_tables(1000, 90), fw, [n] * G, 5.0, 8, precision="fp16", device=dev)
eng.set_state(*bench.teacher_state(G, n, 0, 500))
eng.predict_scores(500)
torch.cuda.synchronize()
lib = _lib.load()
lib.arreau_debug_set_edge_variant.argtypes = [C.c_int]
lib.arreau_debug_set_edge_variant(2)
TILE0 = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lib.arreau_debug_set_tc_profile_tile0.argtypes = [C.c_uint]
assert lib.arreau_debug_set_tc_profile_tile0(TILE0) == 0
buf = torch.zeros(8 * 2 * 32 + 4, dtype=torch.int64, device=dev)
lib.arreau_debug_set_tc_profile.argtypes = [C.c_void_p]
assert lib.arreau_debug_set_tc_profile(buf.data_ptr()) == 0
w = eng.w.t
nep = eng.row_ptr.data_ptr() + 4 * eng.N
for _ in range(2):
    _lib.call("arreau_edge_kernels_f16", eng.dir.data_ptr(), eng.dist.data_ptr(), eng.lattice.data_ptr(),
              eng.crystal_of_atom.data_ptr(), eng.src.data_ptr(), nep, eng.edge_capacity, w["ori"].data_ptr(),
              w["edge_w1_img"].data_ptr(), w["edge_w_img"].data_ptr(), w["b2"].data_ptr(), eng.radius,
              eng.kernels.data_ptr(), eng.stream)
    torch.cuda.synchronize()
lib.arreau_debug_set_tc_profile(None)
raw = buf.cpu().numpy()
p = raw[:8 * 2 * 32].reshape(8, 2, 32)
c0, n0, c1, n1 = raw[8 * 2 * 32:]
print(f'SM clock over tiles 0..256 of CTA 0: {(c1 - c0) / max(n1 - n0, 1):.3f} GHz ({c1 - c0} cycles in {(n1 - n0) / 1e3:.1f} us, {(c1 - c0) / 256:.0f} cycles per tile)')
jobs = ["L0", "G1'", "L1", "G2a'", "L2", "G2b'", "L3", "L4"]
t0 = p[0, 0, 0]
print("cycles relative to the MMA warp's entry into L0 of tile 2; MMA: entry / waits done / issued; EPI: entry / acc in regs / post done")
for it in range(6):
    print(f"tile {it + TILE0}: span to next tile {p[it + 1, 0, 0] - p[it, 0, 0]}")
    for jb, name in enumerate(jobs):
        m = p[it, 0, 3 * jb:3 * jb + 3] - t0
        e = p[it, 1, 3 * jb:3 * jb + 3] - t0
        print(f"   {name:5s} MMA {m[0]:7d} {m[1]:7d} {m[2]:7d}  (wait {m[1] - m[0]:5d} issue {m[2] - m[1]:5d})   "
              f"EPI {e[0]:7d} {e[1]:7d} {e[2]:7d}  (wait+ld {e[1] - e[0]:5d} post {e[2] - e[1]:5d})")
    print(f"   gen' EPI {p[it, 1, 30] - t0:7d} .. {p[it, 1, 31] - t0:7d} ({p[it, 1, 31] - p[it, 1, 30]})")
spans = [p[i + 1, 0, 0] - p[i, 0, 0] for i in range(7)]
print("tile spans", spans, "mean", np.mean(spans))
