"""clock64 timeline of CTA 0 of the ConvNext MLP kernel v2 (8 tiles from TILE0): MMA warp, G-group warp, R-group warp."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from arreau_b200 import _lib  # noqa: E402
from arreau_b200.engine import NUM_ORI, DenoiseEngine  # noqa: E402
from arreau_b200.tables import build_tables  # noqa: E402
from arreau_b200.weights import PonitaWeights  # noqa: E402

dev = torch.device("cuda")
G, n = 1024, 40
TILE0 = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sd, ori, fw = bench.load_weights(n)
eng = DenoiseEngine(PonitaWeights(sd, ori, device=dev), build_tables(1000, 90), fw, [n] * G, 5.0, 8, precision="fp16", device=dev)
eng.set_state(*bench.teacher_state(G, n, 0, 500))
eng.predict_scores(500)
torch.cuda.synchronize()
lib = _lib.load()
lib.arreau_debug_set_tc_profile.argtypes = [C.c_void_p]
lib.arreau_debug_set_tc_profile_tile0.argtypes = [C.c_uint]
buf = torch.zeros(8 * 3 * 32, dtype=torch.int64, device=dev)
assert lib.arreau_debug_set_tc_profile_tile0(TILE0) == 0 and lib.arreau_debug_set_tc_profile(buf.data_ptr()) == 0
w, Z, l = eng.w.t, eng.Z, 2
for _ in range(2):
    _lib.call("arreau_convnext_mlp_f16_pooled", eng.y.data_ptr(), w["mlp_w_img"].data_ptr() + l * 8 * 32768,
              w["mlp_b1"][l].data_ptr(), w["mlp_b2"][l].data_ptr(), w["layer_scale"][l].data_ptr(), eng.N * NUM_ORI,
              eng.h.data_ptr(), w["ori"].data_ptr(), eng.pool[l + 1].data_ptr(), w["readout_v"][l + 1].data_ptr(), Z, eng.stream)
    torch.cuda.synchronize()
lib.arreau_debug_set_tc_profile(None)
lib.arreau_debug_set_tc_profile_tile0(2)
p = buf.cpu().numpy().reshape(8, 3, 32)
t0 = p[0, 0, 8]
print("cycles relative to the MMA warp's entry into GEMM2 slice 0 of the first stamped tile")
for it in range(6):
    m, g, r = p[it, 0] - t0, p[it, 1] - t0, p[it, 2] - t0
    print(f"tile {TILE0 + it}: span {p[it + 1, 0, 8] - p[it, 0, 8]}")
    print("   MMA  G1_j (entry,issued):", [(int(m[2 * j]), int(m[2 * j + 1])) for j in range(4)], " G2_j:", [(int(m[8 + 2 * j]), int(m[9 + 2 * j])) for j in range(4)])
    print("   GELU slice j (entry, waits done, done):", [(int(g[3 * j]), int(g[3 * j + 1]), int(g[3 * j + 2])) for j in range(4)],
          " durations", [int(g[3 * j + 2] - g[3 * j + 1]) for j in range(4)])
    print("   RES  entry, d2 ready, staged, issued:", [int(r[i]) for i in range(4)], " work", int(r[3] - r[1]))
print("tile spans", [int(p[i + 1, 0, 8] - p[i, 0, 8]) for i in range(7)])
