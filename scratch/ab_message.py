"""A/B of the fused fp16 message kernel with / without the L2 prefetch of the next slab block, on a C2-sized (and a
C3-sized) state: same process, alternating, bitwise comparison of y.   python scratch/ab_message.py [G n radius cap]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from arreau_b200 import _lib  # noqa: E402
from arreau_b200.engine import HIDDEN, DenoiseEngine  # noqa: E402
from arreau_b200.tables import build_tables  # noqa: E402
from arreau_b200.weights import PonitaWeights  # noqa: E402

G, n, radius, cap = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1024, 40, 5.0, 8)
dev = torch.device("cuda")
lib = _lib.load()
lib.arreau_debug_set_message_prefetch.argtypes = [C.c_int]
sd, ori, fw = bench.load_weights(n)
eng = DenoiseEngine(PonitaWeights(sd, ori, device=dev), build_tables(1000, 90), fw, [n] * G, radius, cap, precision="fp16", device=dev)
eng.set_state(*bench.teacher_state(G, n, 0, 500))
eng.predict_scores(500)
torch.cuda.synchronize()
w, s = eng.w.t, torch.cuda.current_stream().cuda_stream


def run(l, y):
    frag = w["fiber_frag"].data_ptr() + l * HIDDEN * 32 * 16
    _lib.call("arreau_message_fiber_norm_fused", eng.kernels[l].data_ptr(), eng.h.data_ptr(), eng.row_ptr.data_ptr(),
              eng.src.data_ptr(), frag, w["conv_bias"][l].data_ptr(), w["ln_w"][l].data_ptr(), w["ln_b"][l].data_ptr(), eng.N,
              y.data_ptr(), None, s)


ys = {}
for v in (0, 1, 2, 3):
    lib.arreau_debug_set_message_prefetch(v)
    ys[v] = torch.zeros_like(eng.y)
    run(2, ys[v])
torch.cuda.synchronize()
print("bitwise equal:", all(torch.equal(ys[0].view(torch.int16), ys[v].view(torch.int16)) for v in (1, 2, 3)), "E/N", eng.num_edges() / eng.N)
times = {0: [], 1: [], 2: [], 3: []}
for rep in range(7):
    for v in (0, 1, 2, 3):
        lib.arreau_debug_set_message_prefetch(v)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(4):
            for l in range(5):
                run(l, eng.y)       # 5 different slabs: 6.7 GB, nothing survives in L2
        e1.record()
        torch.cuda.synchronize()
        if rep:
            times[v].append(e0.elapsed_time(e1) * 1e3 / 20)
E, N = eng.num_edges(), eng.N
nbytes = E * 16 * 128 * 2 + 4 * N * 16 * 128 + 12 * E + 2 * N * 16 * 128
for v in (0, 1, 2, 3):
    t = float(np.median(times[v]))
    print(f"prefetch={v}: {t:.1f} us per launch, {nbytes / t / 1e3:.0f} GB/s algorithmic; all {np.round(times[v], 1).tolist()}")
lib.arreau_debug_set_message_prefetch(1)
