"""Same-process A/B of the two ConvNext MLP kernels (arreau_debug_set_mlp_variant): bitwise comparison of h and of the
pooled read-out features, CUDA-event timing, alternating.   python scratch/ab_mlp.py [G n]"""
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

G, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 40)
dev = torch.device("cuda")
lib = _lib.load()
lib.arreau_debug_set_mlp_variant.argtypes = [C.c_int]
sd, ori, fw = bench.load_weights(n)
eng = DenoiseEngine(PonitaWeights(sd, ori, device=dev), build_tables(1000, 90), fw, [n] * G, 5.0, 8, precision="fp16", device=dev)
eng.set_state(*bench.teacher_state(G, n, 0, 500))
eng.predict_scores(500)
torch.cuda.synchronize()
w, Z = eng.w.t, eng.Z
h0 = eng.h.clone()


def run(l, pooled=True):
    if pooled:
        _lib.call("arreau_convnext_mlp_f16_pooled", eng.y.data_ptr(), w["mlp_w_img"].data_ptr() + l * 8 * 32768,
                  w["mlp_b1"][l].data_ptr(), w["mlp_b2"][l].data_ptr(), w["layer_scale"][l].data_ptr(), eng.N * NUM_ORI,
                  eng.h.data_ptr(), w["ori"].data_ptr(), eng.pool[l + 1].data_ptr(), w["readout_v"][l + 1].data_ptr(), Z, eng.stream)
    else:
        _lib.call("arreau_convnext_mlp_f16", eng.y.data_ptr(), w["mlp_w_img"].data_ptr() + l * 8 * 32768,
                  w["mlp_b1"][l].data_ptr(), w["mlp_b2"][l].data_ptr(), w["layer_scale"][l].data_ptr(), eng.N * NUM_ORI,
                  eng.h.data_ptr(), eng.stream)


outs = {}
for pooled in (True, False):
    for v in (1, 2):
        lib.arreau_debug_set_mlp_variant(v)
        eng.h.copy_(h0)
        eng.pool.zero_()
        run(2, pooled)
        torch.cuda.synchronize()
        outs[(pooled, v)] = (eng.h.clone(), eng.pool[3].clone())
    a, b = outs[(pooled, 1)], outs[(pooled, 2)]
    print(f"pooled={pooled}: h bitwise equal {torch.equal(a[0], b[0])}, pool equal {torch.equal(a[1], b[1])}, "
          f"max |dh| {float((a[0] - h0).abs().max()):.3e}")
times = {1: [], 2: []}
for rep in range(7):
    for v in (1, 2):
        lib.arreau_debug_set_mlp_variant(v)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(4):
            for l in range(5):
                run(l)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            times[v].append(e0.elapsed_time(e1) * 1e3 / 20)
flops = 2.0 * 2 * 128 * 512 * eng.N * NUM_ORI
for v in (1, 2):
    t = float(np.median(times[v]))
    print(f"variant {v}: {t:.1f} us per launch, {flops / t / 1e6:.0f} TFLOP/s; all {np.round(times[v], 1).tolist()}")
lib.arreau_debug_set_mlp_variant(2)
