"""A/B of the fp16 message kernels on a C2-sized state: python scratch/ab_message.py"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import bench
from arreau_b200 import _lib
from arreau_b200.engine import DenoiseEngine, HIDDEN
try:
    from arreau_b200.engine import build_conv_tiles
except ImportError:
    build_conv_tiles = None
from arreau_b200.tables import build_tables
from arreau_b200.weights import PonitaWeights
dev = torch.device('cuda')
G, n = 1024, 40
sd, ori, fw = bench.load_weights(n)
eng = DenoiseEngine(PonitaWeights(sd, ori, device=dev), build_tables(1000, 90), fw, [n] * G, 5.0, 8, precision='fp16', device=dev)
eng.set_state(*bench.teacher_state(G, n, 0, 500))
eng.predict_scores(500)
torch.cuda.synchronize()
w, s = eng.w.t, torch.cuda.current_stream().cuda_stream
tiles = torch.as_tensor(build_conv_tiles([n] * G)).to(dev) if build_conv_tiles else None
def run(name, l):
    frag = w["fiber_frag"].data_ptr() + l * HIDDEN * 32 * 16
    args = (w["conv_bias"][l].data_ptr(), w["ln_w"][l].data_ptr(), w["ln_b"][l].data_ptr(), eng.N)
    if name == 'fused':
        _lib.call("arreau_message_fiber_norm_fused", eng.kernels[l].data_ptr(), eng.h.data_ptr(), eng.row_ptr.data_ptr(),
                  eng.src.data_ptr(), frag, *args, eng.y.data_ptr(), None, s)
    else:
        _lib.call("arreau_message_fiber_norm_cached", eng.kernels[l].data_ptr(), eng.h.data_ptr(), eng.row_ptr.data_ptr(),
                  eng.src.data_ptr(), tiles.data_ptr(), int(tiles.shape[0]), frag, *args, eng.y.data_ptr(), None, s)
for rep in range(2):
    for name in (('fused', 'cached') if 'arreau_message_fiber_norm_cached' in _lib.SIGNATURES else ('fused',)):
        for l in range(5): run(name, l)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(4):
            for l in range(5): run(name, l)       # 5 different slabs: 6.7 GB, nothing survives in L2
        e1.record(); torch.cuda.synchronize()
        print(name, 'us per launch', round(e0.elapsed_time(e1) * 1e3 / 20, 1))
