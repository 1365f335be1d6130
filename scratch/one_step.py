"""A few C2-sized denoise steps (fp16 tensor path) for ncu captures: python scratch/one_step.py [steps]"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import bench
from arreau_b200.engine import DenoiseEngine
from arreau_b200.tables import build_tables
from arreau_b200.weights import PonitaWeights
dev = torch.device('cuda')
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
G, n = 1024, 40
sd, ori, fw = bench.load_weights(n)
eng = DenoiseEngine(PonitaWeights(sd, ori, device=dev), build_tables(1000, 90), fw, [n] * G, 5.0, 8, precision='fp16', device=dev)
eng.set_state(*bench.teacher_state(G, n, 0, 500))
t = 500
for i in range(steps):
    eng.draw_noise(1, i)
    eng.step(t); t -= 1
torch.cuda.synchronize()
print('E/N', eng.num_edges() / eng.N)
