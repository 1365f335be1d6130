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
eng.predict_scores(500)
torch.cuda.synchronize()
