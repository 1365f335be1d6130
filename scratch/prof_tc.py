import numpy as np, torch, sys, ctypes as C
sys.path.insert(0, '.')
import bench
from arreau_b200 import _lib
from arreau_b200.weights import PonitaWeights
from arreau_b200.engine import DenoiseEngine
from arreau_b200.tables import build_tables
dev = torch.device('cuda')
G, n = 1024, 40
sd, ori, fw = bench.load_weights(n)
eng = DenoiseEngine(PonitaWeights(sd, ori, device=dev), build_tables(1000, 90), fw, [n] * G, 5.0, 8, precision='fp16', device=dev)
eng.set_state(*bench.teacher_state(G, n, 0, 500))
eng.predict_scores(500); torch.cuda.synchronize()
lib = _lib.load()
flags = 0
buf = torch.zeros(6 * 2 * 16, dtype=torch.int64, device=dev)
lib.arreau_debug_set_tc_profile.argtypes = [C.c_void_p]
assert lib.arreau_debug_set_tc_profile(buf.data_ptr()) == 0
w = eng.w.t
nep = eng.row_ptr.data_ptr() + 4 * eng.N
_lib.call("arreau_edge_kernels_f16", eng.dir.data_ptr(), eng.dist.data_ptr(), eng.lattice.data_ptr(),
          eng.crystal_of_atom.data_ptr(), eng.src.data_ptr(), nep, eng.edge_capacity, w["ori"].data_ptr(),
          w["edge_w1_img"].data_ptr(), w["edge_w_img"].data_ptr(), w["b2"].data_ptr(), eng.radius,
          eng.kernels.data_ptr(), eng.stream)
torch.cuda.synchronize()
p = buf.cpu().numpy().reshape(6, 2, 16)
t0 = p[1, 0, 0]
names_m = ['top', 'a3_ok', 'L0', 'L0x', 'L1', 'G1', 'L2', 'L2x', 'L3', 'G2', 'L4', 'L4x']
names_e = ['top', 'gen', 'attr', 'E0', 'E1', 'Q1', 'E2', 'E3', 'E4', 'Q2']
print('flags', flags)
for it in range(2, 4):
    print(f'tile {it}:')
    print('  MMA :', ' '.join(f'{n}={p[it,0,i]-t0}' for i, n in enumerate(names_m)))
    print('  EPI :', ' '.join(f'{n}={p[it,1,i]-t0}' for i, n in enumerate(names_e)))
