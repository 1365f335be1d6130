import numpy as np, torch, sys, ctypes as C
sys.path.insert(0, '.')
from arreau_b200 import _lib
from arreau_b200.weights import PonitaWeights
from arreau_b200.engine import DenoiseEngine
from arreau_b200.tables import build_tables
from arreau_b200.synthetic import make_crystals
dev = torch.device('cuda')
w = np.load('tests/golden/weights_seed0.npz')
sd = {k: w[k] for k in w.files if k not in ('ori_grid', 'fourier_w')}
pw = PonitaWeights(sd, w['ori_grid'], device=dev)
cr = make_crystals(1024, 40, None, seed=0)
eng = DenoiseEngine(pw, build_tables(1000, 90), w['fourier_w'], cr.num_atoms, 5.0, 8, precision='fp16', device=dev)
eng.set_state(cr.frac, cr.types, cr.lengths, cr.angles)
eng.predict_scores(300); torch.cuda.synchronize()
buf = torch.zeros(6 * 2 * 16, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.arreau_debug_set_tc_profile.argtypes = [C.c_void_p]
assert lib.arreau_debug_set_tc_profile(buf.data_ptr()) == 0
eng.forward(); torch.cuda.synchronize()
p = buf.cpu().numpy().reshape(6, 2, 16)
t0 = p[0, 1, 0]
names_m = ['wait_a1', 'a1_ok', 'g1_issued/wait_a2', 'a2_ok', 'g2_issued/wait_a3', 'a3_ok'] + [f'l{l}_issued' for l in range(5)] + ['L2:xempty_ok', 'L2:w0', 'L2:w1', 'L2:w2', 'L2:w3']
names_e = ['tile_start', 'gen_done', 'x0_full', 'epi1_done', 'd2_full', 'epi2_done'] + [f'epi3_{l}_done' for l in range(5)] + ['L2:xfull', 'L2:ld_done', 'L2:rd_wait', 'L2:bar1', 'L2:bar2']
for it in range(1, 4):
    print(f'tile {it}:')
    print('  MMA :', ' '.join(f'{n}={p[it,0,i]-t0}' for i, n in enumerate(names_m)))
    print('  EPI :', ' '.join(f'{n}={p[it,1,i]-t0}' for i, n in enumerate(names_e)))
