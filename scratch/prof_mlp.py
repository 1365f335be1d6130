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
t = pw.t
_lib.call('arreau_convnext_mlp_f16', eng.y.data_ptr(), t['mlp_w_img'].data_ptr(), t['mlp_b1'][0].data_ptr(), t['mlp_b2'][0].data_ptr(), t['layer_scale'][0].data_ptr(), eng.N * 16, eng.h.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
p = buf.cpu().numpy().reshape(6, 2, 16)
t0 = p[0, 1, 0]
nm = ['wait_a', 'a_ok', 'g1_0', 'g1_1', 'g1_2', 'g1_3', 'g2_0', 'g2_1', 'g2_2', 'g2_3']
ne = ['start', 'E0_rdy', 'E0_done', 'E1_rdy', 'E1_done', 'E2_rdy', 'E2_done', 'E3_rdy', 'E3_done', 'hv_issued', 'd2_full', 'fin_done']
for it in range(1, 4):
    print(f'tile {it}:')
    print('  MMA :', ' '.join(f'{n}={p[it,0,i]-t0}' for i, n in enumerate(nm)))
    print('  EPI :', ' '.join(f'{n}={p[it,1,i]-t0}' for i, n in enumerate(ne)))
