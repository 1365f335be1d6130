import numpy as np, torch, sys, time
sys.path.insert(0, '.')
from arreau_b200 import _lib
from arreau_b200.weights import PonitaWeights, umma_tile_image
from arreau_b200.engine import DenoiseEngine
from arreau_b200.tables import build_tables
dev = torch.device('cuda')
w = np.load('tests/golden/weights_seed0.npz')
sd = {k: w[k] for k in w.files if k not in ('ori_grid', 'fourier_w')}
pw = PonitaWeights(sd, w['ori_grid'], device=dev)
rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
which = sys.argv[1] if len(sys.argv) > 1 else 'mlp'
s = torch.cuda.current_stream().cuda_stream
if which in ('mlp', 'all'):
    for R in (128, 1000, 128 * 300 + 16):
        g = torch.Generator().manual_seed(R)
        y = torch.randn(R, 128, generator=g)
        h0 = torch.randn(R, 128, generator=g)
        tiles = (R + 127) // 128
        ypad = torch.zeros(tiles * 128, 128); ypad[:R] = y
        yimg = torch.cat([umma_tile_image(ypad[t * 128:(t + 1) * 128].numpy()) for t in range(tiles)]).to(dev)
        l = 2
        t = pw.t
        h32 = h0.clone().to(dev); hbf = h0.clone().to(dev)
        yd = y.to(dev)
        _lib.call('arreau_convnext_mlp_f32', yd.data_ptr(), t['mlp_w1_t'][l].data_ptr(), t['mlp_b1'][l].data_ptr(), t['mlp_w2_t'][l].data_ptr(), t['mlp_b2'][l].data_ptr(), t['layer_scale'][l].data_ptr(), R, h32.data_ptr(), s)
        torch.cuda.synchronize()
        _lib.call('arreau_convnext_mlp_f16', yimg.data_ptr(), t['mlp_w_img'].data_ptr() + l * 8 * 32768, t['mlp_b1'][l].data_ptr(), t['mlp_b2'][l].data_ptr(), t['layer_scale'][l].data_ptr(), R, hbf.data_ptr(), s)
        torch.cuda.synchronize()
        d32 = (h32 - h0.to(dev)); dbf = (hbf - h0.to(dev))
        print(f'mlp R={R}: rel err of update fp16 vs fp32 = {rel(dbf, d32):.3e}  max|update|={float(d32.abs().max()):.3f}', flush=True)
if which in ('edge', 'all'):
    st = np.load('tests/golden/steps_c1_T1000.npz')
    eng = DenoiseEngine(pw, build_tables(1000, 90), w['fourier_w'], st['num_atoms'], 5.0, 8, device=dev)
    eng.set_state(st['t500/frac'], st['t500/types'], st['t500/lengths'], st['angles'])
    eng.predict_scores(500); torch.cuda.synchronize()
    E = eng.num_edges(); cap = eng.edge_capacity
    k32 = eng.kernels[:, :E].clone()
    kbf = torch.zeros(5, cap, 16, 128, dtype=torch.float16, device=dev)
    t = pw.t
    _lib.call('arreau_edge_kernels_f16', eng.dir.data_ptr(), eng.dist.data_ptr(), eng.lattice.data_ptr(), eng.crystal_of_atom.data_ptr(), eng.src.data_ptr(), eng.row_ptr.data_ptr() + 4 * eng.N, cap, t['ori'].data_ptr(), t['edge_w1_img'].data_ptr(), t['edge_w_img'].data_ptr(), t['b2'].data_ptr(), 5.0, kbf.data_ptr(), s)
    torch.cuda.synchronize()
    for l in range(5):
        print(f'edge l={l}: E={E} rel err fp16 vs fp32 = {rel(kbf[l, :E].float(), k32[l]):.3e}', flush=True)
    print('rows beyond E untouched:', bool((kbf[:, E:] == 0).all()))
if which in ('fwd', 'all'):
    st = np.load('tests/golden/steps_c1_T1000.npz'); f = np.load('tests/golden/forward_c1_t500.npz')
    eng = DenoiseEngine(pw, build_tables(1000, 90), w['fourier_w'], st['num_atoms'], 5.0, 8, precision='fp16', device=dev)
    eng.set_state(st['t500/frac'], st['t500/types'], st['t500/lengths'], st['angles'])
    score, logits, len0 = eng.predict_scores(500); torch.cuda.synchronize()
    r = lambda a, b: float(np.abs(a.cpu().numpy().astype(np.float64) - b).max() / np.abs(b).max())
    print('fwd fp16: logits', r(logits, f['logits']), 'score', r(score, f['vec_out'][:, 0]), 'len0', r(len0, f['len0']), flush=True)
