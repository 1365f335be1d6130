import numpy as np, torch, sys, types
sys.path.insert(0, '.')
torch.set_default_dtype(torch.float64)
from arreau_b200.engine import DenoiseEngine
from arreau_b200.synthetic import calibrate_length_readout
from arreau_b200.tables import build_tables
from arreau_b200.weights import PonitaWeights
from oracle import restatement as R
dev = torch.device('cuda')
w = np.load('tests/golden/weights_seed0.npz')
sd = {k: w[k] for k in w.files if k not in ('ori_grid', 'fourier_w')}
s = np.load('tests/golden/sample_T11.npz')
n_per, G = int(s['n_per']), int(s['num_crystals'])
sdc = calibrate_length_readout(sd, n_per)
pw = PonitaWeights(sdc, w['ori_grid'], device=dev)
eng = DenoiseEngine(pw, build_tables(11, 90), w['fourier_w'], [n_per]*G, 5.0, 8, device=dev, debug=True)
rel = lambda a, b: float(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-30))
T64 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)
W = R.PonitaWeights({k: T64(v) for k, v in sdc.items()}, T64(w['ori_grid']), 5.0)
tabs = R.DiffusionTables.build(11, 90)
for k, timestep in [(0, 10), (5, 5)]:
    eng.set_state(s['step_frac'][k], s['step_types'][k], s['step_lengths'][k], s['angles'])
    score, logits, len0 = eng.predict_scores(timestep)
    torch.cuda.synchronize()
    frac, ty, le, an = T64(s['step_frac'][k]), torch.as_tensor(s['step_types'][k]), T64(s['step_lengths'][k]), T64(s['angles'])
    na = torch.full((G,), n_per)
    N = G*n_per
    t = torch.full((N,), timestep)
    lat = R.lattice_from_params(le, an)
    rep = lambda a: torch.repeat_interleave(a, na, dim=0)
    tt = tabs.vp_betas[t].view(-1, 1)
    x = torch.cat([torch.nn.functional.one_hot(ty, 90), R.fourier_time_embedding(tt, T64(w['fourier_w'])), rep(na).unsqueeze(-1), rep(le), rep(an), rep((le/na.unsqueeze(-1)).abs())], dim=1)
    vec = torch.cat([frac.unsqueeze(1), rep(lat)], dim=1)
    cart = R.frac_to_cart_coords(frac, lat, na)
    batch = torch.repeat_interleave(torch.arange(G), na)
    ei, co, nimg, dist, direction = R.radius_graph_pbc(cart, lat, na, 5.0, 8)
    E = eng.num_edges()
    print('step', k, 'E', E, ei.shape[1], 'edges equal', np.array_equal(eng.src[:E].cpu().numpy(), ei[0].numpy()))
    ol, ov, og, inter = R.ponita_forward(W, x, vec, ei, dist, direction, lat, batch, G, out_dims=(90,1,0,3), return_intermediates=True)
    print(' x', rel(eng.x.cpu(), x), 'vec', rel(eng.vec.cpu(), vec), 'lat', rel(eng.lattice.cpu(), lat), 'pos', rel(eng.pos.cpu(), cart))
    print(' h0', rel(eng.h_debug[0].cpu(), inter['h0']))
    for l in range(5):
        kb = inter['kernel_basis']
        kern = kb @ W[f'interaction_layers.{l}.conv.kernel.weight'].T
        print(' l', l, 'kern', rel(eng.kernels[l,:E].float().cpu(), kern), 'x1', rel(eng.x1_debug[l].cpu(), inter[f'x1_{l}']), 'x2', rel(eng.x2_debug[l].cpu(), inter[f'x2_{l}']), 'h', rel(eng.h_debug[l+1].cpu(), inter[f'h_{l}']),
              'max|kern|', float(kern.abs().max()), 'max|x1|', float(inter[f'x1_{l}'].abs().max()))
    print(' score', rel(score.cpu(), ov.squeeze(1)), 'logits', rel(logits.cpu(), ol), 'len0', rel(len0.cpu(), og))
    print(' golden score', rel(score.cpu(), s['step_score'][k]), rel(ov.squeeze(1), s['step_score'][k]))
    a = inter['attr']; print(' attr max', a.abs().amax(dim=(0,1)))

# mirror path
f = np.load('tests/golden/forward_c1_t500.npz')
from arreau_b200.ponita.models.ponita import PonitaFiberBundle
m = PonitaFiberBundle((164, 4), 128, 90, 3, 0, 0, 5, output_dim_vec=1, radius=5.0, num_ori=16, basis_dim=256, degree=3, widening_factor=4, layer_scale=1e-6, multiple_readouts=True, ori_grid=w['ori_grid'])
print(m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}))
m = m.to(dev)
g = types.SimpleNamespace(x=torch.as_tensor(f['x'], device=dev), vec=torch.as_tensor(f['vec'], device=dev), edge_index=torch.as_tensor(f['edge_index'], device=dev), dists=torch.as_tensor(f['dist'], device=dev), inter_atom_direction=torch.as_tensor(f['direction'], device=dev), lattice=torch.as_tensor(f['lattice'], device=dev), batch=torch.as_tensor(f['batch'], device=dev))
lo, ve, l0, _, _ = m(g)
print('mirror: logits', rel(lo.cpu(), f['logits']), 'vec', rel(ve.cpu(), f['vec_out']), 'len0', rel(l0.cpu(), f['len0']))
for k in ['basis_fn.1.weight', 'interaction_layers.0.layer_scale', 'read_out_layers.0.weight', 'interaction_layers.2.norm.weight']:
    print(k, rel(m.state_dict()[k].cpu(), sd[k]), m.state_dict()[k].dtype)
