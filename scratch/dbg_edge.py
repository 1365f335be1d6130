import numpy as np, torch, sys
sys.path.insert(0, '.')
torch.set_default_dtype(torch.float64)
from arreau_b200.engine import DenoiseEngine
from arreau_b200.tables import build_tables
from arreau_b200.weights import PonitaWeights
from oracle import restatement as R
dev = torch.device('cuda')
w = np.load('tests/golden/weights_seed0.npz')
sd = {k: w[k] for k in w.files if k not in ('ori_grid', 'fourier_w')}
s = np.load('tests/golden/sample_T11.npz')
n_per, G = int(s['n_per']), int(s['num_crystals'])
pw = PonitaWeights(sd, w['ori_grid'], device=dev)
eng = DenoiseEngine(pw, build_tables(11, 90), w['fourier_w'], [n_per]*G, 5.0, 8, device=dev, debug=True)
T64 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)
W = R.PonitaWeights({k: T64(v) for k, v in sd.items()}, T64(w['ori_grid']), 5.0)
k = 0
eng.set_state(s['step_frac'][k], s['step_types'][k], s['step_lengths'][k], s['angles'])
eng.predict_scores(10); torch.cuda.synchronize()
E = eng.num_edges()
src = eng.src[:E].cpu().long(); dist = eng.dist[:E].cpu(); direction = eng.dir[:E].cpu(); lat = eng.lattice.cpu()
batch = torch.repeat_interleave(torch.arange(G), torch.full((G,), n_per))
ori = W.ori_grid
rel = direction[:, None, :]
inv1 = (rel * ori[None]).sum(-1, keepdim=True)
inv2 = (rel - inv1 * ori[None]).norm(dim=-1, keepdim=True)
lat_e = lat[batch[src]]
cs = [R._cosine_similarity(direction, lat_e[:, m, :]) for m in range(3)]
es = torch.stack([dist, cs[0], cs[1], cs[2]], dim=-1)
attr = torch.cat([inv1, inv2, es[:, None, :].expand(-1, 16, -1)], dim=-1)
def mlp(a, dt):
    h = R.polynomial_features(a.to(dt), 3)
    h = R.gelu(h @ W['basis_fn.1.weight'].to(dt).T + W['basis_fn.1.bias'].to(dt))
    return R.gelu(h @ W['basis_fn.3.weight'].to(dt).T + W['basis_fn.3.bias'].to(dt))
kb64 = mlp(attr, torch.float64) * R.polynomial_cutoff(dist, 5.0)[:, None, None]
kb32 = (mlp(attr, torch.float32) * R.polynomial_cutoff(dist, 5.0).float()[:, None, None]).double()
print('fp32 torch vs fp64 kb', float((kb32-kb64).abs().max()/kb64.abs().max()))
wk = W['interaction_layers.0.conv.kernel.weight']
kern = kb64 @ wk.T
got = eng.kernels[0, :E].float().cpu().double()
err = (got - kern).abs()
print('kern rel', float(err.max()/kern.abs().max()))
pe = err.amax(dim=(1,2)); po = err.amax(dim=(0,2))
print('per-o err', po.numpy().round(4))
worst = torch.argsort(pe, descending=True)[:10]
for e in worst.tolist():
    print('edge', e, 'err', float(pe[e]), 'dist', float(dist[e]), 'dir', direction[e].numpy().round(4), 'attr o0', attr[e,0].numpy().round(4), 'win', float(R.polynomial_cutoff(dist[e:e+1], 5.0)))
print('frac of edges with err>1e-4:', float((pe > 1e-4*kern.abs().max()).double().mean()))
good = torch.argsort(pe)[:5]
for e in good.tolist():
    print('good edge', e, 'err', float(pe[e]), 'dist', float(dist[e]), 'attr o0', attr[e,0].numpy().round(4))
print('pre-activation max', float((R.polynomial_features(attr,3) @ W['basis_fn.1.weight'].T).abs().max()))
