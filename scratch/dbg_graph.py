import numpy as np, torch, sys
sys.path.insert(0, '.')
from arreau_b200.diffusion.diffusion_helpers import radius_graph_pbc
z = np.load('tests/golden/graph_cases.npz')
cases = {}
for k in z.files:
    i, n = k.split('/', 1); cases.setdefault(i, {})[n] = z[k]
dev = torch.device('cuda')
for i in sorted(cases):
    c = cases[i]
    out = radius_graph_pbc(torch.as_tensor(c['cart'], device=dev), torch.as_tensor(c['lattice'], device=dev), torch.as_tensor(c['num_atoms'], device=dev), float(c['radius']), int(c['cap']))
    ei, off, nimg, dist, direction = [o.cpu().numpy() for o in out]
    same_e = ei.shape == c['edge_index'].shape and np.array_equal(ei, c['edge_index'])
    msg = f"{str(c['name']):28s} edges_same={same_e}"
    if same_e:
        nd = (dist != c['dist']).sum(); ndir = (direction != c['direction']).sum()
        msg += f" dist_mismatch={nd}/{dist.size} dir_mismatch={ndir}"
        if nd:
            j = np.nonzero(dist != c['dist'])[0][0]
            d2 = (direction[j]**2)
            msg += f" first: got {dist[j].hex()} ref {c['dist'][j].hex()} dir_eq={np.array_equal(direction[j], c['direction'][j])} d2 {((d2[0]+d2[1])+d2[2]).hex()} sqrt {np.sqrt((d2[0]+d2[1])+d2[2]).hex()} alt {np.sqrt(d2[0]+(d2[1]+d2[2])).hex()}"
    print(msg)
