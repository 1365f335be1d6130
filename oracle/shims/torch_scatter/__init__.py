import torch

def scatter_add(src, index, dim=0, dim_size=None):
    assert dim == 0
    dim_size = int(index.max().item()) + 1 if dim_size is None else dim_size
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)

def scatter_mean(*a, **k):
    raise NotImplementedError
