import torch

def is_sparse(x):
    return False

def is_torch_sparse_tensor(x):
    return False

def to_edge_index(x):
    raise NotImplementedError

def coalesce(*a, **k):
    raise NotImplementedError

def remove_self_loops(*a, **k):
    raise NotImplementedError

def add_self_loops(*a, **k):
    raise NotImplementedError
