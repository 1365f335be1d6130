import inspect
import torch


def global_add_pool(x, batch, size=None):
    """PyG global_add_pool: segment sum over `batch` (ponita.py:152)."""
    size = int(batch.max().item()) + 1 if size is None else size
    out = torch.zeros((size,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    return out.index_add_(0, batch, x)


def radius_graph(*a, **k):
    raise NotImplementedError("torch_cluster radius_graph is not reached on the diffusion path")


class _Inspector:
    def __init__(self, module):
        self.module = module

    def collect_param_data(self, name, coll_dict):
        fn = getattr(self.module, name)
        params = [p for p in inspect.signature(fn).parameters if p != "self"]
        return {p: coll_dict[p] for p in params if p in coll_dict}


class MessagePassing(torch.nn.Module):
    """Stand-in for PyG 2.5.2 MessagePassing with node_dim=0, aggr="add",
    flow="source_to_target": x_j = x[edge_index[0]]; aggregate = scatter-add over
    edge_index[1]; update = identity; no hooks.  Exposes exactly the attributes the
    reference's propagate2 copy touches (ponita/nn/conv.py:188-286)."""

    def __init__(self, aggr="add", node_dim=0, flow="source_to_target", **kwargs):
        super().__init__()
        assert aggr == "add" and node_dim == 0 and flow == "source_to_target"
        self.aggr = aggr
        self.node_dim = node_dim
        self.explain = False
        self.decomposed_layers = 1
        self.fuse = False
        self.inspector = _Inspector(self)
        msg_params = [p for p in inspect.signature(self.message).parameters]
        self._user_args = msg_params
        self._fused_user_args = []
        for name in ("_propagate_forward_pre_hooks", "_propagate_forward_hooks",
                     "_message_forward_pre_hooks", "_message_forward_hooks",
                     "_aggregate_forward_pre_hooks", "_aggregate_forward_hooks",
                     "_message_and_aggregate_forward_pre_hooks",
                     "_message_and_aggregate_forward_hooks"):
            setattr(self, name, {})

    def _check_input(self, edge_index, size):
        return [None, None] if size is None else list(size)

    def _collect(self, args, edge_index, size, kwargs):
        out = {}
        for arg in args:
            if arg.endswith("_j"):
                out[arg] = kwargs[arg[:-2]].index_select(0, edge_index[0])
                if size[0] is None:
                    size[0] = kwargs[arg[:-2]].size(0)
                if size[1] is None:
                    size[1] = kwargs[arg[:-2]].size(0)
            elif arg.endswith("_i"):
                out[arg] = kwargs[arg[:-2]].index_select(0, edge_index[1])
            else:
                out[arg] = kwargs.get(arg)
        out["index"] = edge_index[1]
        out["ptr"] = None
        out["dim_size"] = size[1]
        out["edge_index"] = edge_index
        return out

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        out = torch.zeros((dim_size,) + tuple(inputs.shape[1:]), dtype=inputs.dtype,
                          device=inputs.device)
        return out.index_add_(0, index, inputs)

    def update(self, inputs):
        return inputs

    def message(self, x_j):
        return x_j

    def propagate(self, edge_index, size=None, **kwargs):
        size = self._check_input(edge_index, size)
        coll = self._collect(self._user_args, edge_index, size, kwargs)
        msg = self.message(**self.inspector.collect_param_data("message", coll))
        out = self.aggregate(msg, **self.inspector.collect_param_data("aggregate", coll))
        return self.update(out)
