import torch

class Metric(torch.nn.Module):
    def add_state(self, name, default, dist_reduce_fx=None):
        self.register_buffer(name, default.clone())
