import torch

class LightningModule(torch.nn.Module):
    def save_hyperparameters(self, *a, **k):
        pass
    def log(self, *a, **k):
        pass

class Trainer:
    pass
