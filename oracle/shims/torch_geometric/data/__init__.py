class Data:
    """Attribute bag (PyG Data/Batch are only used as such on the path,
    diffusion_loss.py:330-335)."""
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

class Batch(Data):
    pass
