class BaseTransform:
    def __init__(self):
        pass
    def __call__(self, data):
        return data

class Compose(BaseTransform):
    def __init__(self, transforms):
        self.transforms = transforms
    def __call__(self, data):
        for t in self.transforms:
            data = t(data)
        return data

class RadiusGraph(BaseTransform):
    def __init__(self, *a, **k):
        raise NotImplementedError("RadiusGraph is not reached on the diffusion path")
