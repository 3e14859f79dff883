class DualTransformer2DModel:
    def __init__(self, *a, **k):
        raise RuntimeError("DualTransformer2DModel is unreachable for SD configs")
