class Box:
    """Records its arguments; nothing on the env step path reads them."""

    def __init__(self, low, high, shape=None, dtype=None):
        self.low = low
        self.high = high
        self.shape = shape if shape is not None else getattr(low, "shape", None)
        self.dtype = dtype
