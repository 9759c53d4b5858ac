"""Test stub: tensorboardX is not installed in this image.  bts_main.py only needs SummaryWriter."""


class SummaryWriter:
    def __init__(self, *a, **k):
        self.scalars = []

    def add_scalar(self, tag, value, step=None):
        self.scalars.append((tag, float(value), step))

    def add_image(self, *a, **k):
        pass

    def flush(self):
        pass

    def close(self):
        pass
