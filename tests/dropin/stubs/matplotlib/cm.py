import numpy as np


def get_cmap(name=None):
    def cmapper(v, bytes=False):
        v = np.asarray(v, dtype=np.float32)
        out = np.stack([v, v, v, np.ones_like(v)], -1)
        return (out * 255).astype(np.uint8) if bytes else out
    return cmapper
