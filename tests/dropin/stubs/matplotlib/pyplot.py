"""bts_test.py saves log-depth colour maps with plt.imsave; the stub writes an 8-bit grey PNG with cv2."""
import numpy as np


def imsave(fname, arr, cmap=None, **k):
    import cv2
    a = np.nan_to_num(np.asarray(arr, dtype=np.float32), nan=0.0, posinf=0.0, neginf=0.0)
    lo, hi = float(a.min()), float(a.max())
    g = ((a - lo) / (hi - lo + 1e-12) * 255).astype(np.uint8)
    cv2.imwrite(fname, g)
