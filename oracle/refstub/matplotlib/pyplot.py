"""TEST INFRASTRUCTURE ONLY — `plt.get_cmap(name)` -> callable mapping [0, 1] arrays to RGBA, enough for the reference's
summary writer (src/summary/diffusion_dcbase_summary.py:8, used only with --save_image)."""
import numpy as np


def get_cmap(name="plasma"):
    def cmap(x):
        x = np.clip(np.asarray(x, dtype=np.float64), 0.0, 1.0)
        return np.stack([x, 1.0 - np.abs(2.0 * x - 1.0), 1.0 - x, np.ones_like(x)], axis=-1)
    return cmap
