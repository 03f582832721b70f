"""imageio.imread stand-in (scene/pose_optimizer.py:345): PIL underneath."""
import numpy as np
from PIL import Image


def imread(path, *a, **k):
    return np.array(Image.open(path))


def imwrite(path, arr, *a, **k):
    Image.fromarray(np.asarray(arr)).save(path)
