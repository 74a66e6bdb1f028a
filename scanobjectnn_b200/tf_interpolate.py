"""Drop-in for pointnet2/tf_ops/3d_interpolation/tf_interpolate.py: same function names and argument order."""
from .ops import three_interpolate, three_nn  # noqa: F401
