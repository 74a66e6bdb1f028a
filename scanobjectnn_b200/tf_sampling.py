"""Drop-in for pointnet2/tf_ops/sampling/tf_sampling.py: same function names and argument order, torch CUDA
tensors instead of TF tensors.  (prob_sample is not on the point-set-abstraction path and is not provided.)"""
from .ops import farthest_point_sample, gather_point  # noqa: F401
