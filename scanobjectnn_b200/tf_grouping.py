"""Drop-in for pointnet2/tf_ops/grouping/tf_grouping.py: same function names and argument order."""
from .ops import group_point, knn_point, query_ball_point, select_top_k  # noqa: F401
