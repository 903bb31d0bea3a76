"""Drop-in mirror of the reference's `defense` package (ConvONet/defense/__init__.py:1-3,
ONet/defense/__init__.py) on top of the sm_100a kernels."""
from .SOR import SORDefense
from .pn_utils import (fps_points, farthest_point_sample, index_points, knn_point, query_ball_point,
                       dgcnn_knn)
from .repulsion_loss import RepulsionLoss, repulsion_loss

__all__ = ["SORDefense", "fps_points", "farthest_point_sample", "index_points", "knn_point",
           "query_ball_point", "dgcnn_knn", "RepulsionLoss", "repulsion_loss"]
