"""B200-native (sm_100a) exact-GP hot path of Scale-Mixtures-of-NNGPs: Gram + Cholesky + Student-t LML /
predictive behind the reference's spax kernel / model / likelihood API.  See DESIGN.md."""
from . import _lib, device, nt_kernels, spax
from .device import StackSpec, make_hp
from .nt_kernels import get_mlp_kernel, get_dense_resnet_kernel

__all__ = ["_lib", "device", "nt_kernels", "spax", "StackSpec", "make_hp", "get_mlp_kernel",
           "get_dense_resnet_kernel"]
