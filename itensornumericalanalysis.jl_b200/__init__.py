"""B200-native batched evaluation of ITensorNetworkFunctions (quantics MPS / tree tensor
networks): host-side mirror of the reference's evaluate path on top of libttneval.so.

Import as `import itna_b200` (the directory name carries a dot, see itna_b200.py at the repo
root).  Layout:
  csrc/      CUDA kernels (sm_100a) + the C ABI declared in include/ttneval.h
  julia/     the Julia method + packer a maintainer adds to the reference (ccall binding)
  *.py       Python mirror of the reference interface for this path (test / bench harness)
"""
from .graphs import (NamedGraph, named_grid, named_comb_tree, named_binary_tree,
                     named_path_graph, uniform_tree, vertices, is_tree)
from .indexmaps import (Index, IndsNetwork, RealIndexMap, ComplexIndexMap, IndsNetworkMap,
                        RealIndsNetworkMap, ComplexIndsNetworkMap, continuous_siteinds,
                        real_continuous_siteinds, complex_continuous_siteinds,
                        default_dimension_vertices, digit_siteinds, complex_digit_siteinds)
from .network import Tensor, TensorNetwork, random_tensornetwork, add, multiply
from .itensornetworkfunction import (ITensorNetworkFunction, evaluate, evaluate_indices, batched_ind_values, Plan,
                                     default_contraction_alg)
from .elementary_functions import (const_itn, exp_itn, cosh_itn, sinh_itn, tanh_itn, cos_itn,
                                   sin_itn, rand_itn, delta_p, const_itensornetwork,
                                   exp_itensornetwork, cosh_itensornetwork, sinh_itensornetwork,
                                   tanh_itensornetwork, cos_itensornetwork, sin_itensornetwork,
                                   random_itensornetwork)
from .packer import pack, PackedNetwork
from .ttn_io import save_ttn, load_ttn
from .integration import partial_integrate, integrate
from . import _capi
