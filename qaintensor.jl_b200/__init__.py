"""qaintensor_b200 -- B200-native drop-in for the contraction / truncation hot path of
Qaintensor.jl.  The public names mirror src/Qaintensor.jl:50-101; all arithmetic runs
in ``lib/libqaintensor_cuda.so`` (hand-written sm_100a CUDA behind a C ABI)."""
from . import _lib  # noqa: F401  (fails loudly if the library is not built)
from ._lib import QtnError, launch_count  # noqa: F401
from .tensor_network import (GeneralTensorNetwork, Summation, Tensor, TensorNetwork,  # noqa: F401
                             is_power_two, shift_pair, shift_summation)
from .contract import (ContractionPlan, choose_slices, contract, contract_order, contract_rep,  # noqa: F401
                       ncon, permutedims, search_order)
from .network2graph import (contraction_order, line_graph, network_graph,  # noqa: F401
                            optimize_contraction_order, tree_decomposition_width)
from .svd import contract_svd, svd, svd_trunc  # noqa: F401
from .gates import CircuitGate, circuit_gate, qft_circuit  # noqa: F401
from .tensor_circuit import decompose, tensor_circuit  # noqa: F401
from .mps import (MPS, ClosedMPS, OpenMPS, PeriodicMPS, check_mps, contract_svd_mps, permute, switch)  # noqa: F401
from .mpo import MPO, apply_MPO, extend_MPO  # noqa: F401
from .mps_sim import DeviceMPS, brickwork_layer_sites, tfi_mpo  # noqa: F401
from .native_network import NativeNetwork  # noqa: F401
from . import circuits, gates  # noqa: F401
