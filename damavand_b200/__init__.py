"""damavand_b200 -- B200-native statevector engine behind damavand's `gpu` / `distributed_gpu`
apply methods.  `Circuit` is drop-in compatible with the reference's `damavand.Circuit`."""
from .circuit import Circuit, initialize_mpi
from ._lib import DamavandError

__all__ = ["Circuit", "initialize_mpi", "DamavandError"]
