"""Drop-in import shim: `from damavand import Circuit` resolves to the B200-native engine
(the reference's PyO3 module exposes exactly `Circuit` and `initialize_mpi`, src/lib.rs:21-26)."""
from damavand_b200 import Circuit, initialize_mpi

__all__ = ["Circuit", "initialize_mpi"]
