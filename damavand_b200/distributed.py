"""Rank bootstrap for apply_method="distributed_gpu": one process per GPU.

The reference bootstraps MPI by importing mpi4py in user code (README.md:51-52) or through
``damavand.initialize_mpi()`` (src/lib.rs:16-19) and then talks rsmpi.  Here torch.distributed is
the plumbing (torchrun sets RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*); it only carries the
128-byte NCCL id and a few host-side broadcasts -- amplitudes move through the library's own NCCL
communicator.  If mpi4py happens to be importable and no torch process group exists, its
rank/size are used to create one.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def is_initialized() -> bool:
    try:
        dist = _dist()
    except Exception:
        return False
    return dist.is_available() and dist.is_initialized()


def initialize(backend: Optional[str] = None) -> Tuple[int, int]:
    """Create the torch.distributed process group from the environment if there is none yet.
    Returns (rank, world).  A single process without RANK/WORLD_SIZE is world 1."""
    if is_initialized():
        dist = _dist()
        return dist.get_rank(), dist.get_world_size()
    if "RANK" not in os.environ or "WORLD_SIZE" not in os.environ:
        try:  # optional mpi4py bootstrap, as in the reference's examples
            from mpi4py import MPI  # type: ignore
            comm = MPI.COMM_WORLD
            os.environ.setdefault("RANK", str(comm.Get_rank()))
            os.environ.setdefault("WORLD_SIZE", str(comm.Get_size()))
            os.environ.setdefault("LOCAL_RANK", str(comm.Get_rank()))
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
        except Exception:
            return 0, 1
    world = int(os.environ["WORLD_SIZE"])
    if world == 1:
        return 0, 1
    import torch
    dist = _dist()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend == "nccl":
        torch.cuda.set_device(local_device())
    dist.init_process_group(backend=backend)
    return dist.get_rank(), dist.get_world_size()


def rank_world() -> Tuple[int, int]:
    if is_initialized():
        dist = _dist()
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def local_device() -> int:
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    return int(os.environ.get("RANK", "0"))


def broadcast_bytes(payload: Optional[bytes], nbytes: int, src: int = 0) -> bytes:
    """Broadcast a fixed-size byte string from `src` over the torch process group."""
    rank, world = rank_world()
    if world == 1:
        assert payload is not None
        return payload
    import torch
    dist = _dist()
    on_gpu = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    if rank == src:
        t = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


def broadcast_array(arr: Optional[np.ndarray], shape, dtype, src: int = 0) -> np.ndarray:
    """Broadcast a numpy array (float64 / int64) from `src`."""
    rank, world = rank_world()
    if world == 1:
        return np.asarray(arr, dtype=dtype)
    import torch
    dist = _dist()
    on_gpu = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    if rank == src:
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).to(dev)
    else:
        t = torch.zeros(tuple(shape), dtype=torch.from_numpy(np.zeros(1, dtype=dtype)).dtype, device=dev)
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def barrier() -> None:
    if is_initialized():
        _dist().barrier()
