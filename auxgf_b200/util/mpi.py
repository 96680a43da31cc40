"""Process-parallel helpers of the drivers -- the role ``auxgf/util/mpi.py:25-69`` plays in the reference
(``mpi.rank``, ``mpi.size``, ``mpi.reduce`` = Reduce + Bcast, ``mpi.barrier``), without MPI: one process per GPU
(``torchrun`` / any launcher that sets ``RANK``, ``WORLD_SIZE``, ``LOCAL_RANK``), the collective is the NCCL
communicator owned by ``libagf2_b200.so`` (``agf2_b200_comm_*``), over NVLink / NVSwitch.

The 128-byte NCCL unique id is the only thing exchanged outside NCCL: rank 0 writes it to a file in the temporary
directory of the node (single-node jobs: the 8 GPUs of one box), the other ranks read it.

``use(comm)`` installs another communicator object (``rank``, ``size``, ``reduce(array)``, ``barrier()``): the
CPU-only test-suite injects a gloo-backed one to exercise the multi-rank driver logic without GPUs.
"""
import os
import tempfile
import time

import numpy as np

_comm = None
_seq = 0


class _Serial:
    rank, size = 0, 1

    def reduce(self, m):
        return m

    def barrier(self):
        pass


def _exchange_id(make_id, rank, timeout=300.0):
    """rank 0: make_id() -> bytes, published through a file; others: wait for it."""
    global _seq
    key = "%s_%s_%d_%d" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "none"),
                           os.getppid(), _seq)
    _seq += 1
    path = os.path.join(tempfile.gettempdir(), "agf2_b200_ncclid_" + key)
    if rank == 0:
        uid = make_id()
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(uid)
        os.replace(tmp, path)          # atomic: readers never see a partial file
        return uid, path
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise RuntimeError("timed out waiting for the NCCL unique id of rank 0 (%s)" % path)
        time.sleep(0.02)
    with open(path, "rb") as f:
        return f.read(), path


class NcclComm:
    """The communicator of the library's default context on this process's GPU."""

    def __init__(self, rank, size, local_rank):
        from auxgf_b200.lib import agf2 as lib
        self._lib = lib
        self.rank, self.size = rank, size
        lib.set_device(local_rank)
        try:
            import torch
            torch.cuda.set_device(local_rank)
        except ImportError:
            pass
        uid, path = _exchange_id(lib.comm_unique_id, rank)
        lib.comm_init(uid, rank, size)      # collective: returns once every rank has joined
        lib.comm_barrier()
        if rank == 0:
            try:
                os.remove(path)
            except OSError:
                pass

    def reduce(self, m):
        """Sum over the ranks, result on every rank (Reduce + Bcast of auxgf/util/mpi.py:60-64).  CUDA tensors are
        reduced in place; NumPy arrays travel through the GPU."""
        return self._lib.comm_allreduce(m)

    def barrier(self):
        self._lib.comm_barrier()

    def close(self):
        self._lib.comm_destroy()


def init():
    """Communicator of this process, created on first use from the launcher's environment."""
    global _comm
    if _comm is None:
        size = int(os.environ.get("WORLD_SIZE", "1"))
        if size > 1:
            _comm = NcclComm(int(os.environ["RANK"]), size, int(os.environ.get("LOCAL_RANK", os.environ["RANK"])))
        else:
            _comm = _Serial()
    return _comm


def use(comm):
    """Install a communicator object (test-suite / embedding applications).  Returns the previous one."""
    global _comm
    prev, _comm = _comm, comm
    return prev


def finalize():
    global _comm
    if _comm is not None and hasattr(_comm, "close"):
        _comm.close()
    _comm = None


def rank():
    return init().rank


def size():
    return init().size


def reduce(m):
    return init().reduce(m)


def barrier():
    init().barrier()


def split_int(n, size=None):
    """Contiguous near-equal split of range(n) (auxgf/util/mpi.py:128): list of (start, stop) per rank."""
    size = init().size if size is None else size
    base, rem = divmod(n, size)
    out, lo = [], 0
    for r in range(size):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def reduce_host(m):
    """Convenience for NumPy arrays with any communicator."""
    return np.asarray(reduce(np.ascontiguousarray(m, dtype=np.float64)))
