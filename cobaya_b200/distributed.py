"""
One process per GPU under ``torchrun``: make the reference's own multi-process layer
(``cobaya/mpi.py``) run on top of ``torch.distributed``.

The reference parallelises with one chain per MPI rank and drives everything that is
per-process through ``cobaya.mpi``: root-only writers of ``.checkpoint/.covmat/.progress``
and ``updated.yaml`` (``output.py:432-536``, ``mcmc.py:1045-1078``), the per-rank chain file
``prefix.<rank+1>.txt`` (``mcmc.py:142``), ``mpi_size`` stored for resuming
(``mcmc.py:131-139``), one ``SeedSequence`` child per rank (``sampler.py:369-384``) and the
``ProcessState`` protocol that turns a failure on one rank into ``OtherProcessError`` on the
others (``mpi.py:350-467``).  The B200 engine runs one process per GPU with
``torch.distributed`` (NCCL) as its data plane and there is no ``mpi4py`` in the image, so
without this module every rank would believe it is the root.

``init()`` gives ``cobaya.mpi`` a ``COMM_WORLD`` with the subset of the mpi4py interface the
reference uses -- ``Get_rank/Get_size``, ``bcast/scatter/gather/allgather/barrier``,
``Ibarrier``, and the ``Isend/iprobe/Recv`` triple of ``ProcessState`` -- implemented on the
process group's key-value store (TCPStore): the control plane carries a few small pickles
per checkpoint and never touches the GPUs.  There is then ONE source of truth for
rank/size (``cobaya.mpi``), used by the reference's code and by ``cobaya_b200.plugin``
alike; the moments all-reduce stays on NCCL.

Usage (every rank, before ``cobaya.run.run``)::

    import cobaya_b200.distributed as cbd
    cbd.init()                     # init_process_group + cobaya.mpi on top of it
    from cobaya.run import run
    run(info)
"""

from __future__ import annotations

import os
import pickle
import time

import numpy as np

ANY_SOURCE = -1
_installed = None


class _Done:
    def Test(self):
        return True

    def Wait(self):
        return True


class _BarrierRequest:
    def __init__(self, comm, key):
        self.comm, self.key = comm, key

    def Test(self):
        return int(self.comm.store.add(self.key, 0)) >= self.comm.size

    def Wait(self):
        while not self.Test():
            time.sleep(self.comm.poll)
        return True


class Status:
    """mpi4py.MPI.Status: only the source is used (mpi.py:395-401)."""

    def __init__(self):
        self.source = ANY_SOURCE

    def Get_source(self):
        return self.source


class _Pickle:
    """``MPI.pickle.__init__(dumps, loads)`` is how the reference switches to dill
    (mpi.py:66-72)."""

    def __init__(self, dumps=pickle.dumps, loads=pickle.loads):
        self.dumps, self.loads = dumps, loads


class StoreComm:
    """The part of ``mpi4py.MPI.Comm`` that ``cobaya/mpi.py`` calls, on a c10d store.

    Collectives are numbered by a per-rank call counter (all ranks issue them in the same
    order, as MPI requires); a value is deleted by the last rank that reads it."""

    def __init__(self, store, rank: int, size: int, poll: float = 0.002):
        self.store, self.rank, self.size, self.poll = store, int(rank), int(size), poll
        self._n = 0
        self._sent: dict[tuple[int, int], int] = {}
        self._seen: dict[tuple[int, int], int] = {}
        self.pickle = _Pickle()

    # ---- mpi4py names
    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    # ---- store helpers
    def _wait_get(self, key: str) -> bytes:
        while not self.store.check([key]):
            time.sleep(self.poll)
        return self.store.get(key)

    def _read_shared(self, key: str, readers: int):
        data = self._wait_get(key)
        if int(self.store.add(key + "/r", 1)) >= readers:
            self.store.delete_key(key)
            self.store.delete_key(key + "/r")
        return self.pickle.loads(data)

    def _next(self, what: str) -> str:
        self._n += 1
        return f"{what}/{self._n}"

    # ---- collectives (object semantics of mpi4py's lower-case methods)
    def bcast(self, obj=None, root=0):
        key = self._next("bc")
        if self.rank == root:
            if self.size > 1:
                self.store.set(key, self.pickle.dumps(obj))
            return obj
        return self._read_shared(key, self.size - 1)

    def scatter(self, objs=None, root=0):
        key = self._next("sc")
        if self.rank == root:
            if objs is None or len(objs) != self.size:
                raise ValueError("scatter needs one item per process on the root")
            for r in range(self.size):
                if r != root:
                    self.store.set(f"{key}/{r}", self.pickle.dumps(objs[r]))
            return objs[root]
        return self._read_shared(f"{key}/{self.rank}", 1)

    def gather(self, obj, root=0):
        key = self._next("ga")
        if self.rank != root:
            self.store.set(f"{key}/{self.rank}", self.pickle.dumps(obj))
            return None
        return [obj if r == root else self._read_shared(f"{key}/{r}", 1)
                for r in range(self.size)]

    def allgather(self, obj):
        key = self._next("ag")
        self.store.set(f"{key}/{self.rank}", self.pickle.dumps(obj))
        return [obj if r == self.rank else self._read_shared(f"{key}/{r}", self.size - 1)
                for r in range(self.size)]

    def Ibarrier(self):
        key = self._next("ba")
        self.store.add(key, 1)
        return _BarrierRequest(self, key)

    def barrier(self):
        self.Ibarrier().Wait()

    Barrier = barrier

    # ---- the state messages of ProcessState (mpi.py:376-401): one int per message
    def Isend(self, buf, dest, tag=0):
        seq = self._sent.get((dest, tag), 0)
        self._sent[(dest, tag)] = seq + 1
        self.store.set(f"p2p/{tag}/{self.rank}/{dest}/{seq}", str(int(np.asarray(buf).ravel()[0])))
        return _Done()

    def _pending(self, source, tag):
        srcs = range(self.size) if source == ANY_SOURCE else [source]
        for s in srcs:
            if s == self.rank:
                continue
            seq = self._seen.get((s, tag), 0)
            if self.store.check([f"p2p/{tag}/{s}/{self.rank}/{seq}"]):
                return s, seq
        return None

    def iprobe(self, source=ANY_SOURCE, tag=0):
        return self._pending(source, tag) is not None

    Iprobe = iprobe

    def Recv(self, buf, source=ANY_SOURCE, tag=0, status=None):
        hit = self._pending(source, tag)
        while hit is None:
            time.sleep(self.poll)
            hit = self._pending(source, tag)
        s, seq = hit
        key = f"p2p/{tag}/{s}/{self.rank}/{seq}"
        buf[0] = int(self.store.get(key))
        self.store.delete_key(key)
        self._seen[(s, tag)] = seq + 1
        if status is not None:
            status.source = s

    def Abort(self, code=1):
        os._exit(int(code))


class _MPIModule:
    """Stands where ``from mpi4py import MPI`` would (mpi.py:59-72)."""

    ANY_SOURCE = ANY_SOURCE
    Status = Status

    def __init__(self, comm):
        self.COMM_WORLD = comm
        self.pickle = comm.pickle


def install(store, rank: int, size: int):
    """Point ``cobaya.mpi`` at a :class:`StoreComm`.  Returns the communicator."""
    global _installed
    from cobaya import mpi

    comm = StoreComm(store, rank, size)
    mpi._mpi = _MPIModule(comm)
    mpi._mpi_comm = comm
    mpi._mpi_size = size
    mpi._mpi_rank = rank
    try:
        import dill

        comm.pickle.__init__(dill.dumps, dill.loads)
    except ImportError:
        pass
    _installed = comm
    return comm


def installed():
    return _installed


def init(backend: str | None = None, device: int | None = None):
    """Initialise ``torch.distributed`` from the torchrun environment (if it is not yet)
    and install the communicator into ``cobaya.mpi``.  ``backend`` defaults to NCCL when a
    CUDA device is visible, else gloo (CPU tests).  A single process needs nothing."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 and not dist.is_initialized():
        return None
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        local = int(os.environ.get("LOCAL_RANK", "0")) if device is None else int(device)
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    from torch.distributed.distributed_c10d import _get_default_store

    store = dist.PrefixStore("cobaya_b200_mpi", _get_default_store())
    return install(store, dist.get_rank(), dist.get_world_size())


def world_consistent() -> bool:
    """True unless torch.distributed has several ranks that ``cobaya.mpi`` does not know of."""
    import sys

    dist = sys.modules.get("torch.distributed")  # never import torch just to ask (seconds)
    if dist is None or not (dist.is_available() and dist.is_initialized()) or \
            dist.get_world_size() == 1:
        return True
    from cobaya import mpi

    return mpi.size() == dist.get_world_size() and mpi.rank() == dist.get_rank()
