"""Tensor-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 /
NVSwitch) for bootstrap.

Replaces the reference's GPU GEMM -> .to('cpu') -> deepspeed_comm.all_reduce (oneCCL) ->
.to('cuda') bounce (decoder.py:60-77).  Two data paths:
  * fused (default on GPUs): ``PeerArena`` -- one peer-mapped device arena per rank (CUDA IPC over
    NVLink); the row-parallel projection, its all-reduce and the residual add are ONE kernel
    (``ops.gemm_allreduce`` -> lia_gemm_allreduce_bf16) that exchanges partial tiles through the
    arenas while the remaining tiles are still being computed.
  * plain (``LIA_TP_FUSED=0``, and the CPU/gloo tests): GEMM -> in-place ``all_reduce`` on the
    compute stream -> residual add.
Sharding rule: weights.shard_layer (tensor_parallel.py:30-141).
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib, graphs

_group = None


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's env (RANK/WORLD_SIZE/MASTER_*).
    Returns (rank, world).  A world of 1 needs no process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def set_group(group):
    global _group
    _group = group


def world_size():
    return dist.get_world_size(_group) if dist.is_initialized() else 1


def all_reduce(t):
    """In-place sum over the tensor-parallel group (bf16, like the reference's message dtype)."""
    if dist.is_initialized() and dist.get_world_size(_group) > 1 and os.environ.get("LIA_TP_NO_WAIT", "0") == "0":
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_group)
    return t


def barrier():
    if dist.is_initialized():
        dist.barrier(group=_group)


def max_over_ranks(value, device):
    """Max of a python float over ranks (timing rule: device time, max over ranks)."""
    if not dist.is_initialized() or dist.get_world_size(_group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=_group)
    return float(t.item())


def fused_enabled():
    """Fused projection + all-reduce over peer memory (default) vs GEMM -> NCCL all-reduce -> add."""
    return os.environ.get("LIA_TP_FUSED", "1") != "0"


class _RawCuda:
    """__cuda_array_interface__ carrier: lets torch view arena bytes without owning them."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


def _align(n, a=256):
    return (int(n) + a - 1) // a * a


class PeerArena:
    """One symmetric device arena per rank, mapped into every peer (lia_p2p_*; CUDA IPC).

    Layout (identical on every rank): ``ctl`` (epoch/flags) | receive area (2 parities) | named
    activation buffers the prefill path needs to be remotely writable.  ``handles`` are exchanged
    with ``all_gather_object`` on the (NCCL or gloo) process group -- plumbing only; the data path
    is the kernel's own loads/stores over NVLink."""

    def __init__(self, rank, world, device, recv_bytes, buffers, group=None, exchange=None):
        lib = _lib.load()
        self.rank, self.world, self.device = rank, world, torch.device(device)
        self.ctl_off = 0
        self.recv_off = _align(lib.lia_tp_ctl_bytes())
        self.recv_bytes = _align(recv_bytes)
        off = self.recv_off + 2 * self.recv_bytes
        self.offsets = {}
        for name, nbytes in buffers:
            self.offsets[name] = (off, int(nbytes))
            off = _align(off + nbytes)
        self.nbytes = off
        self.mc = None                           # multicast mapping of the arena (SymmArena)
        self._alloc(group, exchange)
        self._bytes = torch.as_tensor(_RawCuda(self.local, self.nbytes), device=self.device)

    def _alloc(self, group, exchange):
        """cudaMalloc + CUDA IPC (lia_p2p_*): every peer maps this rank's arena; handles travel through torch.distributed."""
        lib = _lib.load()
        rank, world = self.rank, self.world
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_uint8 * _lib.P2P_HANDLE_BYTES)()
        with torch.cuda.device(self.device):
            _lib.check(lib.lia_p2p_alloc(self.nbytes, ctypes.byref(ptr), handle), "lia_p2p_alloc")
        self.local = ptr.value
        self.peers = [None] * world
        self.peers[rank] = self.local
        mine = bytes(handle)
        if exchange is None:
            gathered = [None] * world
            dist.all_gather_object(gathered, mine, group=group)
        else:
            gathered = exchange(mine)            # tests: single-process stand-in
        self._opened = []
        for r in range(world):
            if r == rank:
                continue
            if isinstance(gathered[r], int):     # same-process stand-in: already a device pointer
                self.peers[r] = gathered[r]
                continue
            h = (ctypes.c_uint8 * _lib.P2P_HANDLE_BYTES).from_buffer_copy(gathered[r])
            pp = ctypes.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(lib.lia_p2p_open(h, ctypes.byref(pp)), "lia_p2p_open")
            self.peers[r] = pp.value
            self._opened.append(pp.value)

    def tensor(self, name, shape, dtype=torch.bfloat16):
        off, nbytes = self.offsets[name]
        t = self._bytes[off:off + nbytes].view(dtype)
        n = 1
        for s_ in shape:
            n *= s_
        return t[:n].view(*shape)

    def offset_of(self, t):
        off = t.data_ptr() - self.local
        if not (0 <= off < self.nbytes):
            raise _lib.LiaError("tensor does not live in the peer arena")
        return off

    def args(self, out=None):
        """LiaTpArgs for one call; ``out`` (an arena tensor) is needed for M > 128."""
        a = _lib.LiaTpArgs()
        a.rank, a.world = self.rank, self.world
        for r in range(self.world):
            a.arena[r] = self.peers[r]
        a.ctl_off, a.recv_off, a.recv_bytes = self.ctl_off, self.recv_off, self.recv_bytes
        a.out_off = self.offset_of(out) if out is not None else 0
        a.mc_arena = self.mc
        return a

    def check(self):
        """Raise if a kernel gave up waiting for a peer (synchronises the device)."""
        rc = _lib.load().lia_tp_error(ctypes.byref(self.args()))
        if rc != 0:
            raise _lib.LiaError(_lib.last_error())

    def close(self, sync=True):
        """Unmap the peers and free the local arena.  ``sync``: barrier first so that no peer still
        uses this rank's memory (skipped in __del__, where a collective could hang at exit)."""
        lib = _lib.load()
        if getattr(self, "local", None):
            torch.cuda.synchronize(self.device)
            if sync and dist.is_initialized() and self._opened:
                dist.barrier()
            self._bytes = None
            for p in self._opened:
                lib.lia_p2p_close(p)
            self._opened = []
            lib.lia_p2p_free(self.local)
            self.local = None

    def __del__(self):
        # cudaFree / cudaDeviceSynchronize inside an open graph capture would invalidate it: park the free (graphs.py)
        if getattr(self, "local", None):
            graphs.finalize(lambda: self.close(sync=False))



class SymmArena(PeerArena):
    """The same arena on SYMMETRIC memory bound to an NVLink-switch multicast object: besides every peer's mapping there is
    ONE multicast address behind which all ranks' copies sit, so the prefill exchange can reduce inside the switch
    (``multimem.ld_reduce``) and deliver with one store (``multimem.st``).  Allocation and rendezvous are
    ``torch.distributed._symmetric_memory`` (CUDA VMM + fabric handles + cuMulticast*): plumbing, like the IPC exchange of
    ``PeerArena``; the data path is the kernel's own multimem instructions.  ``available()`` says whether this process group
    can have one (NVSwitch + a driver that exposes multicast); otherwise ``PeerArena`` is used."""

    _probe = None

    @staticmethod
    def available(device):
        if os.environ.get("LIA_TP_SYMM", "1") == "0" or not dist.is_initialized() or dist.get_backend() != "nccl":
            return False
        if SymmArena._probe is None:
            ok = False
            try:
                import torch.distributed._symmetric_memory as symm_mem
                t = symm_mem.empty(4096, dtype=torch.uint8, device=torch.device(device))
                h = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
                ok = int(h.multicast_ptr) != 0
                SymmArena._keep_probe = (t, h)
            except Exception:
                ok = False
            flag = torch.tensor([1 if ok else 0], device=torch.device(device))
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)          # every rank must agree
            SymmArena._probe = bool(flag.item())
        return SymmArena._probe

    def _alloc(self, group, exchange):
        import torch.distributed._symmetric_memory as symm_mem
        self._symm = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=self.device)
        self._symm.zero_()
        torch.cuda.synchronize(self.device)
        self._hdl = symm_mem.rendezvous(self._symm, dist.group.WORLD.group_name)
        self.local = self._symm.data_ptr()
        self.peers = [int(p) for p in self._hdl.buffer_ptrs]
        self.peers[self.rank] = self.local
        self.mc = int(self._hdl.multicast_ptr) or None
        self._opened = []
        dist.barrier()                            # every rank's arena is zeroed before anyone's kernel touches it

    def close(self, sync=True):
        if getattr(self, "local", None):
            torch.cuda.synchronize(self.device)
            if sync and dist.is_initialized():
                dist.barrier()
            self._bytes = None
            self.local = None
            self._hdl = None
            self._symm = None
