"""Peer-memory gradient exchange of the data-parallel step (new component, the reference has no distributed
code): every rank's flat gradient buffer lives in cudaMalloc'ed memory shared through CUDA IPC, and ONE kernel
(nnb_allreduce_adam, csrc/nnb_collective.cu) sums the buffers over NVLink P2P loads and applies the three Adam updates.
torch.distributed is only used to move the 64-byte IPC handles once (plumbing)."""
import ctypes as C
import torch
from . import _lib as L


class _RawCuda:
    """exposes a raw device pointer through __cuda_array_interface__ so that torch can alias it as a tensor"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _ipc_alloc(nbytes):
    ptr = C.c_void_p(); h = (C.c_ubyte * 64)()
    L.check(L.lib.nnb_ipc_alloc(C.c_size_t(nbytes), C.byref(ptr), h), "nnb_ipc_alloc")
    return ptr.value, bytes(h)


def _ipc_open(handle):
    ptr = C.c_void_p(); h = (C.c_ubyte * 64).from_buffer_copy(handle)
    L.check(L.lib.nnb_ipc_open(h, C.byref(ptr)), "nnb_ipc_open")
    return ptr.value


class PeerGradExchange:
    def __init__(self, n_floats, device, group=None):
        import torch.distributed as dist
        assert n_floats % 4 == 0
        self.world = dist.get_world_size(group); self.rank = dist.get_rank(group)
        if self.world > L.MAX_RANKS:
            raise ValueError("peer exchange supports up to %d ranks of one box" % L.MAX_RANKS)
        self.n = n_floats; self.device = device
        with torch.cuda.device(device):
            gptr, gh = _ipc_alloc(n_floats * 4)
            fptr, fh = _ipc_alloc(L.FLAG_PAD_BYTES)
        handles = [None] * self.world
        dist.all_gather_object(handles, (gh, fh), group=group)
        self.grad_ptrs, self.flag_ptrs = [], []
        with torch.cuda.device(device):
            for r, (g_, f_) in enumerate(handles):
                if r == self.rank:
                    self.grad_ptrs.append(gptr); self.flag_ptrs.append(fptr)
                else:
                    self.grad_ptrs.append(_ipc_open(g_)); self.flag_ptrs.append(_ipc_open(f_))
        self._own = (gptr, fptr)
        self.grad = torch.as_tensor(_RawCuda(gptr, n_floats, "<f4"), device=device)        # local accumulate buffer (peers read it)
        self.flags = torch.as_tensor(_RawCuda(fptr, L.FLAG_PAD_BYTES // 4, "<i4"), device=device)
        self.reduced = torch.zeros(n_floats, device=device)                                # summed gradient (what .grad shows)
        dist.barrier(group=group)                    # every rank has mapped every buffer before the first kernel touches them

    def error_flag(self):
        """non-zero after a spin-wait gave up (a peer never arrived)"""
        return int(self.flags[2 * L.MAX_RANKS + 2].item())

    def allreduce_adam(self, segs):
        """segs: list of (p, m, v, offset, count, lr_dev, step_dev, beta1, beta2, eps) with device tensors"""
        a = L.AllreduceAdamArgs()
        for r in range(self.world):
            a.peer_grads[r] = self.grad_ptrs[r]; a.peer_flags[r] = self.flag_ptrs[r]
        a.world, a.rank, a.n_total = self.world, self.rank, self.n
        a.reduced_out = L.ptr(self.reduced)
        a.nsegs = len(segs)
        for q, (p, m, v, off, cnt, lr, step, b1, b2, eps) in enumerate(segs):
            s = a.segs[q]
            s.p, s.m, s.v, s.offset, s.count = L.ptr(p), L.ptr(m), L.ptr(v), int(off), int(cnt)
            s.lr_dev, s.step_dev, s.beta1, s.beta2, s.eps = L.ptr(lr), L.ptr(step), float(b1), float(b2), float(eps)
        L.check(L.lib.nnb_allreduce_adam(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "nnb_allreduce_adam")


class LocalGradExchange(PeerGradExchange):
    """world == 1: the same kernel as the data-parallel exchange, over ONE buffer -- it then is the multi-tensor Adam of the step
    (all parameter groups in one launch instead of five) and zeroes the gradient buffer for the next step.  No IPC, no process group."""

    def __init__(self, n_floats, device):
        assert n_floats % 4 == 0
        self.world = 1; self.rank = 0; self.n = n_floats; self.device = device
        self.grad = torch.zeros(n_floats, device=device)
        self.flags = torch.zeros(L.FLAG_PAD_BYTES // 4, dtype=torch.int32, device=device)
        self.reduced = torch.zeros(n_floats, device=device)
        self.grad_ptrs = [self.grad.data_ptr()]; self.flag_ptrs = [self.flags.data_ptr()]
