"""Gradient all-reduce of the data-parallel run through the thin C ABI of libfa_b200.so
(fa_b200_nccl_{unique_id, init, allreduce_f32, destroy}, include/fa_b200.h) -- no torch.distributed.

One process per GPU.  Rank 0 creates the 128-byte ncclUniqueId and serves it to the other ranks over a plain TCP socket
on MASTER_ADDR : MASTER_PORT + 1 (the rendezvous variables torchrun already exports; SURVEY.md section 5 suggested a
multiprocessing pipe, which needs a common parent -- a socket also works under torchrun)."""
from __future__ import annotations

import ctypes
import os
import socket
import time
from typing import Optional

from .. import _lib

ID_BYTES = 128


def _prefer_loaded_nccl() -> None:
    """PyTorch ships its own libnccl.so.2; two NCCL builds under one SONAME cannot live in one process, so point the C
    side at the copy that torch (already imported by the training loop) will use or has loaded."""
    if os.environ.get("FA_B200_NCCL_LIB"):
        return
    try:
        import nvidia.nccl  # type: ignore
        base = list(nvidia.nccl.__path__)[0]
        cand = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["FA_B200_NCCL_LIB"] = cand
    except Exception:
        pass


def exchange_unique_id(rank: int, world: int, payload: Optional[bytes], addr: Optional[str] = None,
                       port: Optional[int] = None, timeout_s: float = 120.0) -> bytes:
    """Rank 0 hands `payload` (ID_BYTES bytes) to every other rank; returns the payload on all ranks."""
    addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = port or int(os.environ.get("MASTER_PORT", "29500")) + 1
    if world == 1:
        assert payload is not None
        return payload
    if rank == 0:
        assert payload is not None and len(payload) == ID_BYTES
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((addr, port))
        srv.listen(world)
        srv.settimeout(timeout_s)
        try:
            for _ in range(world - 1):
                conn, _ = srv.accept()
                with conn:
                    conn.sendall(payload)
        finally:
            srv.close()
        return payload
    deadline = time.time() + timeout_s
    while True:
        try:
            with socket.create_connection((addr, port), timeout=5.0) as c:
                buf = b""
                while len(buf) < ID_BYTES:
                    chunk = c.recv(ID_BYTES - len(buf))
                    if not chunk:
                        raise ConnectionError("rendezvous closed early")
                    buf += chunk
                return buf
        except (ConnectionRefusedError, ConnectionError, socket.timeout, OSError):
            if time.time() > deadline:
                raise TimeoutError(f"rank {rank}: no ncclUniqueId from {addr}:{port} within {timeout_s} s")
            time.sleep(0.05)


class NcclComm:
    """The process-wide communicator of libfa_b200.so.  `init` is collective; the current CUDA device must be set."""

    def __init__(self):
        self.lib = None
        self.rank, self.world = 0, 1

    def _bind(self):
        if self.lib is None:
            _prefer_loaded_nccl()
            lib = _lib.load_library()
            lib.fa_b200_nccl_unique_id.argtypes = [ctypes.c_void_p]
            lib.fa_b200_nccl_init.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
            lib.fa_b200_nccl_allreduce_f32.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
            lib.fa_b200_nccl_last_error.restype = ctypes.c_char_p
            for f in ("unique_id", "init", "allreduce_f32", "destroy", "world", "rank", "version"):
                getattr(lib, "fa_b200_nccl_" + f).restype = ctypes.c_int
            self.lib = lib
        return self.lib

    def _check(self, rc: int) -> None:
        if rc != _lib.FA_OK:
            raise _lib.FaError(rc, self.lib.fa_b200_nccl_last_error().decode())

    def init(self, rank: int, world: int, addr: Optional[str] = None, port: Optional[int] = None) -> "NcclComm":
        lib = self._bind()
        payload = None
        if rank == 0:
            buf = ctypes.create_string_buffer(ID_BYTES)
            self._check(lib.fa_b200_nccl_unique_id(buf))
            payload = buf.raw
        uid = exchange_unique_id(rank, world, payload, addr, port)
        self._check(lib.fa_b200_nccl_init(ctypes.c_char_p(uid), rank, world))
        self.rank, self.world = rank, world
        return self

    def all_reduce(self, t, average: bool = True, stream: Optional[int] = None) -> None:
        """In-place sum / average of a contiguous fp32 CUDA tensor on `stream` (default: torch's current stream)."""
        import torch
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        self._check(self.lib.fa_b200_nccl_allreduce_f32(t.data_ptr(), t.numel(), 1 if average else 0, stream))

    def destroy(self) -> None:
        if self.lib is not None:
            self._check(self.lib.fa_b200_nccl_destroy())
