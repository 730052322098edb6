"""Multi-GPU batch split / merge (one process per GPU) on top of the ``zafb_dist_*`` C ABI.

The transforms have no exchange step -- every clip is independent -- so a
shard is a contiguous clip range and the only collectives are the ones that move a batch: scatter
the clips from a root rank, gather (or all-gather) the per-rank results, broadcast operators.
They run over NCCL (NVLink 5 / NVSwitch inside the box); results are bitwise independent of the
sharding.

    comm = zaf.dist.Communicator.from_env()          # RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT
    x_shard = comm.scatter(x_dev if comm.rank == 0 else None, n_clips, (ns,), np.float32)
    spec = zaf.stft(x_shard, w, hop)                 # local kernels only
    full = comm.gather(spec, n_clips)                # DeviceArray on root, None elsewhere

The 128-byte NCCL id travels from rank 0 to its peers over a one-shot TCP exchange (plain
sockets, no PyTorch); callers that already have a channel can pass the id to the constructor.
"""
from __future__ import annotations

import ctypes as C
import os
import socket
import time

import numpy as np

from . import _lib
from ._device import DeviceArray, ensure_init, init

ID_BYTES = 128


def shard_range(n_items: int, rank: int, world: int):
    """Rank r of R owns items [floor(r*B/R), floor((r+1)*B/R)) -- same arithmetic as
    ``zafb_dist_shard_range`` (checked against it in tests/test_abi.py)."""
    if not (0 <= rank < world):
        raise ValueError("rank must be in [0, world)")
    return (rank * n_items) // world, ((rank + 1) * n_items) // world


def exchange_id(rank: int, world: int, make_id, addr: str = "127.0.0.1", port: int = 29517, timeout: float = 120.0) -> bytes:
    """Rank 0 calls ``make_id()`` and serves the bytes to its world-1 peers; the others fetch them."""
    if world == 1:
        return make_id()
    if rank == 0:
        payload = make_id()
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((addr, port))
        srv.listen(world)
        srv.settimeout(timeout)
        try:
            for _ in range(world - 1):
                conn, _peer = srv.accept()
                with conn:
                    conn.sendall(len(payload).to_bytes(4, "little") + payload)
        finally:
            srv.close()
        return payload
    deadline = time.time() + timeout
    while True:
        try:
            with socket.create_connection((addr, port), timeout=5.0) as conn:
                buf = b""
                while len(buf) < 4 or len(buf) < 4 + int.from_bytes(buf[:4], "little"):
                    chunk = conn.recv(4096)
                    if not chunk:
                        break
                    buf += chunk
                n = int.from_bytes(buf[:4], "little")
                if len(buf) >= 4 + n:
                    return buf[4:4 + n]
        except OSError:
            pass
        if time.time() > deadline:
            raise TimeoutError(f"rank {rank}: no NCCL id from rank 0 at {addr}:{port}")
        time.sleep(0.05)


def nccl_version() -> int:
    v = C.c_int(0)
    _lib.check(_lib.lib().zafb_dist_nccl_version(C.byref(v)))
    return v.value


def make_unique_id() -> bytes:
    buf = C.create_string_buffer(ID_BYTES)
    _lib.check(_lib.lib().zafb_dist_unique_id(buf))
    return buf.raw


class Communicator:
    """One NCCL communicator over all ranks (one process per GPU)."""

    def __init__(self, rank: int, world: int, unique_id: bytes):
        if len(unique_id) != ID_BYTES:
            raise ValueError(f"unique_id must be {ID_BYTES} bytes")
        ensure_init()
        self.rank, self.world = int(rank), int(world)
        h = C.c_void_p()
        _lib.check(_lib.lib().zafb_dist_init(C.byref(h), unique_id, self.rank, self.world))
        self._h = h

    @classmethod
    def from_env(cls, device=None):
        """Build from the launcher's environment (torchrun's variable names)."""
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        local = int(os.environ.get("LOCAL_RANK", str(rank)))
        init(local if device is None else device)
        addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
        port = int(os.environ.get("ZAFB_DIST_PORT", str(int(os.environ.get("MASTER_PORT", "29500")) + 17)))
        return cls(rank, world, exchange_id(rank, world, make_unique_id, addr, port))

    def close(self):
        if self._h:
            _lib.lib().zafb_dist_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def shard_range(self, n_rows: int):
        return shard_range(n_rows, self.rank, self.world)

    @staticmethod
    def _row_bytes(row_shape, dtype):
        return int(np.prod(row_shape, dtype=np.int64)) * np.dtype(dtype).itemsize

    @staticmethod
    def _sp(stream):
        return stream.ptr if stream is not None else None

    # ------------------------------------------------------------------ collectives
    def broadcast(self, array: DeviceArray, root: int = 0, stream=None) -> DeviceArray:
        _lib.check(_lib.lib().zafb_dist_broadcast(self._h, C.c_void_p(array.ptr), array.nbytes, root, self._sp(stream)))
        return array

    def scatter(self, src, n_rows: int, row_shape, dtype, root: int = 0, stream=None, out=None) -> DeviceArray:
        """Split ``n_rows`` rows held by ``root`` (``src``: DeviceArray of memory shape (n_rows, *row_shape));
        returns this rank's rows [begin, end) as a DeviceArray of memory shape (end-begin, *row_shape)."""
        row_shape = tuple(int(s) for s in row_shape)
        b, e = self.shard_range(n_rows)
        dst = out if out is not None else DeviceArray((e - b,) + row_shape, dtype)
        src_ptr = C.c_void_p(src.ptr) if (self.rank == root and src is not None) else None
        _lib.check(_lib.lib().zafb_dist_scatter_rows(self._h, src_ptr, C.c_void_p(dst.ptr), n_rows,
                                                     self._row_bytes(row_shape, dtype), root, self._sp(stream)))
        return dst

    def gather(self, shard: DeviceArray, n_rows: int, root: int = 0, stream=None, out=None):
        """Merge the per-rank row blocks on ``root``; returns the (n_rows, ...) DeviceArray there, None elsewhere.
        A transposed (frame-major) shard yields a transposed result."""
        row_shape = shard.mem_shape[1:]
        dst = None
        if self.rank == root:
            dst = out if out is not None else DeviceArray((n_rows,) + row_shape, shard.dtype, transposed=shard.transposed)
        _lib.check(_lib.lib().zafb_dist_gather_rows(self._h, C.c_void_p(shard.ptr), C.c_void_p(dst.ptr) if dst else None,
                                                    n_rows, self._row_bytes(row_shape, shard.dtype), root, self._sp(stream)))
        return dst

    def allgather(self, shard: DeviceArray, n_rows: int, stream=None, out=None) -> DeviceArray:
        row_shape = shard.mem_shape[1:]
        dst = out if out is not None else DeviceArray((n_rows,) + row_shape, shard.dtype, transposed=shard.transposed)
        _lib.check(_lib.lib().zafb_dist_allgather_rows(self._h, C.c_void_p(shard.ptr), C.c_void_p(dst.ptr), n_rows,
                                                       self._row_bytes(row_shape, shard.dtype), self._sp(stream)))
        return dst

    def gather_onesided(self, half_shard: DeviceArray, n_clips: int, window_length: int, root: int = 0, stream=None, out=None):
        """Merge per-rank ONE-SIDED frame-major spectra (``zaf.stft(shard, w, hop, onesided=True)``, memory
        (clips, frames, pitch >= N/2+1)) on ``root`` and rebuild the reference's two-sided spectrum there with the device
        mirror kernel: half the NVLink traffic of gathering the two-sided result, bit-identical values.  Returns the
        (n_clips, N, frames) frame-major DeviceArray on root, None elsewhere."""
        n = int(window_length)
        nt, pitch = half_shard.mem_shape[1], half_shard.mem_shape[2]
        staging = None
        if self.rank == root:
            staging = DeviceArray((n_clips, nt, pitch), np.complex64)
        _lib.check(_lib.lib().zafb_dist_gather_rows(self._h, C.c_void_p(half_shard.ptr), C.c_void_p(staging.ptr) if staging else None,
                                                    n_clips, nt * pitch * 8, root, self._sp(stream)))
        if self.rank != root:
            return None
        dst = out if out is not None else DeviceArray((n_clips, nt, n), np.complex64, transposed=True)
        _lib.check(_lib.lib().zafb_spec_mirror_f32(C.c_void_p(staging.ptr), pitch, n_clips * nt, n, C.c_void_p(dst.ptr),
                                                   self._sp(stream)))
        if stream is not None:
            stream.synchronize()
        else:
            from ._device import synchronize
            synchronize()
        staging.free()
        return dst

    # ------------------------------------------------------------------ peer memory
    def map_from_root(self, array, shape, dtype, root: int = 0, transposed: bool = False) -> DeviceArray:
        """Collective.  Every rank receives a DeviceArray of memory shape ``shape`` that ALIASES the root's
        ``array`` (a whole ``zaf.empty`` / ``to_device`` allocation): the root gets ``array`` itself, the peers a CUDA-IPC
        mapping of it.  Passing (a row range of) it as ``out=`` of a transform makes the kernel store its result
        straight into the root's HBM over NVLink -- the merge needs no separate collective.  Release with ``unmap``."""
        from ._device import to_device

        handle = np.zeros(64, np.uint8)
        if self.rank == root:
            _lib.check(_lib.lib().zafb_dist_peer_export(C.c_void_p(array.ptr), handle.ctypes.data))
        hd = to_device(handle)
        self.broadcast(hd, root)
        handle = hd.to_host()
        hd.free()
        if self.rank == root:
            return array
        mapped = C.c_void_p()
        _lib.check(_lib.lib().zafb_dist_peer_open(handle.ctypes.data, C.byref(mapped)))
        view = DeviceArray(shape, dtype, ptr=mapped.value, transposed=transposed)
        view._peer_mapping = mapped.value
        return view

    def unmap(self, view: DeviceArray):
        m = getattr(view, "_peer_mapping", None)
        if m:
            _lib.check(_lib.lib().zafb_dist_peer_close(C.c_void_p(m)))
            view._peer_mapping = None

    @staticmethod
    def rows(array: DeviceArray, begin: int, end: int) -> DeviceArray:
        """Non-owning view of rows [begin, end) of the leading axis of ``array``'s memory."""
        row_bytes = int(np.prod(array.mem_shape[1:], dtype=np.int64)) * array.dtype.itemsize
        return DeviceArray((end - begin,) + tuple(array.mem_shape[1:]), array.dtype, ptr=array.ptr + begin * row_bytes,
                           owner=array, transposed=array.transposed)

    def max(self, value: float, stream=None) -> float:
        """Max over ranks of a host scalar (device-timed durations); synchronises the stream."""
        v = C.c_double(float(value))
        _lib.check(_lib.lib().zafb_dist_max_f64(self._h, C.byref(v), self._sp(stream)))
        return v.value

    def barrier(self, stream=None):
        self.max(0.0, stream)
