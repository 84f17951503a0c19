"""ctypes binding of libauron_b200.so -- the host side that plays the role of Auron's JVM glue
(auron-core/src/main/java/org/apache/auron/jni/AuronCallNativeWrapper.java:78-192): ship a protobuf
TaskDefinition down, pull Arrow batches back through the Arrow C Data Interface, feed FFI-reader
inputs through the export_next_batch upcall.

There is NO CPU fallback: if the CUDA library is missing or no device is present every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable

import pyarrow as pa

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libauron_b200.so")


class ArrowSchemaC(C.Structure):
    _fields_ = [("format", C.c_char_p), ("name", C.c_char_p), ("metadata", C.c_char_p), ("flags", C.c_int64), ("n_children", C.c_int64),
                ("children", C.c_void_p), ("dictionary", C.c_void_p), ("release", C.c_void_p), ("private_data", C.c_void_p)]


class ArrowArrayC(C.Structure):
    _fields_ = [("length", C.c_int64), ("null_count", C.c_int64), ("offset", C.c_int64), ("n_buffers", C.c_int64), ("n_children", C.c_int64),
                ("buffers", C.c_void_p), ("children", C.c_void_p), ("dictionary", C.c_void_p), ("release", C.c_void_p), ("private_data", C.c_void_p)]


_EXPORT_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p, C.c_void_p)
_READ_CB = C.CFUNCTYPE(C.c_int64, C.c_void_p, C.c_char_p, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64)
_RUNNING_CB = C.CFUNCTYPE(C.c_int, C.c_void_p)
_METRIC_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int64)


class ShuffleBlockC(C.Structure):
    _fields_ = [("path", C.c_char_p), ("offset", C.c_int64), ("length", C.c_int64), ("data", C.c_void_p)]


_BLOCK_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p, C.POINTER(ShuffleBlockC))
_IPC_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64)
_FETCH_FAILED_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p, C.c_char_p)


class Callbacks(C.Structure):
    _fields_ = [("user", C.c_void_p), ("export_next_batch", _EXPORT_CB), ("read_fully", _READ_CB), ("is_task_running", _RUNNING_CB),
                ("next_shuffle_block", _BLOCK_CB), ("upcalls_from_any_thread", C.c_int32), ("get_conf", C.c_void_p), ("write_ipc", _IPC_CB), ("fetch_failed", _FETCH_FAILED_CB)]


class AuronError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AuronError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` first")
        L = C.CDLL(LIB_PATH)
        L.auron_b200_call_native.restype = C.c_void_p
        L.auron_b200_call_native.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_int]
        L.auron_b200_schema.argtypes = [C.c_void_p, C.c_void_p]
        L.auron_b200_next_batch.argtypes = [C.c_void_p, C.c_void_p]
        L.auron_b200_finalize_native.argtypes = [C.c_void_p]
        L.auron_b200_finalize_native.restype = None
        L.auron_b200_last_error.restype = C.c_char_p
        L.auron_b200_metrics.argtypes = [C.c_void_p, _METRIC_CB, C.c_void_p]
        L.auron_b200_put_device_batch.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int]
        L.auron_b200_drop_device_resource.argtypes = [C.c_char_p]
        L.auron_b200_drop_device_resource.restype = None
        L.auron_b200_put_device_file.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_int]
        L.auron_b200_drop_device_file.argtypes = [C.c_char_p]
        L.auron_b200_drop_device_file.restype = None
        L.auron_b200_put_host_file.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
        L.auron_b200_drop_host_file.argtypes = [C.c_char_p]
        L.auron_b200_drop_host_file.restype = None
        L.auron_b200_k_hash.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        L.auron_b200_k_partition_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int]
        L.auron_b200_kernel_launches.restype = C.c_int64
        L.auron_b200_time_kernel.restype = C.c_double
        L.auron_b200_time_kernel.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_int]
        _lib = L
    return _lib


EXPORTED_SYMBOLS = [
    "auron_b200_call_native", "auron_b200_schema", "auron_b200_next_batch", "auron_b200_finalize_native", "auron_b200_on_exit",
    "auron_b200_last_error", "auron_b200_metrics", "auron_b200_put_device_batch", "auron_b200_drop_device_resource", "auron_b200_k_hash",
    "auron_b200_put_device_file", "auron_b200_drop_device_file", "auron_b200_put_host_file", "auron_b200_drop_host_file", "auron_b200_nccl_unique_id", "auron_b200_nccl_init", "auron_b200_nccl_finalize",
    "auron_b200_k_partition_ids", "auron_b200_kernel_launches", "auron_b200_time_kernel", "auron_b200_set_hbm_budget",
]


def _err() -> str:
    return (lib().auron_b200_last_error() or b"").decode(errors="replace")


def _export_batch(batch: pa.RecordBatch):
    arr, sch = ArrowArrayC(), ArrowSchemaC()
    batch._export_to_c(C.addressof(arr), C.addressof(sch))
    return arr, sch


def _release(struct):
    if struct.release:
        C.CFUNCTYPE(None, C.c_void_p)(struct.release)(C.addressof(struct))


def kernel_launches() -> int:
    return lib().auron_b200_kernel_launches()


class Task:
    """One native task: callNative -> nextBatch* -> finalizeNative (JniBridge.java:49-55)."""

    def __init__(self, task_definition: bytes, inputs: dict[str, Iterable[pa.RecordBatch]] | None = None, device: int = 0,
                 read_fully=None, shuffle_blocks: dict[str, Iterable] | None = None, ipc_consumers: dict[str, list] | None = None):
        """shuffle_blocks: resource id -> iterable of blocks for IpcReaderExec; a block is (path, offset, length) for a file
        segment or a bytes object for an in-memory buffer (AuronBlockObject.hasFileSegment / hasByteBuffer)."""
        self._inputs = {k: iter(v) for k, v in (inputs or {}).items()}
        self._blocks = {k: iter(v) for k, v in (shuffle_blocks or {}).items()}
        self._block_keep = None
        self._cb_error: BaseException | None = None

        def _next_block(user, rid, out):
            try:
                it = self._blocks.get(rid.decode())
                if it is None:
                    return -1
                blk = next(it, None)
                if blk is None:
                    return 0
                if isinstance(blk, (bytes, bytearray, memoryview)):
                    buf = C.create_string_buffer(bytes(blk), len(blk))
                    self._block_keep = buf                       # valid until the next upcall
                    out[0].path = None
                    out[0].offset = 0
                    out[0].length = len(blk)
                    out[0].data = C.cast(buf, C.c_void_p)
                else:
                    path, offset, length = blk
                    self._block_keep = path.encode()
                    out[0].path = self._block_keep
                    out[0].offset = offset
                    out[0].length = length
                    out[0].data = None
                return 1
            except BaseException as e:  # noqa: BLE001 - must not unwind through C
                self._cb_error = e
                return -1

        def _export_next(user, rid, out_ptr):
            try:
                it = self._inputs.get(rid.decode())
                if it is None:
                    return -1
                batch = next(it, None)
                if batch is None:
                    return 0
                batch._export_to_c(out_ptr)
                return 1
            except BaseException as e:  # noqa: BLE001 - must not unwind through C
                self._cb_error = e
                return -1

        def _read(user, fsid, path, pos, buf, length):
            try:
                if read_fully is not None:
                    data = read_fully(fsid.decode(), path.decode(), pos, length)
                else:
                    with open(path.decode(), "rb") as f:
                        f.seek(pos)
                        data = f.read(length)
                C.memmove(buf, data, len(data))
                return len(data)
            except BaseException as e:  # noqa: BLE001
                self._cb_error = e
                return -1

        def _write_ipc(user, rid, data, length):
            # IpcWriterExec's consumer: ipc_consumers[resource id] is a list that receives every delivered block sequence as bytes
            try:
                sink = (ipc_consumers or {}).get(rid.decode())
                if sink is None:
                    return -1
                sink.append(C.string_at(data, length))
                return 0
            except BaseException as e:  # noqa: BLE001
                self._cb_error = e
                return -1

        self.fetch_failures: list[tuple[str, str]] = []   # (resource id, message) of every fetch_failed upcall

        def _fetch_failed(user, rid, msg):
            self.fetch_failures.append((rid.decode(), msg.decode(errors="replace")))

        self._cbs = Callbacks(None, _EXPORT_CB(_export_next), _READ_CB(_read) if read_fully is not None else _READ_CB(0), _RUNNING_CB(0),
                              _BLOCK_CB(_next_block), 0, None, _IPC_CB(_write_ipc), _FETCH_FAILED_CB(_fetch_failed))
        self._handle = lib().auron_b200_call_native(task_definition, len(task_definition), C.addressof(self._cbs), device)
        if not self._handle:
            raise AuronError(_err())
        sch = ArrowSchemaC()
        if lib().auron_b200_schema(self._handle, C.addressof(sch)) != 0:
            raise AuronError(_err())
        self.schema = pa.Schema._import_from_c(C.addressof(sch))

    def next_batch(self) -> pa.RecordBatch | None:
        arr = ArrowArrayC()
        rc = lib().auron_b200_next_batch(self._handle, C.addressof(arr))
        if rc < 0:
            if self._cb_error is not None:
                raise self._cb_error
            raise AuronError(_err())
        if rc == 0:
            return None
        return pa.RecordBatch._import_from_c(C.addressof(arr), self.schema)

    def __iter__(self):
        while True:
            b = self.next_batch()
            if b is None:
                return
            yield b

    def metrics(self) -> list[tuple[int, str, str, int]]:
        out = []

        def cb(user, depth, op, name, value):
            out.append((depth, op.decode(), name.decode(), value))

        lib().auron_b200_metrics(self._handle, _METRIC_CB(cb), None)
        return out

    def close(self):
        if self._handle:
            lib().auron_b200_finalize_native(self._handle)
            self._handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def explain(task_definition: bytes) -> dict:
    """The operator tree the engine's planner builds for `task_definition`, decoded on no device (auron_b200_explain)."""
    import json
    L = lib()
    L.auron_b200_explain.restype = C.c_int64
    L.auron_b200_explain.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int64]
    n = L.auron_b200_explain(task_definition, len(task_definition), None, 0)
    if n < 0:
        raise AuronError(_err())
    buf = C.create_string_buffer(int(n) + 1)
    L.auron_b200_explain(task_definition, len(task_definition), buf, len(buf))
    return json.loads(buf.value.decode())


def run_task(task_definition: bytes, inputs: dict[str, Iterable[pa.RecordBatch]] | None = None, device: int = 0, read_fully=None,
             shuffle_blocks: dict[str, Iterable] | None = None) -> pa.Table:
    """Execute a TaskDefinition and collect its output stream."""
    with Task(task_definition, inputs, device, read_fully, shuffle_blocks) as t:
        batches = list(t)
        return pa.Table.from_batches(batches, schema=t.schema)


def put_device_batch(resource_id: str, batch: pa.RecordBatch, device: int = 0):
    arr, sch = _export_batch(batch)
    try:
        if lib().auron_b200_put_device_batch(resource_id.encode(), C.addressof(arr), C.addressof(sch), device) != 0:
            raise AuronError(_err())
    finally:
        _release(arr)
        _release(sch)


def set_hbm_budget(bytes_: int, device: int = 0) -> int:
    """HBM budget shared by the spillable operators on `device` (0 or less: the default); returns the budget in force"""
    L = lib()
    L.auron_b200_set_hbm_budget.restype = C.c_int64
    L.auron_b200_set_hbm_budget.argtypes = [C.c_int, C.c_int64]
    return int(L.auron_b200_set_hbm_budget(device, bytes_))


def drop_device_resource(resource_id: str):
    lib().auron_b200_drop_device_resource(resource_id.encode())


def put_device_file(path: str, data, device: int = 0):
    """Make a Parquet file image resident in HBM under `path` (data: any bytes-like object)."""
    import numpy as np
    arr = np.frombuffer(data, dtype=np.uint8)
    if lib().auron_b200_put_device_file(path.encode(), arr.ctypes.data, arr.nbytes, device) != 0:
        raise AuronError(_err())


_host_files = {}


def put_host_file(path: str, buf):
    """Register a host-resident image of `path` (a uint8 torch tensor — pin it for full PCIe rate — or a numpy array).
    The buffer is kept alive until drop_host_file."""
    if hasattr(buf, "data_ptr"):
        ptr, n = buf.data_ptr(), buf.numel() * buf.element_size()
    else:
        import numpy as np

        buf = np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8)) if not isinstance(buf, np.ndarray) else buf
        ptr, n = buf.ctypes.data, buf.nbytes
    if lib().auron_b200_put_host_file(path.encode(), ptr, n) != 0:
        raise AuronError(_err())
    _host_files[path] = buf


def drop_host_file(path: str):
    lib().auron_b200_drop_host_file(path.encode())
    _host_files.pop(path, None)


def drop_device_file(path: str):
    lib().auron_b200_drop_device_file(path.encode())


def nccl_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    if lib().auron_b200_nccl_unique_id(buf) != 0:
        raise AuronError(_err())
    return bytes(buf)


def nccl_init(unique_id: bytes, rank: int, world: int, device: int):
    """Join the engine's NCCL communicator (one process per GPU).  The id comes from rank 0's nccl_unique_id()."""
    buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
    if lib().auron_b200_nccl_init(buf, rank, world, device) != 0:
        raise AuronError(_err())


def nccl_finalize():
    lib().auron_b200_nccl_finalize()


def owner_of_partition(p: int, world: int, num_parts: int) -> int:
    """rank that owns shuffle partition p after the in-box exchange (exchange.cu: contiguous blocks)"""
    return p * world // num_parts


def _kernel_one_column(fn, batch: pa.RecordBatch, cols: list[int], *mid_args, device: int = 0) -> pa.Array:
    arr, sch = _export_batch(batch)
    oa, os_ = ArrowArrayC(), ArrowSchemaC()
    idx = (C.c_int32 * len(cols))(*cols)
    try:
        rc = fn(C.addressof(arr), C.addressof(sch), idx, len(cols), *mid_args, C.addressof(oa), C.addressof(os_), device)
        if rc != 0:
            raise AuronError(_err())
    finally:
        _release(arr)
        _release(sch)
    schema = pa.Schema._import_from_c(C.addressof(os_))
    return pa.RecordBatch._import_from_c(C.addressof(oa), schema).column(0)


def k_hash(batch: pa.RecordBatch, cols: list[int], kind: str = "murmur3", seed: int = 42, device: int = 0) -> pa.Array:
    return _kernel_one_column(lib().auron_b200_k_hash, batch, cols, 0 if kind == "murmur3" else 1, seed, device=device)


def k_partition_ids(batch: pa.RecordBatch, cols: list[int], num_partitions: int, device: int = 0) -> pa.Array:
    return _kernel_one_column(lib().auron_b200_k_partition_ids, batch, cols, num_partitions, device=device)


def time_kernel(kernel: str, resource_id: str, iters: int = 10, arg0: int = 0, device: int = 0) -> float:
    ms = lib().auron_b200_time_kernel(kernel.encode(), resource_id.encode(), iters, arg0, device)
    if ms < 0:
        raise AuronError(_err())
    return ms
