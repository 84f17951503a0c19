"""Driver for tests/jni_mock/mock_jvm.cc: plays the JVM side of the reference's JNI boundary (AuronCallNativeWrapper.loadNextBatch
and friends) against the JNI natives exported by libauron_b200.so."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import pyarrow as pa

from auron_b200 import runtime

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "jni_mock", "mock_jvm.cc")
OUT = os.path.join(HERE, "jni_mock", "_build", "libmock_jvm.so")

EXPORT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)


def build_mock() -> str:
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-pthread", SRC, "-o", OUT], check=True)
    return OUT


_mock = None


def mock():
    global _mock
    if _mock is None:
        M = C.CDLL(build_mock())
        M.mock_new.restype = C.c_void_p
        M.mock_env.restype = C.c_void_p
        M.mock_env.argtypes = [C.c_void_p]
        M.mock_wrapper.restype = C.c_void_p
        M.mock_wrapper.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int, C.c_int]
        M.mock_wrapper_fail_import_after.argtypes = [C.c_void_p, C.c_int]
        M.mock_class.restype = C.c_void_p
        M.mock_class.argtypes = [C.c_void_p, C.c_char_p]
        M.mock_put_exporter.argtypes = [C.c_void_p, C.c_char_p, EXPORT_FN, C.c_void_p]
        M.mock_exporter_closed.argtypes = [C.c_void_p, C.c_char_p]
        M.mock_put_fs_provider.argtypes = [C.c_void_p, C.c_char_p]
        M.mock_add_block.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int64, C.c_int64, C.c_char_p]
        M.mock_blocks_closed.argtypes = [C.c_void_p, C.c_char_p]
        M.mock_set_conf.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        M.mock_set_task_running.argtypes = [C.c_void_p, C.c_int]
        M.mock_take_schema.argtypes = [C.c_void_p, C.c_void_p]
        M.mock_num_batches.argtypes = [C.c_void_p]
        M.mock_take_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        for f in (M.mock_error, M.mock_metrics):
            f.restype = C.c_char_p
            f.argtypes = [C.c_void_p]
        for f in (M.mock_pending_exception, M.mock_trouble):
            f.restype = C.c_char_p
            f.argtypes = [C.c_void_p]
        M.mock_counter.restype = C.c_int64
        M.mock_counter.argtypes = [C.c_void_p, C.c_char_p]
        M.mock_free.argtypes = [C.c_void_p]
        _mock = M
    return _mock


_natives = None


def natives():
    """the four JniBridge natives of the product library (JniBridge.java:49-55)"""
    global _natives
    if _natives is None:
        L = runtime.lib()
        p = "Java_org_apache_auron_jni_JniBridge_"
        call = getattr(L, p + "callNative")
        call.restype = C.c_int64
        call.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        nxt = getattr(L, p + "nextBatch")
        nxt.restype = C.c_uint8
        nxt.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        fin = getattr(L, p + "finalizeNative")
        fin.restype = None
        fin.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        ext = getattr(L, p + "onExit")
        ext.restype = None
        ext.argtypes = [C.c_void_p, C.c_void_p]
        _natives = (call, nxt, fin, ext)
    return _natives


class _CStruct(C.Structure):
    _fields_ = [("raw", C.c_uint8 * 80)]   # ArrowArray is 80 bytes, ArrowSchema 72


class MockJvm:
    """One JVM + one AuronCallNativeWrapper, used the way AuronCallNativeWrapper.java:78-190 uses the natives."""

    def __init__(self, task_definition: bytes, metric_depth: int = 4, metric_fanout: int = 2):
        self.M = mock()
        self.vm = self.M.mock_new()
        self.env = self.M.mock_env(self.vm)
        self.wrapper = self.M.mock_wrapper(self.vm, task_definition, len(task_definition), metric_depth, metric_fanout)
        self.bridge_class = self.M.mock_class(self.vm, b"org/apache/auron/jni/JniBridge")
        self.ptr = 0
        self._keep = []
        self.schema = None

    # ---- resources (JniBridge.putResource on the JVM side)
    def put_exporter(self, resource_id: str, batches: list[pa.RecordBatch], fail_after: int | None = None):
        """AuronArrowFFIExporter over `batches` (ArrowFFIExporter.scala: exportNextBatch fills the ArrowArray at the pointer)"""
        state = {"i": 0}

        def export(_user, out_ptr):
            if fail_after is not None and state["i"] >= fail_after:
                return -1
            if state["i"] >= len(batches):
                return 0
            b = batches[state["i"]]
            state["i"] += 1
            pa.StructArray.from_arrays(b.columns, fields=list(b.schema))._export_to_c(out_ptr)
            return 1

        fn = EXPORT_FN(export)
        self._keep.append(fn)
        self.M.mock_put_exporter(self.vm, resource_id.encode(), fn, None)

    def set_conf(self, key: str, value: str):
        self.M.mock_set_conf(self.vm, key.encode(), value.encode())

    def put_fs_provider(self, resource_id: str):
        self.M.mock_put_fs_provider(self.vm, resource_id.encode())

    def add_block(self, resource_id: str, kind: str, *, path: str = "", offset: int = 0, length: int = 0, data: bytes = b""):
        k = {"file": 0, "direct": 1, "heap": 2, "channel": 3}[kind]
        self._keep.append(data)
        self.M.mock_add_block(self.vm, resource_id.encode(), k, path.encode(), offset, length if k == 0 else len(data), data)

    # ---- the call sequence of AuronCallNativeWrapper
    def call_native(self) -> bool:
        call, _, _, _ = natives()
        self.ptr = call(self.env, self.bridge_class, 1 << 30, None, self.wrapper)
        if self.ptr:
            holder = _CStruct()
            if self.M.mock_take_schema(self.wrapper, C.byref(holder)):
                self.schema = pa.Schema._import_from_c(C.addressof(holder))
        return self.ptr != 0

    def load_next_batch(self) -> pa.RecordBatch | None:
        """loadNextBatch (AuronCallNativeWrapper.java:113-129): None at the end of the stream or on error"""
        _, nxt, _, _ = natives()
        before = self.M.mock_num_batches(self.wrapper)
        if not nxt(self.env, self.bridge_class, self.ptr):
            return None
        assert self.M.mock_num_batches(self.wrapper) == before + 1, "nextBatch returned true without importBatch"
        holder = _CStruct()
        self.M.mock_take_batch(self.wrapper, before, C.byref(holder))
        arr = pa.Array._import_from_c(C.addressof(holder), pa.struct(list(self.schema)))
        return pa.RecordBatch.from_struct_array(arr)

    def close(self):
        _, _, fin, _ = natives()
        if self.ptr:
            fin(self.env, self.bridge_class, self.ptr)
            self.ptr = 0

    def run(self) -> pa.Table:
        assert self.call_native(), self.pending_exception()
        out = []
        while (b := self.load_next_batch()) is not None:
            out.append(b)
        err, pend = self.error(), self.pending_exception()
        self.close()
        assert not err and not pend, (err, pend)
        return pa.Table.from_batches(out, schema=self.schema)

    # ---- observations
    def error(self) -> str:
        return self.M.mock_error(self.wrapper).decode()

    def pending_exception(self) -> str:
        return self.M.mock_pending_exception(self.vm).decode()

    def metrics(self) -> dict[str, int]:
        out = {}
        for line in self.M.mock_metrics(self.wrapper).decode().splitlines():
            k, v = line.rsplit("=", 1)
            out[k] = out.get(k, 0) + int(v)
        return out

    def counter(self, name: str) -> int:
        return self.M.mock_counter(self.vm, name.encode())

    def trouble(self) -> str:
        return self.M.mock_trouble(self.vm).decode()

    def assert_clean(self):
        """nothing leaked and no JNI rule broken once the task is finalized"""
        assert self.trouble() == ""
        assert self.counter("live_globals") == 0, "global references leaked"
        assert self.counter("global_new") > 0
        assert self.counter("frames_open") == 0, "PushLocalFrame without PopLocalFrame"
        assert self.counter("input_wrappers") == self.counter("input_wrappers_closed")
