"""ctypes view of include/rans_b200.h.  No compute happens in Python."""
import ctypes as C
import os

import numpy as np

from .build import LIB_PATH, build

CODER_WORD, CODER_BYTE, CODER_ALIAS, CODER_RANS64 = 0, 1, 2, 3
MEM_HOST, MEM_DEVICE = 0, 1
LANES = 32

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)

# every symbol include/rans_b200.h declares: (name, restype, argtypes)
EXPORTS = [
    ("rb200_count_freqs", C.c_int, [C.c_void_p, C.c_size_t, _u32p]),
    ("rb200_normalize_freqs", C.c_int, [_u32p, _u32p, C.c_uint32]),
    ("rb200_word_tables_build", C.c_int, [_u32p, _u32p, _u32p, _u8p]),
    ("rb200_alias_tables_build", C.c_int, [_u32p, _u32p, _u32p, _u32p, _u32p, _u8p, _u32p]),
    ("rb200_ctx_create", C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    ("rb200_ctx_destroy", None, [C.c_void_p]),
    ("rb200_ctx_set_stream", C.c_int, [C.c_void_p, C.c_void_p]),
    ("rb200_sync", C.c_int, [C.c_void_p]),
    ("rb200_last_cuda_error", C.c_char_p, [C.c_void_p]),
    ("rb200_strerror", C.c_char_p, [C.c_int]),
    ("rb200_version", C.c_int, []),
    ("rb200_launch_count", C.c_uint64, [C.c_void_p]),
    ("rb200_model_create", C.c_int, [C.c_void_p, C.c_int, C.c_uint32, _u32p, C.POINTER(C.c_void_p)]),
    ("rb200_model_destroy", None, [C.c_void_p]),
    ("rb200_chunk_count", C.c_size_t, [C.c_size_t, C.c_uint32]),
    ("rb200_encode_bound", C.c_size_t, [C.c_size_t, C.c_uint32]),
    ("rb200_encode", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_size_t,
                               C.c_void_p, C.POINTER(C.c_size_t), C.c_int]),
    ("rb200_decode", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p,
                               C.c_size_t, C.c_int]),
    ("rb200_histogram", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, _u64p, C.c_int]),
    ("rb200_model_from_data", C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_size_t, C.c_int, _u32p,
                                        C.POINTER(C.c_void_p)]),
    ("rb200_blocks_build_models", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]),
    ("rb200_blocks_encode", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                      C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t), C.c_int]),
    ("rb200_blocks_model_encode", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                            C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t), C.c_int]),
    ("rb200_blocks_decode", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                      C.c_uint32, C.c_void_p, C.c_int]),
    ("rb200_comm_unique_id", C.c_int, [C.c_void_p]),
    ("rb200_comm_create", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    ("rb200_comm_destroy", None, [C.c_void_p]),
    ("rb200_gather_plan", C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, _u64p]),
    ("rb200_gather_blobs", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    ("rb200_container_size", C.c_size_t, [C.c_size_t, C.c_size_t]),
    ("rb200_container_pack", C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_size_t, _u32p, C.c_void_p, C.c_void_p, C.c_size_t,
                                       C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    ("rb200_container_open", C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
]
CONTAINER_CRC_BLOB = 1


class ContainerInfo(C.Structure):
    _fields_ = [("coder", C.c_uint32), ("scale_bits", C.c_uint32), ("chunk_syms", C.c_uint32), ("flags", C.c_uint32),
                ("n_symbols", C.c_uint64), ("n_chunks", C.c_uint64), ("blob_bytes", C.c_uint64),
                ("freqs", _u32p), ("offsets", _u64p), ("blob", _u8p)]


def container_pack(coder, scale_bits, chunk_syms, n, freqs, offsets, blob, flags=0):
    lib = load()
    freqs = np.ascontiguousarray(freqs, dtype=np.uint32)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    total = int(lib.dll.rb200_container_size(offsets.size - 1, blob.size))
    out = np.zeros(total, np.uint8)
    size = C.c_size_t(0)
    lib.check(lib.dll.rb200_container_pack(coder, scale_bits, chunk_syms, n, freqs.ctypes.data_as(_u32p), offsets.ctypes.data,
                                           blob.ctypes.data, blob.size, flags, out.ctypes.data, total, C.byref(size)))
    return out[:size.value]


def container_open(buf):
    """Returns (info dict, freqs, offsets, blob) as numpy views into buf."""
    lib = load()
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    info = ContainerInfo()
    lib.check(lib.dll.rb200_container_open(buf.ctypes.data, buf.size, C.byref(info)))
    base = buf.ctypes.data
    f_off = C.addressof(info.freqs.contents) - base
    o_off = C.addressof(info.offsets.contents) - base
    b_off = (C.addressof(info.blob.contents) - base) if info.blob_bytes else buf.size
    freqs = buf[f_off:f_off + 1024].view(np.uint32)
    offsets = buf[o_off:o_off + 8 * (info.n_chunks + 1)].view(np.uint64)
    blob = buf[b_off:b_off + info.blob_bytes]
    meta = {k: int(getattr(info, k)) for k in ("coder", "scale_bits", "chunk_syms", "flags", "n_symbols", "n_chunks", "blob_bytes")}
    return meta, freqs, offsets, blob


class RansError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rb200 error {code}: {msg}")
        self.code = code


class Lib:
    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing and could not be built; there is no CPU fallback")
        self.path = path
        self.dll = C.CDLL(path)
        for name, res, args in EXPORTS:
            f = getattr(self.dll, name)      # AttributeError if the .so lacks a declared symbol
            f.restype = res
            f.argtypes = args

    def check(self, rc, ctx=None):
        if rc != 0:
            msg = self.dll.rb200_strerror(rc).decode()
            if rc in (-5, -8) and ctx is not None:
                msg += ": " + self.dll.rb200_last_cuda_error(ctx).decode()
            raise RansError(rc, msg)


_LIB = None


def load():
    global _LIB
    if _LIB is None:
        _LIB = Lib()
    return _LIB


def _np_ptr(a):
    return a.ctypes.data


class SymbolStats:
    """Host-side mirror of the reference's SymbolStats (main.cpp:49-57, main_alias.cpp:47-72)."""

    def __init__(self):
        self.freqs = np.zeros(256, np.uint32)
        self.cum_freqs = np.zeros(257, np.uint32)

    def count_freqs(self, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        lib = load()
        lib.check(lib.dll.rb200_count_freqs(_np_ptr(data), data.size, self.freqs.ctypes.data_as(_u32p)))
        return self

    def normalize_freqs(self, target_total):
        lib = load()
        lib.check(lib.dll.rb200_normalize_freqs(self.freqs.ctypes.data_as(_u32p), self.cum_freqs.ctypes.data_as(_u32p),
                                                target_total))
        return self

    def word_tables(self):
        slots = np.zeros(4096, np.uint32)
        s2s = np.zeros(4096, np.uint8)
        lib = load()
        lib.check(lib.dll.rb200_word_tables_build(self.freqs.ctypes.data_as(_u32p), self.cum_freqs.ctypes.data_as(_u32p),
                                                  slots.ctypes.data_as(_u32p), s2s.ctypes.data_as(_u8p)))
        return slots, s2s

    def make_alias_table(self):
        self.divider = np.zeros(256, np.uint32)
        self.slot_adjust = np.zeros(512, np.uint32)
        self.slot_freqs = np.zeros(512, np.uint32)
        self.sym_id = np.zeros(512, np.uint8)
        self.alias_remap = np.zeros(int(self.cum_freqs[256]), np.uint32)
        lib = load()
        lib.check(lib.dll.rb200_alias_tables_build(
            self.freqs.ctypes.data_as(_u32p), self.cum_freqs.ctypes.data_as(_u32p), self.divider.ctypes.data_as(_u32p),
            self.slot_adjust.ctypes.data_as(_u32p), self.slot_freqs.ctypes.data_as(_u32p),
            self.sym_id.ctypes.data_as(_u8p), self.alias_remap.ctypes.data_as(_u32p)))
        return self


class Context:
    def __init__(self, device=0, stream=0):
        self.lib = load()
        h = C.c_void_p()
        self.lib.check(self.lib.dll.rb200_ctx_create(C.byref(h), device, C.c_void_p(stream or 0)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.dll.rb200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        self.lib.check(self.lib.dll.rb200_ctx_set_stream(self.h, C.c_void_p(stream or 0)))

    def sync(self):
        self.lib.check(self.lib.dll.rb200_sync(self.h), self.h)

    @property
    def launches(self):
        return int(self.lib.dll.rb200_launch_count(self.h))

    def model(self, coder, scale_bits, freqs):
        return Model(self, coder, scale_bits, freqs)

    # ---- geometry
    def chunk_count(self, n, chunk_syms):
        return int(self.lib.dll.rb200_chunk_count(n, chunk_syms))

    def encode_bound(self, n, chunk_syms):
        return int(self.lib.dll.rb200_encode_bound(n, chunk_syms))

    # ---- host-buffer calls (numpy in / numpy out)
    def encode_host(self, model, data, chunk_syms, blob_cap=None):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        n_chunks = self.chunk_count(data.size, chunk_syms)
        cap = self.encode_bound(data.size, chunk_syms) if blob_cap is None else blob_cap
        blob = np.zeros(max(cap, 16), np.uint8)
        offsets = np.zeros(n_chunks + 1, np.uint64)
        size = C.c_size_t(0)
        self.lib.check(self.lib.dll.rb200_encode(self.h, model.h, _np_ptr(data), data.size, chunk_syms, _np_ptr(blob), cap,
                                                 _np_ptr(offsets), C.byref(size), MEM_HOST), self.h)
        return blob[:size.value].copy(), offsets

    def decode_host(self, model, blob, offsets, n, chunk_syms):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        out = np.zeros(max(n, 1), np.uint8)
        self.lib.check(self.lib.dll.rb200_decode(self.h, model.h, _np_ptr(blob), blob.size, _np_ptr(offsets), chunk_syms,
                                                 _np_ptr(out), n, MEM_HOST), self.h)
        return out[:n]

    # ---- device-pointer calls (integers = device addresses); asynchronous
    def encode_device(self, model, in_ptr, n, chunk_syms, blob_ptr, blob_cap, offsets_ptr):
        self.lib.check(self.lib.dll.rb200_encode(self.h, model.h, in_ptr, n, chunk_syms, blob_ptr, blob_cap, offsets_ptr,
                                                 None, MEM_DEVICE), self.h)

    def decode_device(self, model, blob_ptr, blob_size, offsets_ptr, chunk_syms, out_ptr, n):
        self.lib.check(self.lib.dll.rb200_decode(self.h, model.h, blob_ptr, blob_size, offsets_ptr, chunk_syms, out_ptr, n,
                                                 MEM_DEVICE), self.h)

    def histogram(self, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        counts = np.zeros(256, np.uint64)
        self.lib.check(self.lib.dll.rb200_histogram(self.h, _np_ptr(data), data.size, counts.ctypes.data_as(_u64p), MEM_HOST),
                       self.h)
        return counts

    def histogram_device(self, ptr, n):
        counts = np.zeros(256, np.uint64)
        self.lib.check(self.lib.dll.rb200_histogram(self.h, ptr, n, counts.ctypes.data_as(_u64p), MEM_DEVICE), self.h)
        return counts

    # ---- per-block models, device pointers (asynchronous)
    def blocks_build_models_device(self, in_ptr, n_blocks, block_size, freqs_ptr):
        self.lib.check(self.lib.dll.rb200_blocks_build_models(self.h, in_ptr, n_blocks, block_size, freqs_ptr, MEM_DEVICE), self.h)

    def blocks_encode_device(self, in_ptr, n_blocks, block_size, freqs_ptr, chunk_syms, blob_ptr, blob_cap, offsets_ptr):
        self.lib.check(self.lib.dll.rb200_blocks_encode(self.h, in_ptr, n_blocks, block_size, freqs_ptr, chunk_syms, blob_ptr,
                                                        blob_cap, offsets_ptr, None, MEM_DEVICE), self.h)

    def blocks_model_encode_device(self, in_ptr, n_blocks, block_size, freqs_ptr, chunk_syms, blob_ptr, blob_cap, offsets_ptr):
        self.lib.check(self.lib.dll.rb200_blocks_model_encode(self.h, in_ptr, n_blocks, block_size, freqs_ptr, chunk_syms, blob_ptr,
                                                              blob_cap, offsets_ptr, None, MEM_DEVICE), self.h)

    def blocks_decode_device(self, blob_ptr, blob_size, offsets_ptr, freqs_ptr, n_blocks, block_size, chunk_syms, out_ptr):
        self.lib.check(self.lib.dll.rb200_blocks_decode(self.h, blob_ptr, blob_size, offsets_ptr, freqs_ptr, n_blocks, block_size,
                                                        chunk_syms, out_ptr, MEM_DEVICE), self.h)

    # ---- per-block models, host buffers
    def blocks_build_models(self, data, n_blocks, block_size):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        freqs = np.zeros((n_blocks, 256), np.uint16)
        self.lib.check(self.lib.dll.rb200_blocks_build_models(self.h, _np_ptr(data), n_blocks, block_size, _np_ptr(freqs),
                                                              MEM_HOST), self.h)
        return freqs

    def blocks_encode_host(self, data, n_blocks, block_size, block_freqs, chunk_syms):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        block_freqs = np.ascontiguousarray(block_freqs, dtype=np.uint16)
        n_chunks = n_blocks * (block_size // chunk_syms)
        cap = self.encode_bound(data.size, chunk_syms)
        blob = np.zeros(max(cap, 16), np.uint8)
        offsets = np.zeros(n_chunks + 1, np.uint64)
        size = C.c_size_t(0)
        self.lib.check(self.lib.dll.rb200_blocks_encode(self.h, _np_ptr(data), n_blocks, block_size, _np_ptr(block_freqs),
                                                        chunk_syms, _np_ptr(blob), cap, _np_ptr(offsets), C.byref(size),
                                                        MEM_HOST), self.h)
        return blob[:size.value].copy(), offsets

    def blocks_model_encode_host(self, data, n_blocks, block_size, chunk_syms):
        """models + encode in one call; returns (blob, offsets, block_freqs)"""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        freqs = np.zeros((n_blocks, 256), np.uint16)
        n_chunks = n_blocks * (block_size // chunk_syms)
        cap = self.encode_bound(data.size, chunk_syms)
        blob = np.zeros(max(cap, 16), np.uint8)
        offsets = np.zeros(n_chunks + 1, np.uint64)
        size = C.c_size_t(0)
        self.lib.check(self.lib.dll.rb200_blocks_model_encode(self.h, _np_ptr(data), n_blocks, block_size, _np_ptr(freqs), chunk_syms,
                                                              _np_ptr(blob), cap, _np_ptr(offsets), C.byref(size), MEM_HOST), self.h)
        return blob[:size.value].copy(), offsets, freqs

    def blocks_decode_host(self, blob, offsets, block_freqs, n_blocks, block_size, chunk_syms):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        block_freqs = np.ascontiguousarray(block_freqs, dtype=np.uint16)
        out = np.zeros(n_blocks * block_size, np.uint8)
        self.lib.check(self.lib.dll.rb200_blocks_decode(self.h, _np_ptr(blob), blob.size, _np_ptr(offsets),
                                                        _np_ptr(block_freqs), n_blocks, block_size, chunk_syms, _np_ptr(out),
                                                        MEM_HOST), self.h)
        return out


class Model:
    def __init__(self, ctx, coder, scale_bits, freqs, _handle=None):
        self.ctx = ctx
        self.coder = coder
        self.scale_bits = scale_bits
        freqs = np.ascontiguousarray(freqs, dtype=np.uint32)
        assert freqs.size == 256
        self.freqs = freqs
        if _handle is not None:
            self.h = _handle
            return
        h = C.c_void_p()
        ctx.lib.check(ctx.lib.dll.rb200_model_create(ctx.h, coder, scale_bits, freqs.ctypes.data_as(_u32p), C.byref(h)), ctx.h)
        self.h = h

    @classmethod
    def from_data(cls, ctx, coder, scale_bits, data=None, device_ptr=None, n=None):
        """rb200_model_from_data: histogram on the GPU + the reference's normalize_freqs + tables.  Pass a host
        array as `data`, or `device_ptr` + `n` for symbols that already live on the device."""
        freqs = np.zeros(256, np.uint32)
        h = C.c_void_p()
        if device_ptr is None:
            data = np.ascontiguousarray(data, dtype=np.uint8)
            ptr, n, kind = _np_ptr(data), data.size, MEM_HOST
        else:
            ptr, kind = device_ptr, MEM_DEVICE
        ctx.lib.check(ctx.lib.dll.rb200_model_from_data(ctx.h, coder, scale_bits, ptr, n, kind, freqs.ctypes.data_as(_u32p),
                                                        C.byref(h)), ctx.h)
        return cls(ctx, coder, scale_bits, freqs, _handle=h)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.dll.rb200_model_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
