"""Minimal read-only HDF5 subset reader for legacy Keras ``.keras``/``.h5`` model files.

The reference loads its surrogates with ``tensorflow.keras.models.load_model``
(reference: src/SurrogateModeling/model.py:28-34).  Neither TensorFlow nor any
HDF5 library exists in this image, so this module parses exactly the feature
subset those files use: superblock v0, v1 object headers with
continuation blocks, old-style groups (symbol-table message -> v1 B-tree -> SNOD
-> local heap), contiguous little-endian IEEE datasets, and variable-length
string attributes stored in global heap collections.

Host-side, load-time only: nothing here is on the sampling hot path.
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    """Raised when a file uses an HDF5 feature outside the supported subset."""


@dataclass
class _Obj:
    addr: int
    messages: List[Tuple[int, bytes, int]] = field(default_factory=list)  # (type, payload, payload_addr)


class H5File:
    """Parsed view of one HDF5 file (groups, datasets and string attributes)."""

    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        if self.buf[:8] != _SIG:
            raise H5FormatError(f"{path}: not an HDF5 file")
        ver = self.buf[8]
        if ver != 0:
            raise H5FormatError(f"{path}: superblock version {ver} unsupported (need 0)")
        so, sl = self.buf[13], self.buf[14]
        if (so, sl) != (8, 8):
            raise H5FormatError("only 8-byte offsets/lengths supported")
        # superblock v0: 8 sig + 8 versions/sizes + 2+2 K + 4 flags, then 4 addresses, then root entry
        base = 24
        self.base_addr = self._u64(base)
        root_entry = base + 32
        self.root_addr = self._u64(root_entry + 8)
        self._gheap_cache: Dict[int, Dict[int, bytes]] = {}

    # -- primitive readers -------------------------------------------------
    def _u8(self, o):
        return self.buf[o]

    def _u16(self, o):
        return struct.unpack_from("<H", self.buf, o)[0]

    def _u32(self, o):
        return struct.unpack_from("<I", self.buf, o)[0]

    def _u64(self, o):
        return struct.unpack_from("<Q", self.buf, o)[0]

    # -- object headers ----------------------------------------------------
    def _read_object(self, addr: int) -> _Obj:
        obj = _Obj(addr)
        if self.buf[addr:addr + 4] == b"OHDR":
            raise H5FormatError("v2 object headers unsupported")
        ver = self._u8(addr)
        if ver != 1:
            raise H5FormatError(f"object header version {ver} unsupported")
        nmsg = self._u16(addr + 2)
        hsize = self._u32(addr + 8)
        blocks = [(addr + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            off, size = blocks.pop(0)
            end = off + size
            while off + 8 <= end and seen < nmsg:
                mtype = self._u16(off)
                msize = self._u16(off + 2)
                payload_addr = off + 8
                payload = self.buf[payload_addr:payload_addr + msize]
                seen += 1
                if mtype == 0x10:  # continuation
                    blocks.append((self._u64(payload_addr), self._u64(payload_addr + 8)))
                else:
                    obj.messages.append((mtype, payload, payload_addr))
                off = payload_addr + msize
        return obj

    # -- groups --------------------------------------------------------------
    def _heap_name(self, heap_addr: int, off: int) -> str:
        if self.buf[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5FormatError("bad local heap signature")
        data_addr = self._u64(heap_addr + 24)
        s = data_addr + off
        e = self.buf.index(b"\x00", s)
        return self.buf[s:e].decode("utf-8")

    def _walk_btree(self, addr: int, heap_addr: int, out: Dict[str, int]):
        if self.buf[addr:addr + 4] != b"TREE":
            raise H5FormatError("bad B-tree signature")
        ntype, level, used = self._u8(addr + 4), self._u8(addr + 5), self._u16(addr + 6)
        if ntype != 0:
            raise H5FormatError("only group B-trees supported")
        p = addr + 24
        for i in range(used):
            child = self._u64(p + 8 + i * 16)  # key_i(8) child_i(8) ...
            if level > 0:
                self._walk_btree(child, heap_addr, out)
            else:
                if self.buf[child:child + 4] != b"SNOD":
                    raise H5FormatError("bad SNOD signature")
                nsym = self._u16(child + 6)
                for j in range(nsym):
                    e = child + 8 + j * 40
                    out[self._heap_name(heap_addr, self._u64(e))] = self._u64(e + 8)

    def children(self, addr: int) -> Dict[str, int]:
        """name -> object-header address for every link of the group at ``addr``."""
        obj = self._read_object(addr)
        for mtype, payload, paddr in obj.messages:
            if mtype == 0x11:
                btree, heap = self._u64(paddr), self._u64(paddr + 8)
                out: Dict[str, int] = {}
                self._walk_btree(btree, heap, out)
                return out
        return {}

    def resolve(self, path: str) -> int:
        addr = self.root_addr
        for part in [p for p in path.split("/") if p]:
            kids = self.children(addr)
            if part not in kids:
                raise KeyError(f"{path}: no member {part!r}")
            addr = kids[part]
        return addr

    # -- datasets ------------------------------------------------------------
    def dataset(self, addr: int) -> np.ndarray:
        obj = self._read_object(addr)
        shape: Optional[Tuple[int, ...]] = None
        dtype = None
        data_addr = None
        data_size = None
        for mtype, payload, paddr in obj.messages:
            if mtype == 0x01:  # dataspace
                ver, rank, flags = payload[0], payload[1], payload[2]
                p = 8 if ver == 1 else 4
                shape = tuple(struct.unpack_from("<Q", payload, p + 8 * i)[0] for i in range(rank))
            elif mtype == 0x03:  # datatype
                cls = payload[0] & 0x0F
                size = struct.unpack_from("<I", payload, 4)[0]
                if cls == 1 and (payload[1] & 1) == 0:
                    dtype = np.dtype(f"<f{size}")
                elif cls == 0 and (payload[1] & 1) == 0:
                    signed = (payload[1] >> 3) & 1
                    dtype = np.dtype(f"<{'i' if signed else 'u'}{size}")
                else:
                    raise H5FormatError(f"datatype class {cls} unsupported for datasets")
            elif mtype == 0x08:  # layout
                ver, cls = payload[0], payload[1]
                if ver != 3 or cls != 1:
                    raise H5FormatError(f"layout v{ver} class {cls} unsupported (need v3 contiguous)")
                data_addr = struct.unpack_from("<Q", payload, 2)[0]
                data_size = struct.unpack_from("<Q", payload, 10)[0]
            elif mtype == 0x0B:
                raise H5FormatError("filtered datasets unsupported")
        if shape is None or dtype is None or data_addr is None:
            raise H5FormatError("object is not a simple dataset")
        n = int(np.prod(shape)) if shape else 1
        if data_addr == _UNDEF:
            return np.zeros(shape, dtype)
        arr = np.frombuffer(self.buf, dtype=dtype, count=n, offset=data_addr + self.base_addr)
        return arr.reshape(shape).copy()

    # -- attributes ------------------------------------------------------------
    def _gheap_obj(self, coll_addr: int, index: int) -> bytes:
        objs = self._gheap_cache.get(coll_addr)
        if objs is None:
            if self.buf[coll_addr:coll_addr + 4] != b"GCOL":
                raise H5FormatError("bad global heap signature")
            csize = self._u64(coll_addr + 8)
            objs = {}
            p = coll_addr + 16
            end = coll_addr + csize
            while p + 16 <= end:
                idx = self._u16(p)
                osize = self._u64(p + 8)
                if idx == 0:
                    break
                objs[idx] = self.buf[p + 16:p + 16 + osize]
                p += 16 + ((osize + 7) // 8) * 8
            self._gheap_cache[coll_addr] = objs
        return objs[index]

    def attrs(self, addr: int) -> Dict[str, object]:
        """String / vlen-string attributes of an object (others are skipped)."""
        out: Dict[str, object] = {}
        obj = self._read_object(addr)
        for mtype, payload, paddr in obj.messages:
            if mtype != 0x0C:
                continue
            ver = payload[0]
            nsize, tsize, ssize = struct.unpack_from("<HHH", payload, 2)
            p = 8
            if ver == 3:
                p = 9
            pad = (lambda x: ((x + 7) // 8) * 8) if ver == 1 else (lambda x: x)
            name = payload[p:p + nsize].split(b"\x00")[0].decode()
            p += pad(nsize)
            dt = payload[p:p + tsize]
            p += pad(tsize)
            ds = payload[p:p + ssize]
            p += pad(ssize)
            data = payload[p:]
            cls = dt[0] & 0x0F
            rank = ds[1]
            dsp = 8 if ds[0] == 1 else 4
            dims = [struct.unpack_from("<Q", ds, dsp + 8 * i)[0] for i in range(rank)]
            count = int(np.prod(dims)) if dims else 1
            vals: List[str] = []
            if cls == 9:  # variable length (string)
                for i in range(count):
                    ln, coll, idx = struct.unpack_from("<IQI", data, i * 16)
                    vals.append(self._gheap_obj(coll, idx)[:ln].decode("utf-8"))
            elif cls == 3:  # fixed-length string
                size = struct.unpack_from("<I", dt, 4)[0]
                for i in range(count):
                    vals.append(data[i * size:(i + 1) * size].split(b"\x00")[0].decode("utf-8"))
            else:
                continue
            out[name] = vals[0] if not dims else vals
        return out


@dataclass
class DenseLayer:
    """One Keras ``Dense`` layer: ``y = act(x @ kernel + bias)``, kernel layout [in, out]."""
    name: str
    kernel: np.ndarray
    bias: np.ndarray
    activation: str


def _model_config(h5: H5File) -> dict:
    try:
        cfg = h5.attrs(h5.root_addr).get("model_config")
        if isinstance(cfg, str):
            return json.loads(cfg)
    except (H5FormatError, KeyError, IndexError, struct.error):
        pass
    # Fallback: byte-search and brace-match the JSON blob.
    i = h5.buf.find(b'{"class_name"')
    if i < 0:
        raise H5FormatError("model_config not found")
    depth = 0
    for j in range(i, len(h5.buf)):
        c = h5.buf[j]
        if c == 0x7B:
            depth += 1
        elif c == 0x7D:
            depth -= 1
            if depth == 0:
                return json.loads(h5.buf[i:j + 1].decode("utf-8"))
    raise H5FormatError("unterminated model_config")


def load_keras_dense_stack(path: str) -> List[DenseLayer]:
    """Read a Keras ``Sequential`` of ``Dense`` layers from a legacy-HDF5 model file.

    Layer order comes from the embedded ``model_config`` (layer names are not
    stable across the reference's files), weights from
    ``/model_weights/<layer>/<layer>/{kernel:0,bias:0}``.
    """
    h5 = H5File(path)
    cfg = _model_config(h5)
    if cfg.get("class_name") != "Sequential":
        raise H5FormatError(f"unsupported model class {cfg.get('class_name')!r}")
    layers_cfg = cfg["config"]["layers"] if isinstance(cfg["config"], dict) else cfg["config"]
    mw = h5.resolve("/model_weights")
    groups = h5.children(mw)
    out: List[DenseLayer] = []
    for lc in layers_cfg:
        if lc["class_name"] == "InputLayer":
            continue
        if lc["class_name"] != "Dense":
            raise H5FormatError(f"unsupported layer class {lc['class_name']!r}")
        c = lc["config"]
        name = c["name"]
        if not c.get("use_bias", True):
            raise H5FormatError("Dense layers without bias unsupported")
        g = groups[name]
        inner = h5.children(g)
        ds = h5.children(inner[name]) if name in inner else inner
        kernel = h5.dataset(ds["kernel:0"])
        bias = h5.dataset(ds["bias:0"])
        if kernel.dtype != np.float32 or bias.dtype != np.float32:
            raise H5FormatError("expected float32 weights")
        if kernel.shape != (kernel.shape[0], c["units"]) or bias.shape != (c["units"],):
            raise H5FormatError(f"{name}: weight shapes disagree with model_config")
        out.append(DenseLayer(name, kernel, bias, c.get("activation", "linear")))
    return out
