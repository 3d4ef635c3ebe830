"""Minimal pure-Python HDF5 reader/writer for Keras weight files (no h5py in this image).

Covers exactly the subset the reference's artefacts use (all written by h5py under
Keras 2.1/2.2): superblock v0, version-1 object headers, symbol-table groups
(B-tree v1 `TREE` + local heap `HEAP` + `SNOD` nodes), contiguous little-endian
fixed-point / IEEE datasets, and string / numeric attributes (fixed-length strings).

Files read through this: the referee decoders `example_notebooks/referee_decoders/nn_d5_*`
(Keras `model.save`), and the agents' `trained_models/**/final_dqn_weights.h5f`
(keras-rl `DQNAgent.save_weights`, cluster_scripts/d5_dp/0.001/Single_Point_Training_Script.py:160).

`write_h5` emits the same subset, so weights trained here load back into Keras
(`model.load_weights`) -- SURVEY section 8(f) rank 1.
"""
import struct

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF
_H5HL_FREE_NULL = 1


class H5File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.buf = f.read()
        if self.buf[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file: %s" % path)
        ver = self.buf[8]
        if ver != 0:
            raise ValueError("only superblock version 0 is supported (got %d)" % ver)
        if self.buf[13] != 8 or self.buf[14] != 8:
            raise ValueError("only 8-byte offsets/lengths are supported")
        # superblock v0: root symbol table entry starts at byte 56; object header address at +8
        self.root = struct.unpack_from("<Q", self.buf, 56 + 8)[0]
        self._cache = {}

    # ---- object header -------------------------------------------------------------
    def _messages(self, addr):
        b = self.buf
        ver, _, nmsg, _refcnt, size = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise ValueError("only version-1 object headers are supported")
        out = []
        blocks = [(addr + 16, size)]
        while blocks and len(out) < nmsg:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                body = pos + 8
                if mtype == 0x10:   # continuation
                    caddr, clen = struct.unpack_from("<QQ", b, body)
                    blocks.append((caddr, clen))
                out.append((mtype, body, msize))
                pos = body + msize
        return out

    def _group_entries(self, btree, heap):
        b = self.buf
        if b[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap")
        heap_data = struct.unpack_from("<Q", b, heap + 24)[0]
        entries = {}

        def walk(node):
            if b[node:node + 4] != b"TREE":
                raise ValueError("bad B-tree node")
            _ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
            pos = node + 24     # after signature, type, level, entries, two sibling pointers
            for i in range(used):
                child = struct.unpack_from("<Q", b, pos + 8 + i * 16)[0]
                if level > 0:
                    walk(child)
                else:
                    if b[child:child + 4] != b"SNOD":
                        raise ValueError("bad symbol node")
                    nsym = struct.unpack_from("<H", b, child + 6)[0]
                    for s in range(nsym):
                        e = child + 8 + s * 40
                        name_off, ohdr = struct.unpack_from("<QQ", b, e)
                        p = heap_data + name_off
                        q = b.index(b"\0", p)
                        entries[b[p:q].decode()] = ohdr
        walk(btree)
        return entries

    def _info(self, addr):
        if addr in self._cache:
            return self._cache[addr]
        info = {"group": None, "shape": None, "dtype": None, "data": None, "attrs": {}}
        b = self.buf
        for mtype, body, msize in self._messages(addr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", b, body)
                info["group"] = (btree, heap)
            elif mtype == 0x01:
                info["shape"] = self._dataspace(body)
            elif mtype == 0x03:
                info["dtype"] = self._datatype(body)[0]
            elif mtype == 0x08:
                info["data"] = self._layout(body)
            elif mtype == 0x0C:
                name, val = self._attribute(body)
                info["attrs"][name] = val
        self._cache[addr] = info
        return info

    def _dataspace(self, body):
        b = self.buf
        ver, rank, flags = struct.unpack_from("<BBB", b, body)
        if ver == 1:
            off = body + 8
        elif ver == 2:
            off = body + 4
        else:
            raise ValueError("dataspace version %d" % ver)
        return tuple(struct.unpack_from("<%dQ" % rank, b, off)) if rank else ()

    def _datatype(self, body):
        b = self.buf
        cls_ver, bf0, _bf1, _bf2, size = struct.unpack_from("<BBBBI", b, body)
        cls = cls_ver & 0x0F
        if cls == 0:      # fixed point
            signed = (bf0 >> 3) & 1
            return np.dtype("<%s%d" % ("i" if signed else "u", size)), 8 + 4
        if cls == 1:      # float
            return np.dtype("<f%d" % size), 8 + 12
        if cls == 3:      # fixed-length string
            return np.dtype("S%d" % size), 8
        if cls == 9:      # variable-length (strings in the global heap): 16-byte descriptors
            return np.dtype("V%d" % size), 8
        raise ValueError("unsupported datatype class %d" % cls)

    def _layout(self, body):
        b = self.buf
        ver = b[body]
        if ver != 3:
            raise ValueError("only data layout version 3 is supported")
        cls = b[body + 1]
        if cls == 1:      # contiguous
            addr, size = struct.unpack_from("<QQ", b, body + 2)
            return ("contiguous", addr, size)
        if cls == 0:      # compact
            size = struct.unpack_from("<H", b, body + 2)[0]
            return ("compact", body + 4, size)
        raise ValueError("chunked datasets are not supported")

    def _attribute(self, body):
        b = self.buf
        ver, _, name_sz, dt_sz, ds_sz = struct.unpack_from("<BBHHH", b, body)
        if ver != 1:
            raise ValueError("attribute version %d" % ver)
        pad = lambda n: (n + 7) & ~7
        pos = body + 8
        name = b[pos:pos + name_sz].split(b"\0")[0].decode()
        pos += pad(name_sz)
        dtype, _ = self._datatype(pos)
        pos += pad(dt_sz)
        shape = self._dataspace(pos)
        pos += pad(ds_sz)
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(b, dtype=dtype, count=n, offset=pos)
        if dtype.kind == "V":
            vals = [self._vlen(pos + i * dtype.itemsize) for i in range(n)]
            return name, (vals if shape else vals[0])
        if dtype.kind == "S":
            vals = [x.split(b"\0")[0].decode(errors="replace") for x in arr]
            return name, (vals if shape else vals[0])
        return name, (arr.reshape(shape).copy() if shape else arr[0])

    def _vlen(self, pos):
        """Variable-length element: (length u32, global-heap collection address u64, object index u32)."""
        b = self.buf
        length, gcol, index = struct.unpack_from("<IQI", b, pos)
        if b[gcol:gcol + 4] != b"GCOL":
            return None
        csize = struct.unpack_from("<Q", b, gcol + 8)[0]
        p, end = gcol + 16, gcol + csize
        while p + 16 <= end:
            idx, _ref, osize = struct.unpack_from("<HH4xQ", b, p)
            if idx == 0:
                break
            if idx == index:
                return b[p + 16:p + 16 + length].decode(errors="replace")
            p += 16 + ((osize + 7) & ~7)
        return None

    # ---- public --------------------------------------------------------------------
    def _resolve(self, path):
        addr = self.root
        for part in [p for p in path.split("/") if p]:
            info = self._info(addr)
            if info["group"] is None:
                raise KeyError(path)
            ents = self._group_entries(*info["group"])
            if part not in ents:
                raise KeyError(path)
            addr = ents[part]
        return addr

    def keys(self, path="/"):
        info = self._info(self._resolve(path))
        return sorted(self._group_entries(*info["group"])) if info["group"] else []

    def attrs(self, path="/"):
        return dict(self._info(self._resolve(path))["attrs"])

    def is_group(self, path):
        return self._info(self._resolve(path))["group"] is not None

    def __getitem__(self, path):
        info = self._info(self._resolve(path))
        if info["data"] is None:
            raise KeyError("%s is not a dataset" % path)
        kind, addr, size = info["data"]
        shape, dtype = info["shape"], info["dtype"]
        n = int(np.prod(shape)) if shape else 1
        if addr == _UNDEF or n == 0:
            return np.zeros(shape, dtype)
        return np.frombuffer(self.buf, dtype=dtype, count=n, offset=addr).reshape(shape).copy()

    def datasets(self, path="/"):
        """All dataset paths under `path`."""
        out = []
        base = path.rstrip("/")
        for k in self.keys(path):
            p = base + "/" + k
            if self.is_group(p):
                out += self.datasets(p)
            else:
                out.append(p)
        return out


# ======================================================================================
# writer (same subset)
# ======================================================================================
class _Writer:
    def __init__(self):
        self.buf = bytearray()

    def tell(self):
        return len(self.buf)

    def align(self, n=8):
        while len(self.buf) % n:
            self.buf.append(0)

    def write(self, b):
        pos = len(self.buf)
        self.buf += b
        return pos


def _pad8(b):
    return b + b"\0" * ((-len(b)) % 8)


def _dt_msg(dtype):
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        size = dtype.itemsize
        if size == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            bits = (0x20, 0x1F, 0x00)
        elif size == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            bits = (0x20, 0x3F, 0x00)
        else:
            raise ValueError("float size")
        return struct.pack("<BBBBI", 0x11, bits[0], bits[1], bits[2], size) + props
    if dtype.kind in "iu":
        size = dtype.itemsize
        return struct.pack("<BBBBI", 0x10, 0x08 if dtype.kind == "i" else 0, 0, 0, size) + struct.pack("<HH", 0, size * 8)
    if dtype.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0, 0, 0, dtype.itemsize)
    raise ValueError("unsupported dtype %s" % dtype)


def _ds_msg(shape):
    rank = len(shape)
    return struct.pack("<BBB5x", 1, rank, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _msg(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _attr_msg(name, value):
    if isinstance(value, str):
        value = np.array(value.encode(), dtype="S%d" % max(1, len(value.encode())))
    elif isinstance(value, (list, tuple)) and value and isinstance(value[0], str):
        enc = [v.encode() for v in value]
        value = np.array(enc, dtype="S%d" % max(1, max(len(e) for e in enc)))
    else:
        value = np.asarray(value)
    nameb = name.encode() + b"\0"
    dt = _dt_msg(value.dtype)
    ds = _ds_msg(value.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nameb), len(dt), len(ds))
    body += _pad8(nameb) + _pad8(dt) + _pad8(ds) + value.tobytes()
    return _msg(0x0C, body)


def _object_header(w, msgs):
    w.align(8)
    total = sum(len(m) for m in msgs)
    pos = w.write(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, total))
    for m in msgs:
        w.write(m)
    return pos


def _write_dataset(w, arr, attrs):
    arr = np.ascontiguousarray(arr)
    if arr.dtype.byteorder == ">":
        arr = arr.astype(arr.dtype.newbyteorder("<"))
    w.align(8)
    data_addr = w.write(arr.tobytes()) if arr.size else _UNDEF
    layout = struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes)
    msgs = [_msg(0x01, _ds_msg(arr.shape)), _msg(0x03, _dt_msg(arr.dtype), flags=1),
            _msg(0x05, struct.pack("<BBBB", 2, 2, 2, 0x00) + b""),     # fill value v2: alloc late, never write, undefined
            _msg(0x08, layout)]
    msgs += [_attr_msg(k, v) for k, v in attrs.items()]
    return _object_header(w, msgs)


def _write_group(w, children, attrs):
    """children: dict name -> object header address (already written)."""
    names = sorted(children)
    # local heap: offset 0 holds the empty string
    heap_data = bytearray(b"\0" * 8)
    offs = {}
    for n in names:
        offs[n] = len(heap_data)
        heap_data += _pad8(n.encode() + b"\0")
    if len(heap_data) < 16:
        heap_data += b"\0" * (16 - len(heap_data))
    w.align(8)
    hd_addr = w.write(bytes(heap_data))
    # free-list head: libhdf5 (H5HL prefix decode) accepts H5HL_FREE_NULL (= 1, "no free block") or an offset inside the data
    # segment; the undefined address is rejected ("bad heap free list").  The segment above has no free block.
    heap_addr = w.write(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), _H5HL_FREE_NULL, hd_addr))
    # symbol nodes: up to 2K=8 entries per SNOD (group leaf node K = 4)
    K = 4
    snods = []
    for i in range(0, max(1, len(names)), 2 * K):
        chunk = names[i:i + 2 * K]
        body = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
        for n in chunk:
            body += struct.pack("<QQI4x16x", offs[n], children[n], 0)
        body += b"\0" * (40 * (2 * K - len(chunk)))
        w.align(8)
        snods.append((w.write(body), chunk))
    # single-level B-tree (internal K = 16 -> up to 32 children): enough for Keras weight files
    if len(snods) > 32:
        raise ValueError("too many children for a single B-tree node")
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), _UNDEF, _UNDEF)
    tree += struct.pack("<Q", 0)
    for addr, chunk in snods:
        tree += struct.pack("<QQ", addr, offs[chunk[-1]] if chunk else 0)
    tree += b"\0" * (16 * (32 - len(snods)))
    w.align(8)
    tree_addr = w.write(tree)
    msgs = [_msg(0x11, struct.pack("<QQ", tree_addr, heap_addr))] + [_attr_msg(k, v) for k, v in attrs.items()]
    return _object_header(w, msgs), tree_addr, heap_addr


def write_h5(path, tree):
    """tree: nested dict.  Keys starting with '@' are attributes of the enclosing group; a value
    that is a dict is a sub-group; an ndarray (or a (ndarray, attrs) tuple) is a dataset."""
    w = _Writer()
    w.write(b"\0" * 96)      # superblock v0 placeholder (56 bytes + 40-byte root symbol table entry)

    def emit(node):
        attrs = {k[1:]: v for k, v in node.items() if k.startswith("@")}
        children = {}
        for k, v in node.items():
            if k.startswith("@"):
                continue
            if isinstance(v, dict):
                children[k] = emit(v)[0]
            elif isinstance(v, tuple):
                children[k] = _write_dataset(w, v[0], v[1])
            else:
                children[k] = _write_dataset(w, np.asarray(v), {})
        return _write_group(w, children, attrs)

    root_addr, tree_addr, heap_addr = emit(tree)
    w.align(8)
    eof = w.tell()
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
    sb += struct.pack("<QQI4xQQ", 0, root_addr, 1, tree_addr, heap_addr)
    assert len(sb) == 96
    w.buf[:96] = sb
    with open(path, "wb") as f:
        f.write(bytes(w.buf))
