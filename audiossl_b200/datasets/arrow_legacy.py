"""The legacy ``pyarrow.serialize`` envelope the reference's LMDB records use
(scripts/dataset_preprocess/dataset2lmdb.py:16-23 ``pa.serialize(obj).to_buffer()``; audiossl/datasets/lmdb.py:32-33,48,57
``pa.deserialize``; pinned pyarrow 6.0.1, setup.cfg:21).  ``serialize`` / ``deserialize`` were removed from pyarrow
(absent from the installed 24.0), so the envelope is restated from the published Arrow C++ sources of that release
(``arrow/python/serialize.cc`` ``SerializedPyObject::WriteTo`` / ``SequenceBuilder``, ``deserialize.cc``
``GetPythonTypes`` / ``GetValue``); the Arrow IPC pieces inside it (record-batch stream, tensor messages) are read and
written by the installed pyarrow's own ``pyarrow.ipc``.  PARITY UNPINNED: no buffer produced by pyarrow <= 6 is
available offline; reader and writer are tested against each other (tests/test_data_cpu.py).

Envelope:
    int32 num_tensors | int32 num_sparse_tensors | int32 num_ndarrays | int32 num_buffers
    (pad to 8)  Arrow IPC stream holding ONE record batch with ONE column "list": a dense union with one element per
                top-level object; the union's child fields are NAMED by the decimal PythonType tag they carry
                (0 NONE, 1 BOOL, 2 INT, 3 PY2INT, 4 BYTES, 5 STRING, 6 HALF_FLOAT, 7 FLOAT, 8 DOUBLE, 9 DATE64, 10 LIST,
                11 DICT, 12 TUPLE, 13 SET, 14 TENSOR, 15 NDARRAY, 16 BUFFER); LIST / TUPLE / SET children are
                list<dense union> (recursively), TENSOR / NDARRAY / BUFFER children are int32 indices into the
                blobs that follow
    (pad to 64) per tensor, then per ndarray: an Arrow Tensor IPC message, each followed by padding to 64
    per buffer: int64 size | bytes
"""
import struct

import numpy as np

NONE, BOOL, INT, PY2INT, BYTES, STRING, HALF_FLOAT, FLOAT, DOUBLE, DATE64, LIST, DICT, TUPLE, SET, TENSOR, NDARRAY, BUFFER = range(17)


def _pa():
    import pyarrow as pa
    return pa


# --------------------------------------------------------------------------------------------- reader
def loads(buf):
    """bytes-like -> the Python object (tuples, lists, ints, floats, bytes, str, bools, None, numpy arrays)."""
    pa = _pa()
    mv = memoryview(buf).cast("B") if not isinstance(buf, memoryview) else buf.cast("B")
    n_tensors, n_sparse, n_ndarrays, n_buffers = struct.unpack_from("<iiii", mv, 0)
    if n_sparse:
        raise NotImplementedError("sparse tensors are not used by the reference's records")
    src = pa.BufferReader(pa.py_buffer(mv))
    src.seek(16)
    reader = pa.ipc.open_stream(src)
    batch = reader.read_next_batch()
    try:
        reader.read_next_batch()  # consume the end-of-stream marker so the position is past the stream
    except StopIteration:
        pass

    def align(k):
        pos = src.tell()
        src.seek((pos + k - 1) // k * k)

    align(64)
    tensors = []
    for _ in range(n_tensors + n_ndarrays):
        tensors.append(pa.ipc.read_tensor(src).to_numpy())
        align(64)
    ndarrays = tensors[n_tensors:]
    buffers = []
    for _ in range(n_buffers):
        size = struct.unpack("<q", src.read(8))[0]
        buffers.append(src.read(size))
    col = batch.column(0)
    out = _decode_union(col, 0, len(col), dict(tensors=tensors[:n_tensors], ndarrays=ndarrays, buffers=buffers))
    return out[0] if len(out) == 1 else out


def _decode_union(arr, start, stop, blobs):
    if stop <= start:  # an empty container: its union may have no children at all
        return []
    tags = [int(arr.type.field(i).name) for i in range(arr.type.num_fields)]
    codes = {c: i for i, c in enumerate(arr.type.type_codes)}
    type_ids = arr.type_codes.to_numpy(zero_copy_only=False)
    offsets = arr.offsets.to_numpy(zero_copy_only=False)
    children = [arr.field(i) for i in range(arr.type.num_fields)]
    out = []
    for i in range(start, stop):
        ci = codes[int(type_ids[i])]
        tag, child, j = tags[ci], children[ci], int(offsets[i])
        if tag == NONE:
            out.append(None)
        elif tag in (BOOL, INT, PY2INT, BYTES, STRING, HALF_FLOAT, FLOAT, DOUBLE, DATE64):
            out.append(child[j].as_py())
        elif tag in (LIST, TUPLE, SET):
            lo, hi = child.offsets[j].as_py(), child.offsets[j + 1].as_py()
            vals = _decode_union(child.values, lo, hi, blobs)
            out.append(vals if tag == LIST else tuple(vals) if tag == TUPLE else set(vals))
        elif tag == DICT:
            lo, hi = child.offsets[j].as_py(), child.offsets[j + 1].as_py()
            st = child.values
            keys = _decode_union(st.field("keys"), lo, hi, blobs)
            vals = _decode_union(st.field("vals"), lo, hi, blobs)
            out.append(dict(zip(keys, vals)))
        elif tag == TENSOR:
            out.append(blobs["tensors"][child[j].as_py()])
        elif tag == NDARRAY:
            out.append(blobs["ndarrays"][child[j].as_py()])
        elif tag == BUFFER:
            out.append(blobs["buffers"][child[j].as_py()])
        else:
            raise NotImplementedError("PythonType tag %d" % tag)
    return out


# --------------------------------------------------------------------------------------------- writer
class _Seq:
    """SequenceBuilder: a dense union whose children appear in order of first use, named by their tag."""

    def __init__(self, blobs):
        self.blobs = blobs
        self.type_ids, self.offsets = [], []
        self.order, self.vals = [], {}

    def _slot(self, tag):
        if tag not in self.vals:
            self.vals[tag] = []
            self.order.append(tag)
        self.type_ids.append(self.order.index(tag))
        self.offsets.append(len(self.vals[tag]))
        return self.vals[tag]

    def append(self, obj):
        if obj is None:
            self._slot(NONE).append(None)
        elif isinstance(obj, (bool, np.bool_)):
            self._slot(BOOL).append(bool(obj))
        elif isinstance(obj, (int, np.integer)):
            self._slot(INT).append(int(obj))
        elif isinstance(obj, (float, np.floating)):
            self._slot(DOUBLE).append(float(obj))
        elif isinstance(obj, bytes):
            self._slot(BYTES).append(obj)
        elif isinstance(obj, str):
            self._slot(STRING).append(obj)
        elif isinstance(obj, np.ndarray):
            self.blobs.append(np.ascontiguousarray(obj))
            self._slot(NDARRAY).append(len(self.blobs) - 1)
        elif isinstance(obj, (list, tuple)):
            sub = _Seq(self.blobs)
            for o in obj:
                sub.append(o)
            self._slot(LIST if isinstance(obj, list) else TUPLE).append(sub)
        else:
            raise NotImplementedError("cannot serialize %r" % type(obj))

    def finish(self):
        pa = _pa()
        children, names = [], []
        for tag in self.order:
            v = self.vals[tag]
            names.append(str(tag))
            if tag == NONE:
                children.append(pa.nulls(len(v)))
            elif tag == BOOL:
                children.append(pa.array(v, pa.bool_()))
            elif tag == INT:
                children.append(pa.array(v, pa.int64()))
            elif tag == DOUBLE:
                children.append(pa.array(v, pa.float64()))
            elif tag == BYTES:
                children.append(pa.array(v, pa.binary()))
            elif tag == STRING:
                children.append(pa.array(v, pa.string()))
            elif tag == NDARRAY:
                children.append(pa.array(v, pa.int32()))
            else:  # LIST / TUPLE: list<dense union> over the concatenation of the sub-sequences
                # the sub-sequences are concatenated through ONE builder so that all lists share one union type
                merged = _Seq(self.blobs)
                bounds = [0]
                for sub in v:
                    merged.extend(sub)
                    bounds.append(len(merged.type_ids))
                children.append(pa.ListArray.from_arrays(pa.array(bounds, pa.int32()), merged.finish()))
        return pa.UnionArray.from_dense(pa.array(self.type_ids, pa.int8()), pa.array(self.offsets, pa.int32()),
                                        children, names)

    def extend(self, other):
        """append the elements of another (unfinished) sequence, preserving their order."""
        for tid, off in zip(other.type_ids, other.offsets):
            tag = other.order[tid]
            self._slot(tag).append(other.vals[tag][off])


def dumps(obj):
    """the bytes ``pa.serialize(obj).to_buffer()`` wrote for the object kinds the reference stores: tuples / lists of
    numpy arrays, bytes, str, int, float."""
    pa = _pa()
    blobs = []
    top = _Seq(blobs)
    top.append(obj)
    col = top.finish()
    batch = pa.RecordBatch.from_arrays([col], ["list"])
    sink = pa.BufferOutputStream()
    sink.write(struct.pack("<iiii", 0, 0, len(blobs), 0))  # 16 bytes: already 8-aligned
    with pa.ipc.new_stream(sink, batch.schema) as w:
        w.write_batch(batch)

    def pad(k):
        rem = sink.tell() % k
        if rem:
            sink.write(b"\0" * (k - rem))

    pad(64)
    for a in blobs:
        pa.ipc.write_tensor(pa.Tensor.from_numpy(a), sink)
        pad(64)
    return sink.getvalue().to_pybytes()
