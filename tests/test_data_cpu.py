"""CPU tests of the input-pipeline formats and of the ``audiossl`` import alias.

The LMDB page layout and the legacy pyarrow envelope are restated from their published formats (liblmdb and
pyarrow <= 6 are not in the image: PARITY UNPINNED, see the module docstrings); these tests check the reader against
the writer, both against hand-decoded layout facts, and the dataset semantics of audiossl/datasets/lmdb.py."""
import os
import struct
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _records(n, rng):
    out = []
    for i in range(n):
        label = np.zeros((1, 5), np.float32)
        label[0, i % 5] = 1.0
        out.append(("clip%05d" % i, (rng.standard_normal((1, 700 + 13 * (i % 9))) * 0.1).astype(np.float32), label))
    return out


def test_lmdb_writer_reader_round_trip_deep_tree(tmp_path):
    from audiossl_b200.datasets.lmdb_format import LMDBReader, write_lmdb
    rng = np.random.RandomState(0)
    items = {}
    for i in range(20000):  # inline values: ~1000 leaf pages, two branch levels
        items[b"k%06d" % (i * 7919 % 20000)] = bytes(rng.randint(0, 256, 150 + i % 40, dtype=np.uint8))
    items[b"big"] = bytes(rng.randint(0, 256, 3 * 4096 + 5, dtype=np.uint8))  # overflow run of 4 pages
    items[b"edge"] = bytes(2038 - 8 - 4)  # exactly nodemax: still inline
    items[b"edge+1"] = bytes(2038 - 8 - 6 + 1)  # one byte over: overflow
    path = str(tmp_path / "t.lmdb")
    st = write_lmdb(path, items)
    assert st["depth"] == 3 and st["entries"] == len(items) and st["overflow"] == 4 + 1
    r = LMDBReader(path)
    assert r.stat()["entries"] == len(items) and r.stat()["depth"] == 3 and r.psize == 4096
    assert r.keys() == sorted(items)
    for k in list(items)[::97] + [b"big", b"edge", b"edge+1"]:
        assert bytes(r.get(k)) == items[k]
    assert r.get(b"k") is None and r.get(b"zzz") is None and r.get(b"") is None
    assert os.path.getsize(path) == (st["last_pg"] + 1) * 4096
    r.close()


def test_lmdb_layout_constants(tmp_path):
    """decode the written file by hand with the published offsets (mdb.c MDB_page / MDB_meta / MDB_node)."""
    from audiossl_b200.datasets.lmdb_format import write_lmdb
    path = str(tmp_path / "one.lmdb")
    write_lmdb(path, {b"b": b"22", b"a": b"1"})
    raw = open(path, "rb").read()
    assert len(raw) == 3 * 4096
    for pg, txn in ((0, 0), (1, 1)):
        off = pg * 4096
        assert struct.unpack_from("<QHH", raw, off) == (pg, 0, 0x08)           # P_META
        assert struct.unpack_from("<II", raw, off + 16) == (0xBEEFC0DE, 1)      # magic, data version
        assert struct.unpack_from("<I", raw, off + 16 + 24)[0] == 4096           # free-db md_pad = page size
        assert struct.unpack_from("<Q", raw, off + 16 + 24 + 96 + 8)[0] == txn   # txnid after last_pgno
    depth, = struct.unpack_from("<H", raw, 4096 + 16 + 24 + 48 + 6)
    entries, root = struct.unpack_from("<QQ", raw, 4096 + 16 + 24 + 48 + 32)
    assert (depth, entries, root) == (1, 2, 2)
    leaf = 2 * 4096
    pgno, _, flags, lower, upper = struct.unpack_from("<QHHHH", raw, leaf)
    assert (pgno, flags, lower) == (2, 0x02, 16 + 2 * 2)
    p0, p1 = struct.unpack_from("<HH", raw, leaf + 16)
    assert upper == min(p0, p1) and p0 % 2 == 0 and p1 % 2 == 0
    lo, hi, nflags, ksize = struct.unpack_from("<HHHH", raw, leaf + p0)  # first pointer = smallest key
    assert (lo, hi, nflags, ksize) == (1, 0, 0, 1) and raw[leaf + p0 + 8:leaf + p0 + 10] == b"a1"
    assert raw[leaf + p1 + 8:leaf + p1 + 11] == b"b22"


def test_lmdb_empty_and_directory_layout(tmp_path):
    from audiossl_b200.datasets.lmdb_format import LMDBReader, write_lmdb
    d = tmp_path / "env"
    d.mkdir()
    write_lmdb(str(d / "data.mdb"), {})
    r = LMDBReader(str(d))  # subdir=True layout
    assert len(r) == 0 and r.keys() == [] and r.get(b"x", b"dflt") == b"dflt"
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.lmdb"
        bad.write_bytes(bytes(8192))
        LMDBReader(str(bad))


def test_arrow_legacy_envelope_round_trip_and_layout():
    from audiossl_b200.datasets import arrow_legacy as A
    import pyarrow as pa
    wav = (np.arange(1 * 1601, dtype=np.float32) * 1e-3).reshape(1, 1601)
    label = np.eye(1, 7, 3, dtype=np.float32)
    buf = A.dumps((wav, label))
    assert struct.unpack_from("<iiii", buf, 0) == (0, 0, 2, 0)  # tensors, sparse, ndarrays, buffers
    got = A.loads(buf)
    assert isinstance(got, tuple) and np.array_equal(got[0], wav) and np.array_equal(got[1], label)
    assert got[0].dtype == np.float32 and got[0].shape == (1, 1601)
    # the record batch is plain Arrow IPC: one column "list", a dense union whose children are named by PythonType tag
    rd = pa.ipc.open_stream(pa.BufferReader(buf[16:]))
    batch = rd.read_next_batch()
    assert batch.schema.names == ["list"] and batch.num_rows == 1
    top = batch.column(0)
    assert pa.types.is_union(top.type) and top.type.mode == "dense" and top.type.field(0).name == str(A.TUPLE)
    inner = top.field(0).values
    assert inner.type.field(0).name == str(A.NDARRAY) and inner.field(0).to_pylist() == [0, 1]
    # key list, count, mixed containers
    keys = [b"Y-0abc", b"Yzz", b"a"]
    assert A.loads(A.dumps(keys)) == keys and A.loads(A.dumps(3)) == 3 and A.loads(A.dumps([])) == []
    assert A.loads(A.dumps((1, [2.5, "x", None], (b"q", True)))) == (1, [2.5, "x", None], (b"q", True))
    f64 = np.linspace(0, 1, 11)
    assert A.loads(A.dumps([f64]))[0].dtype == np.float64


def test_lmdb_dataset_semantics(tmp_path):
    import random
    from audiossl_b200.datasets import LMDBDataset, collate_waveforms, write_dataset
    rng = np.random.RandomState(1)
    recs = _records(40, rng)
    st = write_dataset(str(tmp_path / "train.lmdb"), recs)
    assert st["entries"] == 42
    ds = LMDBDataset(str(tmp_path), "train")
    assert len(ds) == 40 and ds.num_classes == 5 and ds.length == 40
    by_name = {n: (w, l) for n, w, l in recs}
    wav, label = ds[3]
    name = ds.keys[3].decode()
    assert wav.dim() == 1 and torch.equal(wav, torch.from_numpy(by_name[name][0][0])) and label.shape == (5,)
    # transform / target_transform / return_key plumbing
    ds2 = LMDBDataset(str(tmp_path), "train", transform=lambda w: (w[:10] * 2, 7),
                      target_transform=lambda x, y: (x + 1, y * 0), return_key=True)
    (x, seven), y, key = ds2[0]
    assert x.shape == (10,) and seven == 7 and float(y.sum()) == 0 and key == ds2.keys[0]
    # subset + cycle walk through the shuffled key list like the reference
    random.seed(3)
    sub = LMDBDataset(str(tmp_path), "train", subset=16)
    assert len(sub) == 16 and sub.start == 16 and sorted(sub.org_keys) == sorted(k.encode() for k in by_name)
    first = list(sub.keys)
    sub.cycle()
    assert sub.keys == sub.org_keys[16:32] and not set(first) & set(sub.keys)
    sub.cycle()  # wraps around
    assert len(sub.keys) == 16 and sub.start == 0
    # fixed-length host batches for the device transform
    wavs, labels = collate_waveforms([ds[i] for i in range(4)], 750)
    assert wavs.shape == (4, 1, 750) and labels.shape == (4, 5)
    w0 = ds[0][0]
    m = min(750, w0.numel())
    assert torch.equal(wavs[0, 0, :m], w0[:m]) and float(wavs[0, 0, m:].abs().sum()) == 0
    # DataLoader workers re-open the map
    import pickle
    clone = pickle.loads(pickle.dumps(ds))
    assert torch.equal(clone[5][0], ds[5][0])


def test_audiossl_alias_resolves_reference_import_paths():
    code = ("import sys; sys.path.insert(0, %r); import audiossl;"
            "from audiossl.methods.atst.model import ATSTLightningModule as A;"
            "from audiossl_b200.methods.atst.model import ATSTLightningModule as B;"
            "from audiossl.methods.atstframe.model import FrameATSTLightningModule;"
            "from audiossl.methods.atst.transform import ATSTTrainTransform;"
            "from audiossl.transforms.common import MinMax, RandomCrop;"
            "from audiossl.transforms.byol_a import Mixup, RandomResizeCrop;"
            "from audiossl.models.atst.audio_transformer import AST_base;"
            "from audiossl.datasets import LMDBDataset;"
            "import audiossl.methods.atstframe.embedding as E;"
            "assert A is B and E.N_BLOCKS == 12;"
            "\ntry:\n import audiossl.methods.mae\n raise SystemExit(1)\nexcept ModuleNotFoundError:\n print('ok')") % ROOT
    import subprocess
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr
