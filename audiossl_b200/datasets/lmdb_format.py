"""LMDB ``data.mdb`` files without liblmdb: a read-only B+tree page walker and a bulk writer.

The reference stores its training set in LMDB (audiossl/datasets/lmdb.py:23-77 reads with the ``lmdb`` binding 1.3.0;
scripts/dataset_preprocess/dataset2lmdb.py:108-129 writes).  The binding is not in this image, so the on-disk format
is restated here from the published layout of LMDB 0.9 (``mdb.c``: MDB_page / MDB_meta / MDB_db / MDB_node), 64-bit
little-endian, default comparison (memcmp, shorter key first), no DUPSORT, no named sub-databases - exactly what the
reference's files use.  PARITY UNPINNED: no file written by liblmdb is available offline to check against; the
reader and the writer are tested against each other and against the layout constants below (tests/test_data_cpu.py).

Layout facts used (all offsets in bytes):
  page header (16): pgno u64 | pad u16 | flags u16 | lower u16, upper u16  (overflow pages: u32 page count instead)
      flags: BRANCH 0x01, LEAF 0x02, OVERFLOW 0x04, META 0x08
  node pointers: u16 offsets from the page start, at byte 16, ``(lower - 16) / 2`` of them, sorted by key
  node header (8): lo u16 | hi u16 | flags u16 | ksize u16, then the key, then (leaf) the data
      leaf: data size = lo | hi << 16; flag BIGDATA 0x01: the data field is the u64 page number of an overflow run
      branch: child page = lo | hi << 16 | flags << 32; node 0 carries an empty key (= minus infinity)
  meta pages 0 and 1 (after the page header): magic 0xBEEFC0DE u32 | version 1 u32 | address u64 | mapsize u64 |
      2 x MDB_db (48: pad u32 | flags u16 | depth u16 | branch_pages, leaf_pages, overflow_pages, entries, root: u64) |
      last_pgno u64 | txnid u64; the meta with the larger txnid is live; dbs[0] is the free list (its ``pad`` field
      holds the page size), dbs[1] the main database; an empty tree has root = 2^64 - 1
  a value is stored on overflow pages when 8 + ksize + dsize > nodemax = (((psize - 16) / 2) & ~1) - 2
"""
import mmap
import os
import struct

MAGIC, VERSION = 0xBEEFC0DE, 1
P_BRANCH, P_LEAF, P_OVERFLOW, P_META = 0x01, 0x02, 0x04, 0x08
F_BIGDATA = 0x01
PAGEHDR, NODEHDR = 16, 8
P_INVALID = (1 << 64) - 1
_META = struct.Struct("<IIQQ")
_DB = struct.Struct("<IHHQQQQQ")


def data_file(path):
    """``lmdb.open(path, subdir=os.path.isdir(path))`` as the reference does: a directory holds data.mdb."""
    return os.path.join(path, "data.mdb") if os.path.isdir(path) else path


class LMDBReader:
    """read-only view of the main database: ``get(key)``, ``keys()``, ``items()``, ``stat()``."""

    def __init__(self, path):
        self.path = data_file(path)
        self._f = open(self.path, "rb")
        self._m = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        metas = []
        psize = None
        for pg in (0, 1):
            # the page size is not known before a meta page is parsed; page 1 starts at the size page 0 declares
            off = 0 if pg == 0 else psize
            if off is None or off + PAGEHDR + _META.size + 2 * _DB.size + 16 > len(self._m):
                break
            flags = struct.unpack_from("<H", self._m, off + 10)[0]
            magic, version, _, mapsize = _META.unpack_from(self._m, off + PAGEHDR)
            if not (flags & P_META) or magic != MAGIC:
                if pg == 0:
                    raise ValueError("%s is not an LMDB data file (bad magic)" % self.path)
                break
            if version != VERSION:
                raise ValueError("unsupported LMDB data version %d" % version)
            dbs = [_DB.unpack_from(self._m, off + PAGEHDR + _META.size + i * _DB.size) for i in range(2)]
            last_pg, txnid = struct.unpack_from("<QQ", self._m, off + PAGEHDR + _META.size + 2 * _DB.size)
            if psize is None:
                psize = dbs[0][0]
            metas.append(dict(dbs=dbs, last_pg=last_pg, txnid=txnid, mapsize=mapsize))
        meta = max(metas, key=lambda m: m["txnid"])
        self.psize = psize
        self.meta = meta
        _, self.flags, self.depth, self.branch_pages, self.leaf_pages, self.overflow_pages, self.entries, self.root = meta["dbs"][1]

    def close(self):
        self._m.close()
        self._f.close()

    def stat(self):
        return dict(psize=self.psize, depth=self.depth, branch_pages=self.branch_pages, leaf_pages=self.leaf_pages,
                    overflow_pages=self.overflow_pages, entries=self.entries)

    def __len__(self):
        return self.entries

    # ---- page access
    def _page(self, pgno):
        off = pgno * self.psize
        _, _, flags, lower, upper = struct.unpack_from("<QHHHH", self._m, off)
        return off, flags, (lower - PAGEHDR) // 2

    def _node(self, off, i):
        p = off + struct.unpack_from("<H", self._m, off + PAGEHDR + 2 * i)[0]
        lo, hi, flags, ksize = struct.unpack_from("<HHHH", self._m, p)
        return p, lo, hi, flags, ksize

    def _key(self, p, ksize):
        return bytes(self._m[p + NODEHDR:p + NODEHDR + ksize])

    def _value(self, p, lo, hi, flags, ksize):
        size = lo | (hi << 16)
        d = p + NODEHDR + ksize
        if flags & F_BIGDATA:
            pg = struct.unpack_from("<Q", self._m, d)[0]
            start = pg * self.psize + PAGEHDR
            return memoryview(self._m)[start:start + size]
        return memoryview(self._m)[d:d + size]

    def get(self, key, default=None):
        """the value stored under ``key`` as a read-only memoryview into the map (zero copy), or ``default``."""
        if self.root == P_INVALID:
            return default
        key = bytes(key)
        pgno = self.root
        while True:
            off, flags, n = self._page(pgno)
            if flags & P_LEAF:
                lo_i, hi_i = 0, n - 1
                while lo_i <= hi_i:
                    mid = (lo_i + hi_i) // 2
                    p, lo, hi, nflags, ksize = self._node(off, mid)
                    k = self._key(p, ksize)
                    if k == key:
                        return self._value(p, lo, hi, nflags, ksize)
                    if k < key:
                        lo_i = mid + 1
                    else:
                        hi_i = mid - 1
                return default
            # branch: the last node whose key <= search key (node 0 has the empty key)
            lo_i, hi_i = 1, n - 1
            child = 0
            while lo_i <= hi_i:
                mid = (lo_i + hi_i) // 2
                p, lo, hi, nflags, ksize = self._node(off, mid)
                if self._key(p, ksize) <= key:
                    child = mid
                    lo_i = mid + 1
                else:
                    hi_i = mid - 1
            p, lo, hi, nflags, ksize = self._node(off, child)
            pgno = lo | (hi << 16) | (nflags << 32)

    def items(self):
        """(key, value view) in key order."""
        if self.root == P_INVALID:
            return
        stack = [self.root]
        while stack:
            pgno = stack.pop()
            off, flags, n = self._page(pgno)
            if flags & P_LEAF:
                for i in range(n):
                    p, lo, hi, nflags, ksize = self._node(off, i)
                    yield self._key(p, ksize), self._value(p, lo, hi, nflags, ksize)
            else:
                children = []
                for i in range(n):
                    p, lo, hi, nflags, ksize = self._node(off, i)
                    children.append(lo | (hi << 16) | (nflags << 32))
                stack.extend(reversed(children))

    def keys(self):
        return [k for k, _ in self.items()]


def write_lmdb(path, items, psize=4096, mapsize=1 << 30):
    """bulk-write ``items`` ({key bytes: value bytes-like} or iterable of pairs) as a fresh single-file LMDB
    environment (``subdir=False``, what dataset2lmdb.py produces): sorted leaves filled left to right, values larger
    than nodemax on overflow runs, branch levels built bottom-up, both meta pages written."""
    pairs = sorted((bytes(k), v) for k, v in (items.items() if hasattr(items, "items") else items))
    for a, b in zip(pairs, pairs[1:]):
        if a[0] == b[0]:
            raise ValueError("duplicate key %r" % a[0])
    nodemax = (((psize - PAGEHDR) // 2) & ~1) - 2
    pages = {}      # pgno -> bytes (overflow runs are stored under their first page number)
    next_pg = [2]
    counts = dict(branch=0, leaf=0, overflow=0)

    def alloc(n=1):
        pg = next_pg[0]
        next_pg[0] += n
        return pg

    def build_page(flags, nodes):
        """nodes: list of packed node bytes (already even-sized) -> page bytes (nodes packed from the top down)."""
        pgno = alloc()
        buf = bytearray(psize)
        upper = psize
        ptrs = []
        for nd in nodes:
            upper -= len(nd)
            buf[upper:upper + len(nd)] = nd
            ptrs.append(upper)
        lower = PAGEHDR + 2 * len(nodes)
        assert lower <= upper
        struct.pack_into("<QHHHH", buf, 0, pgno, 0, flags, lower, upper)
        struct.pack_into("<%dH" % len(ptrs), buf, PAGEHDR, *ptrs)
        pages[pgno] = bytes(buf)
        return pgno

    def even(b):
        return b + b"\0" if len(b) & 1 else b

    def leaf_node(key, val):
        val = memoryview(val).cast("B") if not isinstance(val, (bytes, bytearray)) else val
        size = len(val)
        if NODEHDR + len(key) + size > nodemax:
            npg = (PAGEHDR - 1 + size) // psize + 1
            pg = alloc(npg)
            run = bytearray(npg * psize)
            struct.pack_into("<QHHI", run, 0, pg, 0, P_OVERFLOW, npg)
            run[PAGEHDR:PAGEHDR + size] = val
            pages[pg] = bytes(run)
            counts["overflow"] += npg
            return even(struct.pack("<HHHH", size & 0xFFFF, size >> 16, F_BIGDATA, len(key)) + key + struct.pack("<Q", pg))
        return even(struct.pack("<HHHH", size & 0xFFFF, size >> 16, 0, len(key)) + key + bytes(val))

    def branch_node(key, child):
        return even(struct.pack("<HHHH", child & 0xFFFF, (child >> 16) & 0xFFFF, child >> 32, len(key)) + key)

    def pack_level(flags, entries, make_node, first_key_empty):
        """entries: [(key, payload)] -> [(first key of page, pgno)]"""
        out, cur, cur_first, used = [], [], None, PAGEHDR
        for key, payload in entries:
            nd = make_node(b"" if (first_key_empty and not cur) else key, payload)
            if cur and used + len(nd) + 2 > psize:
                out.append((cur_first, build_page(flags, cur)))
                cur, cur_first, used = [], None, PAGEHDR
                if first_key_empty:  # this node now opens a branch page: its key becomes the empty "minus infinity"
                    nd = make_node(b"", payload)
            if not cur:
                cur_first = key
            cur.append(nd)
            used += len(nd) + 2
        if cur:
            out.append((cur_first, build_page(flags, cur)))
        return out

    depth, root = 0, P_INVALID
    if pairs:
        level = pack_level(P_LEAF, pairs, leaf_node, False)
        counts["leaf"] = len(level)
        depth = 1
        while len(level) > 1:
            level = pack_level(P_BRANCH, level, branch_node, True)
            counts["branch"] += len(level)
            depth += 1
        root = level[0][1]
    last_pg = next_pg[0] - 1

    def meta_page(pgno, txnid):
        buf = bytearray(psize)
        struct.pack_into("<QHHHH", buf, 0, pgno, 0, P_META, 0, 0)
        _META.pack_into(buf, PAGEHDR, MAGIC, VERSION, 0, mapsize)
        _DB.pack_into(buf, PAGEHDR + _META.size, psize, 0, 0, 0, 0, 0, 0, P_INVALID)  # free list: empty
        if txnid == 0:  # the initial meta of a fresh environment: empty main database
            _DB.pack_into(buf, PAGEHDR + _META.size + _DB.size, 0, 0, 0, 0, 0, 0, 0, P_INVALID)
            struct.pack_into("<QQ", buf, PAGEHDR + _META.size + 2 * _DB.size, 1, 0)
        else:
            _DB.pack_into(buf, PAGEHDR + _META.size + _DB.size, 0, 0, depth, counts["branch"], counts["leaf"],
                          counts["overflow"], len(pairs), root)
            struct.pack_into("<QQ", buf, PAGEHDR + _META.size + 2 * _DB.size, max(last_pg, 1), txnid)
        return bytes(buf)

    with open(path, "wb") as f:
        f.write(meta_page(0, 0))
        f.write(meta_page(1, 1))
        for pg in sorted(pages):
            assert f.tell() == pg * psize
            f.write(pages[pg])
    return dict(depth=depth, entries=len(pairs), last_pg=last_pg, **counts)
