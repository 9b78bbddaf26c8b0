"""TEST INFRASTRUCTURE ONLY (oracle): NumPy restatement of the Faiss file format of the one index
type the reference serialises — ``faiss.write_index(IndexIVFFlat(IndexFlatIP(d), d, nlist,
METRIC_INNER_PRODUCT))`` at reference spectral_library.py:167-181, read back at :490.

Faiss is a third-party dependency, unpinned (reference setup.py:99) and absent here, so this
restates the published serialisation (faiss/impl/index_write.cpp: ``write_index``,
``write_ivf_header``, ``write_index_header``, ``write_direct_map``, ``write_InvertedLists``) —
**parity unpinned**: no reference test or fixture holds an index file. The product's C++ reader /
writer (ann-solo_b200/csrc/faiss_io.cu) is checked against this independent NumPy version in both
directions.

Layout (little-endian): "IwFl" | header(d i32, ntotal i64, 1<<20 i64, 1<<20 i64, is_trained u8,
metric i32) | nlist u64 | nprobe u64 | "IxFI" | header | count u64 | centroids f32 | direct-map type
u8 | n u64 | n*i64 | "ilar" | nlist u64 | code_size u64 | ("full" | nlist u64 | sizes u64[nlist]) or
("sprs" | 2m u64 | (list, size) u64 pairs) | per non-empty list: codes, then ids i64.
"""
from __future__ import annotations

import struct
from typing import List, Optional

import numpy as np


def _header(d: int, ntotal: int, metric: int = 0, is_trained: bool = True) -> bytes:
    return struct.pack("<iqqq?i", d, ntotal, 1 << 20, 1 << 20, is_trained, metric)


def write_ivf_flat(path: str, centroids: np.ndarray, list_ids: List[np.ndarray], list_vecs: List[np.ndarray],
                   ntotal: Optional[int] = None, nprobe: int = 1, sparse_sizes: Optional[bool] = None,
                   metric: int = 0, direct_map: Optional[np.ndarray] = None) -> None:
    centroids = np.ascontiguousarray(centroids, "<f4")
    nlist, d = centroids.shape
    assert len(list_ids) == len(list_vecs) == nlist
    sizes = np.array([len(i) for i in list_ids], "<u8")
    if ntotal is None:
        ntotal = int(sizes.sum())
    with open(path, "wb") as f:
        f.write(b"IwFl")
        f.write(_header(d, ntotal, metric))
        f.write(struct.pack("<QQ", nlist, nprobe))
        f.write(b"IxFI")
        f.write(_header(d, nlist, metric))
        f.write(struct.pack("<Q", centroids.size))
        f.write(centroids.tobytes())
        if direct_map is None:
            f.write(struct.pack("<BQ", 0, 0))
        else:  # DirectMap::Array
            dm = np.ascontiguousarray(direct_map, "<i8")
            f.write(struct.pack("<BQ", 1, dm.size))
            f.write(dm.tobytes())
        f.write(b"ilar")
        f.write(struct.pack("<QQ", nlist, 4 * d))
        non0 = int((sizes > 0).sum())
        if sparse_sizes is None:
            sparse_sizes = not non0 > nlist // 2
        if not sparse_sizes:
            f.write(b"full")
            f.write(struct.pack("<Q", nlist))
            f.write(sizes.tobytes())
        else:
            f.write(b"sprs")
            nz = np.flatnonzero(sizes)
            pairs = np.stack([nz.astype("<u8"), sizes[nz]], axis=1)
            f.write(struct.pack("<Q", pairs.size))
            f.write(np.ascontiguousarray(pairs).tobytes())
        for ids, vecs in zip(list_ids, list_vecs):
            if len(ids) == 0:
                continue
            vecs = np.ascontiguousarray(vecs, "<f4")
            assert vecs.shape == (len(ids), d)
            f.write(vecs.tobytes())
            f.write(np.ascontiguousarray(ids, "<i8").tobytes())


def read_ivf_flat(path: str) -> dict:
    buf = open(path, "rb").read()
    pos = 0

    def take(fmt):
        nonlocal pos
        v = struct.unpack_from("<" + fmt, buf, pos)
        pos += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def fourcc():
        nonlocal pos
        s = buf[pos:pos + 4].decode("ascii")
        pos += 4
        return s

    def arr(dtype, n):
        nonlocal pos
        a = np.frombuffer(buf, dtype, n, pos).copy()
        pos += a.nbytes
        return a

    def header():
        d, ntotal, _, _, trained, metric = take("iqqq?i")
        if metric > 1:
            take("f")
        return d, ntotal, trained, metric

    out = {"fourcc": fourcc()}
    assert out["fourcc"] == "IwFl", out["fourcc"]
    d, ntotal, trained, metric = header()
    nlist, nprobe = take("QQ")
    out.update(d=d, ntotal=ntotal, is_trained=trained, metric=metric, nlist=nlist, nprobe=nprobe)
    out["quantizer_fourcc"] = fourcc()
    qd, qn, _, _ = header()
    cnt = take("Q")
    assert (qd, qn, cnt) == (d, nlist, nlist * d)
    out["centroids"] = arr("<f4", cnt).reshape(nlist, d)
    dm_type = take("B")
    out["direct_map"] = arr("<i8", take("Q"))
    if dm_type == 2:
        arr("<i8", 2 * take("Q"))
    assert fourcc() == "ilar"
    il_nlist, code_size = take("QQ")
    assert il_nlist == nlist and code_size == 4 * d
    enc = fourcc()
    nv = take("Q")
    sizes = np.zeros(nlist, np.int64)
    if enc == "full":
        sizes[:] = arr("<u8", nv)
    else:
        assert enc == "sprs"
        pairs = arr("<u8", nv).reshape(-1, 2)
        sizes[pairs[:, 0].astype(np.int64)] = pairs[:, 1]
    out["size_encoding"] = enc
    out["list_ids"], out["list_vecs"] = [], []
    for n in sizes:
        out["list_vecs"].append(arr("<f4", int(n) * d).reshape(int(n), d))
        out["list_ids"].append(arr("<i8", int(n)))
    out["bytes_parsed"] = pos
    out["file_size"] = len(buf)
    return out
