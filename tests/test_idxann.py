"""Faiss ".idxann" files on the host (no GPU): the C++ parser behind solo_idxann_inspect against
files written by the oracle's independent NumPy restatement of the format (oracle/faiss_io.py)."""
import numpy as np
import pytest

from oracle import faiss_io


def _random_index(rng, n=300, d=24, nlist=16, empty=()):
    x = rng.random((n, d), dtype=np.float32)
    cent = rng.random((nlist, d), dtype=np.float32)
    lists = [l for l in range(nlist) if l not in empty]
    assign = rng.choice(lists, n)
    ids = [np.flatnonzero(assign == l).astype(np.int64) for l in range(nlist)]
    return x, cent, assign, ids, [x[i] for i in ids]


def test_oracle_roundtrip(tmp_path):
    rng = np.random.default_rng(5)
    x, cent, assign, ids, vecs = _random_index(rng)
    p = str(tmp_path / "a.idxann")
    faiss_io.write_ivf_flat(p, cent, ids, vecs, nprobe=7)
    r = faiss_io.read_ivf_flat(p)
    assert r["bytes_parsed"] == r["file_size"]
    assert (r["d"], r["ntotal"], r["nlist"], r["nprobe"], r["metric"]) == (24, 300, 16, 7, 0)
    assert np.array_equal(r["centroids"], cent)
    for l in range(16):
        assert np.array_equal(r["list_ids"][l], ids[l]) and np.array_equal(r["list_vecs"][l], vecs[l])


def test_known_bytes(tmp_path):
    """The exact byte string of a two-list, d = 2 index, spelled out from the Faiss layout."""
    import struct
    p = str(tmp_path / "k.idxann")
    cent = np.array([[1, 0], [0, 1]], np.float32)
    faiss_io.write_ivf_flat(p, cent, [np.array([1]), np.array([0, 2])],
                            [np.array([[.5, .25]], np.float32), np.array([[0, 1], [2, 3]], np.float32)], nprobe=1)
    hdr = lambda d, n: struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, 0)
    want = (b"IwFl" + hdr(2, 3) + struct.pack("<QQ", 2, 1) + b"IxFI" + hdr(2, 2) + struct.pack("<Q", 4) +
            struct.pack("<4f", 1, 0, 0, 1) + struct.pack("<BQ", 0, 0) + b"ilar" + struct.pack("<QQ", 2, 8) +
            b"full" + struct.pack("<QQQ", 2, 1, 2) + struct.pack("<2f", .5, .25) + struct.pack("<q", 1) +
            struct.pack("<4f", 0, 1, 2, 3) + struct.pack("<2q", 0, 2))
    assert open(p, "rb").read() == want


@pytest.mark.parametrize("empty,enc", [((), "full"), (tuple(range(3, 16)), "sprs")])
def test_inspect_matches_oracle_writer(tmp_path, empty, enc):
    from ann_solo_b200.index import inspect_index
    rng = np.random.default_rng(6)
    x, cent, assign, ids, vecs = _random_index(rng, empty=empty)
    p = str(tmp_path / "b.idxann")
    faiss_io.write_ivf_flat(p, cent, ids, vecs, ntotal=310, nprobe=3, direct_map=np.arange(5))
    assert faiss_io.read_ivf_flat(p)["size_encoding"] == enc
    info = inspect_index(p)
    assert info["fourcc"] == "IwFl" and info["quantizer_fourcc"] == "IxFI"
    assert (info["d"], info["ntotal"], info["nlist"], info["nprobe"], info["metric"], info["is_trained"]) == \
           (24, 310, 16, 3, 0, 1)
    assert info["code_size"] == 96 and info["nstored"] == 300
    assert info["max_list_len"] == max(len(i) for i in ids)
    import os
    assert info["bytes_parsed"] == os.path.getsize(p)


def test_inspect_rejects_bad_files(tmp_path):
    from ann_solo_b200.index import inspect_index
    rng = np.random.default_rng(7)
    x, cent, assign, ids, vecs = _random_index(rng)
    p = str(tmp_path / "c.idxann")
    faiss_io.write_ivf_flat(p, cent, ids, vecs)
    raw = open(p, "rb").read()
    with pytest.raises(ValueError, match="cannot open"):
        inspect_index(str(tmp_path / "missing.idxann"))
    open(p, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(ValueError, match="truncated"):
        inspect_index(p)
    open(p, "wb").write(b"IxFI" + raw[4:])  # a flat index, not IVF
    with pytest.raises(ValueError, match="not supported"):
        inspect_index(p)
    open(p, "wb").write(raw.replace(b"ilar", b"ilxx"))
    with pytest.raises(ValueError, match="inverted lists"):
        inspect_index(p)
    open(p, "wb").write(b"xy")
    with pytest.raises(ValueError, match="not a Faiss index"):
        inspect_index(p)
