"""Seeded synthetic spectral libraries and query sets.

There is no network and no data set in the tree, so tests and ``bench.py`` search spectra drawn
from this generator. Spectra are produced already *processed* — what the reference's
``process_spectrum`` (src/ann_solo/spectrum.py:57-119) hands to the hot path: at most 50 peaks,
m/z ascending, rank-scaled intensities (``scaling='rank'``, spectrum.py:104-110) L2-normalised
to float32.

A *peak store* is a dict of contiguous arrays (CSR over spectra)::

    mz      float32 [n_peaks]     inten  float32 [n_peaks]    chg  uint8 [n_peaks]
    off     int64   [n + 1]       prec_mz float64 [n]          prec_z int32 [n]
    valid   uint8   [n]           is_decoy uint8 [n] (library only)
"""
from __future__ import annotations

import numpy as np

PROTON = 1.00728
MAX_RANK = 50
# most frequent non-zero mass differences in the reference's open-search results
# (/root/reference/notebooks/kim2014_stats.ipynb:383-394)
MODS = np.array([57.022, 27.996, 0.994, 15.995, -0.986, 14.016, -17.025, -18.010, 1.988])


def _segments(counts: np.ndarray):
    off = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    seg = np.repeat(np.arange(len(counts), dtype=np.int64), counts)
    pos = np.arange(off[-1], dtype=np.int64) - off[seg]
    return off, seg, pos


def _sorted_uniform(rng, counts, lo, hi_per_seg):
    """m/z ascending and uniform in [lo, hi) per spectrum, as float32, strictly increasing."""
    off, seg, pos = _segments(counts)
    # order statistics of U(0,1): normalised cumulative sums of exponentials
    e = rng.standard_exponential(off[-1] + len(counts))  # one extra gap per spectrum
    eoff = off + np.arange(len(counts) + 1)
    cs = np.cumsum(e)
    start = np.concatenate(([0.0], cs))[eoff[:-1]]
    total = cs[eoff[1:] - 1] - start
    idx = eoff[seg] + pos
    u = (cs[idx] - start[seg]) / total[seg]
    mz = (lo + u * (hi_per_seg[seg] - lo)).astype(np.float32)
    return off, seg, pos, mz


def _fix_increasing(mz, off):
    """Nudge float32 ties so every spectrum is strictly ascending (cheap, vectorised)."""
    for _ in range(4):
        d = np.diff(mz)
        bad = np.flatnonzero(d <= 0) + 1
        bad = bad[~np.isin(bad, off[1:-1])]
        if len(bad) == 0:
            break
        mz[bad] = np.nextafter(mz[bad - 1], np.float32(np.inf)) + np.float32(1e-3)
    return mz


def _rank_intensity(rng, counts, seg):
    """Distinct integer ranks MAX_RANK, MAX_RANK-1, ... randomly permuted, then L2-normalised."""
    n = len(counts)
    key = rng.random((n, MAX_RANK))
    col = np.arange(MAX_RANK)[None, :]
    key[col >= counts[:, None]] = 2.0  # padding sorts last
    order = np.argsort(key, axis=1, kind="stable")
    rank = np.empty((n, MAX_RANK), np.float32)
    np.put_along_axis(rank, order, (MAX_RANK - col).astype(np.float32).repeat(n, 0), axis=1)
    valid = col < counts[:, None]
    r = rank[valid]  # row-major == CSR order
    ss = np.zeros(n, np.float64)
    np.add.at(ss, seg, r.astype(np.float64) ** 2)
    nrm = np.sqrt(ss).astype(np.float32)
    return (r / nrm[seg]).astype(np.float32)


def renormalise_rank(inten_raw: np.ndarray, off: np.ndarray) -> np.ndarray:
    """Re-rank (most intense -> MAX_RANK) and L2-normalise each spectrum of a CSR array."""
    counts = np.diff(off)
    seg = np.repeat(np.arange(len(counts)), counts)
    order = np.lexsort((-inten_raw, seg))  # by spectrum, then intensity descending
    pos = np.arange(len(seg)) - off[seg]
    rank = np.empty(len(seg), np.float32)
    rank[order] = (MAX_RANK - pos).astype(np.float32)
    ss = np.zeros(len(counts), np.float64)
    np.add.at(ss, seg, rank.astype(np.float64) ** 2)
    nrm = np.sqrt(ss).astype(np.float32)
    return (rank / nrm[seg]).astype(np.float32)


def make_library(n_targets: int, decoy_fraction: float = 0.5, seed: int = 1, decoy_seed: int = 2,
                 charges=(2, 3, 4), charge_p=(0.5, 0.4, 0.1)) -> dict:
    """Targets + decoys (decoys = copies of random targets with 70 % of fragment m/z re-drawn,
    same precursor), rows in id order: targets first, then decoys."""
    rng = np.random.default_rng(seed)
    z = rng.choice(np.asarray(charges, np.int32), size=n_targets, p=charge_p).astype(np.int32)
    mass = np.clip(rng.lognormal(mean=np.log(1600.0), sigma=0.45, size=n_targets), 700.0, 4000.0)
    prec_mz = mass / z + PROTON
    counts = rng.integers(20, MAX_RANK + 1, size=n_targets)
    hi = np.minimum(2000.0, mass)
    off, seg, pos, mz = _sorted_uniform(rng, counts, 100.0, hi)
    mz = _fix_increasing(mz, off)
    inten = _rank_intensity(rng, counts, seg)
    chg = rng.choice(np.array([0, 1, 2], np.uint8), size=off[-1], p=(0.2, 0.65, 0.15)).astype(np.uint8)
    lib = dict(mz=mz, inten=inten, chg=chg, off=off, prec_mz=prec_mz.astype(np.float64), prec_z=z,
               valid=np.ones(n_targets, np.uint8), is_decoy=np.zeros(n_targets, np.uint8))
    n_decoys = int(round(n_targets * decoy_fraction))
    if n_decoys:
        drng = np.random.default_rng(decoy_seed)
        src = drng.integers(0, n_targets, size=n_decoys)
        dec = take_spectra(lib, src)
        redraw = drng.random(len(dec["mz"])) < 0.7
        dseg = np.repeat(np.arange(n_decoys), np.diff(dec["off"]))
        dhi = hi[src][dseg]
        new_mz = (100.0 + drng.random(len(dec["mz"])) * (dhi - 100.0)).astype(np.float32)
        dec["mz"] = np.where(redraw, new_mz, dec["mz"])
        sort_store_by_mz(dec)
        dec["is_decoy"][:] = 1
        lib = concat_stores([lib, dec])
    return lib


def take_spectra(store: dict, rows: np.ndarray) -> dict:
    rows = np.asarray(rows, np.int64)
    counts = (store["off"][rows + 1] - store["off"][rows]).astype(np.int64)
    off, seg, pos = _segments(counts)
    src = store["off"][rows][seg] + pos
    out = dict(mz=store["mz"][src].copy(), inten=store["inten"][src].copy(), chg=store["chg"][src].copy(),
               off=off, prec_mz=store["prec_mz"][rows].copy(), prec_z=store["prec_z"][rows].copy(),
               valid=store["valid"][rows].copy())
    if "is_decoy" in store:
        out["is_decoy"] = store["is_decoy"][rows].copy()
    return out


def concat_stores(stores) -> dict:
    out = {}
    for k in ("mz", "inten", "chg", "prec_mz", "prec_z", "valid", "is_decoy"):
        if all(k in s for s in stores):
            out[k] = np.concatenate([s[k] for s in stores])
    counts = np.concatenate([np.diff(s["off"]) for s in stores])
    out["off"] = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=out["off"][1:])
    return out


def sort_store_by_mz(store: dict) -> None:
    """Sort every spectrum's peaks by m/z in place (keeps intensity / charge attached)."""
    counts = np.diff(store["off"])
    seg = np.repeat(np.arange(len(counts)), counts)
    order = np.lexsort((store["mz"], seg))
    for k in ("mz", "inten", "chg"):
        store[k] = store[k][order]
    store["mz"] = _fix_increasing(store["mz"], store["off"])


def split_by_charge(lib: dict) -> dict:
    """{charge: (store, rows)} — rows are the positions in `lib`; the per-charge row order is the
    library's id order, which is what spec_info['charge'][z]['id'] holds in the reference
    (src/ann_solo/reader.py:180-183)."""
    out = {}
    for z in np.unique(lib["prec_z"]):
        rows = np.flatnonzero(lib["prec_z"] == z)
        out[int(z)] = (take_spectra(lib, rows), rows)
    return out


def make_queries(lib: dict, n_queries: int, seed: int = 3, related_fraction: float = 0.6,
                 charges=(2, 3, 4), charge_p=(0.5, 0.4, 0.1)) -> dict:
    """60 % derived from a random library target (m/z jitter N(0, 0.004), 20 % of peaks dropped,
    20 % noise peaks added, half of them modified: precursor shifted by dm/z and a random suffix
    of the fragments shifted by dm/(fragment charge)); 40 % unrelated random spectra."""
    rng = np.random.default_rng(seed)
    targets = np.flatnonzero(lib["is_decoy"] == 0) if "is_decoy" in lib else np.arange(len(lib["prec_mz"]))
    n_rel = int(round(n_queries * related_fraction))
    src = rng.choice(targets, size=n_rel)
    rel = take_spectra(lib, src)
    seg = np.repeat(np.arange(n_rel), np.diff(rel["off"]))
    pos = np.arange(len(seg)) - rel["off"][seg]
    cnt = np.diff(rel["off"])[seg]
    mz = rel["mz"].astype(np.float64) + rng.normal(0.0, 0.004, size=len(seg))
    # modification on half of the related queries
    modded = rng.random(n_rel) < 0.5
    dm = np.where(modded, rng.choice(MODS, size=n_rel), 0.0)
    cut = rng.integers(0, np.diff(rel["off"]) + 1)  # fragments at positions >= cut carry the mod
    frag_z = np.maximum(rel["chg"].astype(np.float64), 1.0)
    mz = np.where(pos >= cut[seg], mz + dm[seg] / frag_z, mz)
    prec_mz = rel["prec_mz"] + dm / rel["prec_z"]
    keep = rng.random(len(seg)) >= 0.2
    keep[rel["off"][:-1]] = True  # never drop a spectrum entirely
    raw_int = rel["inten"].astype(np.float64)
    # noise peaks: 20 % of the original count, random m/z, low-to-mid intensity
    n_noise = np.maximum((np.diff(rel["off"]) * 0.2).astype(np.int64), 0)
    noff, nseg, _ = _segments(n_noise)
    nmz = 100.0 + rng.random(noff[-1]) * (np.minimum(2000.0, (rel["prec_mz"] - PROTON) * rel["prec_z"])[nseg] - 100.0)
    nint = rng.random(noff[-1]) * np.median(raw_int)
    all_seg = np.concatenate([seg[keep], nseg])
    all_mz = np.concatenate([mz[keep], nmz])
    all_int = np.concatenate([raw_int[keep], nint])
    ok = (all_mz >= 50.0) & (all_mz <= 2005.0)
    all_seg, all_mz, all_int = all_seg[ok], all_mz[ok], all_int[ok]
    # cap at MAX_RANK peaks per spectrum (keep the most intense), then sort by m/z
    order = np.lexsort((-all_int, all_seg))
    all_seg, all_mz, all_int = all_seg[order], all_mz[order], all_int[order]
    counts = np.bincount(all_seg, minlength=n_rel)
    o2 = np.zeros(n_rel + 1, np.int64)
    np.cumsum(counts, out=o2[1:])
    p2 = np.arange(len(all_seg)) - o2[all_seg]
    top = p2 < MAX_RANK
    all_seg, all_mz, all_int = all_seg[top], all_mz[top], all_int[top]
    order = np.lexsort((all_mz, all_seg))
    all_seg, all_mz, all_int = all_seg[order], all_mz[order], all_int[order]
    counts = np.bincount(all_seg, minlength=n_rel)
    roff = np.zeros(n_rel + 1, np.int64)
    np.cumsum(counts, out=roff[1:])
    rel_q = dict(mz=_fix_increasing(all_mz.astype(np.float32), roff), inten=renormalise_rank(all_int, roff),
                 chg=np.zeros(len(all_mz), np.uint8), off=roff, prec_mz=prec_mz.astype(np.float64),
                 prec_z=rel["prec_z"].copy(), valid=np.ones(n_rel, np.uint8))
    n_unrel = n_queries - n_rel
    stores = [rel_q]
    truth = [src]
    if n_unrel > 0:
        unrel = make_library(n_unrel, decoy_fraction=0.0, seed=seed + 1000, charges=charges, charge_p=charge_p)
        unrel.pop("is_decoy")
        unrel["chg"][:] = 0
        stores.append(unrel)
        truth.append(np.full(n_unrel, -1, np.int64))
    q = concat_stores(stores)
    perm = rng.permutation(n_queries)
    out = take_spectra(q, perm)
    out["truth"] = np.concatenate(truth)[perm]
    return out
