"""mzML query files on the host (no GPU): the native scanner behind solo_mzml_count / solo_mzml_read
against the oracle's stdlib-XML restatement of what reference reader.py:659-741 takes from a file."""
import math

import numpy as np
import pytest

from oracle import mzml_io


def _spectra(n=30, seed=5):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        k = int(rng.integers(0, 80))
        mz = np.sort(rng.uniform(50, 2000, k))
        if i % 7 == 3:
            mz = mz[::-1].copy()        # unsorted on disk
        s = dict(id=f"controllerType=0 controllerNumber=1 scan={100 + i}", ms_level=2, mz=mz,
                 intensity=rng.gamma(0.7, 1000.0, k), prec_mz=float(rng.uniform(300, 1500)),
                 bits=64 if i % 2 else 32, zlib=i % 3 == 0)
        if i % 4 == 0:
            s["ms_level"] = 1           # survey scans are skipped but counted in `index`
            del s["prec_mz"]
        if i % 5 != 1:
            s["charge"] = int(rng.integers(1, 5))
        elif i % 2:
            s["possible_charge"] = 3
        if i % 6:
            s["rt"] = float(rng.uniform(0, 120))
        if i == 7:
            s["id"] = "index=7"
        if i == 9:
            s["id"] = "spectrum_without_a_number"     # reference: ValueError -> skipped with a warning
        if i == 11:
            s["id"] = "scan=11 merged"                 # int('11 merged') fails too
        if i == 13:
            s["extra_precursor"] = True
        if i == 15:
            del s["prec_mz"]                           # MS2 without a selected ion
            s.pop("charge", None)
            s.pop("possible_charge", None)
        out.append(s)
    return out


def test_native_scanner_equals_oracle_reader(tmp_path):
    from ann_solo_b200.reader import read_mzml_store
    p = str(tmp_path / "run.mzml")
    spectra = _spectra()
    mzml_io.write_mzml(p, spectra)
    want = mzml_io.read_mzml(p)
    got = read_mzml_store(p)
    assert len(want) == len(got["prec_mz"]) and 15 < len(want) < 25
    assert got["n_skipped"] == 3                      # ids 9 and 11, no selected ion at 15
    for i, w in enumerate(want):
        b, e = got["off"][i], got["off"][i + 1]
        assert got["identifier"][i] == w["identifier"] and got["index"][i] == w["index"]
        assert got["prec_mz"][i] == w["prec_mz"] and got["prec_z"][i] == w["prec_z"]
        assert (math.isnan(got["rt"][i]) and math.isnan(w["rt"])) or got["rt"][i] == w["rt"]
        assert np.array_equal(got["mz64"][b:e], w["mz"]) and np.array_equal(got["inten"][b:e], w["inten"])
        assert (np.diff(got["mz64"][b:e]) >= 0).all()
    # against what was written
    kept = [i for i, s in enumerate(spectra) if s["ms_level"] == 2 and i not in (9, 11, 15)]
    assert got["index"].tolist() == kept
    assert got["identifier"][kept.index(7)] == "7" and got["identifier"][kept.index(13)] == "113"
    assert got["prec_z"][kept.index(13)] == spectra[13].get("charge", 0)      # the second precursor is ignored
    for j, i in enumerate(kept):
        s = spectra[i]
        b, e = got["off"][j], got["off"][j + 1]
        mz = np.sort(np.asarray(s["mz"], np.float32 if s["bits"] == 32 else np.float64).astype(np.float64))
        assert np.array_equal(got["mz64"][b:e], mz)
        assert got["prec_mz"][j] == s["prec_mz"]


def test_read_mzml_objects_like_the_reference(tmp_path):
    from ann_solo_b200.reader import read_mzml, read_query_file
    p = str(tmp_path / "small.mzml")
    rng = np.random.default_rng(42)
    spectra = [dict(id=f"scan={nr}", ms_level=2, mz=np.arange(1, 9) * 100.0, intensity=np.ones(8),
                    prec_mz=450.5 + nr, charge=int(rng.choice([2, 3])), rt=1.5) for nr in (17, 111)]
    mzml_io.write_mzml(p, spectra)
    got = list(read_mzml(p))
    assert [s.identifier for s in got] == ["17", "111"] and [s.index for s in got] == [0, 1]
    for s in got:
        assert s.precursor_charge in (2, 3) and s.retention_time == 1.5 and not s.is_processed
        assert s.mz.dtype == np.float64 and s.intensity.dtype == np.float32 and len(s.mz) == 8
    assert len(list(read_query_file(p))) == 2
    (tmp_path / "x.mzxml").write_text("<mzXML/>")
    assert list(read_query_file(str(tmp_path / "x.mzxml"))) == []


def test_errors(tmp_path):
    from ann_solo_b200.reader import read_mzml_store
    p = str(tmp_path / "bad.mzml")
    good = dict(id="scan=1", ms_level=2, mz=[100.0, 200.0], intensity=[1.0, 2.0], prec_mz=500.0, charge=2)
    mzml_io.write_mzml(p, [good])
    text = open(p).read()
    open(p, "w").write(text.replace("</spectrum>", ""))
    with pytest.raises(ValueError, match="not closed"):
        read_mzml_store(p)
    open(p, "w").write(text.replace('defaultArrayLength="2"', 'defaultArrayLength="3"'))
    with pytest.raises(ValueError, match="defaultArrayLength"):
        read_mzml_store(p)
    open(p, "w").write(text.replace("MS:1000576", "MS:1002312"))
    with pytest.raises(ValueError, match="numpress"):
        read_mzml_store(p)
    b64 = text[text.find("<binary>") + 8:text.find("</binary>")]
    open(p, "w").write(text.replace(b64, "!!" + b64[2:], 1))
    with pytest.raises(ValueError, match="base64"):
        read_mzml_store(p)
    open(p, "w").write("")
    assert len(read_mzml_store(p)["prec_mz"]) == 0


def test_mzxml_scanner_equals_oracle_reader(tmp_path):
    """reference reader.py:743-811; MS2 scans nested inside their survey scan, as mzXML writes them."""
    from ann_solo_b200.reader import read_mzxml, read_mzxml_store, read_query_file
    rng = np.random.default_rng(8)
    scans, num = [], 0

    def ms2():
        nonlocal num
        num += 1
        k = int(rng.integers(0, 70))
        s = dict(num=num, ms_level=2, mz=np.sort(rng.uniform(50, 2000, k)), intensity=rng.gamma(0.7, 1000.0, k),
                 prec_mz=float(rng.uniform(300, 1500)), precision=64 if num % 2 else 32, zlib=num % 3 == 0)
        if num % 4:
            s["charge"] = int(rng.integers(1, 5))
        if num % 5:
            s["rt"] = float(rng.uniform(0, 7200))
        if num == 6:
            del s["prec_mz"]                          # MS2 without precursorMz: skipped
        if num == 8:
            s["mz"] = s["mz"][::-1].copy()
        return s

    for _ in range(5):
        num += 1
        survey = dict(num=num, ms_level=1, mz=np.sort(rng.uniform(300, 1500, 20)), intensity=rng.random(20), rt=1.0)
        survey["children"] = [ms2() for _ in range(3)]
        scans.append(survey)
    scans.append(ms2())                               # a top-level MS2 scan
    p = str(tmp_path / "run.mzxml")
    mzml_io.write_mzxml(p, scans)
    want = mzml_io.read_mzxml(p)
    got = read_mzxml_store(p)
    assert len(want) == len(got["prec_mz"]) == 15 and got["n_skipped"] == 1
    for i, w in enumerate(want):
        b, e = got["off"][i], got["off"][i + 1]
        assert got["identifier"][i] == w["identifier"] and got["index"][i] == w["index"]
        assert got["prec_mz"][i] == w["prec_mz"] and got["prec_z"][i] == w["prec_z"]
        assert (math.isnan(got["rt"][i]) and math.isnan(w["rt"])) or got["rt"][i] == pytest.approx(w["rt"], rel=1e-15)
        assert np.array_equal(got["mz64"][b:e], w["mz"]) and np.array_equal(got["inten"][b:e], w["inten"])
    objs = list(read_mzxml(p))
    assert [o.identifier for o in objs] == [w["identifier"] for w in want]
    assert objs[0].index == 1 and objs[0].mz.dtype == np.float64       # the survey scan is index 0
    assert len(list(read_query_file(p))) == 15
