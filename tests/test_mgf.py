"""MGF query files on the host (no GPU): the native parser behind solo_mgf_count / solo_mgf_read
against the oracle's pure-Python restatement, on MassIVE-KB style entries like the reference's own
test writes (src/tests/query_reader_test.py:41-66) and on the other fields read_mgf uses."""
import math

import numpy as np
import pytest

from oracle import mgf_io


def _entries(n=25, seed=2):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        k = int(rng.integers(0, 60))
        mz = rng.uniform(50, 2000, k)
        if i % 3:
            mz = np.sort(mz)            # every third spectrum is written unsorted
        s = dict(prec_mz=float(rng.uniform(300, 1500)), mz=mz, intensity=rng.gamma(0.7, 1000.0, k))
        if i % 2 == 0:
            s["seq"] = "PEPTIDEK+15.995"[: 5 + i % 10]
        if i % 4 != 3:
            s["charge"] = ["2+", "3+", "2+ and 3+", "1-", "4"][i % 5]
        if i % 5 == 0:
            s["title"] = f'run1.scan{i} File:"a=b.raw"'      # '=' inside the value
        elif i % 5 == 1:
            s["scan"] = str(1000 + i)
        if i % 6 == 0:
            s["rt"] = float(rng.uniform(0, 7200))
        if i % 7 == 0:
            s["decoy"] = True
        if i % 8 == 0:
            s["extra"] = ["# comment inside", "INSTRUMENT=ESI-QUAD-TOF", ""]
            s["pepmass_intensity"] = True
        if i % 9 == 0:
            s["peak_charge"] = True
        out.append(s)
    return out


def test_native_parser_equals_oracle_reader(tmp_path):
    from ann_solo_b200.reader import read_mgf_store
    p = str(tmp_path / "q.mgf")
    entries = _entries()
    mgf_io.write_mgf(p, entries)
    want = mgf_io.read_mgf(p)
    got = read_mgf_store(p)
    assert len(want) == len(entries) == len(got["prec_mz"])
    for i, w in enumerate(want):
        b, e = got["off"][i], got["off"][i + 1]
        assert got["identifier"][i] == w["identifier"] and got["seq"][i] == w["seq"]
        assert got["prec_mz"][i] == w["prec_mz"] and got["prec_z"][i] == w["prec_z"]
        assert bool(got["is_decoy"][i]) == w["is_decoy"]
        assert (math.isnan(got["rt"][i]) and math.isnan(w["rt"])) or got["rt"][i] == w["rt"]
        assert np.array_equal(got["mz64"][b:e], w["mz"]) and np.array_equal(got["inten"][b:e], w["inten"])
        assert (np.diff(got["mz64"][b:e]) >= 0).all()
    # what was written is what is read (float repr round-trips exactly)
    assert got["prec_mz"].tolist() == [s["prec_mz"] for s in entries]
    assert got["prec_z"].tolist() == [0 if "charge" not in s else {"2+": 2, "3+": 3, "2+ and 3+": 2, "1-": -1, "4": 4}[
        s["charge"]] for s in entries]
    assert got["identifier"][2] == "3" and got["identifier"][1] == "1001" and got["identifier"][0].startswith("run1.scan0")


def test_read_mgf_objects_like_the_reference(tmp_path):
    """reference reader.py:884-911 + its test (three MassIVE-KB style spectra are read)."""
    from ann_solo_b200.reader import read_mgf, read_query_file
    p = str(tmp_path / "small.mgf")
    rng = np.random.default_rng(42)
    entries = [dict(seq=pep, prec_mz=400.0 + 10 * i, charge=f"{rng.choice([2, 3])}+", mz=np.arange(1, 8) * 100.0 + i,
                    intensity=np.ones(7)) for i, pep in enumerate(["LESLIEK", "PEPTIDEK", "HPYLEDR"])]
    mgf_io.write_mgf(p, entries, header=())
    spectra = list(read_mgf(p))
    assert len(spectra) == 3
    for i, s in enumerate(spectra, 1):
        assert s.index == i and s.identifier == str(i) and not s.is_processed and not s.is_decoy
        assert s.precursor_charge in (2, 3) and s.retention_time is None and s.peptide == entries[i - 1]["seq"]
        assert s.mz.dtype == np.float64 and s.intensity.dtype == np.float32 and len(s.mz) == 7
    assert len(list(read_query_file(p))) == 3
    with pytest.raises(FileNotFoundError):
        read_query_file(str(tmp_path / "missing.mgf"))
    with pytest.raises(FileNotFoundError):
        read_query_file(str(tmp_path / "small.txt"))
    (tmp_path / "x.mzxml").write_text("<mzXML/>")
    assert list(read_query_file(str(tmp_path / "x.mzxml"))) == []       # mzML / mzXML are read natively too


def test_errors(tmp_path):
    from ann_solo_b200.reader import read_mgf_store
    p = tmp_path / "bad.mgf"
    p.write_text("BEGIN IONS\nPEPMASS=500.1\n100.0 1.0\n")
    with pytest.raises(ValueError, match="without END IONS"):
        read_mgf_store(str(p))
    p.write_text("BEGIN IONS\nTITLE=a\n100.0 1.0\nEND IONS\n")
    with pytest.raises(ValueError, match="no PEPMASS"):
        read_mgf_store(str(p))
    p.write_text("BEGIN IONS\nPEPMASS=500.1\n100.0 abc\nEND IONS\n")
    with pytest.raises(ValueError, match="expected a peak"):
        read_mgf_store(str(p))
    p.write_text("")
    assert len(read_mgf_store(str(p))["prec_mz"]) == 0


def test_mgf_as_a_spectral_library(tmp_path):
    """reference reader.py:34 (_supported_extensions) / :283-284: an MGF file as the library; the reader
    surface (spec_info per charge in file order, raw spectra, ProForma peptides) needs no GPU."""
    from ann_solo_b200.reader import SpectralLibraryReader, _mgf_seq_to_proforma
    # reference reader.py:837-866
    assert _mgf_seq_to_proforma("PEPTIDEK") == "PEPTIDEK"
    assert _mgf_seq_to_proforma("AC+57.021DM+15.995K") == "AC[+57.021]DM[+15.995]K"
    assert _mgf_seq_to_proforma("+42.011ACDK") == "[+42.011]-ACDK"
    rng = np.random.default_rng(3)
    entries = []
    for i in range(12):
        k = int(rng.integers(12, 40))
        entries.append(dict(title=f"lib{i}", seq=["PEPTIDEK", "AC+57.021DK", "+42.011LESLIEK"][i % 3],
                            prec_mz=float(rng.uniform(300, 900)), charge=f"{2 + i % 2}+",
                            mz=np.sort(rng.uniform(100, 1500, k)), intensity=rng.gamma(0.7, 1000.0, k), decoy=i % 4 == 0))
    p = str(tmp_path / "lib.mgf")
    mgf_io.write_mgf(p, entries)
    r = SpectralLibraryReader(p, "0123456789")
    assert sorted(r.spec_info["charge"]) == [2, 3]
    assert r.spec_info["charge"][2]["id"].tolist() == [f"lib{i}" for i in range(0, 12, 2)]
    assert r.spec_info["charge"][3]["precursor_mz"].dtype == np.float32
    assert np.array_equal(r.spec_info["charge"][3]["precursor_mz"],
                          np.array([e["prec_mz"] for e in entries[1::2]], np.float32))
    s = r.read_spectrum("lib4")
    assert s.peptide == "AC[+57.021]DK" and s.is_decoy and s.precursor_charge == 2 and not s.is_processed
    assert np.array_equal(s.mz, entries[4]["mz"].astype(np.float32)) and all(a is None for a in s.annotation)
    assert [x.identifier for x in r.read_all_spectra()] == [f"lib{i}" for i in range(12)]
    del entries[5]["charge"]
    mgf_io.write_mgf(p, entries)
    with pytest.raises(ValueError, match="CHARGE"):
        SpectralLibraryReader(p)
