"""``.splib`` ingestion on the host (no GPU): the native parser behind solo_splib_count / solo_splib_read
against the oracle's pure-Python restatement of reference parsers.pyx:41-186 on synthetic files."""
import numpy as np
import pytest

from oracle import splib_io

ANNOTATIONS = ["b2/0.01", "y3^2/0.02", "y10-18/0.1", "a4/0.0,b4-28/0.1", "?", "p^2/0.3", "y7^3i/0.0", "b12^2-17/0.2",
               "IWA/0.1", "y1/-0.01", "b/0.1", "y15^12/0.0", ""]


def _library(n=60, seed=1):
    rng = np.random.default_rng(seed)
    specs = []
    for i in range(n):
        k = int(rng.integers(0, 40))
        specs.append(dict(id=1000 + 7 * i, peptide="PEPT" + "KR"[i % 2] * (i % 6 + 1), charge=int(rng.integers(1, 6)),
                          prec_mz=float(rng.uniform(300, 1200)), mz=np.sort(rng.uniform(100, 1900, k)),
                          intensity=rng.uniform(1, 1e4, k),
                          annotations=[ANNOTATIONS[int(j)] for j in rng.integers(0, len(ANNOTATIONS), k)],
                          decoy=i % 3 == 0, mods="|1|3,C,Carbamidomethyl" if i % 2 else ""))
    return specs


def test_annotation_charges_follow_the_reference_parser():
    # parsers.pyx:160-186: a/b/y ions only; '<type><index>/' -> 1; '^z' -> z; everything else unannotated
    want = dict(zip(ANNOTATIONS, [1, 2, 0, 1, 0, 0, 3, 2, 0, 1, 0, 12, 0]))
    for a, c in want.items():
        assert splib_io.parse_annotation(a.encode()) == c, a


def test_native_parser_equals_oracle_reader(tmp_path):
    from ann_solo_b200 import parsers
    specs = _library()
    p = str(tmp_path / "lib.splib")
    offsets = splib_io.write_splib(p, specs)
    want, got = splib_io.read_splib(p), parsers.read_splib(p)
    for k in ("id", "prec_z", "prec_mz", "is_decoy", "file_offset", "off", "mz", "inten", "chg"):
        assert got[k].dtype == want[k].dtype and np.array_equal(got[k], want[k]), k
    assert got["peptide"] == want["peptide"] == [s["peptide"] for s in specs]
    assert list(got["file_offset"]) == offsets
    assert np.array_equal(got["mz"], np.concatenate([s["mz"] for s in specs]).astype(np.float32))
    assert got["is_decoy"].tolist() == [int(s["decoy"]) for s in specs]


def test_splib_parser_object_surface(tmp_path):
    """SplibParser.read_spectrum keeps the reference's (spectrum, offset) / StopIteration contract."""
    from ann_solo_b200.parsers import SplibParser
    specs = _library(12, seed=3)
    p = str(tmp_path / "lib.splib")
    offsets = splib_io.write_splib(p, specs)
    parser = SplibParser(p.encode())
    parser.seek_first_spectrum()
    seen = []
    while True:
        try:
            spectrum, off = parser.read_spectrum()
        except StopIteration:
            break
        seen.append(off)
        s = specs[len(seen) - 1]
        assert spectrum.identifier == str(s["id"]) and spectrum.peptide == s["peptide"]
        assert spectrum.precursor_charge == s["charge"] and spectrum.precursor_mz == s["prec_mz"]
        assert spectrum.is_decoy == s["decoy"] and spectrum.mz.dtype == np.float32
        chg = [0 if a is None else a.charge for a in spectrum.annotation]
        assert chg == [splib_io.parse_annotation(a.encode()) for a in s["annotations"]]
    assert seen == offsets
    spectrum, off = parser.read_spectrum(offsets[5])
    assert off == offsets[5] and spectrum.identifier == str(specs[5]["id"])


def test_errors(tmp_path):
    from ann_solo_b200 import parsers
    with pytest.raises(FileNotFoundError):
        parsers.read_splib(str(tmp_path / "missing.splib"))
    p = str(tmp_path / "lib.splib")
    offsets = splib_io.write_splib(p, _library(5, seed=4))
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:offsets[-1] + 9])   # inside the last spectrum's name line: no precursor m/z follows
    with pytest.raises(ValueError, match="truncated|claims|malformed"):
        parsers.read_splib(p)


def test_spectral_library_reader_surface_needs_no_gpu(tmp_path):
    """reader.SpectralLibraryReader over a .splib: spec_info (reference reader.py:180-189), raw
    read_spectrum / read_all_spectra, errors — the parts that never touch the device."""
    from ann_solo_b200.reader import SpectralLibraryReader
    specs = _library(40, seed=6)
    p = str(tmp_path / "lib.splib")
    splib_io.write_splib(p, specs)
    r = SpectralLibraryReader(p, "0123456789abcdef")
    by_charge = {}
    for s in specs:
        by_charge.setdefault(s["charge"], []).append(s)
    assert sorted(r.spec_info["charge"]) == sorted(by_charge)
    for z, group in by_charge.items():
        info = r.spec_info["charge"][z]
        assert info["id"].tolist() == [str(s["id"]) for s in group]                    # file order
        assert info["precursor_mz"].dtype == np.float32
        assert np.array_equal(info["precursor_mz"], np.array([s["prec_mz"] for s in group], np.float32))
    s0 = r.read_spectrum(str(specs[3]["id"]))
    assert s0.identifier == str(specs[3]["id"]) and s0.peptide == specs[3]["peptide"] and not s0.is_processed
    assert np.array_equal(s0.mz, specs[3]["mz"].astype(np.float32)) and s0.is_decoy == specs[3]["decoy"]
    assert [x.identifier for x in r.read_all_spectra()] == [str(s["id"]) for s in specs]
    with r as same:
        assert same is r and r.get_version() == "null"
    with pytest.raises(FileNotFoundError):
        SpectralLibraryReader(str(tmp_path / "nope.splib"))
    (tmp_path / "lib.msp").write_text("Name: X/2\n")
    with pytest.raises(FileNotFoundError, match="Unrecognized file format"):
        SpectralLibraryReader(str(tmp_path / "lib.msp"))


SPTXT = """### a SpectraST text library
Name: AAAC[160]DEFGK/2
LibID: 0
MW: 1012.45
PrecursorMZ: 506.2321
Status: Normal
FullName: K.AAAC[160]DEFGK.L/2 (HCD)
Comment: AvePrecursorMz=506.5 Mods=1/3,C,Carbamidomethyl Parent=506.232 Remark=_NONE_
NumPeaks: 5
201.1234\t1200.5\tb2/0.01\t
175.1190\t800.0\ty1/-0.00\t
330.1660\t95.5\ty3-18/0.02,b4^2/0.1\t
402.2000\t10.0\t?\t
250.6000\t55.0\ty5^2/0.00\t

Name: PEPTIDER/3
LibID: 1
Comment: Parent=319.8231 Mods=0 Spec=Consensus Remark=DECOY_shuffle
Num Peaks: 3
100.0\t1.0\tp^3/0.0
200.0\t2.0\tIWA/0.0
300.0\t3.0\ta2^2i/0.1
"""


def test_sptxt_library_known_answers(tmp_path):
    """reference reader.py:324-418 (+ :565-597 annotations, :300-322 ProForma), hand-worked answers."""
    from ann_solo_b200.parsers import read_sptxt, sptxt_annotation_charge, sptxt_seq_to_proforma
    from ann_solo_b200.reader import SpectralLibraryReader
    p = tmp_path / "lib.sptxt"
    p.write_text(SPTXT)
    st = read_sptxt(str(p))
    assert st["id"] == ["1", "2"] and st["prec_z"].tolist() == [2, 3]
    assert st["prec_mz"].tolist() == [506.2321, 319.8231]                 # PrecursorMZ:, else Parent=
    assert st["is_decoy"].tolist() == [0, 1] and st["off"].tolist() == [0, 5, 8]
    # peaks come back m/z-ascending (MsmsSpectrum orders them), annotations follow their peaks
    assert np.array_equal(st["mz"][:5], np.array([175.1190, 201.1234, 250.6, 330.166, 402.2], np.float32))
    assert np.array_equal(st["inten"][:5], np.array([800.0, 1200.5, 55.0, 95.5, 10.0], np.float32))
    assert st["chg"].tolist() == [1, 1, 2, 1, 0, 3, 0, 2]                 # y3-18 -> abs(-1) = 1; '?' and IWA -> none
    assert st["peptide"] == ["AAAC[160][Carbamidomethyl]DEFGK", "PEPTIDER"]
    assert [sptxt_annotation_charge(a) for a in ("b2/0.01", "y10-18/0.1", "y5^2/0.0", "p^3/0.0", "a2^2i/0.1", "?", "IWA/0", "")] == \
        [1, 1, 2, 3, 2, 0, 0, 0]
    assert sptxt_seq_to_proforma("ACDK", ["0,A,Acetyl", "1,C,Carbamidomethyl"]) == "A[Acetyl]C[Carbamidomethyl]DK"
    r = SpectralLibraryReader(str(p))
    assert sorted(r.spec_info["charge"]) == [2, 3] and r.spec_info["charge"][3]["id"].tolist() == ["2"]
    s = r.read_spectrum("1")
    assert s.precursor_mz == 506.2321 and [None if a is None else a.charge for a in s.annotation] == [1, 1, 2, 1, None]
    (tmp_path / "bad.sptxt").write_text("Name: AAK/2\nPrecursorMZ: 300.1\n100.0\t1.0\n")
    with pytest.raises(ValueError, match="NumPeaks"):
        read_sptxt(str(tmp_path / "bad.sptxt"))
