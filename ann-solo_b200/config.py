"""Hot-path configuration keys, same names and defaults as the reference's configargparse
singleton (src/ann_solo/config.py:71-216). Only the keys the hot path reads are kept; the
reference's CLI/ini parsing is out of scope and can feed this object
through ``config.update(vars(namespace))``.
"""
from __future__ import annotations


class Config:
    _defaults = dict(
        # preprocessing (config.py:71-117)
        resolution=None, min_mz=11, max_mz=2010, remove_precursor=False, remove_precursor_tolerance=0,
        min_intensity=0.01, min_peaks=10, min_mz_range=250, max_peaks_used=50, max_peaks_used_library=50,
        scaling="rank",
        # search (config.py:125-151)
        precursor_tolerance_mass=20.0, precursor_tolerance_mode="ppm",
        precursor_tolerance_mass_open=300.0, precursor_tolerance_mode_open="Da",
        fragment_mz_tolerance=0.02, allow_peak_shifts=True, fdr=0.01,
        # ANN (config.py:172-216)
        mode="ann", bin_size=0.04, hash_len=800, num_candidates=1024, batch_size=16384, num_list=256,
        num_probe=128, no_gpu=False,
    )

    def __init__(self, **kw):
        self._ns = dict(self._defaults)
        self.update(kw)

    def update(self, kw):
        for k, v in kw.items():
            self._ns[k] = v

    def __getattr__(self, name):
        try:
            return self.__dict__["_ns"][name]
        except KeyError:
            raise AttributeError(name)

    def __getitem__(self, name):
        return self._ns[name]


config = Config()
