"""The streaming loop (SoloEngine.search_stream: copies of neighbouring batches on a copy stream, two query slots)
returns exactly what the synchronous one-batch call returns, batch after batch — ragged batch sizes, alternating
charges, pinned and pageable host buffers, an empty batch in between (reference loop: spectral_library.py:301-306)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_search_stream_equals_batch_by_batch(engine, synth):
    import torch
    from ann_solo_b200.engine import SoloEngine
    lib = synth.make_library(9000, seed=31, decoy_seed=32)
    per_charge = synth.split_by_charge(lib)
    queries = synth.make_queries(lib, 1500, seed=33)
    for z in (2, 3):
        store, _ = per_charge[z]
        engine.load_library(z, store)
        engine.ivf_train_library(z, 64, iters=2, seed=4)
        engine.ivf_add_library(z)
    params = SoloEngine.make_params(True, 128, 16, 500.0, "Da", 0.02, True, max_pairs=50)
    by_z = {z: np.flatnonzero(queries["prec_z"] == z) for z in (2, 3)}
    sizes = [257, 1, 0, 300, 64, 129]
    batches, keep = [], []
    pos = {2: 0, 3: 0}
    for i, n in enumerate(sizes):
        z = 2 + i % 2
        rows = by_z[z][pos[z]:pos[z] + n]
        pos[z] += n
        q = synth.take_spectra(queries, rows)
        if i % 3 == 0:   # page-locked buffers: the copies really run asynchronously
            for key in ("mz", "inten", "off", "prec_mz"):
                t = torch.from_numpy(np.ascontiguousarray(q[key])).pin_memory()
                keep.append(t)
                q[key] = t.numpy()
        batches.append((z, q))
    want = [{k: v.copy() for k, v in engine.search_batch(z, params, q).items()} for z, q in batches]
    for rep in range(2):
        got = list(engine.search_stream(params, [(z, q, None) for z, q in batches]))
        assert [g[0] for g in got] == [b[0] for b in batches]
        for (z, res), w, (_, q) in zip(got, want, batches):
            n = len(q["prec_mz"])
            for key in ("best_row", "n_pairs", "n_cand"):
                assert np.array_equal(res[key][:n], w[key][:n]), key
            assert np.array_equal(res["score"][:n].view(np.uint64), w["score"][:n].view(np.uint64))
            for i in range(n):
                m = int(w["n_pairs"][i])
                assert np.array_equal(res["pairs"][i, :m], w["pairs"][i, :m])
    # the synchronous API still works on the default slot afterwards
    again = engine.search_batch(batches[0][0], params, batches[0][1])
    assert np.array_equal(again["best_row"], want[0]["best_row"])
