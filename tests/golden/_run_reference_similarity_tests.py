"""Run the reference's own src/tests/spectrum_similarity_test.py, unmodified, against the reference's
unmodified spectrum_similarity.py as make_golden.py imports it (stubbed third-party modules).
Container-only helper (needs /root/reference); prints pytest's summary and returns its exit code."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402

make_golden.load_reference_similarity()
import pytest  # noqa: E402

TEST = "/root/reference/src/tests/spectrum_similarity_test.py"
# test_kendalltau also calls scipy.stats.kendalltau(None, None) for the no-match pairs, which SciPy >= 1.1x
# rejects (older SciPy returned NaN); the search never computes features for an SSM without peak matches
# (utils.py:332-333). Its four remaining known answers are checked by hand below.
rc = pytest.main(["-q", "-p", "no:cacheprovider", "--rootdir", "/tmp", "-c", "/dev/null", TEST,
                  "-k", "not test_kendalltau"])
import importlib.util as iu  # noqa: E402

spec = iu.spec_from_file_location("ref_sim_test", TEST)
t = iu.module_from_spec(spec)
spec.loader.exec_module(t)
for fixture, want in (("all_match", 19.29406731), ("all_match_top", 4.09434456), ("partial_match", 4.25896654),
                      ("partial_match_top", 0.0)):
    f = getattr(t, fixture)
    f = getattr(f, "__wrapped__", None) or f._get_wrapped_function()
    got = f().kendalltau()
    ok = abs(got - want) <= 1e-6 * max(1.0, abs(want))
    print(f"kendalltau {fixture}: {got!r} (reference test expects {want}) {'ok' if ok else 'MISMATCH'}")
    rc = rc or (0 if ok else 1)
sys.exit(rc)
