"""CPU oracle for the pyg2p interpolation hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, with numpy + scipy.spatial.cKDTree, the arithmetic of the
reference functions listed in SURVEY.md §8(a).  It is imported only by `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs,
always as the checker or the timed CPU baseline - never by `pyg2p_b200` itself
(the product path has no CPU fallback and raises if the CUDA library is missing).

Third-party arithmetic that is not under /root/reference: `scipy.spatial.cKDTree`
(reference pins scipy>=1.6.0 / 1.10.1; here scipy 1.18.1) - used directly as the
k-NN oracle with the reference's own parameters (`leafsize=30`, `workers`), and
`numexpr.evaluate` (libm sin/cos) - restated with numpy, whose float64 sin/cos are
bit-identical to libm in this image (checked on 2M values).

Parity pin: the oracle is checked (tests/test_oracle_golden.py) against golden
vectors produced by RUNNING THE REFERENCE'S OWN CODE (`ScipyInterpolation`,
`Interpolator._interpolate_scipy_append_mv`, `Converter`, `Aggregator`, `Corrector`)
imported from /root/reference with shims for its missing third-party imports
(tests/golden/make_golden.py), and against the reference's golden intertable
`tests/data/tbl_pf10slhf_550800_scipy_nearest.npy.gz`.
"""
