"""The CPU oracle (oracle/fem_oracle.c) is pinned bit-for-bit to golden vectors produced by the
unmodified reference (oracle/gen_golden.py)."""
import numpy as np
import pytest

from util import bits_equal, golden_names, load_golden, make_oracle, oracle_module


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_histories_bitwise(name):
    g = load_golden(name)
    o = make_oracle(g)
    done = 0
    for n in [int(s) for s in g["steps"]]:
        o.run(n - done)
        done = n
        for q in range(g["P"]):
            assert bits_equal(o.d0(q), g[f"hist_{n}_r{q}"]), (name, n, q)
    o.close()


@pytest.mark.parametrize("name", [n for n in golden_names() if len(load_golden(n)["nosync_steps"])])
def test_oracle_matches_reference_unsynchronised_branch(name):
    """MODEL=True branch of Dynamic_solver.py:22 — no syn_cpus."""
    g = load_golden(name)
    o = make_oracle(g)
    done = 0
    for n in [int(s) for s in g["nosync_steps"]]:
        o.run(n - done, model=True)
        done = n
        for q in range(g["P"]):
            assert bits_equal(o.d0(q), g[f"nosync_{n}_r{q}"]), (name, n, q)
    o.close()


def test_oracle_known_answers():
    """Known answers from the reference repo / survey probes: golden dt (Results/plotter.py:25) and the
    serial displacement norm after 2000 steps."""
    g = load_golden("beam_coarse_P1")
    assert float(g["dt"]) == 0.00024784067462642383
    o = make_oracle(g)
    o.run(2000)
    assert np.linalg.norm(o.d0(0)) == 0.10200279842135908
    o.close()


def test_oracle_csr_matvec_is_scipy_order():
    """scipy's csr_matvec (the arithmetic behind Dynamic_solver.py:12) == sequential mul/add."""
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    A = sp.random(2000, 2000, density=0.02, format="csr", random_state=5, dtype=np.float64)
    A.data[:] = rng.standard_normal(A.nnz) * 10.0 ** rng.integers(-6, 6, A.nnz)
    x = rng.standard_normal((2000, 1))
    y = oracle_module().csr_matvec(A.indptr, A.indices, A.data, x)
    assert bits_equal(y, A.dot(x))


@pytest.mark.parametrize("name,model", [("beam_coarse_P1", False), ("beam_coarse_P2", False), ("beam_coarse_P3", False), ("beam_coarse_P2", True)])
def test_numpy_step_baseline_matches_reference_histories_bitwise(name, model):
    """oracle/numpy_step.py — the reference's numpy/scipy statement sequence on P single-threaded processes (the
    numpy-scipy CPU baseline of bench.py) — reproduces the golden histories of the unmodified reference."""
    import os
    import sys
    from util import ROOT
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy_step
    g = load_golden(name)
    steps = [int(s) for s in (g["nosync_steps"] if model else g["steps"])]
    n = [s for s in steps if s <= 100][-1]
    secs, states = numpy_step.run_ranks(g["ranks"], len(g["points"]), g["dt"], float(g["alpha"]), n, model=model, want_state=True)
    for q in range(g["P"]):
        assert bits_equal(states[q], g[f"{'nosync' if model else 'hist'}_{n}_r{q}"]), (name, q)
    assert secs > 0
