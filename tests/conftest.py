import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# seeds of SURVEY.md §8(d) config 2
KEYGEN_SEED = 0xB20000A1
DATA_SEED = 0xB20000D1


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import mk_oracle
    mk_oracle.build()
    return mk_oracle


@pytest.fixture(scope="session")
def toy_keys(oracle):
    """Small ring (N = 64) so the exact schoolbook back-end runs whole bootstraps in milliseconds."""
    prm = dict(n=24, N=64, k=2, l=2, bgbit=7, t=3, basebit=3, sigma_lwe=2.0 ** -18, sigma_gsw=2.0 ** -40, sigma_ks=2.0 ** -18)
    return oracle.KeySet(prm, seed=11, nthreads=4)


@pytest.fixture(scope="session")
def keys2(oracle):
    """2-party default parameters (mk_api.jl:32-38), oracle keygen with the fixed seed."""
    return oracle.KeySet(oracle.PARAMS_2PARTY, seed=KEYGEN_SEED, nthreads=os.cpu_count() or 8)


def make_engine(ks):
    """GPU engine loaded with an oracle key set's raw key arrays (the same bytes on both sides)."""
    import torus_fhe_b200 as T
    p = ks.prm
    sp = T.SchemeParameters_3gen(p.n, p.sigma_lwe, p.N, 1, False, p.l, p.bgbit, p.sigma_gsw, p.t, p.basebit, p.sigma_ks, p.k)
    eng = T.Engine(sp, device=0)
    eng.load_keys([ks.bsk[i] for i in range(p.k)], [ks.ksk[i] for i in range(p.k)])
    return eng


@pytest.fixture(scope="session")
def engine2(keys2):
    eng = make_engine(keys2)
    yield eng
    eng.close()


@pytest.fixture
def rng():
    return np.random.default_rng(0xB200)
