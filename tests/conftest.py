import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def tiny_golden():
    return dict(np.load(os.path.join(GOLDEN, "tiny_n48_m160_q3.npz")))


@pytest.fixture(scope="session")
def c1_golden():
    return dict(np.load(os.path.join(GOLDEN, "c1_n281_m3093_q1.npz")))


@pytest.fixture(scope="session")
def ingest_golden():
    return dict(np.load(os.path.join(GOLDEN, "ingest_n37_m90.npz")))


def c1_inputs(g):
    """Re-generate the inputs of the C1 golden case exactly as oracle/gen_golden.py did."""
    from gapit_b200 import synth
    n, m, seed, npc = int(g["n"]), int(g["m"]), int(g["seed"]), int(g["npc"])
    G = synth.make_genotypes(n, m, seed=seed, threads=1)
    G[3, :] = 0
    G[7, :] = 2
    G[11, :] = 1
    G[20, :] = G[19, :]
    Y = synth.make_phenotypes(G, seed=seed)[:, 0]
    X0 = synth.make_covariates(G, npc=npc)
    return G, Y, X0


@pytest.fixture(scope="session")
def gpu_ctx():
    import gapit_b200 as gb
    ctx = gb.Context(0)
    yield ctx
    ctx.close()
