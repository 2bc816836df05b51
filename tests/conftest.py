import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA GPU (runs the sm_100a kernels)")


BACKENDS = [pytest.param("oracle", id="oracle"),
            pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


def _make_backend(kind, flags=0, max_agents=1024, **kw):
    if kind == "oracle":
        from oracle.oracle_py import OracleBackend

        return OracleBackend(flags=flags)
    from landmass_b200._native import NativeBackend

    return NativeBackend(max_agents=max_agents, flags=flags, **kw)


@pytest.fixture(params=BACKENDS)
def backend_kind(request):
    return request.param


@pytest.fixture
def make_backend(backend_kind):
    """Factory: make_backend(flags=0, max_agents=1024) -> backend of the parametrised kind."""
    made = []

    def factory(flags=0, max_agents=1024, **kw):
        b = _make_backend(backend_kind, flags=flags, max_agents=max_agents, **kw)
        made.append(b)
        return b

    yield factory
    for b in made:
        b.close()


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle_py

    oracle_py.build()
    return oracle_py
