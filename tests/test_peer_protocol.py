"""The peer-memory panel broadcast (gogp_b200/csrc/peer_bcast.hpp, unmodified) with the ranks as host threads and
the CUDA runtime calls it makes replaced by host stand-ins (tests/peer/peer_host.cc): what is left is the host-level
protocol -- posted / acked counters, the barrier, event-slot reuse, the root overwriting its buffer as soon as the
broadcast returns -- on 2, 4, 6 and 8 ranks, and once more under ThreadSanitizer.  No GPU needed."""
import os
import subprocess

import pytest

from tests.conftest import ROOT

SRC = os.path.join(ROOT, "tests", "peer", "peer_host.cc")
CUDA_INC = "/usr/local/cuda/include"


def _build(name, extra):
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, name)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-I" + CUDA_INC] + extra + ["-o", exe, SRC, "-lrt"],
                       capture_output=True, text=True)
    return exe, r


@pytest.mark.parametrize("world,pr", [(2, 2), (2, 1), (4, 2), (6, 3), (8, 4), (8, 2)])
def test_broadcast_protocol_on_host_threads(world, pr):
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed")
    exe, r = _build("peer_host", [])
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, str(world), str(pr), "300"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, "rank/code %d: %s" % (r.returncode, r.stderr)


def test_broadcast_protocol_under_thread_sanitizer():
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed")
    exe, r = _build("peer_host_tsan", ["-fsanitize=thread"])
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr[-200:])
    r = subprocess.run([exe, "8", "4", "60"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "data race" not in r.stdout + r.stderr, (r.stdout + r.stderr)[-2000:]
