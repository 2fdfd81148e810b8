"""Whole-job checks of the multi-rank schedule on the CPU (tests/_schedule_sim.py): no deadlock, every
collective matched on every participant, and no race anywhere — for the measured NCCL transport and for
the experimental symmetric-memory transport, whose receive-slot count is exactly what this verifies."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _schedule_sim import simulate  # noqa: E402
from dense_linear_app_b200.tiles import TileDesc, TileMatrix  # noqa: E402


def maker(P, Q, nt, b=16):
    N = nt * b
    return lambda r: TileMatrix(TileDesc(b, b, b * b, N, N, 0, 0, N, N, P, Q), r, "cpu")


@pytest.mark.parametrize("P,Q,nt", [(1, 2, 7), (2, 1, 6), (2, 2, 9), (2, 4, 13), (3, 2, 8), (2, 4, 3)])
@pytest.mark.parametrize("lookahead", [True, False])
def test_nccl_schedule_whole_job(P, Q, nt, lookahead):
    g, problems, races = simulate(maker(P, Q, nt), P * Q, lookahead=lookahead, transport="nccl")
    assert problems == []
    assert races == []
    assert sum(k == "bcast-done" for k in g.kind) > 0


@pytest.mark.parametrize("P,Q,nt", [(1, 2, 9), (2, 2, 11), (2, 4, 17), (2, 4, 5)])
def test_symm_transport_whole_job(P, Q, nt):
    """Peer copies + flags with Q+P+2 receive slots: matched flags, no deadlock, and in particular no
    remote write into a slot that a slower rank is still reading."""
    g, problems, races = simulate(maker(P, Q, nt), P * Q, lookahead=True, transport="symm")
    assert problems == []
    assert races == []
    assert sum(k == "peer-copy" for k in g.kind) > 0 and sum(k == "put" for k in g.kind) > 0


def test_symm_transport_needs_more_than_two_slots():
    """With only the two slots the NCCL path uses, one-sided writes DO race with a lagging reader —
    the reason the symmetric transport allocates Q+P+2 (and the proof that the checker sees it)."""
    g, problems, races = simulate(maker(2, 4, 17), 8, lookahead=True, transport="symm", nslots=2)
    assert problems == []
    assert any(reg[1] == "P" for reg, _, _ in races)
