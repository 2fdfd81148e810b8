"""Whole-job checks of the multi-rank schedule on the CPU (tests/_schedule_sim.py): no deadlock, every
collective matched on every participant, and no race anywhere — for the NCCL transport and for the
peer-push transport (copy-engine pushes, flag words, credits), including back-to-back factorizations
through the same receive slots."""
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


@pytest.mark.parametrize("P,Q,nt", [(1, 2, 9), (2, 2, 11), (2, 4, 17), (2, 4, 5), (3, 2, 10), (2, 1, 6)])
@pytest.mark.parametrize("runs", [1, 2])
def test_peer_transport_whole_job(P, Q, nt, runs):
    """Pushes + flags + credits: every wait has its post, no deadlock, and no remote write into a
    slot (panel or L_kk) that a slower rank is still reading — also across two factorizations."""
    g, problems, races = simulate(maker(P, Q, nt), P * Q, runs=runs, lookahead=True, transport="peer")
    assert problems == []
    assert races == []
    assert sum(k == "peer-copy" for k in g.kind) > 0 and sum(k == "flagwait" for k in g.kind) > 0


def test_peer_transport_two_slots_are_safe_with_credits_only():
    """Two receive slots are race-free because the sender waits for the reader's credit; without
    the credits the one-sided writes DO race with a lagging reader (proof that the checker sees it)."""
    g, problems, races = simulate(maker(2, 4, 17), 8, lookahead=True, transport="peer", nslots=2)
    assert problems == [] and races == []
    g, problems, races = simulate(maker(2, 4, 17), 8, lookahead=True, transport="peer", nslots=2, credits=False)
    assert any(reg[1] in ("P", "D") for reg, _, _ in races)
