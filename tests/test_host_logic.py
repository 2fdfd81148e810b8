"""Host-side logic that needs no GPU: grid/layout, DAG order and payloads, generators, descriptor
validation, CLI parsing, CSV contract."""
import json
import os

import numpy as np
import pytest

from dense_linear_app_b200 import bench_sweep, dag, v6_test
from dense_linear_app_b200.grid import LocalLayout, ProcessGrid, panel_slots
from dense_linear_app_b200.tiles import TileDesc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ---- grid ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("P,Q,nt", [(1, 1, 5), (1, 2, 7), (2, 2, 9), (2, 4, 16), (3, 2, 11), (2, 4, 3)])
def test_layout_partitions_lower_tiles(P, Q, nt):
    g = ProcessGrid(P, Q)
    seen = {}
    for r in range(g.size):
        lay = LocalLayout(nt, g, r)
        idx = [lay.index(i, j) for i, j in lay.tiles()]
        assert idx == list(range(lay.ntiles))            # storage order is dense and ascending
        for i, j in lay.tiles():
            assert g.owner(i, j) == r and i >= j
            assert (i, j) not in seen
            seen[(i, j)] = r
        for j in lay.cols:                               # owned tiles of a panel column are contiguous
            rows = list(lay.rows_in_col(j, j))
            if rows:
                ids = [lay.index(i, j) for i in rows]
                assert ids == list(range(ids[0], ids[0] + len(ids)))
    assert len(seen) == nt * (nt + 1) // 2


def test_grid_for_world():
    assert [(g.P, g.Q) for g in map(ProcessGrid.for_world, (1, 2, 4, 8))] == [(1, 1), (1, 2), (2, 2), (2, 4)]


def test_panel_slots_group_by_owner_row():
    slot, groups = panel_slots(10, 3, 2)
    assert sorted(slot) == list(range(3, 10))
    assert sorted(slot.values()) == list(range(7))
    for p, first, cnt in groups:
        rows = [i for i in range(3, 10) if i % 3 == p]
        assert [slot[i] for i in rows] == list(range(first, first + cnt))


# ---- DAG -------------------------------------------------------------------------------------------
def test_dag_counts_and_order():
    N, B = 20, 4
    tasks = list(dag.build_dag(N, B))
    nb = 5
    counts = dag.task_counts(N, B)
    assert counts == {"POTRF": nb, "TRSM": 10, "SYRK": 10, "GEMM": 10}
    for op, n in counts.items():
        assert sum(t.op == op for t in tasks) == n
    # reference order inside wave 0 (C1:278-333)
    w0 = [t for t in tasks if t.k == 0]
    assert [t.op for t in w0[:5]] == ["POTRF", "TRSM", "TRSM", "TRSM", "TRSM"]
    assert [(t.op, t.out) for t in w0[5:8]] == [("SYRK", (1, 1)), ("GEMM", (2, 1)), ("SYRK", (2, 2))]
    # dependencies: every read tile was last written by an earlier task or is an input
    done = set()
    for t in tasks:
        if t.op == "TRSM":
            assert ("POTRF", t.deps[1]) in done
        if t.op in ("SYRK", "GEMM"):
            for d in t.deps[1:]:
                assert ("TRSM", d) in done
        done.add((t.op, t.out))
    # flops add up to N^3/3 in units of B^3
    units = counts["POTRF"] / 3 + counts["TRSM"] + counts["SYRK"] + 2 * counts["GEMM"]
    assert abs(units - nb ** 3 / 3) < 1e-9


def test_ragged_dag_rounds_up():
    assert dag.task_counts(10, 4)["POTRF"] == 3


def test_payloads_match_reference_schema():
    assert json.loads(dag.make_payload_potrf("a", 4)) == {"op": "POTRF", "B": 4, "in": "a"}
    assert json.loads(dag.make_payload_trsm("l", "a", 4)) == {"op": "TRSM", "B": 4, "inL": "l", "inA": "a"}
    assert json.loads(dag.make_payload_syrk("c", "a", 4)) == {"op": "SYRK", "B": 4, "inC": "c", "inA": "a"}
    assert json.loads(dag.make_payload_gemm("c", "x", "y", 4)) == {"op": "GEMM", "B": 4, "inC": "c", "inAi": "x",
                                                                  "inAj": "y"}
    assert list(json.loads(dag.make_payload_gemm("c", "x", "y", 4))) == ["op", "B", "inC", "inAi", "inAj"]
    assert dag.block_id_from_ij(3, 1) == "blk/3/1"
    t = dag.TileTask("TRSM", 0, (2, 0), ((2, 0), (0, 0)))
    payload, deps = dag.payload_for(t, {(2, 0): "A20", (0, 0): "L00"}, 8)
    assert json.loads(payload)["inL"] == "L00" and deps == ["L00", "A20"]   # v1 routing, not v2's heuristic


def test_run_waves_with_oracle_executor(oracle):
    """The client loop against a CPU executor (oracle) reproduces dpotrf: DAG + payload routing."""
    N, B = 24, 8
    A = dag.enforce_strict_diag_dominance(dag.make_spd_like_chameleon(N))
    nb = N // B
    blocks = {dag.block_id_from_ij(i, j): dag.extract_block(A, B, i, j).tobytes(order="F")
              for i in range(nb) for j in range(i + 1)}

    def submit_one(payload, deps):
        p = json.loads(payload)
        t = lambda name: np.frombuffer(deps[p[name]]).reshape(B, B).T.copy(order="F")  # noqa: E731
        if p["op"] == "POTRF":
            a = t("in"); assert oracle.potrf_tile(a) == 0
        elif p["op"] == "TRSM":
            a = t("inA"); oracle.trsm_tile(t("inL"), a)
        elif p["op"] == "SYRK":
            a = t("inC"); oracle.syrk_tile(t("inA"), a)
        else:
            a = t("inC"); oracle.gemm_tile(t("inAi"), t("inAj"), a)
        return a.tobytes(order="F")

    out = dag.run_waves(N, B, blocks, submit_one)
    L = np.zeros((N, N))
    for i in range(nb):
        for j in range(i + 1):
            L[i * B:(i + 1) * B, j * B:(j + 1) * B] = np.frombuffer(out[dag.block_id_from_ij(i, j)]).reshape(B, B).T
    L = np.tril(L)
    assert np.linalg.norm(L @ L.T - A) / np.linalg.norm(A) < 1e-15


# ---- generators --------------------------------------------------------------------------------------
def test_mt19937_64_matches_libstdcxx_golden():
    g = json.load(open(os.path.join(GOLD, "mt19937_64.json")))
    v = dag.MT19937_64(g["seed"]).uniform(1000)
    assert [float.fromhex(h) for h in g["first8_hex"]] == list(v[:8])
    assert [float.fromhex(h) for h in g["last8_of_1000_hex"]] == list(v[-8:])
    assert int(dag.MT19937_64(42).raw(700)[-1]) == g["seed42_raw_700th"]


def test_v1_normal_generator_matches_libstdcxx_golden():
    """generate_random_B_block (C1:102-108): std::normal_distribution on mt19937_64(42), bit for bit."""
    g = json.load(open(os.path.join(GOLD, "normal_dist.json")))
    d = dag.NormalDist(dag.MT19937_64(g["seed"]))
    v = [d() * g["scale"] for _ in range(2000)]
    assert [float.fromhex(h) for h in g["first8_hex"]] == v[:8]
    assert [float.fromhex(h) for h in g["last8_of_2000_hex"]] == v[-8:]
    # the block helper continues ONE stream across calls, like the reference's function-level static
    d2 = dag.NormalDist(dag.MT19937_64(42))
    b0, b1 = dag.generate_random_B_block(4, 0.1, d2), dag.generate_random_B_block(4, 0.1, d2)
    assert list(b0) == v[:16] and list(b1) == v[16:32]


def v1_symmetric_matrix(blocks, N, B):
    """The symmetric matrix the v1 tiles stand for: lower tiles as given, diagonal tiles' LOWER triangle."""
    A = np.zeros((N, N))
    for (i, j), blob in blocks.items():
        t = blob.reshape(B, B).T          # the worker reads the blob column-major (W1:212-227)
        A[i * B:(i + 1) * B, j * B:(j + 1) * B] = np.tril(t) if i == j else t
    return A + np.tril(A, -1).T


def test_v1_blocks_have_nonsymmetric_diagonal_tiles_and_factor_with_the_oracle(oracle):
    """C1:189-192: N(0, 0.1^2) tiles, +B on the diagonal of diagonal tiles.  The diagonal tiles are not
    symmetric — running the client DAG on them is the reference's own evidence that POTRF and SYRK read
    the lower triangle only."""
    N, B = 24, 8
    blocks = dag.make_blocks_v1(N, B)
    d00 = blocks[(0, 0)].reshape(B, B)
    assert not np.allclose(d00, d00.T) and abs(d00[0, 0] - B) < 1.0
    A = v1_symmetric_matrix(blocks, N, B)
    named = {dag.block_id_from_ij(i, j): v.tobytes() for (i, j), v in blocks.items()}

    def submit_one(payload, deps):
        p = json.loads(payload)
        t = lambda name: np.frombuffer(deps[p[name]]).reshape(B, B).T.copy(order="F")  # noqa: E731
        if p["op"] == "POTRF":
            a = t("in"); assert oracle.potrf_tile(a) == 0
        elif p["op"] == "TRSM":
            a = t("inA"); oracle.trsm_tile(t("inL"), a)
        elif p["op"] == "SYRK":
            a = t("inC"); oracle.syrk_tile(t("inA"), a)
        else:
            a = t("inC"); oracle.gemm_tile(t("inAi"), t("inAj"), a)
        return a.tobytes(order="F")

    out = dag.run_waves(N, B, named, submit_one)
    nb = N // B
    L = np.zeros((N, N))
    for i in range(nb):
        for j in range(i + 1):
            L[i * B:(i + 1) * B, j * B:(j + 1) * B] = np.frombuffer(out[dag.block_id_from_ij(i, j)]).reshape(B, B).T
    L = np.tril(L)
    assert np.linalg.norm(L @ L.T - A) / np.linalg.norm(A) < 1e-15


def test_make_spd_like_chameleon():
    N = 12
    A = dag.make_spd_like_chameleon(N)
    assert np.array_equal(A, A.T)
    v = dag.MT19937_64(12345).uniform(N * (N + 1) // 2)
    assert A[0, 0] == v[0] + 100.0 and A[1, 0] == v[1] and A[1, 1] == v[N] + 100.0   # column by column
    off = np.abs(A).sum(1) - np.abs(np.diag(A))
    B = dag.enforce_strict_diag_dominance(dag.make_spd_like_chameleon(N, bump=0.0))
    assert np.all(np.diag(B) > np.abs(B).sum(1) - np.abs(np.diag(B)))
    assert np.all(np.diag(A) > off)
    U = dag.make_spd_like_chameleon(N, uplo="U")
    assert np.array_equal(U, U.T) and U[0, 1] == v[1]


def test_extract_block_zero_pads():
    A = np.arange(25.0).reshape(5, 5)
    blk = dag.extract_block(A, 4, 1, 1)
    assert blk.flags.f_contiguous and blk[0, 0] == A[4, 4] and blk.sum() == A[4, 4]
    assert np.array_equal(dag.extract_block(A, 4, 0, 0), A[:4, :4])


def test_load_params():
    assert (dag.load_params([], {}).N, dag.load_params([], {}).B) == (12, 4)
    p = dag.load_params(["--N=64", "--B=16"], {"CHOLESKY_N": "7"})
    assert (p.N, p.B) == (64, 16)
    p = dag.load_params(["32", "8"], {})
    assert (p.N, p.B) == (32, 8)
    assert dag.load_params([], {"CHOLESKY_N": "40", "CHOLESKY_B": "x"}).B == 4
    with pytest.raises(ValueError):
        dag.load_params(["--N=0"], {})


# ---- descriptor / CLI ----------------------------------------------------------------------------------
def test_tiledesc_validation():
    TileDesc.square(1000, 128).validate()
    TileDesc.one_block(4).validate()
    with pytest.raises(ValueError):
        TileDesc(128, 128, 100, 1000, 1000, 0, 0, 1000, 1000).validate()      # bsiz < mb*nb
    with pytest.raises(ValueError):
        TileDesc(128, 128, 128 * 128, 1000, 1000, 8, 0, 1000, 1000).validate()  # sub-matrix outside
    with pytest.raises(ValueError):
        TileDesc(128, 64, 128 * 64, 1000, 1000, 0, 0, 1000, 1000).validate()   # non-square tiles


def test_v6_test_usage_and_atoi(capsys):
    assert v6_test.main(["v6_test", "1", "2"]) == 1
    assert "Usage:" in capsys.readouterr().err
    assert [v6_test._atoi(s) for s in ("42", " 7x", "abc", "-3")] == [42, 7, 0, -3]


def test_sweep_contract():
    assert bench_sweep.HEADER.strip() == "timestamp,scheduler,mapping,ncpu,ngpu,N,NB,run_idx,ms,exit_code,gflops,rel_error"
    out = "[setup] x\nN = 1000, NB = 128\nTime: 0.003 s\nPerformance: 101.25 Gflop/s\n||A - LL^T||_inf / ||A||_inf = 9.21e-16\n"
    assert bench_sweep.parse_metrics(out) == (101.25, 9.21e-16)
    assert bench_sweep.parse_metrics("garbage") == (-1.0, -1.0)
    assert bench_sweep.driver_argv(0, 1, 1000, 128, 1, 1, 42) == ["0", "1", "1000", "128", "128", "128", "16384", "1000",
                                                                  "1000", "0", "0", "1000", "1000", "1", "1", "42"]
    assert bench_sweep.NS == (1000, 5000, 8000, 12000, 16000) and bench_sweep.NBS[0] == 128 and bench_sweep.REPEATS == 8


def test_v3_named_argument_cli_checks():
    """v3_script_cholesky_x_arg_gpt.c:131-199: all 20 options required, strict geometry, same messages."""
    import io
    from dense_linear_app_b200 import v3_cli
    good = ("--N 3000 --NB 256 --ncpu 4 --ngpu 1 --mat none --dtyp d --mb 256 --nb 256 --bsiz 65536 --lm 3000 "
            "--ln 3000 --i 0 --j 0 --m 3000 --n 3000 --p 1 --q 1 --bump 3000 --uplo L --seed 51").split()

    def run(args):
        err = io.StringIO()
        code, a = v3_cli.parse(["v3"] + args, err)
        return code, a, err.getvalue()

    code, a, _ = run(good)
    assert code is None and a["N"] == 3000 and a["bsiz"] == 65536 and a["bump"] == 3000.0 and a["seed"] == 51
    assert a["dtyp"] == "d" and a["uplo"] == "L" and not a["mat_user"]
    code, a, _ = run(["--N=3000"] + good[2:])
    assert code is None and a["N"] == 3000                         # --name=value form
    assert run(good[:-2])[0] == 1 and "all options are required" in run(good[:-2])[2]
    assert run(["--help"])[0] == 0
    assert run(["--bogus", "1"] + good)[0] == 1

    def with_(name, val):
        out = list(good)
        out[out.index("--" + name) + 1] = val
        return out

    assert "invalid --dtyp q" in run(with_("dtyp", "q"))[2]
    assert "invalid --uplo X" in run(with_("uplo", "X"))[2]
    assert "must be >0" in run(with_("p", "0"))[2]
    assert "--bsiz < mb*nb (bsiz=100 mb=256 nb=256)" in run(with_("bsiz", "100"))[2]
    assert "invalid offsets i=3000 j=0 (lm=3000 ln=3000)" in run(with_("i", "3000"))[2]
    assert "submatrix (i=8,m=3000) outside lm=3000" in run(with_("i", "8"))[2]
    code, a, msg = run(with_("bump", "0"))
    assert code is None and "bump==0" in msg
    assert run(with_("uplo", "u"))[1]["uplo"] == "U" and run(with_("dtyp", "2"))[1]["dtyp"] == "z"
    # unsupported-but-valid choices are refused before any CUDA work
    assert v3_cli.main(["v3"] + with_("uplo", "B")) == 1 and v3_cli.main(["v3"] + with_("dtyp", "s")) == 1
