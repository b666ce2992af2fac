"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI,
against the oracle on seeded inputs, against the committed reference fixtures, and --
at BASELINE.json's full sizes -- through size-independent properties.

Tolerances (BASELINE.json north_star): row_ptr and col_idx bit-exact; values 1e-12
relative for fp64 and 1e-5 for fp32 (accumulation order differs from the oracle's)."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
import mh_spgemm_b200  # noqa: F401
from mh_spgemm_b200 import api, generators as G
from mh_spgemm_b200.csr import CSR

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


def sha(a, dt):
    return hashlib.sha256(np.ascontiguousarray(a, dt).tobytes()).hexdigest()


def assert_matches(orc, C, Cp, Cc, Cv):
    assert np.array_equal(C.ptr.astype(np.int64), Cp), "row_ptr differs"
    assert np.array_equal(C.col, Cc), "col_idx differs"
    bad, first = orc.compare(C.M, (C.ptr, C.col, C.val), (Cp, Cc, Cv), RTOL[C.val.dtype])
    assert bad == 0, f"{bad} values out of tolerance, first at {first}"


INPUTS = {
    "tiny": lambda: (G.uniform_random(64, 64, 300, seed=1), None),
    "rect": lambda: (G.uniform_random(200, 300, 2000, seed=2), G.uniform_random(300, 5000, 9000, seed=3)),
    "poisson32": lambda: (G.poisson2d(32), None),
    "fem": lambda: (G.fem3d(4, 4, 10, 3, seed=5), None),
    "fem_perturbed": lambda: (G.fem3d_perturbed(4, 4, 10, 3, drop=0.10, seed=5), None),
    "rmat14": lambda: (G.rmat(14, 16000, 60000, seed=6), None),
    "dense_rows": lambda: (G.with_dense_rows(G.uniform_random(3000, 3000, 30000, seed=8), 6, 1500, seed=9), None),
    "banded": lambda: (G.banded_random(5000, 12, 300, seed=10), None),
    "road": lambda: (G.road_grid(60), None),
    "one_row": lambda: (CSR(1, 1, [0, 1], [0], [2.0]), None),
    "empty": lambda: (CSR(5, 5, np.zeros(6, np.int32), [], []), None),
    "ragged": lambda: (CSR(3, 3, [0, 0, 3, 3], [0, 1, 2], [1.0, 2.0, 3.0]), None),
}


@pytest.mark.parametrize("force", [(0, 0), (1, 1), (2, 2), (1, 2), (2, 1)])
@pytest.mark.parametrize("name", list(INPUTS))
def test_spgemm_matches_oracle(tool, orc, name, force):
    """Every accumulator path (auto / dense bitmap+window / hash) gives the same CSR."""
    A, B = INPUTS[name]()
    B = A if B is None else B
    tool.set_option("force_sym_path", force[0])
    tool.set_option("force_num_path", force[1])
    try:
        C = tool.spgemm_host(A, B)
    finally:
        tool.set_option("force_sym_path", 0)
        tool.set_option("force_num_path", 0)
    Cp, Cc, Cv = orc.spgemm(A, B)
    assert_matches(orc, C, Cp, Cc, Cv)
    assert tool.stats["intprod"] == orc.intprod(A, B)
    assert tool.stats["gpu_launches"] > 0


@pytest.mark.parametrize("name", ["poisson32", "fem", "rmat14", "dense_rows"])
def test_fp32(tool, orc, name):
    A, B = INPUTS[name]()
    A = A.astype(np.float32)
    C = tool.spgemm_host(A, A)
    Cp, Cc, Cv = orc.spgemm(A, A)
    assert C.val.dtype == np.float32
    assert_matches(orc, C, Cp, Cc, Cv)


def test_poisson_is_exact(tool, orc):
    """BASELINE configs[0]: Poisson 256^2, A*A is exact in fp64 -> values bit-for-bit."""
    A = G.poisson2d(256)
    C = tool.spgemm_host(A, A)
    Cp, Cc, Cv = orc.spgemm(A, A)
    assert C.nnz == 846852
    assert np.array_equal(C.ptr.astype(np.int64), Cp) and np.array_equal(C.col, Cc)
    assert np.array_equal(C.val, Cv)


@pytest.mark.parametrize("name", list(cases.SMALL))
def test_matches_reference_fixture_full(tool, name):
    """Bit-exact structure against what the reference's own kernels produced on a B200."""
    A, B = cases.SMALL[name]()
    B = A if B is None else B
    ref = np.load(os.path.join(GOLD, f"ref_{name}.npz"))
    C = tool.spgemm_host(A, B)
    assert np.array_equal(C.ptr, ref["ptr"]) and np.array_equal(C.col, ref["col"])
    np.testing.assert_allclose(C.val, ref["val"], rtol=1e-12, atol=0)
    # family 1 against the reference's mask matrix
    tp, tc, tm = tool.mask_matrix_B(B.M, B.N, api.DeviceArray(B.ptr), api.DeviceArray(B.col))
    assert np.array_equal(tp, ref["tileptr"]) and np.array_equal(tc, ref["tilecol"])
    assert np.array_equal(tm, ref["tilemask"])


@pytest.mark.parametrize("name", ["dense_rows", "rmat_s14", "fem_small", "F_cant_like", "R_webbase_like"])
def test_matches_reference_fixture_checksums(tool, name):
    """Full-size BASELINE configs[1] and [2] against checksums of the reference's output."""
    A, B = cases.LARGE[name]()
    B = A if B is None else B
    meta = json.load(open(os.path.join(GOLD, f"ref_{name}.json")))
    C = tool.spgemm_host(A, B)
    assert C.nnz == meta["nnz"]
    assert sha(C.ptr, np.int32) == meta["sha_ptr"]
    assert sha(C.col, np.int32) == meta["sha_col"]
    assert float(C.val.sum()) == pytest.approx(meta["sum_val"], rel=1e-10)
    w = (np.arange(C.val.size, dtype=np.int64) % 97 + 1).astype(np.float64)
    assert float((C.val * w).sum()) == pytest.approx(meta["sum_weighted"], rel=1e-10)
    tp, tc, tm = tool.mask_matrix_B(B.M, B.N, api.DeviceArray(B.ptr), api.DeviceArray(B.col))
    assert sha(tp, np.int32) == meta["sha_tileptr"]
    assert sha(tc, np.int32) == meta["sha_tilecol"]
    assert sha(tm, np.uint32) == meta["sha_tilemask"]


def test_live_reference_same_box():
    """C's structure bit-exact against the reference's own GPU implementation run here, now
    (separate process: a fault in the reference must not take the test session down)."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libmhref.so")):
        pytest.skip("oracle/_ref not built")
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); import mh_spgemm_b200\n"
        "from mh_spgemm_b200 import api, generators as G; from oracle import Reference\n"
        "A = G.fem3d(6, 6, 20, 3, seed=77)\n"
        "C = api.Tool(0).spgemm_host(A, A); R = Reference().spgemm(A, A)\n"
        "assert np.array_equal(C.ptr, R['ptr']) and np.array_equal(C.col, R['col'])\n"
        "np.testing.assert_allclose(C.val, R['val'], rtol=1e-12, atol=0); print('LIVE-OK', C.nnz)\n" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert "LIVE-OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


def test_stage_outputs(tool, orc):
    """Families 1 and 2 stage by stage: mask matrix, per-row products / tile-flop / column
    span, and the bins (a stable partition of the rows that agrees with the ladder)."""
    A = G.rmat(13, 8000, 40000, seed=12)
    dAp, dAc = api.DeviceArray(A.ptr), api.DeviceArray(A.col)
    tp, tc, tm = tool.mask_matrix_B(A.M, A.N, dAp, dAc)
    otp, otc, otm = orc.mask_matrix(A)
    assert np.array_equal(tp, otp) and np.array_equal(tc, otc) and np.array_equal(tm, otm)
    dCp, nnzC = tool.symbolic(A.M, A.N, A.N, dAp, dAc, dAp, dAc)
    Cp = orc.symbolic(A, A)
    assert nnzC == Cp[-1] and np.array_equal(dCp.numpy().astype(np.int64), Cp)
    info = tool.row_info(A.M)
    ip = orc.row_intprod(A, A)
    assert np.array_equal(info[:, 0], ip)
    assert np.array_equal(info[:, 1], orc.row_tileflop(A, otp))
    # column span of every non-empty C row
    nz = np.diff(Cp) > 0
    _, Cc, _ = orc.spgemm(A, A)
    first = Cc[Cp[:-1][nz]]
    last = Cc[Cp[1:][nz] - 1]
    assert np.array_equal(info[nz, 2], first) and np.array_equal(info[nz, 3], last)
    for which, key in ((0, "sym_bins"), (1, "num_bins")):
        nb, bins, off = tool.bins(which, A.M)
        assert off[0] == 0 and off[nb] == A.M
        assert np.array_equal(np.sort(bins), np.arange(A.M)), "bins must be a permutation of the rows"
        for b in range(nb):
            seg = bins[off[b]:off[b + 1]]
            assert np.all(np.diff(seg) > 0), "row ids ascending inside a bin (stable partition)"
        assert np.array_equal(np.diff(off[:nb + 1]), list(tool.stats[key].values())[:nb])
    nb, bins, off = tool.bins(1, A.M)
    assert np.array_equal(np.sort(bins[off[0]:off[1]]), np.nonzero(~nz)[0]), "bin 0 == empty C rows"


def test_symbolic_numeric_contract(tool, orc):
    """Two separately callable phases; numeric re-runnable with new values on one pattern."""
    A = G.fem3d(5, 5, 12, 3, seed=21)
    dAp, dAc, dAv = api.DeviceArray(A.ptr), api.DeviceArray(A.col), api.DeviceArray(A.val)
    dCp, nnzC = tool.symbolic(A.M, A.N, A.N, dAp, dAc, dAp, dAc)
    Cp, Cc, Cv = orc.spgemm(A, A)
    assert nnzC == Cp[-1]
    dCc, dCv = tool.numeric(dAv, dAv, nnzC)
    assert np.array_equal(dCc.numpy()[:nnzC], Cc)
    np.testing.assert_allclose(dCv.numpy()[:nnzC], Cv, rtol=1e-12, atol=0)
    A2 = CSR(A.M, A.N, A.ptr, A.col, A.val * 3.0 - 1.0)
    dAv2 = api.DeviceArray(A2.val)
    dCc2, dCv2 = tool.numeric(dAv2, dAv2, nnzC)
    _, Cc2, Cv2 = orc.spgemm(A2, A2)
    assert np.array_equal(dCc2.numpy()[:nnzC], Cc2)
    bad, _ = orc.compare(A.M, (Cp, Cc2, dCv2.numpy()[:nnzC]), (Cp, Cc2, Cv2), 1e-12)
    assert bad == 0


def test_numeric_before_symbolic_is_an_error():
    t = api.Tool(0)
    d = api.DeviceArray(np.zeros(4))
    with pytest.raises(api.MhbError):
        t.numeric_into(d, d, d, d)
    t.release()


def test_special_values_stay_structural(tool, orc):
    """Cancellation to 0.0, NaN and Inf products remain stored entries (no value test on the
    accumulate path, inc/numeric.cuh:237-241); the window's 'unset' marker never collides."""
    A = CSR(3, 3, [0, 2, 3, 4], [0, 1, 1, 2], [1.0, -1.0, np.inf, np.nan])
    B = CSR(3, 3, [0, 1, 2, 3], [0, 0, 2], [1.0, 1.0, 0.0])
    for force in (1, 2):
        tool.set_option("force_num_path", force)
        C = tool.spgemm_host(A, B)
        tool.set_option("force_num_path", 0)
        assert C.ptr.tolist() == [0, 1, 2, 3] and C.col.tolist() == [0, 0, 2]
        assert C.val[0] == 0.0 and np.isinf(C.val[1]) and np.isnan(C.val[2])


def test_full_size_properties(tool):
    """Size-independent properties at BASELINE's full sizes (no oracle needed): sorted and
    duplicate-free rows, row sums equal A*(B*1) (linearity), pattern symmetric for a
    symmetric-pattern input, and idempotence of a repeated call."""
    for A in (G.fem3d(), G.rmat()):
        C = tool.spgemm_host(A, A)
        assert C.is_canonical()
        S = A.to_scipy()
        want = S @ (S @ np.ones(A.N))
        got = np.add.reduceat(np.append(C.val, 0.0), np.minimum(C.ptr[:-1], C.nnz))
        got[np.diff(C.ptr) == 0] = 0.0
        np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-9)
        C2 = tool.spgemm_host(A, A)
        assert np.array_equal(C.ptr, C2.ptr) and np.array_equal(C.col, C2.col)
        np.testing.assert_allclose(C.val, C2.val, rtol=1e-12, atol=0)
    # FEM pattern is symmetric -> so is the pattern of A*A
    A = G.fem3d(6, 6, 30, 3, seed=9)
    C = tool.spgemm_host(A, A)
    P = C.to_scipy()
    P.data[:] = 1
    assert (P != P.T).nnz == 0


def test_cpp_shim_reference_style_driver():
    """A main()-style C++ driver against include/mhb_compat.hpp (CSR / Tool / Timing /
    MH_spgemm with the reference's signature) runs and passes its CSR::operator== check."""
    exe = os.path.join(ROOT, "tests", "cpp", "compat_driver")
    if not os.path.exists(exe):
        from importlib import import_module
        import_module("mh_spgemm_b200.build").build_compat_driver()
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "pass" in p.stdout, p.stdout[-1500:] + p.stderr[-1500:]


def test_row_sharded_slices_concatenate(tool, orc):
    """The multi-GPU decomposition on one device: A split into row blocks balanced by
    intermediate products, each block multiplied on its own, slices concatenated by
    row_ptr offset == the unsharded product (no collective is involved in the data path)."""
    from mh_spgemm_b200 import distributed as D
    A = G.rmat(14, 16000, 60000, seed=6)
    full = tool.spgemm_host(A, A)
    b = D.partition_rows(D.row_work(A, A), 4)
    slices = []
    for g in range(4):
        blk = A.rows(int(b[g]), int(b[g + 1]))
        C = tool.spgemm_host(blk, A)
        slices.append((C.ptr, C.col, C.val))
    gp, gc, gv = D.concat_slices(slices)
    assert np.array_equal(gp, full.ptr.astype(np.int64)) and np.array_equal(gc, full.col)
    np.testing.assert_allclose(gv, full.val, rtol=1e-12, atol=0)


def test_sliced_product_for_int32_overflow(tool, orc):
    """nnz(C) beyond int32 is handled by row slices with local int32 row_ptr and int64 slice
    offsets; exercised here with an artificially small cap."""
    from mh_spgemm_b200 import distributed as D
    A = G.fem3d(4, 4, 30, 3, seed=8)
    slices, offs = tool.spgemm_sliced(A, A, cap=200_000)
    assert len(slices) > 3
    Cp, Cc, Cv = orc.spgemm(A, A)
    gp, gc, gv = D.concat_slices([(c.ptr, c.col, c.val) for _, _, c in slices])
    assert np.array_equal(gp, Cp) and np.array_equal(gc, Cc)
    assert np.array_equal(offs, Cp[[s[0] for s in slices] + [A.M]])
    np.testing.assert_allclose(gv, Cv, rtol=1e-12, atol=0)


def test_overflow_is_reported_not_truncated():
    """nnz(C) above the int32 contract (simulated with the nnz_limit knob) is a status, and
    the handle stays usable."""
    t = api.Tool(0)
    A = G.poisson2d(24)
    t.set_option("nnz_limit", 100)
    with pytest.raises(api.MhbError) as e:
        t.spgemm_host(A, A)
    assert e.value.code == 3 and "shard" in str(e.value)
    t.set_option("nnz_limit", 0)
    assert t.spgemm_host(A, A).nnz > 100
    t.release()


def test_cli_driver_matches_reference_report(tmp_path, orc):
    """`spgemm <file.mtx>` equivalent (src/main.cu:74-217): same report lines, right nnz, and
    the AAT mode (A * A^T through the host transpose of src/utils.cpp:20-46)."""
    from mh_spgemm_b200.mmio import read_mtx, write_mtx
    A = G.uniform_random(300, 200, 2500, seed=3)
    sq = G.fem3d(3, 3, 8, 2, seed=4)
    pa, ps = tmp_path / "rect.mtx", tmp_path / "fem.mtx"
    write_mtx(str(pa), A)
    write_mtx(str(ps), sq, symmetric_lower=False)
    env = dict(os.environ, PYTHONPATH=ROOT)
    run = lambda *a: subprocess.run([sys.executable, "-m", "mh_spgemm_b200.cli", *a], capture_output=True,  # noqa: E731
                                    text=True, timeout=300, env=env, cwd=ROOT)
    p = run(str(ps), "--out", str(tmp_path / "c.mtx"))
    assert p.returncode == 0, p.stdout + p.stderr
    Cp, Cc, Cv = orc.spgemm(sq, sq)
    assert f"C.nnz = {Cp[-1]}" in p.stdout and f"SpGEMM intermediate result = {orc.intprod(sq, sq)}" in p.stdout
    assert "MH-SpGEMM runtime is" in p.stdout and "Gflops is" in p.stdout and "calculate_C_nnz" in p.stdout
    C, _ = read_mtx(str(tmp_path / "c.mtx"))
    assert np.array_equal(C.ptr, Cp) and np.array_equal(C.col, Cc)
    np.testing.assert_allclose(C.val, Cv, rtol=1e-12)
    p = run(str(pa))
    assert "C=AA must have rowA = colA. Exit." in p.stdout
    # AAT: B = A^T through the device transpose; the WHOLE product is compared, not its size
    p = run(str(pa), "--aat", "--out", str(tmp_path / "aat.mtx"), "--write", str(tmp_path / "data"))
    Tp, Tc, Tv = orc.spgemm(A, orc.transpose(A))
    assert p.returncode == 0 and f"C.nnz = {Tp[-1]}" in p.stdout, p.stdout + p.stderr
    C, _ = read_mtx(str(tmp_path / "aat.mtx"))
    assert C.M == A.M and C.N == A.M
    assert np.array_equal(C.ptr, Tp) and np.array_equal(C.col, Tc)
    np.testing.assert_allclose(C.val, Tv, rtol=1e-12)
    # WRITE (src/main.cu:201-213): one "%.2f" Gflops line appended per run
    lines = open(tmp_path / "data" / "Gflops_MH-SpGEMM.csv").read().split()
    assert len(lines) == 1 and float(lines[0]) > 0
    # process.sh: a list of names resolved under a root directory, missing files skipped with a warning
    os.makedirs(tmp_path / "matrix" / "fem")
    os.replace(ps, tmp_path / "matrix" / "fem" / "fem.mtx")
    (tmp_path / "list.txt").write_text("fem\nnot_there\n")
    p = run("--list", str(tmp_path / "list.txt"), "--root", str(tmp_path / "matrix"), "--write", str(tmp_path / "data"))
    assert p.returncode == 0, p.stdout + p.stderr
    assert "Total matrices to process: 2" in p.stdout and "Warning: File not found" in p.stdout
    assert "All listed matrices processed successfully." in p.stdout
    assert len(open(tmp_path / "data" / "Gflops_MH-SpGEMM.csv").read().split()) == 2


def test_cusparse_cross_check():
    """Third oracle: cuSPARSE SpGEMM driven exactly as the reference drives
    it (inc/cusparse_spgemm.cuh:30-88, through oracle/_ref); its rows are compared as sorted
    (col, val) sets because cuSPARSE does not promise column order."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libmhref.so")):
        pytest.skip("oracle/_ref not built")
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); import mh_spgemm_b200\n"
        "from mh_spgemm_b200 import api, generators as G; from oracle import Reference\n"
        "A = G.rmat(13, 8000, 40000, seed=12)\n"
        "C = api.Tool(0).spgemm_host(A, A); R = Reference().cusparse(A, A)\n"
        "assert R['nnz'] == C.nnz and np.array_equal(R['ptr'], C.ptr)\n"
        "rows = np.repeat(np.arange(A.M), np.diff(C.ptr)); o = np.lexsort((R['col'], rows))\n"
        "assert np.array_equal(R['col'][o], C.col)\n"
        "np.testing.assert_allclose(R['val'][o], C.val, rtol=1e-12, atol=0); print('CUSPARSE-OK', C.nnz)\n" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert "CUSPARSE-OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.parametrize("alg", ["default", "alg1", "alg2", "alg3"])
def test_cusparse_algorithms_cross_check(alg):
    """cuSPARSE SpGEMM through our own harness (oracle/cusparse_check.cu) with the memory-bounded
    algorithms the reference never uses (ALG2 / ALG3 with a chunk fraction).
    Runs in a subprocess: a cuSPARSE failure must not poison this process's CUDA context."""
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); import mh_spgemm_b200\n"
        "from mh_spgemm_b200 import api, generators as G; from oracle import CuSparse\n"
        "A = G.rmat(13, 8000, 40000, seed=12); B = G.fem3d(4, 4, 10, 3, seed=2)\n"
        "t = api.Tool(0); cs = CuSparse()\n"
        "for X in (A, B):\n"
        "    C = t.spgemm_host(X, X); R = cs.spgemm(X, X, alg=%r, chunk_fraction=0.3)\n"
        "    assert R['nnz'] == C.nnz and np.array_equal(R['ptr'], C.ptr) and np.array_equal(R['col'], C.col)\n"
        "    np.testing.assert_allclose(R['val'], C.val, rtol=1e-12, atol=0)\n"
        "print('CUSPARSE-ALG-OK')\n" % (ROOT, alg))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert "CUSPARSE-ALG-OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


def test_global_memory_fallbacks_wide_matrix(tool, orc):
    """Columns beyond the 1.8 M that a shared-memory bitmap covers and rows beyond the largest
    shared-memory tables: the symbolic tile hash and the numeric hash both run from the
    global-memory pool (the reference's 'global row' paths, inc/MH_spgemm.cuh:254-260,376-393)."""
    rng = np.random.default_rng(5)
    K, N = 48, 3_000_000
    cols = np.concatenate([np.sort(rng.choice(N, 20_000, replace=False)) for _ in range(K)])
    B = CSR(K, N, np.arange(K + 1) * 20_000, cols, rng.random(cols.size) + 0.5)
    A = CSR(6, K, np.arange(7) * K, np.tile(np.arange(K), 6), rng.random(6 * K) + 0.5)
    C = tool.spgemm_host(A, B)
    st = tool.stats
    assert st["sym_bins"]["H_GLOBAL"] == 6 and st["num_bins"]["H_GLOBAL"] == 6
    Cp, Cc, Cv = orc.spgemm(A, B)
    assert_matches(orc, C, Cp, Cc, Cv)


@pytest.mark.parametrize("shape", [(0, 5, 5), (5, 0, 5), (4, 4, 1), (1, 1, 1)])
def test_degenerate_shapes(tool, orc, shape):
    M, K, N = shape
    rng = np.random.default_rng(1)
    def rnd(m, n):
        if m == 0 or n == 0:
            return CSR(m, n, np.zeros(m + 1, np.int32), [], [])
        return G.uniform_random(m, n, max(1, m * n // 2), seed=int(rng.integers(1 << 30)))
    A, B = rnd(M, K), rnd(K, N)
    C = tool.spgemm_host(A, B)
    Cp, Cc, Cv = orc.spgemm(A, B)
    assert_matches(orc, C, Cp, Cc, Cv)


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_random_shapes(tool, orc, seed):
    """Seeded fuzz over shapes, densities and row-length skew (rectangular A*B, empty rows and
    columns, a few very long rows), every accumulator path, both value types."""
    rng = np.random.default_rng(1000 + seed)
    M, K, N = (int(rng.integers(1, 400)), int(rng.integers(1, 400)), int(rng.choice([7, 60, 500, 5000, 70000])))
    def rand(m, n, density, long_rows):
        nnz = max(1, int(m * n * density))
        r, c = rng.integers(0, m, nnz), rng.integers(0, n, nnz)
        for _ in range(long_rows):  # a few rows that touch many columns
            rr = int(rng.integers(0, m))
            extra = rng.choice(n, size=min(n, int(rng.integers(1, 600))), replace=False)
            r, c = np.concatenate([r, np.full(extra.size, rr)]), np.concatenate([c, extra])
        return CSR.from_coo(m, n, r, c, rng=rng)
    A = rand(M, K, float(rng.choice([0.002, 0.02, 0.2])), int(rng.integers(0, 3)))
    B = rand(K, N, float(rng.choice([0.0005, 0.01, 0.1])) if N > 500 else 0.2, int(rng.integers(0, 3)))
    if seed % 2:
        A, B = A.astype(np.float32), B.astype(np.float32)
    Cp, Cc, Cv = orc.spgemm(A, B)
    for force in ((0, 0), (1, 1), (2, 2)):
        tool.set_option("force_sym_path", force[0])
        tool.set_option("force_num_path", force[1])
        try:
            C = tool.spgemm_host(A, B)
        finally:
            tool.set_option("force_sym_path", 0)
            tool.set_option("force_num_path", 0)
        assert_matches(orc, C, Cp, Cc, Cv)


@pytest.mark.parametrize("opts", [dict(compact_rows=0), dict(compact_rows=0, row_twins=1), dict(compact_rows=1),
                                  dict(compact_rows=1, sym_twins=0), dict(compact_rows=0, sym_twins=0)])
@pytest.mark.parametrize("alias", [True, False])
def test_window_kernel_variants_on_twin_rows(orc, opts, alias):
    """Multi-dof FEM input (twin rows in A and B): the dense-window kernel with B-twin folding,
    the dense A-row-twin kernel and the rank-mapped compact kernel must agree with the oracle,
    with B aliasing A (twin flags shared) and with B a separate copy (flags recomputed)."""
    t = api.Tool(0)
    for k, v in opts.items():
        t.set_option(k, v)
    A = G.fem3d(5, 4, 14, 3, seed=31)
    B = A if alias else CSR(A.M, A.N, A.ptr.copy(), A.col.copy(), A.val.copy() * 0.5 + 0.25)
    C = t.spgemm_host(A, B)
    Cp, Cc, Cv = orc.spgemm(A, B)
    assert_matches(orc, C, Cp, Cc, Cv)
    nb = t.stats["num_bins"]
    assert (nb["WIN_COMPACT"] > 0) == bool(opts.get("compact_rows", 1))
    # rows whose twins are broken up (every third row emptied) still come out right
    keep = np.ones(A.M, bool)
    keep[1::3] = False
    rows = np.repeat(np.arange(A.M), np.diff(A.ptr))
    A2 = CSR.from_coo(A.M, A.N, rows[keep[rows]], A.col[keep[rows]], A.val[keep[rows]])
    C2 = t.spgemm_host(A2, B)
    Cp2, Cc2, Cv2 = orc.spgemm(A2, B)
    assert_matches(orc, C2, Cp2, Cc2, Cv2)
    t.release()


@pytest.mark.parametrize("name", list(cases.TRANSPOSE))
def test_device_transpose_matches_reference(tool, orc, name):
    """mhb_transpose_* (stable radix sort by column) against the vectors recorded from the
    reference's host transpose, bit-for-bit including the values; fp32 too."""
    A = cases.TRANSPOSE[name]()
    ref = np.load(os.path.join(GOLD, f"ref_transpose_{name}.npz"))
    T = tool.transpose(A)
    assert (T.M, T.N) == (A.N, A.M)
    assert np.array_equal(T.ptr, ref["ptr"]) and np.array_equal(T.col, ref["col"]) and np.array_equal(T.val, ref["val"])
    assert tool.stats["gpu_launches"] > 0
    T32 = tool.transpose(A.astype(np.float32))
    assert T32.val.dtype == np.float32 and np.array_equal(T32.col, ref["col"])
    assert np.array_equal(T32.val, ref["val"].astype(np.float32))


def test_device_transpose_edges_and_involution(tool, orc):
    """Empty matrix, empty rows / columns, one long column (a hub), and (A^T)^T == A on a
    skewed 260 K-nonzero input (4096-item radix tiles: many blocks, three passes)."""
    E = CSR(4, 6, np.zeros(5, np.int32), [], [])
    T = tool.transpose(E)
    assert (T.M, T.N, T.nnz) == (6, 4, 0) and np.array_equal(T.ptr, np.zeros(7, np.int32))
    A = G.with_dense_rows(G.rmat(17, 120_000, 260_000, seed=44), 3, 30_000, seed=45)
    T = tool.transpose(A)
    O = orc.transpose(A)
    assert np.array_equal(T.ptr, O.ptr) and np.array_equal(T.col, O.col) and np.array_equal(T.val, O.val)
    TT = tool.transpose(T)
    assert np.array_equal(TT.ptr, A.ptr) and np.array_equal(TT.col, A.col) and np.array_equal(TT.val, A.val)


@pytest.mark.parametrize("name", ["rect_300x200", "wide_40x70000", "fem_3x3x6x2"])
def test_aat_product_full(tool, orc, name):
    """The reference's second standard workload, C = A * A^T (AAT, inc/common.h:37): device
    transpose feeding the SpGEMM, structure bit-exact and values to 1e-12 against the oracle;
    C is symmetric in structure."""
    A = cases.TRANSPOSE[name]()
    C = tool.spgemm_host(A, tool.transpose(A))
    Cp, Cc, Cv = orc.spgemm(A, orc.transpose(A))
    assert_matches(orc, C, Cp, Cc, Cv)
    P = C.to_scipy()
    P.data[:] = 1
    assert (P != P.T).nnz == 0


def test_hash_probe_counter(orc):
    """Option count_probes == the reference's HASH_CONFLICT statistic (inc/common.h:18): zero
    when off, zero for a dense-path input, positive for hashed rows, and the product is
    unchanged by counting."""
    t = api.Tool(0)
    A = G.rmat(14, 16000, 60000, seed=6)
    Cp, Cc, Cv = orc.spgemm(A, A)
    C0 = t.spgemm_host(A, A)
    assert t.stats["hash_probes"] == 0 and t.stats["sym_hash_probes"] == 0
    t.set_option("count_probes", 1)
    t.set_option("force_sym_path", 2)
    t.set_option("force_num_path", 2)
    C1 = t.spgemm_host(A, A)
    st = t.stats
    assert_matches(orc, C1, Cp, Cc, Cv)
    assert np.array_equal(C0.col, C1.col)
    assert st["hash_probes"] > 0 and st["sym_hash_probes"] > 0
    assert st["hash_probes"] < 4 * st["intprod"]  # fill <= 5/8 with linear probing: a few probes per product
    t.spgemm_host(A, A)  # the counter restarts with every call (CAS order may move it a little)
    assert 0 < t.stats["hash_probes"] < 2 * st["hash_probes"]
    t.set_option("force_sym_path", 1)
    t.set_option("force_num_path", 1)
    F = G.fem3d(4, 4, 10, 3, seed=5)
    t.spgemm_host(F, F)
    assert t.stats["hash_probes"] == 0 and t.stats["sym_hash_probes"] == 0
    t.release()


def test_shape_and_dtype_mismatch_are_rejected(tool):
    A = G.uniform_random(50, 60, 300, seed=1)
    B = G.uniform_random(50, 60, 300, seed=2)
    with pytest.raises(ValueError):
        tool.spgemm_host(A, B)  # A.N != B.M
    with pytest.raises(TypeError):
        tool.spgemm_host(A, A.transpose().astype(np.float32))


@pytest.mark.parametrize("name", cases.SUITE12)
def test_suite_analog_matches_reference(tool, name):
    """BASELINE configs[3]: C = A*A on the synthetic analog of every 16matrix.txt shape that
    fits a test run; row_ptr / col_idx must be bit-exact (SHA-256) with what the reference's
    own kernels produced on a B200 for the same input (tests/golden/ref_suite_*.json, written
    by make_golden.py), values by checksum, rows sorted and duplicate-free."""
    meta = json.load(open(os.path.join(GOLD, f"ref_suite_{name}.json")))
    A, _ = cases.SUITE[name]()
    assert (A.M, A.nnz) == (meta["M"], meta["nnzA"])
    C = tool.spgemm_host(A, A)
    assert C.nnz == meta["nnz"]
    assert sha(C.ptr, np.int32) == meta["sha_ptr"]
    assert sha(C.col, np.int32) == meta["sha_col"]
    assert float(C.val.sum()) == pytest.approx(meta["sum_val"], rel=1e-10)
    w = (np.arange(C.val.size, dtype=np.int64) % 97 + 1).astype(np.float64)
    assert float((C.val * w).sum()) == pytest.approx(meta["sum_weighted"], rel=1e-10)
    assert C.is_canonical()


@pytest.mark.parametrize("name", cases.SUITE_LARGE)
def test_suite_large_analog(tool, name):
    """The four large shapes (16-24 M rows / up to 1.75 G nnz(C)): minutes of host-side
    generation and tens of GB of pinned memory, so opt-in (MHB_SLOW=1)."""
    if os.environ.get("MHB_SLOW") != "1":
        pytest.skip("set MHB_SLOW=1 to run the four large suite analogs")
    A, _ = cases.SUITE[name]()
    C = tool.spgemm_host(A, A, copy=False)
    path = os.path.join(GOLD, f"ref_suite_{name}.json")
    if os.path.exists(path):  # the reference faults on two of them
        meta = json.load(open(path))
        assert C.nnz == meta["nnz"] and sha(C.ptr, np.int32) == meta["sha_ptr"] and sha(C.col, np.int32) == meta["sha_col"]
    S = A.to_scipy()
    want = S @ (S @ np.ones(A.N))
    got = np.add.reduceat(np.append(C.val, 0.0), np.minimum(C.ptr[:-1], C.nnz))
    got[np.diff(C.ptr) == 0] = 0.0
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-9)


def _free_port():
    import socket
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        return so.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_spgemm_multi_rank(world):
    """The multi-GPU path behind the C ABI (mhb_shard_*), one PROCESS per rank: B row-sharded,
    peer-mapped windows over CUDA IPC, one-sided halo pull and slice sizes, three steps with
    changing values, the sliced (int32-overflow) form -- every rank's slice of C bit-exact on
    structure and to 1e-12 on values against the host oracle, slice offsets equal to the
    oracle's global row_ptr.  Ranks share a device when the box has fewer GPUs than ranks."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "shard_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    # the ranks print concurrently: count the tokens, not the lines
    assert p.returncode == 0 and p.stdout.count("SHARD-OK") == 5 * world, p.stdout[-3000:] + p.stderr[-3000:]


def test_speculative_symbolic_launch_and_miss(orc):
    """A call with the shape of the previous one launches its symbolic kernels from that call's
    bin sizes (no mid-pipeline host read).  Same pattern again: speculative, no miss.  Same
    SHAPE (M, K, N, nnz) but a pattern that lands in bins the previous call never populated:
    the guess must be detected as a miss and the result must still be exact."""
    t = api.Tool(0)
    A1 = G.poisson2d(64)  # every row in the tiny bins
    assert (A1.M, A1.nnz) == (4096, 20224)
    Cp, Cc, Cv = orc.spgemm(A1, A1)
    assert_matches(orc, t.spgemm_host(A1, A1), Cp, Cc, Cv)
    assert t.stats["speculative_launches"] == 0
    assert_matches(orc, t.spgemm_host(A1, A1), Cp, Cc, Cv)
    st = t.stats
    assert st["speculative_launches"] == 1 and st["speculative_misses"] == 0
    # 100 rows of 200 nonzeros that all select each other (20 000 products per C row) + 224 singletons
    rng = np.random.default_rng(7)
    rows, cols = [], []
    for r in range(100):
        c = np.concatenate([np.arange(100), 100 + np.sort(rng.choice(3996, 100, replace=False))])
        rows.append(np.full(200, r)), cols.append(c)
    rows.append(np.arange(100, 324)), cols.append(rng.integers(0, 4096, 224))
    A2 = CSR.from_coo(4096, 4096, np.concatenate(rows), np.concatenate(cols), rng=rng)
    assert (A2.M, A2.nnz) == (A1.M, A1.nnz)
    C2 = t.spgemm_host(A2, A2)
    Cp2, Cc2, Cv2 = orc.spgemm(A2, A2)
    assert_matches(orc, C2, Cp2, Cc2, Cv2)
    st = t.stats
    assert st["speculative_launches"] == 2 and st["speculative_misses"] == 1
    # and back: the big-row bins of A2 are launched for A1 and find nothing to do
    assert_matches(orc, t.spgemm_host(A1, A1), Cp, Cc, Cv)
    assert t.stats["speculative_misses"] == 1
    t.set_option("speculate", 0)
    assert_matches(orc, t.spgemm_host(A2, A2), Cp2, Cc2, Cv2)
    assert t.stats["speculative_launches"] == 3
    t.release()


def _into(t, A, B, dC, cap_arrays):
    """mhb_spgemm_into_* on device copies of A, B; returns (nnz, C as host CSR) using dC = (ptr, col, val)."""
    dA = [api.DeviceArray(x) for x in (A.ptr, A.col, A.val)]
    dB = dA if B is A else [api.DeviceArray(x) for x in (B.ptr, B.col, B.val)]
    nnz = t.spgemm_into(A.M, A.N, B.N, dA[0], dA[1], dA[2], dB[0], dB[1], dB[2], dC[0], dC[1], dC[2])
    return nnz, CSR(A.M, B.N, dC[0].numpy()[:A.M + 1], dC[1].numpy()[:nnz], dC[2].numpy()[:nnz])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["fem", "rmat14", "poisson32", "dense_rows", "rect"])
def test_fused_call_into_caller_buffers(orc, name, dtype):
    """mhb_spgemm_into_*: the first call on a handle runs the ordinary two-read path, a second call of
    the same shape runs with ONE host synchronisation (both phases launched from the previous call's
    bin sizes, verified by the device-side gate) -- same CSR either way, also with new values."""
    A, B = INPUTS[name]()
    A = CSR(A.M, A.N, A.ptr, A.col, A.val.astype(dtype))
    B = A if B is None else CSR(B.M, B.N, B.ptr, B.col, B.val.astype(dtype))
    Cp, Cc, Cv = orc.spgemm(A, B)
    t = api.Tool(0)
    cap = int(Cp[-1]) + 7
    dC = (api.DeviceArray(np.zeros(A.M + 1, np.int32)), api.DeviceArray(np.zeros(cap, np.int32)),
          api.DeviceArray(np.zeros(cap, dtype)))
    nnz, C = _into(t, A, B, dC, cap)
    assert nnz == Cp[-1] and t.stats["fused_calls"] == 0
    assert_matches(orc, C, Cp, Cc, Cv)
    A2 = CSR(A.M, A.N, A.ptr, A.col, (A.val * 2 - 1).astype(dtype))
    B2 = A2 if B is A else B
    nnz, C = _into(t, A2, B2, dC, cap)
    st = t.stats
    assert st["fused_calls"] == 1 and st["speculative_misses"] == 0
    Cp2, Cc2, Cv2 = orc.spgemm(A2, B2)
    assert_matches(orc, C, Cp2, Cc2, Cv2)
    assert st["intprod"] == orc.intprod(A, B) and st["gpu_launches"] > 0 and t.timing.Numeric > 0
    # the same call in its two halves (begin queues, end waits and verifies)
    dA = [api.DeviceArray(x) for x in (A.ptr, A.col, A.val)]
    dB = dA if B is A else [api.DeviceArray(x) for x in (B.ptr, B.col, B.val)]
    t.spgemm_into_begin(A.M, A.N, B.N, dA[0], dA[1], dA[2], dB[0], dB[1], dB[2], dC[0], dC[1], dC[2])
    assert t.spgemm_into_end() == Cp[-1] and t.stats["fused_calls"] == 2
    assert_matches(orc, CSR(A.M, B.N, dC[0].numpy()[:A.M + 1], dC[1].numpy()[:nnz], dC[2].numpy()[:nnz]), Cp, Cc, Cv)
    with pytest.raises(api.MhbError):
        t.spgemm_into_end()  # no begin outstanding
    # pattern reuse after a fused call
    dA2v = api.DeviceArray(A.val)
    dBv = dA2v if B is A else api.DeviceArray(B.val)
    t.numeric_into(dA2v, dBv, dC[1], dC[2])
    np.testing.assert_allclose(dC[2].numpy()[:nnz], Cv, rtol=RTOL[np.dtype(dtype)] * 10, atol=0)
    t.release()


def test_fused_call_capacity_and_miss(orc):
    """Too-small C arrays: MHB_ERR_CAPACITY with row_ptr and nnz valid and nothing written (ordinary and
    fused path).  Same shape, different pattern: the device-side gate stops the numeric kernels and the
    call is redone the ordinary way."""
    t = api.Tool(0)
    A1 = G.poisson2d(64)
    Cp, Cc, Cv = orc.spgemm(A1, A1)
    n1 = int(Cp[-1])
    small = (api.DeviceArray(np.zeros(A1.M + 1, np.int32)), api.DeviceArray(np.full(100, -5, np.int32)),
             api.DeviceArray(np.full(100, -5.0)))
    for k in range(2):  # k = 0: ordinary path, k = 1: fused path with the capacity bit of the gate
        with pytest.raises(api.MhbError) as ei:
            _into(t, A1, A1, small, 100)
        assert ei.value.code == api.ERR_CAPACITY and ei.value.nnzC == n1
        assert np.array_equal(small[0].numpy().astype(np.int64), Cp)
        assert (small[1].numpy() == -5).all() and (small[2].numpy() == -5.0).all()
    assert t.stats["fused_calls"] == 1
    # 100 rows of 200 nonzeros that select each other + 224 singletons: same (M, K, N, nnz) as A1
    rng = np.random.default_rng(7)
    rows, cols = [], []
    for r in range(100):
        c = np.concatenate([np.arange(100), 100 + np.sort(rng.choice(3996, 100, replace=False))])
        rows.append(np.full(200, r)), cols.append(c)
    rows.append(np.arange(100, 324)), cols.append(rng.integers(0, 4096, 224))
    A2 = CSR.from_coo(4096, 4096, np.concatenate(rows), np.concatenate(cols), rng=rng)
    assert (A2.M, A2.nnz) == (A1.M, A1.nnz)
    Cp2, Cc2, Cv2 = orc.spgemm(A2, A2)
    cap = max(n1, int(Cp2[-1]))
    dC = (api.DeviceArray(np.zeros(A1.M + 1, np.int32)), api.DeviceArray(np.zeros(cap, np.int32)),
          api.DeviceArray(np.zeros(cap)))
    nnz, C = _into(t, A1, A1, dC, cap)
    assert_matches(orc, C, Cp, Cc, Cv)
    misses = t.stats["speculative_misses"]
    nnz, C = _into(t, A2, A2, dC, cap)  # same shape, bins never populated before: gate -> redo
    assert_matches(orc, C, Cp2, Cc2, Cv2)
    assert t.stats["speculative_misses"] == misses + 1
    nnz, C = _into(t, A2, A2, dC, cap)  # now fused
    assert_matches(orc, C, Cp2, Cc2, Cv2)
    assert t.stats["speculative_misses"] == misses + 1
    nnz, C = _into(t, A1, A1, dC, cap)  # back: A2's big-row kernels find nothing to do
    assert_matches(orc, C, Cp, Cc, Cv)
    t.release()


@pytest.mark.parametrize("chunks", [2, 3, 8])
@pytest.mark.parametrize("name", ["fem", "rmat14", "dense_rows", "poisson32", "rect", "ragged", "fem_perturbed"])
def test_host_path_row_chunks(orc, name, chunks):
    """mhb_spgemm_host_* with the numeric phase cut into row chunks whose download overlaps the next
    chunk's computation (forced on for small inputs): same CSR as the single-piece path."""
    A, B = INPUTS[name]()
    B = A if B is None else B
    t = api.Tool(0)
    t.set_option("row_chunks", chunks)
    t.set_option("row_chunk_bytes", 1)
    Cp, Cc, Cv = orc.spgemm(A, B)
    for _ in range(2):  # second call: speculative symbolic in front of the chunked numeric
        assert_matches(orc, t.spgemm_host(A, B), Cp, Cc, Cv)
    t.release()


def test_one_outstanding_begin_per_handle(orc):
    """A second call on a handle whose mhb_spgemm_into_begin has not been ended is refused instead of
    trampling the workspace of the queued one; end still delivers the first result."""
    A = G.fem3d(4, 4, 10, 3, seed=5)
    Cp, Cc, Cv = orc.spgemm(A, A)
    t = api.Tool(0)
    dA = [api.DeviceArray(x) for x in (A.ptr, A.col, A.val)]
    dC = (api.DeviceArray(count=A.M + 1, dtype=np.int32), api.DeviceArray(count=int(Cp[-1]), dtype=np.int32),
          api.DeviceArray(count=int(Cp[-1]), dtype=np.float64))
    args = (A.M, A.N, A.N, dA[0], dA[1], dA[2], dA[0], dA[1], dA[2], dC[0], dC[1], dC[2])
    assert t.spgemm_into(*args) == Cp[-1]          # first call of the shape: ordinary path
    t.spgemm_into_begin(*args)                      # steady state: queued, not waited for
    with pytest.raises(api.MhbError):
        t.spgemm_into_begin(*args)
    with pytest.raises(api.MhbError):
        t.symbolic(A.M, A.N, A.N, dA[0], dA[1], dA[0], dA[1])
    assert t.spgemm_into_end() == Cp[-1]
    assert_matches(orc, CSR(A.M, A.N, dC[0].numpy(), dC[1].numpy(), dC[2].numpy()), Cp, Cc, Cv)
    t.release()


@pytest.mark.parametrize("name", ["fem", "rmat14", "dense_rows", "poisson32", "road", "ragged", "empty", "one_row"])
def test_mask_builder_modes_agree(orc, name):
    """The three builders of B's mask matrix -- round 1's five-kernel chain (0), the one-pass chained scan
    (1) and the two passes around an ordinary scan (2) -- give the same tileptr / tilecol / tilemask, and the
    SpGEMM on top of each is the oracle's."""
    A, B = INPUTS[name]()
    B = A if B is None else B
    Cp, Cc, Cv = orc.spgemm(A, B)
    got = {}
    for mode in (0, 1, 2):
        t = api.Tool(0)
        t.set_option("mask_onepass", mode)
        dBp, dBc = api.DeviceArray(B.ptr), api.DeviceArray(B.col)
        got[mode] = t.mask_matrix_B(B.M, B.N, dBp, dBc)
        assert_matches(orc, t.spgemm_host(A, B), Cp, Cc, Cv)
        t.release()
    for mode in (0, 2):
        for a, b in zip(got[1], got[mode]):
            assert np.array_equal(np.asarray(a), np.asarray(b)), f"mask builder mode {mode} differs from mode 1"


def test_sharded_spgemm_binds_the_callers_stream(orc):
    """ShardedSpGEMM (the torch.distributed layer) on torch's DEFAULT stream without any pre-binding by the
    caller: B's image is produced by torch ops on that stream right before the step, the C arrays come
    from torch's allocator, and the handle must run on the same stream (legacy default stream = handle 0,
    which Tool.set_stream maps to cudaStreamLegacy).  ADVICE r1: the class used to leave the handle on
    its own stream, unordered against the exchange."""
    import torch
    from mh_spgemm_b200.distributed import ShardedSpGEMM, b_views, pack_b
    A = G.fem3d(4, 4, 16, 3, seed=61)
    Cp, Cc, Cv = orc.spgemm(A, A)
    dev = torch.device("cuda", 0)
    t = api.Tool(0)
    sh = ShardedSpGEMM(t, 0, 1, dev)
    packed, _ = pack_b(A)
    ap, ac, av = (torch.from_numpy(x).to(dev) for x in (A.ptr, A.col, A.val))
    for step in range(3):
        # the "exchange": B's image is rewritten on the default stream just before the multiply (scaled values)
        Bbuf = torch.zeros_like(packed, device=dev)
        Bbuf.copy_(packed.to(dev))
        bp, bc, bv = b_views(Bbuf, A.M, A.nnz, torch.float64)
        bv.mul_(float(step + 1))
        cp, cc, cv, off, tot = sh.step((A.M, ap, ac, av), Bbuf, A.M, A.N, A.nnz, torch.float64)
        assert off == 0 and tot == Cp[-1]
        got = CSR(A.M, A.N, cp.cpu().numpy(), cc.cpu().numpy(), cv.cpu().numpy())
        assert_matches(orc, got, Cp, Cc, Cv * float(step + 1))
    t.release()


def test_cpp_shard_driver_two_ranks():
    """Two forked C++ processes drive mhb_shard_* through the C ABI alone (peer windows, fused call in two
    halves, device-side size post) and check their slices against a host Gustavson inside the driver."""
    from importlib import import_module
    exe = import_module("mh_spgemm_b200.build").build_shard_driver()
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.count("SHARD-CPP-OK") == 2, p.stdout[-2000:] + p.stderr[-2000:]
