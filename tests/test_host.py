"""CPU tests of the host side: the C-ABI library loads and exports what include/gapitcuda.h
declares, fails loudly without a GPU, and its host-only entry point (REML) matches the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from gapit_b200 import _lib, shard, synth
from oracle import gapit_oracle as go

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "gapitcuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gapit_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/gapitcuda.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.SIGNATURES) == names


def test_no_cpu_fallback():
    lib = _lib.load()
    if lib.gapit_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.gapit_ctx_create(0, C.byref(h))
    assert rc == -3                                   # GAPIT_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.gapit_last_error()
    import gapit_b200 as gb
    with pytest.raises(gb.GapitCudaError):
        gb.Context(0)
    with pytest.raises(gb.GapitCudaError):
        gb.GAPIT_kinship_VanRaden(np.zeros((4, 3)), verbose=False)


def test_product_path_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/: not the package, not the tools,
    not the R shim."""
    for top in ("gapit_b200", "tools", "r", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".R")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in txt and "from oracle" not in txt, f
    # bench.py: the oracle appears only inside the two CPU legs and the in-run parity check (as the checker)
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    users = set()
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for n in ast.walk(fn):
            if isinstance(n, ast.ImportFrom) and n.module and n.module.startswith("oracle"):
                users.add(fn.name)
    assert users <= {"cpu_scan_sample", "run_reference", "parity_report"}, users


def test_reml_host_code_matches_oracle():
    import gapit_b200 as gb
    for seed, npc in [(3, 0), (4, 2), (9, 1)]:
        G = synth.make_genotypes(90, 300, seed=seed, threads=1)
        Y = synth.make_phenotypes(G, seed=seed)[:, 0]
        X0 = synth.make_covariates(G, npc=npc)
        K = go.kinship_vanraden(G.T)
        eR = go.emma_eigen_R_wo_Z(K, X0)
        ref = go.emma_REMLE(Y, X0, K, eig_R=eR)
        got = gb.GAPIT_emma_REMLE(Y, X0, K, eig_R=eR)
        for k in ("REML", "delta", "ve", "vg"):
            assert got[k] == pytest.approx(ref[k], rel=1e-10), k


def test_reml_from_eigL_alone_matches_the_eigR_route():
    """gapit_reml_eigL evaluates emma.delta.REML.{LL,dLL}.wo.Z (emma.txt:145-164) from eig.L only; it must land on
    the optimum the reference's eig.R route (GAPIT.emma.REMLE.R:21-54) finds."""
    import gapit_b200 as gb
    for seed, n, m, npc in [(3, 90, 300, 0), (4, 120, 500, 2), (9, 200, 900, 5), (11, 64, 200, 1)]:
        G = synth.make_genotypes(n, m, seed=seed, threads=1)
        Y = synth.make_phenotypes(G, seed=seed)[:, 0]
        X0 = synth.make_covariates(G, npc=npc)
        K = go.kinship_vanraden(G.T)
        ref = go.emma_REMLE(Y, X0, K)
        got = gb.GAPIT_emma_REMLE(Y, X0, K, eig_L=go.emma_eigen_L_wo_Z(K))
        for k in ("REML", "delta", "ve", "vg"):
            assert got[k] == pytest.approx(ref[k], rel=1e-10), (k, seed)
    # many traits on the host's threads: identical to the one-trait calls
    Yt = synth.make_phenotypes(G, ntraits=7, seed=5)
    eL = go.emma_eigen_L_wo_Z(K)
    V = eL["vectors"]
    P = V.T @ Yt
    multi = gb.reml_from_projections_multi(eL["values"], P, V.T @ X0)
    for t in range(7):
        one = gb.reml_from_projections(eL["values"], P[:, t], V.T @ X0)
        assert multi[t] == one
    # collinear covariates: X'X singular is an error, not a crash
    lib = _lib.load()
    lam = np.linspace(2.0, 0.0, 10)
    X = np.ones((10, 2))
    out = _lib.RemlOut()
    rc = lib.gapit_reml_eigL(lam.ctypes.data_as(C.c_void_p), lam.ctypes.data_as(C.c_void_p),
                             np.asfortranarray(X).ctypes.data_as(C.c_void_p), 10, 2, 100, -10.0, 10.0, 1e-10, C.byref(out))
    assert rc == _lib_err("GAPIT_ERR_NUMERIC")


def _lib_err(name):
    return {"GAPIT_ERR_ARG": -1, "GAPIT_ERR_CUDA": -2, "GAPIT_ERR_NO_DEVICE": -3, "GAPIT_ERR_DATA": -4,
            "GAPIT_ERR_NUMERIC": -5}[name]


def test_reml_boundary_optimum():
    # pure-noise phenotype on a near-identity kinship: the optimum sits on the grid boundary (:37-44)
    lib = _lib.load()
    rng = np.random.default_rng(0)
    nq = 60
    lam = np.full(nq, 1e-3) + rng.uniform(0, 1e-4, nq)
    etas = rng.standard_normal(nq)
    out = _lib.RemlOut()
    rc = lib.gapit_reml(lam.ctypes.data_as(C.c_void_p), etas.ctypes.data_as(C.c_void_p), nq, 100, -10.0, 10.0, 1e-10,
                        C.byref(out))
    assert rc == 0
    ref = go.emma_REMLE(etas, np.zeros((nq, 0)), None, eig_R={"values": lam, "vectors": np.eye(nq)}) if False else None
    assert out.delta > 0 and np.isfinite(out.REML)
    # bad arguments are reported, not crashed on
    assert lib.gapit_reml(None, None, 0, 100, -10.0, 10.0, 1e-10, C.byref(out)) == -1


def test_packed2_host_layout():
    """Packed2.from_dosages: taxon i of marker k in byte i//4 of row k at bits 2(i%4).., NA = 3."""
    import gapit_b200 as gb
    rng = np.random.default_rng(5)
    G = rng.integers(0, 3, size=(7, 13)).astype(np.int8)      # (m, n)
    G[2, 5] = -1
    P = gb.Packed2.from_dosages(G, marker_major=True)
    assert P.data.shape == (7, 4) and P.n == 13 and P.m == 7
    for k in range(7):
        for i in range(13):
            code = (int(P.data[k, i // 4]) >> (2 * (i % 4))) & 3
            assert code == (3 if G[k, i] < 0 else G[k, i])
    assert (int(P.data[0, 3]) >> 2) == 0                       # bits past the last taxon are zero
    Pf = gb.Packed2.from_dosages(np.where(G.T < 0, np.nan, G.T.astype(float)))
    assert np.array_equal(Pf.data, P.data)
    Pi = gb.Packed2.from_dosages(np.where(G < 0, np.iinfo(np.int32).min, G.astype(np.int32)), marker_major=True)
    assert np.array_equal(Pi.data, P.data)
    # the threaded packer against a numpy restatement on a bigger ragged shape, every thread count
    G2 = rng.integers(-1, 3, size=(1001, 517)).astype(np.int8)
    pad = np.zeros((1001, 520), dtype=np.uint8)
    pad[:, :517] = np.where(G2 < 0, 3, G2)
    q = pad.reshape(1001, 130, 4)
    ref = (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8)
    for threads in (1, 3, 0):
        assert np.array_equal(gb.Packed2.from_dosages(G2, marker_major=True, threads=threads).data, ref)
    bad = G2.astype(np.float64)
    bad[5, 7] = 2.5
    with pytest.raises(gb.GapitCudaError):
        gb.Packed2.from_dosages(bad, marker_major=True)


def test_hapmap_index_host_code():
    """gapit_hapmap_index: row offsets, n, m, genotype width; CRLF and missing heading; malformed rows are errors."""
    lib = _lib.load()
    G = synth.make_genotypes(7, 30, seed=2, threads=1)
    for bit, crlf, heading in [(2, False, True), (1, True, True), (2, True, False)]:
        text = synth.make_hapmap_text(G, bit=bit, crlf=crlf, heading=heading)
        buf = np.frombuffer(text, dtype=np.uint8)
        n, m, b, h = C.c_int64(), C.c_int64(), C.c_int(), C.c_int64()
        off = np.empty(30, dtype=np.int64)
        rc = lib.gapit_hapmap_index(buf.ctypes.data_as(C.c_void_p), len(text), int(heading), C.byref(n), C.byref(m),
                                    C.byref(b), off.ctypes.data_as(C.c_void_p), 30, C.byref(h))
        assert rc == 0 and (n.value, m.value, b.value) == (7, 30, bit)
        rows = synth.hapmap_rows(text)[1 if heading else 0:]
        for k in (0, 13, 29):
            got = text[off[k]: off[k] + 7 * (bit + 1) - 1].decode().split("\t")
            assert got == rows[k][11:]
        assert (h.value >= 0) == heading
        if heading:
            assert text[h.value:h.value + 9] == b"taxa00000"
    bad = text.replace(b"\t", b" ", 3)   # fewer than 12 tab-separated columns in the first row
    buf = np.frombuffer(bad, dtype=np.uint8)
    assert lib.gapit_hapmap_index(buf.ctypes.data_as(C.c_void_p), len(bad), 0, C.byref(n), C.byref(m), C.byref(b), None, 0,
                                  None) == _lib_err("GAPIT_ERR_DATA")
    ragged = text[:-6]                   # last row one genotype short
    buf = np.frombuffer(ragged, dtype=np.uint8)
    assert lib.gapit_hapmap_index(buf.ctypes.data_as(C.c_void_p), len(ragged), 0, C.byref(n), C.byref(m), C.byref(b), None,
                                  0, None) == _lib_err("GAPIT_ERR_DATA")


def test_r_shim_type_checks_against_the_header():
    """r/gapitcuda_R.c cannot be built here (no R), but it must stay in step with include/gapitcuda.h: compile it for
    syntax and types against a minimal stand-in for R's headers (tests/rstub)."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "rstub"),
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "r", "gapitcuda_R.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # every .Call target the R replacement bodies use is defined by the shim
    import re
    shim = open(os.path.join(ROOT, "r", "gapitcuda_R.c")).read()
    defined = set(re.findall(r"^SEXP (gapit_R_\w+)\(", shim, flags=re.M))
    for fn in os.listdir(os.path.join(ROOT, "r")):
        if fn.endswith(".R"):
            for call in re.findall(r'\.Call\("(gapit_R_\w+)"', open(os.path.join(ROOT, "r", fn)).read()):
                assert call in defined, (fn, call)


def test_gram_owned_rows_partition_the_block_rows():
    """gapit_gram_owned_rows (reduce-scatter form of the fused kinship exchange): the ranks' row ranges are disjoint,
    ordered, 256-aligned and cover the Gram buffer's block-rows exactly once."""
    lib = _lib.load()
    for n in (257, 700, 5000, 20000, 50000):
        T = (n + 255) // 256
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                lo, hi = C.c_int64(), C.c_int64()
                assert lib.gapit_gram_owned_rows(n, world, r, C.byref(lo), C.byref(hi)) == 0
                assert lo.value % 256 == 0 and lo.value <= hi.value <= lib.gapit_gram_ld(n)
                covered += list(range(lo.value // 256, (hi.value + 255) // 256))
            assert covered == list(range(T)) or (world > T and sorted(set(covered)) == list(range(T))), (n, world)
    lo, hi = C.c_int64(), C.c_int64()
    assert lib.gapit_gram_owned_rows(700, 2, 2, C.byref(lo), C.byref(hi)) == -1     # rank out of range


def test_marker_shards_cover_in_order():
    for m, w in [(10, 3), (500000, 8), (7, 8), (3093, 2)]:
        parts = [shard.marker_shard(m, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == m
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1


def test_synth_is_deterministic_and_thread_independent():
    a = synth.make_genotypes(64, 9000, seed=5, threads=1)
    b = synth.make_genotypes(64, 9000, seed=5, threads=4)
    assert np.array_equal(a, b)
    assert set(np.unique(a)) <= {0, 1, 2}
    c = a.astype(np.int64).sum(1)
    assert ((c == 0) | (c == 128)).sum() >= 1          # monomorphic columns are injected
    assert (np.all(a[1:] == a[:-1], axis=1)).sum() >= 1  # so are duplicated adjacent columns


# ---- compression glue at GN = n / GN = 1 -------------------------------------------
def _compress_case(n, seed, n_pheno):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(n, 3 * n))
    K = A @ A.T / (3 * n)
    names = [f"L{rng.integers(10**6):06d}" for _ in range(n)]
    assert len(set(names)) == n
    zcols = [names[i] for i in rng.permutation(n)[:n_pheno]]          # phenotyped taxa, in a shuffled order
    zrows = list(zcols)
    Zm = np.eye(n_pheno)
    return names, K, zrows, zcols, Zm


@pytest.mark.parametrize("n,n_pheno,seed", [(17, 17, 1), (23, 15, 2), (40, 40, 3)])
def test_compress_glue_short_circuit_equals_the_clustered_route(n, n_pheno, seed):
    """GN = n: dist + hclust + cutree(k = n) is the identity grouping (numbered by first appearance = row order) and
    KG = KI; GN = 1: one group, KG = mean(KI).  The mirror must reproduce the oracle's restatement of
    GAPIT.Compress / GAPIT.Block / GAPIT.ZmatrixCompress without clustering or the dense Z %*% DS."""
    import gapit_b200.api as api
    from oracle import gapit_oracle as go
    names, K, zrows, zcols, Zm = _compress_case(n, seed, n_pheno)
    for GN in (n, 1):
        GA_o, KG_o = go.compress(names, K, GN=GN)
        GA, KG = api.GAPIT_Compress(names, K, GN=GN)
        assert GA == GA_o
        assert KG.shape == KG_o.shape and np.allclose(KG, KG_o, rtol=1e-14, atol=0)
        if GN == n:
            assert np.array_equal(KG, K)                              # a mean of one element is that element
        GAU_o, KW_o, KO_o, KWO_o = go.block(zcols, GA_o, KG_o)
        GAU, KW, KO, KWO = api.GAPIT_Block(zcols, GA, KG)
        assert GAU == GAU_o
        for a, b in ((KW, KW_o), (KO, KO_o), (KWO, KWO_o)):
            assert a.shape == b.shape and np.allclose(a, b, rtol=1e-14, atol=0)
        rt_o, Z_o = go.zmatrix_compress(zrows, zcols, Zm, GAU_o)
        rt, Z = api.GAPIT_ZmatrixCompress(zrows, zcols, Zm, GAU)
        assert rt == rt_o and np.array_equal(Z, Z_o)
        if GN == n:
            # the incidence matrix is a permutation and K = KW is the kinship of the phenotyped taxa in GA order:
            # Z KW Z' is the kinship of the sorted phenotyped taxa, i.e. nothing was compressed
            idx = {t: i for i, t in enumerate(names)}
            want = K[np.ix_([idx[t] for t in rt], [idx[t] for t in rt])]
            assert np.allclose(Z @ KW @ Z.T, want, rtol=1e-14, atol=0)
        else:
            assert Z.shape == (n_pheno, 1) and np.all(Z == 1.0)


def test_compress_glue_other_group_counts_stay_with_the_reference():
    import gapit_b200.api as api
    names, K, *_ = _compress_case(12, 5, 12)
    with pytest.raises(NotImplementedError):
        api.GAPIT_Compress(names, K, GN=4)


def test_compress_oracle_general_grouping_properties():
    """The restated clustered route at 1 < GN < n: cutree numbering by first appearance, KG = block means."""
    from oracle import gapit_oracle as go
    names, K, *_ = _compress_case(30, 7, 30)
    GA, KG = go.compress(names, K, GN=6)
    g = np.array([x for _, x in GA])
    assert g[0] == 1 and g.max() == 6
    firsts = [int(np.argmax(g == k)) for k in range(1, 7)]
    assert firsts == sorted(firsts)
    for a in range(6):
        for b in range(6):
            assert KG[a, b] == pytest.approx(K[np.ix_(g == a + 1, g == b + 1)].mean(), rel=1e-12)


# ---- host-side pieces of the file formats: no GPU needed -------------------------------
def test_bed_header_and_text_indexing_on_the_host():
    import ctypes as C
    from gapit_b200 import _lib
    lib = _lib.load()
    m, off = C.c_int64(), C.c_int64()
    bed = bytes([0x6C, 0x1B, 0x01]) + bytes(3 * 7)                  # n = 10 samples -> 3 bytes per variant, 7 variants
    assert lib.gapit_bed_header(bed, len(bed), 10, C.byref(m), C.byref(off)) == 0 and (m.value, off.value) == (7, 3)
    assert lib.gapit_bed_header(bed, len(bed), 9, C.byref(m), C.byref(off)) == 0 and m.value == 7     # ceil(9/4) = 3 too
    assert lib.gapit_bed_header(bed, len(bed), 16, C.byref(m), C.byref(off)) != 0                     # 21 is no multiple of 4
    assert b"whole number" in lib.gapit_last_error()
    assert lib.gapit_bed_header(bytes([0x6C, 0x1B, 0x00]) + bytes(21), 24, 10, C.byref(m), None) != 0
    assert b"sample-major" in lib.gapit_last_error()
    assert lib.gapit_bed_header(b"abc" + bytes(21), 24, 10, C.byref(m), None) != 0
    assert b"magic" in lib.gapit_last_error()
    text = b"taxa m1 m2 m3\nA 0 1 2\r\nB\t2\t1\tNA\n\nC,1,1,0"       # final line without newline, one blank line
    nl = C.c_int64()
    assert lib.gapit_text_index_lines(text, len(text), None, 0, C.byref(nl)) == 0 and nl.value == 5
    offs = (C.c_int64 * (nl.value + 1))()
    assert lib.gapit_text_index_lines(text, len(text), offs, nl.value + 1, C.byref(nl)) == 0
    assert list(offs) == [0, 14, 23, 32, 33, len(text)]
    assert lib.gapit_text_index_lines(text, len(text), offs, 3, C.byref(nl)) != 0                      # array too small
    n, mm = C.c_int64(), C.c_int64()
    assert lib.gapit_numtext_shape(text, len(text), offs, nl.value, 1, C.byref(n), C.byref(mm)) == 0
    assert (n.value, mm.value) == (4, 3)      # the blank line in the middle is a row (and will be rejected as ragged)
    t2 = b"taxa m1 m2\nA 0 1\nB 2 2\n\n\n"
    assert lib.gapit_text_index_lines(t2, len(t2), None, 0, C.byref(nl)) == 0
    offs2 = (C.c_int64 * (nl.value + 1))()
    lib.gapit_text_index_lines(t2, len(t2), offs2, nl.value + 1, C.byref(nl))
    assert lib.gapit_numtext_shape(t2, len(t2), offs2, nl.value, 1, C.byref(n), C.byref(mm)) == 0
    assert (n.value, mm.value) == (2, 2)      # trailing blank lines are not taxa


def test_oracle_file_format_readers():
    from oracle import gapit_oracle as go
    n, rows = 6, [0b11_10_01_00, 0b00_11]      # sample 0: 00 (hom A1), 1: 01 (missing), 2: 10 (het), 3: 11 (hom A2), 4: 11, 5: 00
    bed = bytes([0x6C, 0x1B, 0x01] + rows)
    assert go.read_bed(bed, n, "A1").T.tolist() == [[2, -1, 1, 0, 0, 2]]
    assert go.read_bed(bed, n, "A2").T.tolist() == [[0, -1, 1, 2, 2, 0]]
    GD, GT, names = go.read_numeric_text(b"taxa\tm1\tm2\nA 0 1.0\nB,NA,2\n")
    assert GT == ["A", "B"] and names == ["m1", "m2"] and GD[0].tolist() == [0.0, 1.0] and np.isnan(GD[1, 0]) and GD[1, 1] == 2.0


def test_bench_workloads_are_the_named_baseline_configs():
    """bench.py's headline shape and the full-size legs are BASELINE.json's configs[1..4], not smaller stand-ins."""
    import json
    import sys
    import importlib.util
    from gapit_b200 import fullsize
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        cfgs = json.load(f)["configs"]
    nums = [[int(x.replace(",", "")) for x in re.findall(r"\d[\d,]*\d", str(c)) if int(x.replace(",", "")) >= 100] for c in cfgs]
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    old = sys.argv
    sys.argv = ["bench.py"]
    try:
        args = bench.parse()
    finally:
        sys.argv = old
    assert [args.n, args.m] == nums[1][:2] and args.gpus == 1 and args.warmup >= 3          # configs[1]: 5,000 x 500,000
    c3, c4, c5 = fullsize.CONFIGS["c3"], fullsize.CONFIGS["c4"], fullsize.CONFIGS["c5"]
    assert [c3["n"], c3["m"]] == nums[2][:2] and c3["T"] == 1 and c3["scan"]                 # 20,000 x 1,000,000
    assert [c4["n"], c4["m"]] == nums[3][:2] and not c4["scan"]                               # 50,000 x 1,000,000, kinship only
    assert [c5["n"], c5["m"], c5["T"]] == nums[4][:3] and c5["scan"]                          # 10,000 x 2,000,000 x 100
    assert bench.METRIC.startswith("SNPs tested/sec") and bench.UNIT == "SNPs/s"


def test_kinship_quotient_by_reciprocal_and_residual_is_the_rounded_division():
    """The kinship epilogue (csrc/kinship.cu, vanraden_epilogue_lower_kernel) forms K_ij = num / (D/2) as
    q = num * r, r = fl(1 / (D/2)), then q + fma(-q, D/2, num) * r -- the refinement step an fp64 division ends with.
    Restated in exact rational arithmetic (an fma is one rounding of the exact value): the result is the correctly
    rounded quotient, so the cheaper sequence changes nothing in K."""
    from fractions import Fraction
    import random
    rnd = random.Random(20160301)
    for _ in range(20000):
        n = rnd.choice([281, 5000, 20000, 50000])
        m = rnd.randint(1000, 2_000_000)
        D = float(rnd.randint(n * m // 10, n * m))
        numd = float(rnd.randint(-int(2 * D), int(2 * D)))
        Dh = 0.5 * D
        r = 1.0 / Dh
        q = numd * r
        e = float(Fraction(numd) - Fraction(q) * Fraction(Dh))
        v = float(Fraction(q) + Fraction(e) * Fraction(r))
        assert v == float(Fraction(numd) / Fraction(Dh))


def test_gram_launch_is_cut_into_whole_rounds():
    """gapit_gram_plan (host arithmetic, the plan launch_gram uses): every job of a cross-product launch is equally long,
    so the launch lasts ceil(jobs * splits / CTA pairs) rounds.  BASELINE configs[1] on a 148-SM part: 110 pair jobs, 2
    marker splits, 3 rounds with 99 % of the pairs busy (the earlier "at least six waves" rule took 5 splits: 7.4 -> 8
    rounds, 93 %); and no shape is ever cut worse than with a single split."""
    lib = _lib.load()

    def plan(n, m, sms=148):
        j, s, r = C.c_int(), C.c_int(), C.c_int()
        assert lib.gapit_gram_plan(n, m, sms, C.byref(j), C.byref(s), C.byref(r)) == 0
        return j.value, s.value, r.value

    jobs, splits, rounds = plan(5000, 500000)
    assert (jobs, splits, rounds) == (110, 2, 3) and jobs * splits / (74 * rounds) > 0.98
    jobs, splits, rounds = plan(50000, 125000)                      # configs[3], one of 8 marker shards
    assert splits == 1 and jobs * splits / (74 * rounds) > 0.99
    for n in (281, 600, 1000, 3000, 5000, 10000, 20000):
        for m in (1536, 3093, 40000, 200000, 1000000, 2000000):
            jobs, splits, rounds = plan(n, m)
            kblocks = (m + 127) // 128
            assert splits >= 1 and (splits == 1 or kblocks // splits >= 32)
            assert rounds == -(-jobs * splits // 74)
            one = -(-jobs // 74) * (kblocks + 4.0)
            assert rounds * (kblocks / splits + 4.0) <= one * 1.0001, (n, m, splits)
    assert lib.gapit_gram_plan(0, 10, 148, None, None, None) != 0
