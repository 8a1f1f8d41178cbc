"""Drop-in check of the class API: the reference's OWN test programs and example, compiled UNMODIFIED from
/root/reference against channelflow_b200/host (tests/refprogs/Makefile), must pass at their own tolerances.

  -m gpu        tests/_refprogs/gpu/*   linked with the product libraries, run on the B200
  -m "not gpu"  tests/_refprogs/emu/*   linked with the CPU emulation build of the same CUDA sources (a fast subset)

The binaries are built in the container that has /root/reference and travel to the GPU box (tests/_refprogs is
git-ignored, not gpurun-ignored); nothing here reads /root/reference at run time.  Test matrix = the reference's ctest
registrations (tests/CMakeLists.txt:57-111)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROGS = os.path.join(ROOT, "tests", "_refprogs")
REF = "/root/reference"


def _build(flavour):
    if os.path.isdir(REF):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "refprogs"), flavour, "-j8"], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-4000:]


def _write_os_data(d):
    """tests/data/os_*.asc|cmplx (Orr-Sommerfeld eigenfunction, 65 modes, Re 7500) from the committed fixture."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "os_eig.npz"))
    os.makedirs(d, exist_ok=True)
    two_pi = "6.2831853071795862"
    for name, arr, nd in (("os_ueig10_65.asc", g["ueig"], 3), ("os_peig10_65.asc", g["peig"], 1)):
        with open(os.path.join(d, name), "w") as f:
            f.write("%% %d 65 1 0 %s %s -1 1 P\n" % (nd, two_pi, two_pi))
            for row in arr:
                f.write(" ".join(repr(float(x)) for x in row) + "\n")
    with open(os.path.join(d, "os_omega10_65.cmplx"), "w") as f:
        f.write("%r %r\n" % (float(g["omega"][0]), float(g["omega"][1])))


def _run(flavour, prog, args, tmp_path, timeout):
    exe = os.path.join(PROGS, flavour, prog)
    if not os.path.exists(exe):
        pytest.skip("%s not built (needs /root/reference at build time)" % exe)
    run = tmp_path / "run"
    run.mkdir()
    _write_os_data(str(tmp_path / "data"))
    r = subprocess.run([exe] + list(args), cwd=str(run), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    return r


# the reference's ctest list; the DNS programs take (algorithm, mean constraint[, base/fluctuation]) switches
UNIT = [("tridiagTest", ()), ("chebyTest", ()), ("laminarTest", ()), ("helmholtzTest", ()), ("tausolverTest", ()),
        ("poissonTest", ()), ("pressureTest", ()), ("ioTest", ())]
ALGS = ["--cnfe1", "--cnab2", "--cnrk2", "--smrk2", "--sbdf2", "--sbdf3", "--sbdf4"]
DNS_ZERO = [("dnsZeroTest", (a, c)) for a in ALGS for c in ("--bulkv", "--gradp")]
DNS_PARA = [("dnsParabolaTest", (a, f, c)) for a in ALGS for f in ("--fluc", "--base") for c in ("--bulkv", "--gradp")]
DNS_SIN = [("dnsSinusoidTest", (a, f, c)) for a in ("--cnrk2", "--sbdf3") for f in ("--zero", "--parab") for c in ("--bulkv", "--gradp")]


def _id(p):
    return p[0] + "".join(p[1])


@pytest.mark.gpu
@pytest.mark.parametrize("prog", UNIT + DNS_ZERO + DNS_PARA + DNS_SIN, ids=_id)
def test_reference_program_passes_on_gpu(prog, tmp_path):
    _build("gpu")
    r = _run("gpu", prog[0], prog[1], tmp_path, 600)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-1500:])
    assert "pass" in r.stderr


@pytest.mark.gpu
def test_orr_sommerfeld_program_matches_reference_number(tmp_path):
    """dnsOrrsommTest is commented out of the reference's ctest list: run as shipped (no arguments) it ends with
    FinalError 1.0844443950667e-07 against its own bound 1e-07 in the reference build as well (oracle/_ref/bin).  The
    drop-in build must print the same number."""
    _build("gpu")
    r = _run("gpu", "dnsOrrsommTest", (), tmp_path, 900)
    line = [l for l in r.stdout.splitlines() if "FinalError ==" in l]
    assert line, r.stdout[-2000:]
    err = float(line[-1].split("FinalError ==")[1].split()[0])
    assert abs(err - 1.0844443950667e-07) < 1e-12, err


@pytest.mark.gpu
def test_couette_example_runs_on_gpu(tmp_path):
    """examples/couette.cpp: the reference's documented minimal program (DNS of a perturbed Couette flow)."""
    _build("gpu")
    r = _run("gpu", "couette", (), tmp_path, 900)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-1500:])
    _compare_couette(r.stdout)


def _couette_series(text):
    vals = {}
    for key in ("CFL", "L2Norm(u)", "Ubulk"):
        vals[key] = [float(l.split("==")[1]) for l in text.splitlines() if l.strip().startswith(key + " ==")]
    return vals


def _compare_couette(out, upto=None):
    """Every printed time unit against the reference's own run of the same program (tests/golden/couette_ref.txt,
    6 printed digits): 30 time units = 960 SBDF3 steps of a perturbed Couette flow."""
    ref = _couette_series(open(os.path.join(ROOT, "tests", "golden", "couette_ref.txt")).read())
    got = _couette_series(out)
    n = len(ref["L2Norm(u)"]) if upto is None else upto
    assert len(got["L2Norm(u)"]) >= n and n >= 1
    for key in ("CFL", "L2Norm(u)"):
        np.testing.assert_allclose(got[key][:n], ref[key][:n], rtol=2e-5)
    np.testing.assert_allclose(got["Ubulk"][:n], ref["Ubulk"][:n], rtol=2e-4, atol=1e-9)


EMU_FAST = [("tridiagTest", ()), ("chebyTest", ()), ("laminarTest", ()), ("helmholtzTest", ()), ("poissonTest", ()),
            ("dnsZeroTest", ("--sbdf3", "--gradp")), ("pressureTest", ()), ("ioTest", ())]


@pytest.mark.parametrize("prog", EMU_FAST, ids=_id)
def test_reference_program_passes_on_emulation(prog, tmp_path):
    from tests import parity
    parity.emu_lib()
    _build("emu")
    r = _run("emu", prog[0], prog[1], tmp_path, 900)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-1500:])
    assert "pass" in r.stderr


@pytest.mark.slow
@pytest.mark.skipif(not os.environ.get("CF_SLOW_TESTS"), reason="70 s on the CPU emulation; set CF_SLOW_TESTS=1 (the GPU run covers it)")
def test_tausolver_program_passes_on_emulation(tmp_path):
    from tests import parity
    parity.emu_lib()
    _build("emu")
    r = _run("emu", "tausolverTest", (), tmp_path, 1800)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-1500:])


# ---------------------------------------------------------------------------------------------- simulateflow, unchanged
ORACLE_BIN = os.path.join(ROOT, "oracle", "_ref", "bin")


def _simulateflow_pair(flavour, tmp_path, grid, T, timeout, symms=False):
    """BASELINE configs[0]: the reference's programs/simulateflow.cpp and tools/randomfield.cpp, compiled UNMODIFIED against
    the drop-in headers, next to the same two programs of the compiled reference (oracle/_ref/bin) on the same command line:
    plane Couette Re 400, SBDF3, rotational, 2/3 dealiasing, dt = 0.02.  Returns (relative L2 difference of the saved final
    fields, the two stdout logs)."""
    import channelflow_b200 as cf
    from tests import parity
    for exe in ("simulateflow", "randomfield"):
        if not os.path.exists(os.path.join(PROGS, flavour, exe)) or not os.path.exists(os.path.join(ORACLE_BIN, exe)):
            pytest.skip("simulateflow/randomfield binaries not built (need /root/reference at build time)")
    Nx, Ny, Nz = grid
    rf = ["-Nx", str(Nx), "-Ny", str(Ny), "-Nz", str(Nz), "-lx", "1", "-lz", "0.5", "-sd", "1", "-s", "0.4", "-m", "0.2", "u0"]
    sim = ["-R", "400", "-T", str(T), "-dt", "0.02", "-vdt", "false", "-dT", "1", "-l2", "-cfl", "-dv"]
    if symms:  # confine the flow to the shift-reflect subspace, projecting every time unit (dns.cpp:152-156)
        sim += ["-symms", "sigma.asc", "-symmpi", "1"]
    sim += ["u0"]
    logs = {}
    for side, bindir in (("ref", ORACLE_BIN), ("new", os.path.join(PROGS, flavour))):
        d = tmp_path / side
        d.mkdir()
        open(str(d / "sigma.asc"), "w").write("% 1\n1 1 1 -1 0.5 0\n")
        r = subprocess.run([os.path.join(bindir, "randomfield")] + rf, cwd=str(d), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
        assert r.returncode == 0, r.stdout[-2000:]
        r = subprocess.run([os.path.join(bindir, "simulateflow")] + sim, cwd=str(d), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
        assert r.returncode == 0, r.stdout[-3000:]
        logs[side] = r.stdout
        assert os.path.exists(str(d / "data" / ("u%d.ff" % T))), os.listdir(str(d / "data"))
    lib = parity.gpu_lib() if flavour == "gpu" else parity.emu_lib()

    def load(path):
        h = lib.L.cf_field_load(path.encode())
        return cf.FlowField(lib, Nx, Ny, Nz, 3, 2 * np.pi, np.pi, handle=h).get()
    out = {}
    for name in ("u0", "data/u%d" % T):
        a, b = load(str(tmp_path / "ref" / name)), load(str(tmp_path / "new" / name))
        out[name] = float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(a.ravel()))
    return out, logs


def _diag_lines(log, key):
    return [float(l.split("==")[1].split()[0]) for l in log.splitlines() if l.strip().startswith(key + " ==")]


@pytest.mark.gpu
def test_simulateflow_with_symmetry_projection(tmp_path):
    _build("gpu")
    d, logs = _simulateflow_pair("gpu", tmp_path, (32, 33, 32), 2, 900, symms=True)
    assert d["data/u2"] < 1e-9, d
    assert _diag_lines(logs["new"], "L2Norm(u)") == _diag_lines(logs["ref"], "L2Norm(u)")


@pytest.mark.gpu
def test_simulateflow_runs_unchanged_c1(tmp_path):
    """north_star: `simulateflow` runs unchanged on top of the B200 path; C1 = 32x33x32, 100 steps: the saved field agrees with
    the reference run to <= 1e-9 (parity gate after 100 steps), the initial field from `randomfield` to round-off."""
    _build("gpu")
    d, logs = _simulateflow_pair("gpu", tmp_path, (32, 33, 32), 2, 900)
    print("simulateflow C1:", d)
    assert d["u0"] < 1e-14 and d["data/u2"] < 1e-9, d
    for key in ("L2Norm(u)", "CFL"):
        np.testing.assert_allclose(_diag_lines(logs["new"], key), _diag_lines(logs["ref"], key), rtol=1e-5)


def test_simulateflow_runs_unchanged_emulation(tmp_path):
    from tests import parity
    parity.emu_lib()
    _build("emu")
    d, logs = _simulateflow_pair("emu", tmp_path, (16, 17, 16), 1, 900, symms=True)
    assert d["u0"] < 1e-14 and d["data/u1"] < 1e-11, d
    assert _diag_lines(logs["new"], "L2Norm(u)") == _diag_lines(logs["ref"], "L2Norm(u)")


# ------------------------------------------------------------------- the golden pair through the reference's own programs
def _write_golden_ff(lib, d):
    """data/uinit.ff, data/ufinal.ff (reference tests/data/u{init,final}.nc, 48x35x48) from the committed fixture, written by
    this package's own .ff writer (the reference reads and writes .ff when it is built without NetCDF)."""
    from oracle import refcf
    from tests import parity
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_pair.npz"))
    geo = [int(g["Nx"]), int(g["Ny"]), int(g["Nz"]), 3, float(g["Lx"]), float(g["Lz"]), float(g["a"]), float(g["b"])]
    os.makedirs(d, exist_ok=True)
    for name in ("uinit", "ufinal"):
        ur = refcf.RefField(*geo).load_padded_physical(g[name])
        ur.make_spectral()
        parity.to_gpu(lib, ur).save(os.path.join(d, name))


@pytest.mark.gpu
def test_time_integration_program_golden_pair(tmp_path):
    """tests/timeIntegrationTest.cpp, unmodified: uinit -> 440 SBDF3 steps -> L2Dist to ufinal below its own 1e-13."""
    from tests import parity
    _build("gpu")
    _write_golden_ff(parity.gpu_lib(), str(tmp_path / "data"))
    exe = os.path.join(PROGS, "gpu", "timeIntegrationTest")
    r = subprocess.run([exe], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-1500:])
    assert "pass" in r.stderr


@pytest.mark.gpu
def test_reference_benchmark_tool(tmp_path):
    """tools/benchmark.cpp -- the reference's own timing harness (mean wall time per 40-step time unit of the golden-pair
    run, first unit discarded) -- unmodified on the B200, next to the compiled reference on one host core."""
    from tests import parity
    _build("gpu")
    _write_golden_ff(parity.gpu_lib(), str(tmp_path))
    res = {}
    for side, exe in (("b200", os.path.join(PROGS, "gpu", "benchmark")), ("reference_1core", os.path.join(ROOT, "oracle", "_ref", "bin", "benchmark"))):
        if not os.path.exists(exe):
            pytest.skip("benchmark binary not built")
        r = subprocess.run([exe, "-d", str(tmp_path)], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:]
        avg = [l for l in r.stdout.splitlines() if "Average time/timeunit" in l]
        dist = [l for l in r.stdout.splitlines() if "L2Dist" in l]
        res[side] = (float(avg[-1].split(":")[1].strip().rstrip("s")), dist[-1].strip() if dist else "")
    print("reference benchmark tool (s per time unit of 40 steps, 48x35x48):", res)
    assert res["b200"][0] < res["reference_1core"][0]


# ---------------------------------------------------------------------------------------------- the reference's tools
TOOL_CHAIN = [("randomfield", "-Nx {Nx} -Ny {Ny} -Nz {Nz} -lx 1 -lz 0.5 -sd 1 u0"), ("perturbfield", "-sd 3 -m 0.05 u0 u1"),
              ("symmetryop", "-sx -ax 0.25 u1 u2"), ("addfields", "-lc 0.3 u1 0.7 u2 u3"), ("pressure", "-nl rot u3 p3"),
              ("fieldconvert", "u3 u3copy.ff")]


def _tool_chain(flavour, tmp_path, grid):
    """tools/{randomfield,perturbfield,symmetryop,addfields,pressure,fieldconvert}.cpp, unmodified, chained on the drop-in build
    and on the compiled reference with the same command lines; every output file compared."""
    import channelflow_b200 as cf
    from tests import parity
    Nx, Ny, Nz = grid
    for side, bindir in (("ref", ORACLE_BIN), ("new", os.path.join(PROGS, flavour))):
        d = tmp_path / side
        d.mkdir()
        for exe, args in TOOL_CHAIN:
            path = os.path.join(bindir, exe)
            if not os.path.exists(path):
                pytest.skip("%s not built (needs /root/reference at build time)" % path)
            r = subprocess.run([path] + args.format(Nx=Nx, Ny=Ny, Nz=Nz).split(), cwd=str(d), stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                               text=True, timeout=600)
            assert r.returncode == 0, (exe, r.stdout[-2000:])
    # NetCDF out and back in through the unmodified fieldconvert on the drop-in build (the compiled reference has no NetCDF)
    for args in ("u3 u3nc.nc", "u3nc.nc u3back.ff"):
        r = subprocess.run([os.path.join(PROGS, flavour, "fieldconvert")] + args.split(), cwd=str(tmp_path / "new"), stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=600)
        assert r.returncode == 0, (args, r.stdout[-2000:])
    lib = parity.gpu_lib() if flavour == "gpu" else parity.emu_lib()
    out = {}
    a = [cf.FlowField(lib, Nx, Ny, Nz, 3, 2 * np.pi, np.pi, handle=lib.L.cf_field_load(str(tmp_path / "new" / n).encode())).get()
         for n in ("u3", "u3back")]
    out["u3_via_nc"] = float(np.linalg.norm((a[0] - a[1]).ravel()) / np.linalg.norm(a[0].ravel()))
    for name, Nd in (("u1", 3), ("u2", 3), ("u3", 3), ("p3", 1), ("u3copy", 3)):
        arrs = []
        for side in ("ref", "new"):
            h = lib.L.cf_field_load(str(tmp_path / side / name).encode())
            arrs.append(cf.FlowField(lib, Nx, Ny, Nz, Nd, 2 * np.pi, np.pi, handle=h).get())
        out[name] = float(np.linalg.norm((arrs[0] - arrs[1]).ravel()) / np.linalg.norm(arrs[0].ravel()))
    return out


@pytest.mark.gpu
def test_reference_tools_chain_on_gpu(tmp_path):
    _build("gpu")
    r = _tool_chain("gpu", tmp_path, (32, 33, 32))
    assert max(r.values()) < 1e-13, r


def test_reference_tools_chain_on_emulation(tmp_path):
    from tests import parity
    parity.emu_lib()
    _build("emu")
    r = _tool_chain("emu", tmp_path, (16, 17, 16))
    assert max(r.values()) < 1e-13, r


# ------------------------------------------------------------------------------------------------ field2vector program
def _vector2field_program(flavour, tmp_path, lib):
    """tests/vector2fieldTest.cpp, unmodified (Eigen::VectorXd from the stand-in header host/compat/Eigen/Dense, Eigen3 is not
    installed here): field -> vector -> field -> vector round trips of the golden-pair initial field, own tolerance 2e-16."""
    _write_golden_ff(lib, str(tmp_path / "data"))
    exe = os.path.join(PROGS, flavour, "vector2fieldTest")
    if not os.path.exists(exe):
        pytest.skip("vector2fieldTest not built")
    r = subprocess.run([exe], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0 and "pass" in r.stderr, (r.stdout[-2000:], r.stderr[-800:])


@pytest.mark.gpu
def test_vector2field_program_on_gpu(tmp_path):
    from tests import parity
    _build("gpu")
    _vector2field_program("gpu", tmp_path, parity.gpu_lib())


def test_vector2field_program_on_emulation(tmp_path):
    from tests import parity
    lib = parity.emu_lib()
    _build("emu")
    _vector2field_program("emu", tmp_path, lib)
