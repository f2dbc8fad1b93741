"""CPU-side checks of the product's host code: the .ini reader (value-identical to the reference incl. float
truncation), the deck renderer, and that the C-ABI library loads and exports every symbol the header declares."""
import ctypes as C
import os
import re

import pytest

import euler2d_kokkos_b200 as e2d
import oracle
from euler2d_kokkos_b200 import _lib
from euler2d_kokkos_b200.decks import DECKS, deck_text
from util import both_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("deck", list(DECKS))
def test_product_parser_equals_oracle_parser(deck):
    hp, op = both_params(deck)
    for k, v in op.as_dict().items():
        assert getattr(hp, k) == v, k


def test_float_truncation_and_derived_values():
    hp = e2d.HydroParams.from_string(deck_text("implode"))
    assert hp.gamma0 == float.fromhex("0x1.aa7efap+0")        # 1.666 read through strtof
    assert hp.cfl == float.fromhex("0x1.99999ap-1")
    assert hp.smallr == float.fromhex("0x1.b7cdfep-34")       # default 1e-10 is a float argument too
    assert hp.smallp == float.fromhex("0x1.c587529038bb1p-68")
    assert hp.gamma6 == float.fromhex("0x1.99a955b0f383dp-1")
    assert (hp.isize, hp.jsize, hp.imax, hp.jmax, hp.ghostWidth) == (260, 132, 259, 131, 2)
    assert hp.dx == 2.0 ** -7 and hp.implementationVersion == 0 and hp.enableOutput == 1
    assert hp.outputPrefix == "test_implode" and hp.outputDir == "./"


INI_QUIRKS = """
; comment line
# another comment
[RUN]
TEND = 0.25   ; inline comment needs the blank before the semicolon
nstepmax=0x20
noutput=-1
[mesh]
nx=12;not-a-comment
ny = 7
xmax=2.5
boundary_type_xmin=3
[hydro]
problem=four_quadrant
riemann=bogus
gamma0=1.4
   1.5
smallr=
[OTHER]
implementationVersion=1.9
"""


def test_ini_quirks_match_the_reference_reader(tmp_path):
    """Case-insensitive keys, hex integers, ';' needs leading whitespace, continuation lines replace the value,
    empty values fall back to the default, implementationVersion goes float -> int."""
    hp = e2d.HydroParams.from_string(INI_QUIRKS)
    f = tmp_path / "q.ini"
    f.write_text(INI_QUIRKS)
    op = oracle.params_from_ini(str(f))
    for k, v in op.as_dict().items():
        assert getattr(hp, k) == v, k
    assert hp.tEnd == 0.25 and hp.nStepmax == 32 and hp.enableOutput == 0
    assert hp.nx == 12 and hp.ny == 7 and hp.boundary_type_xmin == 3
    assert hp.gamma0 == 1.5                    # the indented line replaced 1.4
    assert hp.smallr == float.fromhex("0x1.b7cdfep-34")
    assert hp.riemannSolverType == 0 and hp.implementationVersion == 1
    hp2 = e2d.HydroParams.from_ini(str(f))     # file path and string path agree
    assert hp2.raw.as_dict() == hp.raw.as_dict()
    if oracle.ref_available():                 # and so does the reference's own reader
        r = oracle.ref_run(str(f), nstep=0, dump=False)["meta"]
        assert (r["nx"], r["ny"]) == (12, 7) and float.fromhex(r["gamma0_hex"]) == 1.5
        assert float.fromhex(r["dx_hex"]) == hp.dx and float.fromhex(r["tend_hex"]) == hp.tEnd


def test_missing_file_gives_defaults_and_io_status():
    p = e2d.Params()
    assert e2d.lib().e2d_params_from_ini(b"/no/such/file.ini", C.byref(p)) == 2
    assert (p.nx, p.ny, p.nStepmax, p.problemType) == (2, 2, 1000, 0)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "euler2d_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(e2d_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 35
    L = C.CDLL(e2d.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/euler2d_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b"euler2d_b200" in e2d.lib().e2d_version() and b"sm_100a" in e2d.lib().e2d_version()


def test_struct_layout_matches_the_header():
    """sizeof(e2d_params) as ctypes sees it must match the C side (checked through a round trip)."""
    hp = e2d.HydroParams.from_string(deck_text("shocked_bubble"))
    q = hp.raw.copy()
    q.nx, q.ny = 100, 50
    e2d.check(e2d.lib().e2d_params_init(C.byref(q)))
    assert (q.isize, q.jsize) == (104, 54) and q.dx == (q.xmax - q.xmin) / 100
    assert q.honourRiemannSolver == 0 and q.outputPrefix == b"test_shocked_bubble"
    assert q.shock_loc == hp.shock_loc  # fields after the edited ones are intact


def test_no_gpu_means_loud_failure():
    if e2d.lib().e2d_device_count() > 0:
        pytest.skip("a GPU is present")
    hp = e2d.HydroParams.from_string(deck_text("implode"))
    with pytest.raises(e2d.E2dError, match="no CPU fallback"):
        e2d.HydroRun(hp)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "euler2d_kokkos_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "liboracle" not in text and "euler2d_oracle" not in text, f


def test_save_npy_is_a_valid_numpy_file(tmp_path):
    """e2d_save_npy (host only) writes what cnpy::npy_save writes for a rank-1 double view: numpy reads it back
    bit for bit, and the header is 16-byte aligned."""
    import ctypes as C

    import numpy as np

    import euler2d_kokkos_b200 as e2d

    for n in (0, 1, 7, 100, 12345):
        a = np.random.default_rng(n).normal(size=n)
        path = tmp_path / f"a{n}.npy"
        rc = e2d.lib().e2d_save_npy(os.fsencode(str(path)), a.ctypes.data_as(C.POINTER(C.c_double)), n)
        assert rc == 0
        b = np.load(path)
        assert b.dtype == np.float64 and b.shape == (n,) and np.array_equal(a.view(np.uint64), b.view(np.uint64))
        raw = open(path, "rb").read()
        hl = raw[8] + 256 * raw[9]
        assert raw[:8] == b"\x93NUMPY\x01\x00" and (10 + hl) % 16 == 0 and raw[10 + hl - 1:10 + hl] == b"\n"


def test_save_npy_matches_cnpy_bytes(tmp_path):
    """When the compiled reference is here: its sedov_blast_radial_distances.npy (written by cnpy) and ours are the
    same bytes."""
    import ctypes as C

    import numpy as np

    import euler2d_kokkos_b200 as e2d
    import oracle
    from euler2d_kokkos_b200.decks import write_deck

    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    import subprocess
    ini = write_deck(str(tmp_path / "s.ini"), "sedov_blast_2d", mesh__nx=40, mesh__ny=40, blast__radius=0.06,
                     run__nOutput=-1)
    subprocess.run([oracle.ref_binary(prefer_kokkos=False), ini, "--nstep", "2", "--radial"], check=True,
                   capture_output=True, cwd=tmp_path, env=dict(os.environ, OMP_NUM_THREADS="1"))
    ref_bytes = open(tmp_path / "sedov_blast_radial_distances.npy", "rb").read()
    d = np.load(tmp_path / "sedov_blast_radial_distances.npy")
    mine = tmp_path / "mine.npy"
    assert e2d.lib().e2d_save_npy(os.fsencode(str(mine)), d.ctypes.data_as(C.POINTER(C.c_double)), d.size) == 0
    assert open(mine, "rb").read() == ref_bytes


def test_configmap_typed_getters_follow_the_reference():
    """config/ConfigMap.cpp:32-74 + config/inih/INIReader.cpp: floats through strtof, integers through strtol base 0,
    booleans from the words the reference accepts, keys case-insensitive, last assignment wins, defaults otherwise."""
    import numpy as np

    from euler2d_kokkos_b200 import ConfigMap, HydroParams

    text = "[hydro]\ngamma0=1.666\ncfl = 0.8 ; trailing comment\nniter_riemann=0x10\n[RUN]\ntEnd=0.1\ntend=0.25\n" \
           "[output]\noutputPrefix=abc\n[other]\nflag=yes\nnoflag=off\n"
    cm = ConfigMap(text=text)
    assert cm.getFloat("hydro", "gamma0", 0.0) == float(np.float32(1.666))
    assert cm.getFloat("hydro", "cfl", 0.0) == float(np.float32(0.8))
    assert cm.getInteger("hydro", "niter_riemann", 3) == 16
    assert cm.getFloat("run", "tEnd", 0.0) == 0.25 and cm.getFloat("RUN", "TEND", 0.0) == 0.25
    assert cm.getString("output", "outputPrefix", "output") == "abc"
    assert cm.getString("output", "outputDir", "./") == "./"
    assert cm.getBool("other", "flag", False) is True and cm.getBool("other", "noflag", True) is False
    assert cm.getBool("other", "absent", True) is True and cm.getFloat("nope", "x", 2.5) == 2.5
    cm.setFloat("mesh", "nx", 48)
    cm.setBool("other", "flag", False)
    assert cm.getInteger("mesh", "nx", 0) == 48 and cm.getBool("other", "flag", True) is False
    # HydroParams::setup(ConfigMap&) == the path-based reader
    a, b = HydroParams(), HydroParams.from_string(text + "[mesh]\nnx=48\n[other]\nflag=false\n")
    a.setup(cm)
    assert bytes(a.raw) == bytes(b.raw)
    # a missing file: ParseError() = -1, an empty map, defaults everywhere (the reference never checks it)
    missing = ConfigMap("/nonexistent/file.ini")
    assert missing.ParseError() == -1 and missing.getFloat("hydro", "gamma0", 1.4) == float(np.float32(1.4))


def test_profile_switch_is_off_by_default_and_cheap():
    import ctypes as C

    L = _lib.lib()
    n, d = C.c_ulonglong(), C.c_int()
    assert L.e2d_profile_stats(C.byref(n), C.byref(d)) == 0
    L.e2d_profile_push(b"x")
    L.e2d_profile_pop()
    n2 = C.c_ulonglong()
    L.e2d_profile_stats(C.byref(n2), C.byref(d))
    assert n2.value == n.value and d.value == 0


def test_compat_headers_compile_a_reference_style_host_program(tmp_path):
    """A host written against the reference's headers (same includes, names and constructor arguments as src/main.cpp)
    compiles and links against the library through include/euler2d_compat — no GPU is needed to build it."""
    import subprocess

    src = tmp_path / "host.cpp"
    src.write_text('''
#include "kokkos_shared.h"
#include "HydroBaseFunctor.h"
#include "ComputeRadialProfileFunctor.h"
#include "HydroParams.h"
#include "HydroRun.h"
#include "real_type.h"
#include "Timer.h"
int main(int argc, char * argv[])
{
  using device = Kokkos::Device<Kokkos::DefaultExecutionSpace, Kokkos::DefaultExecutionSpace::memory_space>;
  using real_t = euler2d::real_t;
  Kokkos::initialize(argc, argv);
  Timer total_timer;
  ConfigMap configMap(argc > 1 ? argv[1] : "none.ini");
  euler2d::HydroParams params = euler2d::HydroParams();
  params.setup(configMap);
  if (argc > 2)
  { // only with a GPU
    euler2d::HydroRun<device> * hydro = new euler2d::HydroRun<device>(params, configMap);
    real_t dt = hydro->compute_dt(0);
    hydro->make_boundaries(hydro->U);
    Kokkos::Profiling::pushRegion("main_loop");
    hydro->godunov_unsplit(0, dt);
    Kokkos::Profiling::popRegion();
    if (params.problemType == euler2d::PROBLEM_BLAST and params.blast_total_energy_inside > 0)
      euler2d::ComputeRadialProfileFunctor<device>::apply(params, hydro->U);
    delete hydro;
  }
  Kokkos::finalize();
  return params.nx == 256 ? 0 : 1;
}
''')
    exe = tmp_path / "host"
    libdir = os.path.join(ROOT, "euler2d_kokkos_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-I" + os.path.join(ROOT, "include", "euler2d_compat"), "-o",
                           str(exe), str(src), "-L" + libdir, "-leuler2d_b200", "-Wl,-rpath," + libdir])
    ini = tmp_path / "d.ini"
    from euler2d_kokkos_b200.decks import deck_text

    ini.write_text(deck_text("implode"))
    assert subprocess.run([str(exe), str(ini)], capture_output=True).returncode == 0
