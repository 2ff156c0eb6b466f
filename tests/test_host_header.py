"""The drop-in boundary: this repo's C++23 include/QuBLAS.h compiled by g++ exactly like user code.

The SAME generated translation units that produced the golden vectors from the reference header
(oracle/ref_harness_gen.py) are compiled against this repo's header instead; every result AND every
deduced result type must equal what the reference produced.  Host-only cases (scalar Qmul/Qadd/Qsub,
converting ctor, variadic Qreduce, scalar TFComplexMul) run on the CPU box; tensor cases
(Qreduce(tensor), the Qgemul composition) go through libqb200.so and need the GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness_gen as gen  # noqa: E402  (test infrastructure: the case emitter only)

GOLD = os.path.join(ROOT, "tests", "golden")
LINK = ["-L" + os.path.join(ROOT, "qublas_b200"), "-lqb200", "-Wl,-rpath," + os.path.join(ROOT, "qublas_b200")]


def _load():
    cases = []
    for kind in ("scalar", "reduce", "tf", "gemul", "cgemul", "fromdouble", "bcm", "rc"):
        with open(os.path.join(GOLD, "ref_%s.json" % kind)) as f:
            cases += json.load(f)["cases"]
    return cases


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    cases = _load()
    out = str(tmp_path_factory.mktemp("hdr"))
    run = gen.build_harness(cases, os.path.join(ROOT, "include"), out, "hdr", n_tu=8, extra_flags=LINK)
    return cases, run


def _host_only(c):
    return c["kind"] in ("scalar", "tf", "cgemul", "fromdouble", "bcm", "rc", "rcgemul") or (c["kind"] == "reduce" and c["form"] == "variadic")


def test_header_scalar_paths_match_reference(harness):
    cases, run = harness
    n = 0
    for idx, c in enumerate(cases):
        if not _host_only(c):
            continue
        for row, want in zip(c["inputs"], c["outputs"]):
            assert run(idx, row) == want, (c["kind"], {k: v for k, v in c.items() if k not in ("inputs", "outputs")}, row)
            n += 1
    assert n > 12000


@pytest.mark.gpu
def test_header_tensor_paths_match_reference(harness):
    cases, run = harness
    n = 0
    for idx, c in enumerate(cases):
        if _host_only(c):
            continue
        for row, want in zip(c["inputs"], c["outputs"]):
            assert run(idx, row) == want, (c["kind"], c.get("len"), c.get("M"), c.get("K"))
            n += 1
    assert n > 200


def _build_readme(tmp_path):
    exe = str(tmp_path / "readme_example")
    subprocess.check_call(["g++", "-std=c++23", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "readme_example.cpp"), "-o", exe] + LINK)
    return exe


def _parse(out):
    return {ln.split()[0]: [int(v) for v in ln.split()[1:] if v.lstrip("-").isdigit()] for ln in out.splitlines() if ln and ln.split()[0] != "runtime_error:"}


def test_readme_example_compiles_and_scalar_part_runs(tmp_path):
    """readme.md:20-110 compiles unchanged; scalar results equal the reference's; without a GPU the
    tensor part fails loudly (std::runtime_error -> exit code 3), it never falls back to the CPU."""
    exe = _build_readme(tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode in (0, 3), p.stdout + p.stderr
    lines = p.stdout.splitlines()
    assert lines[0] == "a 256 q1 8 q2 2 elem 80 h1[3,0,1] 64"
    assert lines[1] == "red3 272"            # Qreduce<type1>(q1, q2, q1, q2) = 34.0
    assert lines[2] == "cprod 0 0"           # reference: tests/golden/ref_tf.json readme case
    assert lines[3] == "bits 000010000000101000000000110000001000000000000100000010010000"   # reference output, readme.md:118-121
    assert lines[4] == "bitsz 8 16 24 32 40 48"
    if p.returncode == 3:
        assert "no CUDA device" in p.stdout and "no CPU fallback" in p.stdout


@pytest.mark.gpu
def test_readme_example_on_gpu(tmp_path):
    exe = _build_readme(tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    r = _parse(p.stdout)
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("red1")][0].split()
    assert int(line[1]) == 0 and int(line[3]) == 7                       # readme.md:57,61: red1 = 0, red2 raw 7
    assert r["m3"] == [192, 448, 192, 448, 448, 448, 0, 0, 192, 0, 0, 0, 448, 0, 0, 0]      # SURVEY 3.5, TransposedA=true
    assert r["m4"] == [64, 64, 128, 448, 448, 320, 448, 0, 448, 0, 192, 256, 64, 192, 256, 384]
    assert r["cvec"] == [0, 0, 0, 0, 0, 0]


def test_bitstream_and_printing_match_reference(tmp_path):
    """BitStream import/export (QuBLAS.h:4545-4827), toString and operator<< on scalars, complex values and tensors:
    tests/cpp/bitstream_example.cpp prints the same text as when compiled against the reference header
    (tests/golden/ref_bitstream.txt, written by `python oracle/ref_harness_gen.py --text`)."""
    exe = _build_cpp(tmp_path, "bitstream_example")
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    with open(os.path.join(ROOT, "tests", "golden", "ref_bitstream.txt")) as f:
        assert p.stdout == f.read()


def test_anus_poly_and_approx_match_reference(tmp_path):
    """ANUS::Qpoly / Qapprox / Qapprox_s (QuBLAS.h:4829-4897) over a sweep of 222 inputs: same raw words and same
    deduced result types as the reference header (tests/golden/ref_anus.txt)."""
    exe = _build_cpp(tmp_path, "anus_example")
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    with open(os.path.join(ROOT, "tests", "golden", "ref_anus.txt")) as f:
        assert p.stdout == f.read()


def _build_cpp(tmp_path, name):
    exe = str(tmp_path / name)
    subprocess.check_call(["g++", "-std=c++23", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe] + LINK)
    return exe


def test_batched_and_complex_example_compiles(tmp_path):
    """Qbatch / QgemulBatched, complex Qgemul with TFComplexMul and with the default BasicComplexMul, a heap-backed
    Qreduce: compiles as user code; on a box without GPU it stops at the first tensor call with std::runtime_error."""
    exe = _build_cpp(tmp_path, "batched_example")
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode in (0, 3), p.stdout + p.stderr


@pytest.mark.gpu
def test_batched_and_complex_example_on_gpu(tmp_path):
    exe = _build_cpp(tmp_path, "batched_example")
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "batched Qgemul: OK" in p.stdout and "complex Qgemul TFComplexMul: OK" in p.stdout
    assert "complex Qgemul BasicComplexMul: OK" in p.stdout and "Qreduce 4096: raw" in p.stdout
    assert "element-wise Qadd/Qsub/Qmul: OK" in p.stdout and "device-resident sweep: OK" in p.stdout


def test_anus_tensor_example_compiles(tmp_path):
    """ANUS::Qpoly / Qapprox / Qtable on tensors compile as user code; without a GPU the first tensor call throws."""
    exe = _build_cpp(tmp_path, "anus_tensor_example")
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode in (0, 3), p.stdout + p.stderr


@pytest.mark.gpu
def test_anus_tensor_example_on_gpu(tmp_path):
    """The tensor forms (GPU) give the same raw words as the scalar forms, which ref_anus.txt pins to the reference."""
    exe = _build_cpp(tmp_path, "anus_tensor_example")
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "tensor ANUS: OK" in p.stdout and "rsqrt(4.0) raw 128 sqrt(4.0) raw 512" in p.stdout
