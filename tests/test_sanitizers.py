"""The host half of the product path -- scene front end, every lower(), and what vk_scene_upload does before it touches the
device (validator, 4-wide collapse, flat-program builder: vk_scene_check) -- compiled with AddressSanitizer +
UndefinedBehaviorSanitizer and driven by tests/sanitize_host.cpp over every scene and over thousands of corrupted scene
descriptions held in exact-size heap arrays.  Any return code is fine; a sanitizer report (or a leak) fails the test.

This is the CPU stand-in for SURVEY §5's sanitizer row: compute-sanitizer is closed on this pool, the device queues check their
own protocol in a debug build (tests/test_zz_warpq_selfcheck_gpu.py).  The oracle -- what every parity test trusts -- gets the
same treatment (tests/sanitize_oracle.cpp).  The host run found one real defect when it was written: records
NOT reachable from the root were never range-checked, but the layout planner walks whole arrays (Validator::run now checks
every node, transform and medium record)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = "/usr/local/cuda/bin/nvcc"
# float-cast-overflow is not part of gcc's -fsanitize=undefined: it catches float -> integer casts of NaN / out-of-range values, whose C++
# meaning (undefined) differs from Rust's saturating `as` that the restatements must reproduce
SAN = ["-fsanitize=address", "-fsanitize=undefined", "-fsanitize=float-cast-overflow", "-fno-omit-frame-pointer",
       "-fno-sanitize-recover=undefined", "-fno-sanitize-recover=float-cast-overflow"]


@pytest.fixture(scope="module")
def sanitized(tmp_path_factory):
    if not (shutil.which("g++") and os.path.exists(NVCC)):
        pytest.skip("needs g++ and nvcc")
    out = tmp_path_factory.mktemp("asan")
    inc = ["-I" + os.path.join(ROOT, "include")]
    procs = []
    for src in ("vecchio_b200/host/vecchio.cpp", "vecchio_b200/host/scene.cpp", "vecchio_b200/host/capi.cpp", "tests/sanitize_host.cpp",
                "tests/sanitize_oracle.cpp", "oracle/oracle.cpp"):
        obj = str(out / (os.path.basename(src) + ".o"))
        procs.append((obj, subprocess.Popen(["g++", "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-fopenmp", *SAN, *inc, "-c", os.path.join(ROOT, src), "-o", obj],
                                            stderr=subprocess.PIPE, text=True)))
    # the planner is a header compiled by nvcc in the product (vk_relayout.cu has no device code): same compiler here
    xsan = [a for f in SAN for a in ("-Xcompiler", f)]
    obj = str(out / "vk_relayout.o")
    procs.append((obj, subprocess.Popen([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-g", "--expt-relaxed-constexpr",
                                         *xsan, *inc, "-c", os.path.join(ROOT, "vecchio_b200/csrc/vk_relayout.cu"), "-o", obj],
                                        stderr=subprocess.PIPE, text=True)))
    objs = []
    for obj, p in procs:
        _, err = p.communicate(timeout=900)
        assert p.returncode == 0, err[-3000:]
        objs.append(obj)
    host = [o for o in objs if os.path.basename(o) in ("vecchio.cpp.o", "scene.cpp.o", "capi.cpp.o")]
    exe = str(out / "sanitize_host")
    r = subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fsanitize=address", "-Xcompiler", "-fsanitize=undefined",
                        "-o", exe, str(out / "sanitize_host.cpp.o"), str(out / "vk_relayout.o"), *host, "-lz"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    exe_oracle = str(out / "sanitize_oracle")
    r = subprocess.run(["g++", "-fsanitize=address", "-fsanitize=undefined", "-fopenmp", "-o", exe_oracle, str(out / "sanitize_oracle.cpp.o"),
                        str(out / "oracle.cpp.o"), *host, "-lz"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe, str(out), exe_oracle


@pytest.mark.parametrize("salt", [0, 1])
def test_host_path_is_clean_under_asan_and_ubsan(sanitized, salt):
    exe, tmp, _ = sanitized
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([exe, os.path.join(ROOT, "assets"), os.path.join(tmp, f"frame_{salt}.ppm"), "1200", str(salt)],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-4000:])
    assert "sanitize_host ok" in r.stdout and "ERROR" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-4000:]
    words = r.stdout.split()
    cameras, refused, accepted = int(words[2]), int(words[4]), int(words[8])
    assert cameras == 2 * (671 + 721 + 671 + 12)  # two seeds x (three turntables: random spheres, bowser, the lightless cover; twelve fixed cameras)
    assert refused > 3000 and accepted > 1000  # both sides of the validator are exercised: most corruptions are caught, many are harmless


def test_oracle_is_clean_under_asan_and_ubsan(sanitized):
    """The checker itself (oracle/oracle.cpp): every scene rebuilt from its lowered description, harvested and degenerate rays
    intersected (also with injected medium variates), small frames with HEAD's and the legacy integrator, a sample slice,
    a lens, and orc_eval_batch over every material / texture / light with NaN, infinite and extreme inputs."""
    _, _, exe = sanitized
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1", OMP_NUM_THREADS="2")
    r = subprocess.run([exe, os.path.join(ROOT, "assets")], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-4000:])
    assert "sanitize_oracle ok" in r.stdout and "ERROR" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-4000:]
