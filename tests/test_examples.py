"""The C++ example over the C ABI (examples/sptrans-benchmark-trans.cc, the harness shape of the reference's
src/sandbox/benchmark_trans/atlas-benchmark-trans.cc) compiles against include/sptrans_b200.h with a plain g++ -- no
CUDA headers, no torch -- links against the product library, and refuses loudly without a device."""
import os
import subprocess
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_benchmark_example_compiles_and_refuses_without_gpu():
    from atlas_b200 import _lib

    src = os.path.join(REPO, "examples", "sptrans-benchmark-trans.cc")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "bench")
        cmd = ["/usr/bin/g++", "-O1", "-std=c++17", "-I", os.path.join(REPO, "include"), src, "-o", exe,
               "-L", os.path.join(REPO, "atlas_b200"), "-lsptrans_b200", "-Wl,-rpath," + os.path.join(REPO, "atlas_b200"),
               "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout
        r = subprocess.run([exe, "--grid", "O16", "--nscalar", "2", "--niter", "2", "--dirtrans"], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
        if _lib.lib.sptrans_device_count() == 0:
            assert r.returncode == 2 and "no CPU fallback" in r.stdout, r.stdout
        else:
            assert r.returncode == 0 and "invtrans[min]" in r.stdout and "round trip" in r.stdout, r.stdout
