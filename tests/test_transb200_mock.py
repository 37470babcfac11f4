"""The C++ adaptor include/atlas_b200/TransB200.h compiled against a mock of atlas's TransImpl/TransFactory
(tests/cpu/mock_atlas): signature drift is a compile error; at run time it must either work (GPU) or refuse
loudly (no GPU)."""
import os
import subprocess
import tempfile

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_and_run():
    src = os.path.join(REPO, "tests", "cpu", "test_transb200_mock.cc")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "t")
        cmd = ["/usr/bin/g++", "-O1", "-std=c++17", "-I", os.path.join(REPO, "include"), "-I", os.path.join(REPO, "tests", "cpu", "mock_atlas"),
               "-o", exe, src, "-L", os.path.join(REPO, "atlas_b200"), "-lsptrans_b200", "-Wl,-rpath," + os.path.join(REPO, "atlas_b200"),
               "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout
        r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return r.returncode, r.stdout


def test_adaptor_compiles_and_refuses_without_gpu():
    rc, out = _build_and_run()
    assert rc == 0, out


@pytest.mark.gpu
def test_adaptor_runs_on_gpu():
    rc, out = _build_and_run()
    assert rc == 0 and "invtrans err" in out, out


def test_plugin_translation_unit_registers_backend():
    """plugin/atlas-b200/src/B200Plugin.cc (the atlas::Plugin a maintainer builds) compiles against
    the mock headers and its static initialisers register the plugin and the "b200" Trans backend."""
    src = os.path.join(REPO, "tests", "cpu", "test_plugin_mock.cc")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "t")
        cmd = ["/usr/bin/g++", "-O1", "-std=c++17", "-I", os.path.join(REPO, "include"), "-I", os.path.join(REPO, "tests", "cpu", "mock_atlas"),
               "-o", exe, src, "-L", os.path.join(REPO, "atlas_b200"), "-lsptrans_b200", "-Wl,-rpath," + os.path.join(REPO, "atlas_b200"),
               "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout
        r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0 and "backend registered: 1" in r.stdout, r.stdout

