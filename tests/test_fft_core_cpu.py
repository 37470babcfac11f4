"""Compiles and runs tests/cpu/test_fft_core.cc: the GPU Fourier kernels' index/phase algebra
(atlas_b200/csrc/fft_core.cuh is __host__ __device__) executed on the CPU against a naive DFT."""
import os
import subprocess
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fft_core_on_cpu():
    src = os.path.join(REPO, "tests", "cpu", "test_fft_core.cc")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "t")
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-o", exe, src], check=True)
        r = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stdout
        assert "ALL OK" in r.stdout


def test_fft2_kernel_bodies_under_thread_emulation():
    """The v2 (register-tiled) Fourier kernel bodies of atlas_b200/csrc/fft2_core.cuh, compiled for the host and run
    with one OS thread per CUDA thread (tests/cpu/test_fft2_emul.cc): every block-level radix, both directions,
    the filter-table construction, against a naive DFT."""
    src = os.path.join(REPO, "tests", "cpu", "test_fft2_emul.cc")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "t2")
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++20", "-pthread", "-o", exe, src], check=True)
        r = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stdout
        assert "ALL OK" in r.stdout


def test_push_work_queue_protocol_on_host_threads():
    """The work queue through which the blocks of the sharded direct Fourier stage share the shipping of completed latitude
    pairs (finish_pair_and_push, fourier.cu), modelled with std::atomic on a thread pool: every chunk shipped exactly once."""
    src = os.path.join(REPO, "tests", "cpu", "test_push_queue.cc")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "t3")
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-pthread", "-o", exe, src], check=True)
        r = subprocess.run([exe], stdout=subprocess.PIPE, text=True, timeout=300)
        assert r.returncode == 0, r.stdout
        assert "ALL OK" in r.stdout
