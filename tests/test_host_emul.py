"""CPU check of the CUDA algorithm's arithmetic and of the product's host epilogue.

tests/host_emul.cpp runs the per-lane bit-plane functions the kernels are built from
(simdjson_b200/csrc/sjb200_bits.cuh) in the kernels' tile/warp/lane decomposition and the
product's finish logic (sjb200_finish.cpp) against the oracle, on seeded adversarial inputs.
No GPU involved; the GPU parity tests live in test_gpu_parity.py."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_emulation_matches_oracle(tmp_path):
    exe = str(tmp_path / "host_emul")
    src = [os.path.join(ROOT, "tests", "host_emul.cpp"), os.path.join(ROOT, "simdjson_b200", "csrc", "sjb200_finish.cpp")]
    inc = ["-I", os.path.join(ROOT, "simdjson_b200", "csrc"), "-I", os.path.join(ROOT, "oracle")]
    subprocess.check_call(["gcc", "-O2", "-c", os.path.join(ROOT, "oracle", "sj_oracle.c"), "-o", str(tmp_path / "o.o")])
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", *inc, *src, str(tmp_path / "o.o"), "-o", exe])
    out = subprocess.run([exe, "12000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-4000:]
    assert "host emulation OK" in out.stdout
