"""Replays the CUDA kernels' per-thread pass functions on the CPU
(tests/cpu_emul/emul_ntt.cpp compiles hexl-fpga_b200/csrc/ntt_core.cuh as plain
C++) and checks them against the oracle: index math, swizzle, twiddle indexing,
exact wrap-around arithmetic on garbage inputs, and the generic-modulus divisor
arithmetic of the dyadic kernel."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kernel_pass_structure_matches_oracle(tmp_path):
    import oracle_binding as ob

    ob.oracle()
    exe = str(tmp_path / "emul_ntt")
    subprocess.run(["g++", "-O2", "-std=c++17", "-x", "c++", os.path.join(ROOT, "tests/cpu_emul/emul_ntt.cpp"),
                    "-o", exe, "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                    "-Wl,-rpath," + os.path.join(ROOT, "oracle")], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "EMUL OK" in out, out[-2000:]
