"""CPU: ria_b200/csrc/rn_math.h restates the glibc float functions the reference calls (atan2f,
sinf, cosf, sincosf).  The same header compiles for the device; here it is compiled for the host
and compared bit for bit with the container's libm."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_restated_glibc_functions_match_libm_bitwise(tmp_path):
    exe = tmp_path / "rn_math_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", str(exe),
                    os.path.join(ROOT, "tests", "rn_math_check.cpp")], check=True)
    out = subprocess.run([str(exe), "3000000"], check=True, capture_output=True, text=True).stdout.split()
    res = {out[i]: int(out[i + 1]) for i in range(0, len(out), 2)}
    assert res["n"] == 3000000
    for k in ("bad_sin", "bad_cos", "bad_sincos", "bad_atan2", "bad_large", "special_bad", "bad_log"):
        assert res[k] == 0, res
