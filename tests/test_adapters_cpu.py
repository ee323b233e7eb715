"""CPU: the C++ adapters compile stand-alone and, where the reference tree is present, against
the reference's own headers with RIA_WITH_ULTRA (ICodec inheritance)."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include "ria_b200_adapters.hpp"
int main() {
    // no GPU here: only check the types line up; constructing a Context would throw
    ria::LDPCCodec* c = nullptr; (void)c;
#ifdef RIA_WITH_ULTRA
    ultra::fec::ICodec* base = c; (void)base;   // usable wherever CodecFactory hands out an ICodec
#endif
    static_assert(ria::LDPCCodec::CODEWORD_BITS == 648, "");
#ifdef RIA_WITH_ULTRA
    static_assert(std::is_base_of<ultra::fec::ICodec, ria::LDPCCodec>::value, "must be an ICodec");
#endif
    return ria::LDPCCodec::getRecommendedIterations(RIA_R1_2) == 80 ? 0 : 1;
}
'''


def _compile(extra):
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.cpp")
        open(src, "w").write(SRC)
        exe = os.path.join(d, "t")
        cmd = ["g++", "-std=c++20", "-I", os.path.join(ROOT, "include")] + extra + [
            src, "-o", exe] + (["/root/reference/src/fec/ldpc_encoder.cpp"] if "-DRIA_WITH_ULTRA" in extra else []) + [
            os.path.join(ROOT, "ria_b200", "libria_b200.so"),
            "-Wl,-rpath," + os.path.join(ROOT, "ria_b200")]
        subprocess.run(cmd, check=True, capture_output=True)
        subprocess.run([exe], check=True)


def test_adapters_compile_standalone(ria_lib):
    _compile([])


def test_adapters_compile_against_reference_headers(ria_lib):
    if not os.path.isdir("/root/reference/src/fec"):
        pytest.skip("reference tree not present")
    _compile(["-DRIA_WITH_ULTRA", "-I/root/reference/include", "-I/root/reference/src"])


def test_waveform_dropins_are_iwaveforms_and_the_harness_links(ria_lib):
    """ria::OFDMChirpWaveform / ria::MCDPSKWaveform derive from the reference's waveform classes (hence from
    ultra::IWaveform) and ria::createWaveform hands them out as WaveformPtr: the harness that runs them on the GPU
    (tests/test_waveform_dropin_gpu.py) compiles and links against the reference objects."""
    if not os.path.isdir("/root/reference/src/waveform"):
        pytest.skip("reference tree not present")
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "harness"], check=True)
    assert os.path.exists(os.path.join(ROOT, "oracle", "_ref", "waveform_harness"))
