"""GPU: the ultra::IWaveform drop-ins (include/ria_b200_adapters.hpp, RIA_WITH_ULTRA) next to the reference's
own OFDMChirpWaveform / MCDPSKWaveform on identical samples.

oracle/_ref/waveform_harness is tests/waveform_harness.cpp compiled against the reference's headers and objects
(oracle/Makefile `harness`, built where /root/reference exists; the binary travels to the GPU box like the other
checkers).  It drives both objects through detectSync / detectDataSync / setFrequencyOffset /
setAbsoluteTrainingPosition / process / getSoftBits / status getters / reset the way a StreamingDecoder does,
for chirp-acquired and connected-mode frames of both waveforms."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "waveform_harness")


def test_iwaveform_dropins_match_the_reference_waveforms(ria_lib):
    if not os.path.exists(HARNESS):
        if not os.path.isdir("/root/reference"):
            pytest.skip("oracle/_ref/waveform_harness was not built and /root/reference is absent")
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "harness"], check=True)
    r = subprocess.run([HARNESS], capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-4000:]
    assert r.stdout.strip().splitlines()[-1].startswith("PASS")
