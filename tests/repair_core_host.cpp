// Host build of ria_b200/csrc/frame_repair_core.h for tests/test_repair_core_cpu.py (test
// infrastructure): the same source the device runs, checked against libstdc++'s std::sort and
// against the unmodified reference's decodeFixedFrame.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../ria_b200/csrc/frame_repair_core.h"

extern "C" {

void rc_sort(float* key, uint16_t* val, int n) { ria_repair::libstdcxx_sort(key, val, n); }

// the library's own std::sort on the reference's element type and comparator (frame_v2.cpp:1725, 1751-1752)
void rc_std_sort(float* key, uint16_t* val, int n) {
    struct SuspectBit { size_t frame_bit; float abs_llr; };
    std::vector<SuspectBit> v(n);
    for (int i = 0; i < n; ++i) v[i] = {val[i], key[i]};
    std::sort(v.begin(), v.end(), [](const SuspectBit& a, const SuspectBit& b) { return a.abs_llr < b.abs_llr; });
    for (int i = 0; i < n; ++i) { key[i] = v[i].abs_llr; val[i] = static_cast<uint16_t>(v[i].frame_bit); }
}

static ria_repair::Frame make_frame(uint8_t* data, int bpc) {
    ria_repair::Frame f;
    for (int c = 0; c < 4; ++c) f.cw[c] = data + c * bpc;
    f.bpc = bpc;
    return f;
}

int rc_frame_valid(uint8_t* data, int bpc) {
    uint8_t tmp[ria_repair::kMaxFrameBytes];
    return ria_repair::frame_valid(make_frame(data, bpc), tmp) ? 1 : 0;
}

// data [4][bpc] in/out, soft [4][648]; returns 1 when a bit-flip search recovered the frame
int rc_repair_bitflips(uint8_t* data, int bpc, const float* soft) {
    using namespace ria_repair;
    std::vector<uint16_t> deltas(kMaxBits), val(kMaxBits);
    std::vector<float> key(kMaxBits);
    uint8_t frame[kMaxFrameBytes], trial[kMaxFrameBytes];
    Scratch s{deltas.data(), key.data(), val.data(), frame, trial};
    return repair_bitflips(make_frame(data, bpc), soft, s) ? 1 : 0;
}

}  // extern "C"
