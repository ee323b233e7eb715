// TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" shim over the UNMODIFIED reference classes, compiled together with the
// reference sources where they lie under /root/reference (see oracle/Makefile) into
// oracle/_ref/libria_ref.so.  It exists so that tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs can call the reference's own implementation of the
// hot path through ctypes.  Nothing in ria_b200/ may link or load it.
//
// Every function here only marshals plain pointers into the reference's public API; no
// algorithm is restated in this file.

#include "ultra/types.hpp"
#include "ultra/fec.hpp"
#include "ultra/logging.hpp"

#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

using namespace ultra;

extern "C" {

// Silence the reference's INFO logging (hot functions log under a global
// mutex; any CPU timing must run at ERROR level).
void ref_quiet(void) { setLogLevel(LogLevel::ERROR); }

// ---------------------------------------------------------------------------------------------
// LDPC  (include/ultra/fec.hpp:21-81)
// ---------------------------------------------------------------------------------------------

// LDPCEncoder::encode (src/fec/ldpc_encoder.cpp:193-257). Returns number of coded bytes.
int ref_ldpc_encode(int rate, const uint8_t* data, int len, uint8_t* out, int out_cap) {
    LDPCEncoder enc(static_cast<CodeRate>(rate));
    Bytes coded = enc.encode(ByteSpan(data, static_cast<size_t>(len)));
    int n = static_cast<int>(coded.size());
    if (n > out_cap) return -n;
    std::memcpy(out, coded.data(), coded.size());
    return n;
}

struct RefLdpcDecoder {
    LDPCDecoder dec;
    explicit RefLdpcDecoder(CodeRate r) : dec(r) {}
};

void* ref_ldpc_decoder_new(int rate, int max_iter, float factor) {
    auto* d = new RefLdpcDecoder(static_cast<CodeRate>(rate));
    d->dec.setMaxIterations(max_iter);
    d->dec.setMinSumFactor(factor);
    return d;
}
void ref_ldpc_decoder_free(void* h) { delete static_cast<RefLdpcDecoder*>(h); }

// LDPCDecoder::decodeSoft (src/fec/ldpc_decoder.cpp:284-429) on one span of n_llr floats.
// Returns number of output bytes; *ok = lastDecodeSuccess(), *iters = lastIterations().
int ref_ldpc_decode_soft(void* h, const float* llr, int n_llr, uint8_t* out, int out_cap,
                         int* ok, int* iters) {
    auto* d = static_cast<RefLdpcDecoder*>(h);
    Bytes b = d->dec.decodeSoft(std::span<const float>(llr, static_cast<size_t>(n_llr)));
    *ok = d->dec.lastDecodeSuccess() ? 1 : 0;
    *iters = d->dec.lastIterations();
    int n = static_cast<int>(b.size());
    if (n > out_cap) return -n;
    if (n) std::memcpy(out, b.data(), b.size());
    return n;
}

// Batch convenience for the CPU baseline: n_cw independent codewords of 648 LLRs each,
// decoded one by one with the same decoder object (exactly what decodeFixedFrame does,
// src/protocol/frame_v2.cpp:1359-1385).  out is [n_cw][out_stride].
void ref_ldpc_decode_batch(void* h, const float* llr, int n_cw, uint8_t* out, int out_stride,
                           uint8_t* ok, int32_t* iters) {
    auto* d = static_cast<RefLdpcDecoder*>(h);
    for (int c = 0; c < n_cw; ++c) {
        Bytes b = d->dec.decodeSoft(std::span<const float>(llr + static_cast<size_t>(c) * 648, 648));
        ok[c] = d->dec.lastDecodeSuccess() ? 1 : 0;
        iters[c] = d->dec.lastIterations();
        size_t n = std::min(b.size(), static_cast<size_t>(out_stride));
        std::memset(out + static_cast<size_t>(c) * out_stride, 0, out_stride);
        std::memcpy(out + static_cast<size_t>(c) * out_stride, b.data(), n);
    }
}

}  // extern "C"
