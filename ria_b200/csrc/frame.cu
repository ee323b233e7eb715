// Fixed 4-codeword frame decode (first pass of v2::decodeFixedFrame) and the fused OFDM receive
// chain entry points.
//
//   ria_frame_decode_batch_dev : FrameInterleaver::deinterleave + ChannelInterleaver::deinterleave
//                                (fused into the LDPC kernel's load, ldpc.cu) -> 4 x decodeSoft ->
//                                reassemble + v2::parseHeader + CRC-16 (this file)
//   ria_ofdm_rx_frames_dev/host: OFDM presynced demod (ofdm.cu) -> frame decode, one call
//
// Reference: src/protocol/frame_v2.cpp:1335-1385, 1548-1556 (decodeFixedFrame first pass),
// :115-128 (CRC-16/CCITT-FALSE), :1195-1253 (parseHeader), :555-600 (frame CRC),
// src/protocol/frame_v2.hpp:222-228 (isControlFrame), :676-692 (bytes per codeword).

#include <cstdlib>

#include "ofdm_tables.h"

namespace ria {

namespace {

constexpr int kFrameBits = 4 * RIA_LDPC_N;     // FrameInterleaver::TOTAL_FRAME_BITS = 2592
constexpr int kInfoStride = 72;                // per-codeword info bytes in the scratch buffer

}  // namespace

// LDPCCodec::getRecommendedIterations, src/fec/ldpc_codec.hpp:86-96
int recommended_ldpc_iterations(int rate) {
    switch (rate) {
        case RIA_R3_4: return 60;
        case RIA_R2_3: return 70;
        case RIA_R1_2: return 80;
        case RIA_R1_3: return 60;
        case RIA_R1_4: return 50;
        default: return 50;
    }
}

namespace {
int recommended_iterations(int rate) { return recommended_ldpc_iterations(rate); }

// CRC-16/CCITT-FALSE (frame_v2.cpp:115-128), one table lookup per byte; the 256-entry table is built
// in shared memory by the CTA (entry b = the register after byte b from a zero start).
__device__ __forceinline__ void crc16_build_table(uint16_t* tab) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) {
        unsigned r = static_cast<unsigned>(b) << 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) r = ((r & 0x8000u) ? ((r << 1) ^ 0x1021u) : (r << 1)) & 0xFFFFu;
        tab[b] = static_cast<uint16_t>(r);
    }
    __syncthreads();
}
__device__ __forceinline__ uint16_t crc16_dev(const uint16_t* __restrict__ tab, const uint8_t* d, int len) {
    unsigned crc = 0xFFFFu;
    for (int i = 0; i < len; ++i) crc = ((crc << 8) & 0xFFFFu) ^ tab[((crc >> 8) ^ d[i]) & 0xFFu];
    return static_cast<uint16_t>(crc);
}

__device__ __forceinline__ bool is_control_frame(uint8_t t) {
    return t == 0x10 || t == 0x11 || t == 0x16 || t == 0x17 || t == 0x20 || t == 0x21 || t == 0x15 || t == 0x40;
}

// Header fields and CRC flags of one reassembled frame `h` (4 x bpc bytes, failed codewords zeroed):
// v2::parseHeader on codeword 0 (needs >= 20 bytes and a decoded CW0, reassemble():1030-1051) and the
// frame CRC of DataFrame::deserialize (:590-596).
__device__ __forceinline__ void parse_frame(const uint16_t* __restrict__ crc_tab, const uint8_t* h, int bpc, ria_frame_status& st) {
    if (!(st.cw_ok[0] && bpc >= 20)) return;
    const uint16_t magic = static_cast<uint16_t>((h[0] << 8) | h[1]);
    if (magic != 0x554C) return;
    st.type = h[2];
    st.seq = static_cast<uint16_t>((h[4] << 8) | h[5]);
    st.src_hash = (static_cast<uint32_t>(h[6]) << 16) | (static_cast<uint32_t>(h[7]) << 8) | h[8];
    st.dst_hash = (static_cast<uint32_t>(h[9]) << 16) | (static_cast<uint32_t>(h[10]) << 8) | h[11];
    if (is_control_frame(st.type)) {
        const uint16_t rx = static_cast<uint16_t>((h[18] << 8) | h[19]);
        if (rx == crc16_dev(crc_tab, h, 18)) { st.header_valid = 1; st.total_cw = 1; st.payload_len = 0; }
        return;
    }
    st.total_cw = h[12];
    st.payload_len = static_cast<uint16_t>((h[13] << 8) | h[14]);
    const uint16_t rx = static_cast<uint16_t>((h[15] << 8) | h[16]);
    if (rx == crc16_dev(crc_tab, h, 15)) st.header_valid = 1;
    // Frame CRC over CodewordStatus::reassemble()'s view of the four chunks (:960-985), not over their plain
    // concatenation: a chunk 1..3 that starts with 0xD5 is taken for a marked codeword (DATA_CW_MARKER,
    // :974) and loses its first two bytes, so such a frame comes out short or shifted and fails -- the same
    // rule ria_repair::reassemble applies, so the first-pass-only and the full decode report the same flags.
    // Every codeword that contributes bytes must have decoded.
    const int expected = 17 + st.payload_len + 2;
    if (st.header_valid && expected <= 4 * bpc) {
        bool have = true;
        int n = 0;
        unsigned crc = 0xFFFFu, stored = 0;
        for (int c = 0; c < 4 && n < expected; ++c) {
            if (!st.cw_ok[c]) have = false;
            const uint8_t* chunk = h + c * bpc;
            const int skip = (c > 0 && bpc >= 2 && chunk[0] == 0xD5) ? 2 : 0;
            const int take = min(expected - n, bpc - skip);
            for (int b = 0; b < take; ++b, ++n) {
                const uint8_t v = chunk[skip + b];
                if (n < expected - 2) crc = ((crc << 8) & 0xFFFFu) ^ crc_tab[((crc >> 8) ^ v) & 0xFFu];
                else stored = (stored << 8) | v;
            }
        }
        if (have && n == expected) st.frame_crc_ok = (stored == crc) ? 1 : 0;
    }
}

// Reassemble the 4 x bytes_per_cw info bytes (failed codewords stay zero, CodewordStatus::data is only
// filled on success), parse the header and check the CRCs.  Byte work over 288 B in / 280 B out per
// frame: a warp takes 32 consecutive frames, moves their bytes through padded shared-memory rows with
// coalesced 32-bit accesses (rows of 73 / 61 / 11 words: every lane on its own bank), and each lane
// parses its frame out of shared memory.  (One thread per frame straight on global memory was LSU-bound:
// 32 scattered byte accesses per instruction.)
constexpr int kStatusWarps = 2;
constexpr int kInWords = kInfoStride;             // 4 x 72 B = 72 words per frame
constexpr int kInRow = kInWords + 1, kOutRow = 61, kStRow = 11;
static_assert(sizeof(ria_frame_status) == 40, "status rows are copied as 10 words");

__global__ void __launch_bounds__(kStatusWarps * 32)
frame_status_kernel(const uint8_t* __restrict__ info, const uint8_t* __restrict__ ok,
                    const int32_t* __restrict__ iters, const uint8_t* __restrict__ attempt,
                    const uint8_t* __restrict__ repair, long long n_frames, int bpc,
                    uint8_t* __restrict__ data, ria_frame_status* __restrict__ status) {
    __shared__ uint16_t crc_tab[256];
    __shared__ uint32_t in_rows[kStatusWarps][32 * kInRow];
    __shared__ uint32_t out_rows[kStatusWarps][32 * kOutRow];
    __shared__ uint32_t st_rows[kStatusWarps][32 * kStRow];
    crc16_build_table(crc_tab);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long f0 = (blockIdx.x * static_cast<long long>(kStatusWarps) + warp) * 32;
    if (f0 >= n_frames) return;
    const int nv = static_cast<int>(n_frames - f0 < 32 ? n_frames - f0 : 32);
    uint32_t* in = in_rows[warp];
    uint32_t* outw = out_rows[warp];
    uint32_t* stw = st_rows[warp];
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(info + f0 * 4 * kInfoStride);
        for (int g = lane; g < nv * kInWords; g += 32) in[(g / kInWords) * kInRow + (g % kInWords)] = __ldcs(src + g);
    }
    __syncwarp();
    if (lane < nv) {
        const long long f = f0 + lane;
        const uint8_t* row = reinterpret_cast<const uint8_t*>(in + lane * kInRow);
        uint8_t* h = reinterpret_cast<uint8_t*>(outw + lane * kOutRow);
        ria_frame_status st;
        memset(&st, 0, sizeof st);
        const uchar4 okv = reinterpret_cast<const uchar4*>(ok)[f];
        const int4 itv = reinterpret_cast<const int4*>(iters)[f];
        const uint8_t oks[4] = {okv.x, okv.y, okv.z, okv.w};
        const int its[4] = {itv.x, itv.y, itv.z, itv.w};
        uchar4 atv = make_uchar4(0, 0, 0, 0);
        if (attempt) atv = reinterpret_cast<const uchar4*>(attempt)[f];
        const uint8_t ats[4] = {atv.x, atv.y, atv.z, atv.w};
        bool all = true;
        for (int c = 0; c < 4; ++c) {
            st.cw_ok[c] = oks[c];
            st.cw_iters[c] = its[c];
            all = all && oks[c];
            if (attempt && oks[c] && ats[c] >= 1 && ats[c] <= 38) {
                st.ladder_cw_mask |= static_cast<uint8_t>(1u << c);
                if (ats[c] > st.ladder_max_attempt) st.ladder_max_attempt = ats[c];
            }
            for (int b = 0; b < bpc; ++b) h[c * bpc + b] = oks[c] ? row[c * kInfoStride + b] : 0;
        }
        st.all_ok = all ? 1 : 0;
        if (repair) st.fp_repair = repair[f];
        parse_frame(crc_tab, h, bpc, st);
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(&st);
#pragma unroll
        for (int w = 0; w < 10; ++w) stw[lane * kStRow + w] = sw[w];
    }
    __syncwarp();
    {
        uint32_t* dst = reinterpret_cast<uint32_t*>(data + f0 * 4 * bpc);       // 4 * bpc bytes = bpc words per frame
        for (int g = lane; g < nv * bpc; g += 32) __stcs(dst + g, outw[(g / bpc) * kOutRow + (g % bpc)]);
        uint32_t* sd = reinterpret_cast<uint32_t*>(status + f0);
        for (int g = lane; g < nv * 10; g += 32) sd[g] = stw[(g / 10) * kStRow + (g % 10)];
    }
}

// the same with one thread per frame on global memory: for caller buffers that are not 4-byte aligned
__global__ void frame_status_bytes_kernel(const uint8_t* __restrict__ info, const uint8_t* __restrict__ ok,
                                          const int32_t* __restrict__ iters, const uint8_t* __restrict__ attempt,
                                          const uint8_t* __restrict__ repair, long long n_frames, int bpc,
                                          uint8_t* __restrict__ data, ria_frame_status* __restrict__ status) {
    __shared__ uint16_t crc_tab[256];
    crc16_build_table(crc_tab);
    const long long f = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (f >= n_frames) return;
    uint8_t* out = data + f * 4 * bpc;
    ria_frame_status st;
    memset(&st, 0, sizeof st);
    bool all = true;
    for (int c = 0; c < 4; ++c) {
        const long long cw = f * 4 + c;
        st.cw_ok[c] = ok[cw];
        st.cw_iters[c] = iters[cw];
        all = all && ok[cw];
        if (attempt && ok[cw] && attempt[cw] >= 1 && attempt[cw] <= 38) {
            st.ladder_cw_mask |= static_cast<uint8_t>(1u << c);
            if (attempt[cw] > st.ladder_max_attempt) st.ladder_max_attempt = attempt[cw];
        }
        for (int b = 0; b < bpc; ++b) out[c * bpc + b] = ok[cw] ? info[cw * kInfoStride + b] : 0;
    }
    st.all_ok = all ? 1 : 0;
    if (repair) st.fp_repair = repair[f];
    parse_frame(crc_tab, out, bpc, st);
    status[f] = st;
}

int bytes_per_codeword(int rate) {
    int k = 0;
    if (ria_ldpc_params(rate, &k, nullptr, nullptr) != RIA_OK) return -1;
    return k / 8;
}

// scratch layout for n frames: info [4n][64] | ok [4n] | iters [4n] (+ llr [n][llr_stride] for the chain)
struct Scratch {
    uint8_t* info; uint8_t* ok; int32_t* iters; float* llr; int32_t* n_llr;
    uint8_t* attempt; int* fail_list;           // retry ladder / repair only
    uint8_t* repair;
};

int carve_scratch(ria_ctx* ctx, int64_t n_frames, int llr_stride, Scratch& s) {
    const size_t n_cw = static_cast<size_t>(n_frames) * 4;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
    const size_t o_info = take(n_cw * kInfoStride);
    const size_t o_ok = take(n_cw);
    const size_t o_it = take(n_cw * 4);
    const size_t o_llr = take(static_cast<size_t>(n_frames) * llr_stride * 4);
    const size_t o_nl = take(static_cast<size_t>(n_frames) * 4);
    const bool ladder = (ctx->decode_flags & RIA_DECODE_RETRY_LADDER) != 0;
    const bool repair = (ctx->decode_flags & RIA_DECODE_FP_REPAIR) != 0;
    const size_t o_at = take(ladder ? n_cw : 0);
    const size_t o_fl = take((ladder || repair) ? static_cast<size_t>(n_frames) * 4 : 0);
    const size_t o_rp = take(repair ? static_cast<size_t>(n_frames) : 0);
    int rc = ensure_scratch(ctx, off);
    if (rc != RIA_OK) return rc;
    unsigned char* b = static_cast<unsigned char*>(ctx->scratch);
    s.info = b + o_info; s.ok = b + o_ok; s.iters = reinterpret_cast<int32_t*>(b + o_it);
    s.llr = reinterpret_cast<float*>(b + o_llr); s.n_llr = reinterpret_cast<int32_t*>(b + o_nl);
    s.attempt = ladder ? b + o_at : nullptr;
    s.fail_list = (ladder || repair) ? reinterpret_cast<int*>(b + o_fl) : nullptr;
    s.repair = repair ? b + o_rp : nullptr;
    return RIA_OK;
}

int frame_decode_impl(ria_ctx* ctx, int rate, int use_ci, int bits_per_symbol, const float* soft_dev,
                      int32_t soft_stride, int64_t n_frames, const Scratch& s, uint8_t* data_dev,
                      ria_frame_status* status_dev) {
    const int bpc = bytes_per_codeword(rate);
    if (bpc < 0) return set_error(ctx, RIA_E_INVAL, "frame: bad rate %d", rate);
    int step = 0;
    if (use_ci) {
        if (bits_per_symbol <= 0) return set_error(ctx, RIA_E_INVAL, "frame: bits_per_symbol must be > 0");
        step = channel_interleaver_step(bits_per_symbol, RIA_LDPC_N);
    }
    int rc = ldpc_launch(ctx, rate, recommended_iterations(rate), 0.9375f, soft_dev, n_frames * 4,
                         1, soft_stride, step, s.info, kInfoStride, s.ok, s.iters);
    if (rc != RIA_OK) return rc;
    if (s.attempt) {
        // retry ladder (frame_v2.cpp:1389-1546) on the frames with a failed codeword
        rc = ldpc_retry_launch(ctx, rate, recommended_iterations(rate), soft_dev, n_frames, 1, soft_stride, step,
                               s.info, kInfoStride, s.ok, s.iters, s.attempt, s.fail_list);
        if (rc != RIA_OK) return rc;
    }
    if (s.repair) {
        // false-positive repair (frame_v2.cpp:1558-1916) on the frames that decoded but do not verify
        rc = frame_repair_launch(ctx, rate, recommended_iterations(rate), soft_dev, n_frames, soft_stride, step,
                                 s.info, kInfoStride, s.ok, s.repair, s.fail_list);
        if (rc != RIA_OK) return rc;
    }
    time_begin(ctx, KK_FRAME_STATUS);
    const bool aligned = ((reinterpret_cast<uintptr_t>(data_dev) | reinterpret_cast<uintptr_t>(status_dev)) & 3) == 0 && bpc <= 60;
    if (aligned) {
        const int per_cta = kStatusWarps * 32;
        const unsigned blocks = static_cast<unsigned>((n_frames + per_cta - 1) / per_cta);
        frame_status_kernel<<<blocks, per_cta, 0, ctx->stream>>>(s.info, s.ok, s.iters, s.attempt, s.repair, n_frames, bpc, data_dev, status_dev);
    } else {
        const int threads = 128;
        const unsigned blocks = static_cast<unsigned>((n_frames + threads - 1) / threads);
        frame_status_bytes_kernel<<<blocks, threads, 0, ctx->stream>>>(s.info, s.ok, s.iters, s.attempt, s.repair, n_frames, bpc, data_dev, status_dev);
    }
    time_end(ctx);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return RIA_OK;
}

}  // namespace
}  // namespace ria

extern "C" int ria_frame_decode_batch_dev(ria_ctx* ctx, int rate, int use_channel_interleave,
                                          int bits_per_symbol, const float* soft_dev, int32_t soft_stride,
                                          int64_t n_frames, uint8_t* data_dev, ria_frame_status* status_dev) {
    using namespace ria;
    if (!ctx) return RIA_E_INVAL;
    if (n_frames < 0) return set_error(ctx, RIA_E_INVAL, "frame: negative size");
    if (n_frames == 0) return RIA_OK;
    if (!soft_dev || !data_dev || !status_dev) return set_error(ctx, RIA_E_INVAL, "frame: null buffer");
    if (soft_stride < kFrameBits) return set_error(ctx, RIA_E_INVAL, "frame: need >= 2592 soft bits per frame");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch s{};
    int rc = carve_scratch(ctx, n_frames, 0, s);
    if (rc != RIA_OK) return rc;
    return frame_decode_impl(ctx, rate, use_channel_interleave, bits_per_symbol, soft_dev, soft_stride, n_frames,
                             s, data_dev, status_dev);
}

extern "C" int ria_frame_decode_batch_host(ria_ctx* ctx, int rate, int use_channel_interleave,
                                           int bits_per_symbol, const float* soft, int32_t soft_stride,
                                           int64_t n_frames, uint8_t* data, ria_frame_status* status) {
    using namespace ria;
    if (!ctx) return RIA_E_INVAL;
    if (n_frames < 0) return set_error(ctx, RIA_E_INVAL, "frame: negative size");
    if (n_frames == 0) return RIA_OK;
    if (!soft || !data || !status) return set_error(ctx, RIA_E_INVAL, "frame: null buffer");
    if (soft_stride < kFrameBits) return set_error(ctx, RIA_E_INVAL, "frame: need >= 2592 soft bits per frame");
    const int bpc = bytes_per_codeword(rate);
    if (bpc < 0) return set_error(ctx, RIA_E_INVAL, "frame: bad rate %d", rate);
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t chunk = n_frames < 65536 ? n_frames : 65536;      // staging sized to the call, not to the maximum
    const int32_t dstride = (kFrameBits + 3) & ~3;
    const size_t in_b = static_cast<size_t>(chunk) * dstride * sizeof(float);
    const size_t out_b = static_cast<size_t>(chunk) * (4 * bpc + sizeof(ria_frame_status));
    int rc = ensure_stage(ctx, 0, in_b + out_b + 512, 0);
    if (rc != RIA_OK) return rc;
    cudaStream_t s = ctx->stream;
    unsigned char* base = static_cast<unsigned char*>(ctx->stage_dev[0]);
    float* d_soft = reinterpret_cast<float*>(base);
    ria_frame_status* d_st = reinterpret_cast<ria_frame_status*>(base + in_b);
    uint8_t* d_data = reinterpret_cast<uint8_t*>(d_st + chunk);
    for (int64_t off = 0; off < n_frames; off += chunk) {
        const int64_t n = (n_frames - off < chunk) ? (n_frames - off) : chunk;
        RIA_CUDA(ctx, cudaMemcpy2DAsync(d_soft, static_cast<size_t>(dstride) * 4, soft + off * soft_stride,
                                        static_cast<size_t>(soft_stride) * 4, static_cast<size_t>(kFrameBits) * 4,
                                        static_cast<size_t>(n), cudaMemcpyHostToDevice, s));
        rc = ria_frame_decode_batch_dev(ctx, rate, use_channel_interleave, bits_per_symbol, d_soft, dstride, n, d_data, d_st);
        if (rc != RIA_OK) return rc;
        RIA_CUDA(ctx, cudaMemcpyAsync(data + off * 4 * bpc, d_data, static_cast<size_t>(n) * 4 * bpc, cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaMemcpyAsync(status + off, d_st, static_cast<size_t>(n) * sizeof(ria_frame_status), cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaStreamSynchronize(s));
    }
    return RIA_OK;
}

extern "C" int ria_ofdm_rx_frames_dev(ria_ctx* ctx, const ria_modem_config* cfg, int rate,
                                      int use_channel_interleave,
                                      const float* samples_dev, int64_t frame_stride, int32_t frame_len,
                                      const float* cfo_hz_dev, const float* phase_dev, int64_t n_frames,
                                      uint8_t* data_dev, ria_frame_status* status_dev, float* snr_db_dev) {
    using namespace ria;
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_frames < 0) return set_error(ctx, RIA_E_INVAL, "rx: negative size");
    if (n_frames == 0) return RIA_OK;
    if (!samples_dev || !data_dev || !status_dev) return set_error(ctx, RIA_E_INVAL, "rx: null buffer");
    if (const char* err = ofdm_config_error(*cfg)) return set_error(ctx, RIA_E_UNSUPPORTED, "ofdm: %s", err);
    const int sym = ofdm_symbol_samples(*cfg);
    const int nd = ria_ofdm_data_carriers(cfg);
    const int bps = nd * ofdm_bits_per_carrier(cfg->modulation);
    const int n_sym = frame_len / sym;
    const int n_llr = (n_sym > 2 ? n_sym - 2 : 0) * bps;
    if (n_llr < kFrameBits)
        return set_error(ctx, RIA_E_INVAL, "rx: frame_len %d yields %d soft bits, a 4-codeword frame needs 2592", frame_len, n_llr);
    const int llr_stride = (n_llr + 3) & ~3;
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch s{};
    int rc = carve_scratch(ctx, n_frames, llr_stride, s);
    if (rc != RIA_OK) return rc;
    rc = ria_ofdm_presynced_batch_dev(ctx, cfg, samples_dev, frame_stride, frame_len, cfo_hz_dev, phase_dev, n_frames,
                                      s.llr, llr_stride, s.n_llr, snr_db_dev, nullptr, nullptr);
    if (rc != RIA_OK) return rc;
    return frame_decode_impl(ctx, rate, use_channel_interleave, bps, s.llr, llr_stride, n_frames, s, data_dev, status_dev);
}

extern "C" int ria_ofdm_rx_frames_host(ria_ctx* ctx, const ria_modem_config* cfg, int rate,
                                       int use_channel_interleave,
                                       const float* samples, int64_t frame_stride, int32_t frame_len,
                                       const float* cfo_hz, const float* phase, int64_t n_frames,
                                       uint8_t* data, ria_frame_status* status, float* snr_db) {
    using namespace ria;
    if (!ctx || !cfg) return RIA_E_INVAL;
    if (n_frames < 0 || frame_len <= 0 || frame_stride < frame_len) return set_error(ctx, RIA_E_INVAL, "rx: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!samples || !data || !status) return set_error(ctx, RIA_E_INVAL, "rx: null buffer");
    const int bpc = bytes_per_codeword(rate);
    if (bpc < 0) return set_error(ctx, RIA_E_INVAL, "rx: bad rate %d", rate);
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    // Chunked pipeline on the context stream with two staging buffers; the H2D of chunk c+1 is
    // issued on the copy stream so it overlaps the kernels of chunk c.
    // 440 MB per upload at 8192 frames (smaller chunks measured slower: 0.86 vs 0.91 M frames/s); a small
    // call (the batch-1 adapters) only stages what it brings
    const int64_t chunk = n_frames < 8192 ? n_frames : 8192;
    const size_t in_b = static_cast<size_t>(chunk) * frame_len * sizeof(float);
    const size_t aux_b = static_cast<size_t>(chunk) * 8;                       // cfo + phase
    const size_t out_b = static_cast<size_t>(chunk) * (4 * bpc + sizeof(ria_frame_status) + 4);
    for (int b = 0; b < 2; ++b) {
        int rc = ensure_stage(ctx, b, in_b + aux_b + out_b + 1024, 0);
        if (rc != RIA_OK) return rc;
    }
    cudaStream_t s = ctx->stream, cs = ctx->copy_stream;
    if (const char* err = ofdm_config_error(*cfg)) return set_error(ctx, RIA_E_UNSUPPORTED, "ofdm: %s", err);
    const int sym = ofdm_symbol_samples(*cfg), cp = ofdm_cyclic_prefix(*cfg);
    static const bool skip_cp = [] { const char* e = std::getenv("RIA_H2D_FULL_SYMBOLS"); return !(e && e[0] == '1'); }();
    // stage_ev[0..1]: "compute of buffer b finished"  stage_ev[2..3]: "H2D of buffer b finished"
    int buf = 0;
    for (int64_t off = 0; off < n_frames; off += chunk, buf ^= 1) {
        const int64_t n = (n_frames - off < chunk) ? (n_frames - off) : chunk;
        unsigned char* base = static_cast<unsigned char*>(ctx->stage_dev[buf]);
        float* d_samp = reinterpret_cast<float*>(base);
        float* d_cfo = reinterpret_cast<float*>(base + in_b);
        float* d_ph = d_cfo + chunk;
        ria_frame_status* d_st = reinterpret_cast<ria_frame_status*>(base + in_b + aux_b);
        float* d_snr = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(d_st) + static_cast<size_t>(chunk) * sizeof(ria_frame_status));
        uint8_t* d_data = reinterpret_cast<uint8_t*>(d_snr + chunk);
        // buffer reuse: wait until the kernels + D2H that used it two chunks ago are done
        RIA_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->stage_ev[buf], 0));
        if (frame_stride == frame_len && frame_len % sym == 0 && skip_cp) {
            // Only the FFT window of each symbol is ever read by the demodulator (the cyclic prefix and the
            // guard are skipped by the kernels), so only the windows cross PCIe: one 2-D copy whose rows are
            // the symbols of the chunk (pitch = symbol, width = FFT size, starting at the prefix length).
            RIA_CUDA(ctx, cudaMemcpy2DAsync(d_samp + cp, static_cast<size_t>(sym) * 4, samples + off * frame_stride + cp,
                                            static_cast<size_t>(sym) * 4, static_cast<size_t>(cfg->fft_size) * 4,
                                            static_cast<size_t>(n) * (frame_len / sym), cudaMemcpyHostToDevice, cs));
        } else if (frame_stride == frame_len) {
            RIA_CUDA(ctx, cudaMemcpyAsync(d_samp, samples + off * frame_stride, static_cast<size_t>(n) * frame_len * sizeof(float),
                                          cudaMemcpyHostToDevice, cs));
        } else {
            RIA_CUDA(ctx, cudaMemcpy2DAsync(d_samp, static_cast<size_t>(frame_len) * 4, samples + off * frame_stride,
                                            static_cast<size_t>(frame_stride) * 4, static_cast<size_t>(frame_len) * 4,
                                            static_cast<size_t>(n), cudaMemcpyHostToDevice, cs));
        }
        if (cfo_hz) RIA_CUDA(ctx, cudaMemcpyAsync(d_cfo, cfo_hz + off, static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice, cs));
        if (phase) RIA_CUDA(ctx, cudaMemcpyAsync(d_ph, phase + off, static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice, cs));
        RIA_CUDA(ctx, cudaEventRecord(ctx->stage_ev[2 + buf], cs));
        RIA_CUDA(ctx, cudaStreamWaitEvent(s, ctx->stage_ev[2 + buf], 0));
        int rc = ria_ofdm_rx_frames_dev(ctx, cfg, rate, use_channel_interleave, d_samp, frame_len, frame_len,
                                        cfo_hz ? d_cfo : nullptr, phase ? d_ph : nullptr, n, d_data, d_st, d_snr);
        if (rc != RIA_OK) return rc;
        RIA_CUDA(ctx, cudaMemcpyAsync(data + off * 4 * bpc, d_data, static_cast<size_t>(n) * 4 * bpc, cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaMemcpyAsync(status + off, d_st, static_cast<size_t>(n) * sizeof(ria_frame_status), cudaMemcpyDeviceToHost, s));
        if (snr_db) RIA_CUDA(ctx, cudaMemcpyAsync(snr_db + off, d_snr, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, s));
        RIA_CUDA(ctx, cudaEventRecord(ctx->stage_ev[buf], s));
    }
    RIA_CUDA(ctx, cudaStreamSynchronize(s));
    return RIA_OK;
}


// ---------------------------------------------------------------------------------------------
// PING energy test of StreamingDecoder::decodeCurrentFrame (src/gui/modem/streaming_decoder.cpp:1127-1160, 1219-1229)
// ---------------------------------------------------------------------------------------------
namespace ria {
namespace {

// One warp per frame: the samples are squared by all lanes (the product is exact per sample) and summed in sample order:
// lane 0 walks the training region, lane 1 the data region behind it, both out of a shared staging tile.
constexpr int kPingWarps = 4;
constexpr int kPingTile = 1024;

__global__ void __launch_bounds__(kPingWarps * 32)
ping_energy_kernel(const float* __restrict__ frames, long long stride, int frame_len, int training_skip, long long n,
                   float* __restrict__ out) {
    __shared__ __align__(16) float tile[kPingWarps][kPingTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* t = tile[warp];
    for (long long f = static_cast<long long>(blockIdx.x) * kPingWarps + warp; f < n; f += static_cast<long long>(gridDim.x) * kPingWarps) {
        const float* x = frames + f * stride;
        const int train_len = min(training_skip, frame_len);
        const int check_len = min(frame_len - train_len, 5000);
        float sums[2] = {0.0f, 0.0f};
        for (int region = 0; region < 2; ++region) {
            const int base = region ? train_len : 0;
            const int len = region ? check_len : train_len;
            float acc = 0.0f;
            for (int o = 0; o < len; o += kPingTile) {
                const int m = min(kPingTile, len - o);
                for (int i = lane; i < m; i += 32) { const float v = x[base + o + i]; t[i] = __fmul_rn(v, v); }
                __syncwarp();
                if (lane == 0) {
                    int i = 0;
                    for (; i + 4 <= m; i += 4) {
                        const float4 q = *reinterpret_cast<const float4*>(t + i);
                        acc = __fadd_rn(acc, q.x); acc = __fadd_rn(acc, q.y); acc = __fadd_rn(acc, q.z); acc = __fadd_rn(acc, q.w);
                    }
                    for (; i < m; ++i) acc = __fadd_rn(acc, t[i]);
                }
                __syncwarp();
            }
            sums[region] = acc;
        }
        if (lane == 0) {
            const float training_rms = train_len > 0 ? __fsqrt_rn(__fdiv_rn(sums[0], static_cast<float>(train_len))) : 0.0f;
            const float rms = check_len > 0 ? __fsqrt_rn(__fdiv_rn(sums[1], static_cast<float>(check_len))) : 0.0f;
            const float ratio = (training_rms > 0.001f) ? __fdiv_rn(rms, training_rms) : 0.0f;
            float4 r = make_float4(training_rms, rms, ratio, ratio < 0.6f ? 1.0f : 0.0f);
            *reinterpret_cast<float4*>(out + f * 4) = r;
        }
    }
}

}  // namespace
}  // namespace ria

extern "C" int ria_ping_energy_batch_dev(ria_ctx* ctx, const float* frames_dev, int64_t frame_stride, int32_t frame_len,
                                         int32_t training_skip, int64_t n_frames, float* out_dev) {
    using namespace ria;
    if (!ctx) return RIA_E_INVAL;
    if (n_frames < 0 || frame_len < 0 || frame_stride < frame_len || training_skip < 0)
        return set_error(ctx, RIA_E_INVAL, "ping energy: bad sizes");
    if (n_frames == 0) return RIA_OK;
    if (!frames_dev || !out_dev) return set_error(ctx, RIA_E_INVAL, "ping energy: null buffer");
    if (reinterpret_cast<uintptr_t>(out_dev) & 15) return set_error(ctx, RIA_E_INVAL, "ping energy: output must be 16-byte aligned");
    RIA_CUDA(ctx, cudaSetDevice(ctx->device));
    long long grid = (n_frames + kPingWarps - 1) / kPingWarps;
    const long long cap = static_cast<long long>(ctx->sm_count) * 8;
    if (grid > cap) grid = cap;
    ping_energy_kernel<<<static_cast<unsigned>(grid), kPingWarps * 32, 0, ctx->stream>>>(frames_dev, frame_stride, frame_len,
                                                                                          training_skip, n_frames, out_dev);
    RIA_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return RIA_OK;
}
