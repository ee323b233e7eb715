// Serial core of the "LDPC false positive recovery" of v2::decodeFixedFrame
// (src/protocol/frame_v2.cpp:1558-1916), host/device.  Everything here is integer / byte logic that
// the reference runs on one frame at a time; the device runs it on one lane per frame
// (frame_repair.cu) and tests/repair_core_check.cpp runs the same source on the host against the
// unmodified reference.
//
// Restated from the reference (file:line in each function):
//   CodewordStatus::reassemble / reassembleCodewords  frame_v2.cpp:1029-1063, 959-989
//   parseHeader                                       frame_v2.cpp:1195-1253
//   ControlFrame/DataFrame::deserialize (validity)    frame_v2.cpp:401-432, 555-600
//   CRC-16/CCITT-FALSE                                frame_v2.cpp:115-128
//   std::sort (libstdc++ bits/stl_algo.h introsort: median-of-3 partition, depth limit
//   2*floor(log2 n) with heap-sort fallback, threshold-16 final insertion sort) -- the suspect
//   list is sorted by |LLR| alone, so the order of equal keys is the library's, and it decides
//   which 30 suspects are searched.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define RC_HD __host__ __device__ inline
// crc16 / parse_header are real calls on the device: inlined into callers that hold the frame bytes in
// local or shared memory, nvcc 12.9 -O3 (sm_100a) produced a header check that failed on valid headers
// (observed on B200; the same source on global-memory bytes, and every host build, was correct).
#define RC_HD_NOINLINE __host__ __device__ __noinline__
#else
#define RC_HD static inline
#define RC_HD_NOINLINE static
#endif

namespace ria_repair {

constexpr int kMaxCwBytes = 60;          // bytes_per_cw <= 60 (R3/4; R5/6 has 67 -> unsupported here)
constexpr int kMaxFrameBytes = 4 * kMaxCwBytes;
constexpr int kMaxBits = kMaxFrameBytes * 8;

RC_HD_NOINLINE unsigned crc16(const uint8_t* d, int len) {
    unsigned crc = 0xFFFFu;                     // 32-bit register, masked: same values as the uint16_t original
    for (int i = 0; i < len; ++i) {
        crc ^= static_cast<unsigned>(d[i]) << 8;
        for (int b = 0; b < 8; ++b)
            crc = ((crc & 0x8000u) ? ((crc << 1) ^ 0x1021u) : (crc << 1)) & 0xFFFFu;
    }
    return crc;
}
// one zero byte through the (linear part of the) CRC register
RC_HD uint16_t crc_shift8(uint16_t s16) {
    unsigned s = s16;
    for (int b = 0; b < 8; ++b)
        s = ((s & 0x8000u) ? ((s << 1) ^ 0x1021u) : (s << 1)) & 0xFFFFu;
    return static_cast<uint16_t>(s);
}

RC_HD bool is_control_type(uint8_t t) {        // isControlFrame, frame_v2.hpp:222-228
    return t == 0x10 || t == 0x11 || t == 0x16 || t == 0x17 || t == 0x20 || t == 0x21 || t == 0x15 || t == 0x40;
}

struct Header { bool valid; bool is_control; int payload_len; };

// parseHeader on `len` bytes (needs >= 20)
RC_HD_NOINLINE Header parse_header(const uint8_t* d, int len) {
    Header h{false, false, 0};
    if (len < 20) return h;
    if (d[0] != 0x55 || d[1] != 0x4C) return h;
    h.is_control = is_control_type(d[2]);
    if (h.is_control) {
        const unsigned rx = (static_cast<unsigned>(d[18]) << 8) | d[19];
        if (rx != crc16(d, 18)) return h;
        h.payload_len = 0;
    } else {
        h.payload_len = (d[13] << 8) | d[14];
        const unsigned rx = (static_cast<unsigned>(d[15]) << 8) | d[16];
        if (rx != crc16(d, 15)) return h;
    }
    h.valid = true;
    return h;
}

struct Frame {
    uint8_t* cw[4];        // info bytes of the four codewords (status.data[c]), bpc valid bytes each
    int bpc;
};

// CodewordStatus::reassemble with all four codewords decoded; returns the length (0 = empty)
RC_HD int reassemble(const Frame& f, uint8_t* out) {
    const Header h = parse_header(f.cw[0], f.bpc);
    if (!h.valid) return 0;
    const int expected = h.is_control ? 20 : 17 + h.payload_len + 2;
    int n = 0;
    for (int i = 0; i < 4; ++i) {
        const int remaining = expected - n;
        if (remaining == 0) break;
        int skip = 0, avail = f.bpc;
        if (i > 0 && f.bpc >= 2 && f.cw[i][0] == 0xD5) { skip = 2; avail = f.bpc - 2; }   // DATA_CW_MARKER quirk (:974)
        const int to_copy = remaining < avail ? remaining : avail;
        for (int b = 0; b < to_copy; ++b) out[n + b] = f.cw[i][skip + b];
        n += to_copy;
    }
    return n;
}

// verifyFrame (:1583-1589) == the frame_valid test (:1566-1576) on an assembled frame
RC_HD bool verify_frame(const uint8_t* a, int len) {
    if (len == 0) return false;
    const Header h = parse_header(a, len);
    if (!h.valid) return false;
    if (h.is_control) return true;            // ControlFrame::deserialize repeats parseHeader's checks
    if (len < 19) return false;
    const int expected = 17 + h.payload_len + 2;
    if (len < expected) return false;
    const unsigned rx = (static_cast<unsigned>(a[expected - 2]) << 8) | a[expected - 1];
    return rx == crc16(a, expected - 2);
}

RC_HD bool frame_valid(const Frame& f, uint8_t* tmp) { return verify_frame(tmp, reassemble(f, tmp)); }

// ---------------------------------------------------------------------------------------------
// libstdc++ std::sort on (key, val) pairs ordered by key alone  (bits/stl_algo.h, GCC 13)
// ---------------------------------------------------------------------------------------------
struct SortView { float* key; uint16_t* val; };
RC_HD void sv_swap(const SortView& s, int a, int b) {
    const float k = s.key[a]; s.key[a] = s.key[b]; s.key[b] = k;
    const uint16_t v = s.val[a]; s.val[a] = s.val[b]; s.val[b] = v;
}
RC_HD void sv_move(const SortView& s, int dst, int src) { s.key[dst] = s.key[src]; s.val[dst] = s.val[src]; }

RC_HD void sv_unguarded_linear_insert(const SortView& s, int last) {
    const float k = s.key[last]; const uint16_t v = s.val[last];
    int next = last - 1;
    while (k < s.key[next]) { sv_move(s, last, next); last = next; --next; }
    s.key[last] = k; s.val[last] = v;
}
RC_HD void sv_insertion_sort(const SortView& s, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (s.key[i] < s.key[first]) {
            const float k = s.key[i]; const uint16_t v = s.val[i];
            for (int j = i; j > first; --j) sv_move(s, j, j - 1);          // move_backward(first, i, i + 1)
            s.key[first] = k; s.val[first] = v;
        } else {
            sv_unguarded_linear_insert(s, i);
        }
    }
}
RC_HD void sv_push_heap(const SortView& s, int first, int hole, int top, float k, uint16_t v) {
    int parent = (hole - 1) / 2;
    while (hole > top && s.key[first + parent] < k) {
        sv_move(s, first + hole, first + parent);
        hole = parent;
        parent = (hole - 1) / 2;
    }
    s.key[first + hole] = k; s.val[first + hole] = v;
}
RC_HD void sv_adjust_heap(const SortView& s, int first, int hole, int len, float k, uint16_t v) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (s.key[first + child] < s.key[first + child - 1]) --child;
        sv_move(s, first + hole, first + child);
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        sv_move(s, first + hole, first + child - 1);
        hole = child - 1;
    }
    sv_push_heap(s, first, hole, top, k, v);
}
// __partial_sort(first, last, last): make_heap + sort_heap
RC_HD void sv_heap_sort(const SortView& s, int first, int last) {
    const int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        for (;;) {
            const float k = s.key[first + parent]; const uint16_t v = s.val[first + parent];
            sv_adjust_heap(s, first, parent, len, k, v);
            if (parent == 0) break;
            --parent;
        }
    }
    while (last - first > 1) {
        --last;
        const float k = s.key[last]; const uint16_t v = s.val[last];
        sv_move(s, last, first);
        sv_adjust_heap(s, first, 0, last - first, k, v);
    }
}
RC_HD void sv_move_median_to_first(const SortView& s, int result, int a, int b, int c) {
    const float ka = s.key[a], kb = s.key[b], kc = s.key[c];
    if (ka < kb) {
        if (kb < kc) sv_swap(s, result, b);
        else if (ka < kc) sv_swap(s, result, c);
        else sv_swap(s, result, a);
    } else if (ka < kc) sv_swap(s, result, a);
    else if (kb < kc) sv_swap(s, result, c);
    else sv_swap(s, result, b);
}
RC_HD int sv_unguarded_partition(const SortView& s, int first, int last, int pivot) {
    for (;;) {
        while (s.key[first] < s.key[pivot]) ++first;
        --last;
        while (s.key[pivot] < s.key[last]) --last;
        if (!(first < last)) return first;
        sv_swap(s, first, last);
        ++first;
    }
}
RC_HD void libstdcxx_sort(float* key, uint16_t* val, int n) {
    if (n <= 0) return;
    const SortView s{key, val};
    // __introsort_loop, recursion on the right part replaced by an explicit stack
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) ++lg;
    int stack_first[48], stack_last[48], stack_depth[48];
    int sp = 0;
    stack_first[sp] = 0; stack_last[sp] = n; stack_depth[sp] = 2 * lg; ++sp;
    while (sp > 0) {
        --sp;
        int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
        while (last - first > 16) {
            if (depth == 0) { sv_heap_sort(s, first, last); break; }
            --depth;
            const int mid = first + (last - first) / 2;
            sv_move_median_to_first(s, first, first + 1, mid, last - 1);
            const int cut = sv_unguarded_partition(s, first + 1, last, first);
            // the reference recurses into [cut, last) first, then continues with [first, cut): the two
            // ranges are disjoint, so the order in which they are finished does not change the result
            stack_first[sp] = cut; stack_last[sp] = last; stack_depth[sp] = depth; ++sp;
            last = cut;
        }
    }
    // __final_insertion_sort
    if (n > 16) {
        sv_insertion_sort(s, 0, 16);
        for (int i = 16; i != n; ++i) sv_unguarded_linear_insert(s, i);
    } else {
        sv_insertion_sort(s, 0, n);
    }
}

// ---------------------------------------------------------------------------------------------
// Bit-flip searches (:1591-1843).  Returns true when the frame was recovered (f.cw modified).
//   soft       [4][648] de-interleaved soft bits of the four codewords
//   scratch    deltas u16[kMaxBits] | key f32[kMaxBits] | val u16[kMaxBits] | frame u8[kMaxFrameBytes] | trial u8[kMaxFrameBytes]
// ---------------------------------------------------------------------------------------------
struct Scratch { uint16_t* deltas; float* key; uint16_t* val; uint8_t* frame; uint8_t* trial; };

RC_HD void flip_frame_bit(const Frame& f, int p) {           // fixBit (:1707-1713)
    const int fb = p / 8, cw = fb / f.bpc, cb = fb % f.bpc;
    if (cw < 4) f.cw[cw][cb] ^= static_cast<uint8_t>(1u << (p % 8));
}

RC_HD bool repair_header(const Frame& f, const Scratch& s) {
    // Case 1 (:1591-1645): reassemble() is empty, i.e. the header in CW0 does not parse.
    uint8_t* d = f.cw[0];
    const int nbits = f.bpc * 8;
    // effect of flipping bit p of CW0 on (stored_hcrc ^ calc_hcrc): bits of bytes 0..14 through the
    // CRC (linear), bits of bytes 15/16 directly; everything else leaves the header check alone
    uint16_t* hd = s.deltas;
    {
        uint16_t st[8];
        for (int b = 0; b < 8; ++b) {           // bit b of the LAST CRC'd byte (14): value << 8 into the register
            uint16_t r = static_cast<uint16_t>((1u << b) << 8);
            st[b] = crc_shift8(r);
        }
        for (int B = 14; B >= 0; --B) {
            for (int b = 0; b < 8; ++b) { hd[B * 8 + b] = st[b]; st[b] = crc_shift8(st[b]); }
        }
        for (int b = 0; b < 8; ++b) { hd[15 * 8 + b] = static_cast<uint16_t>((1u << b) << 8); hd[16 * 8 + b] = static_cast<uint16_t>(1u << b); }
    }
    const uint16_t base_syn = static_cast<uint16_t>(((static_cast<unsigned>(d[15]) << 8) | d[16]) ^ crc16(d, 15));
    const uint16_t magic0 = static_cast<uint16_t>((d[0] << 8) | d[1]);
    auto magic_after = [&](int p) -> uint16_t {          // change of the magic word by flipping bit p
        if (p < 8) return static_cast<uint16_t>((1u << p) << 8);
        if (p < 16) return static_cast<uint16_t>(1u << (p - 8));
        return 0;
    };
    auto hsyn = [&](int p) -> uint16_t { return p < 17 * 8 ? hd[p] : 0; };
    // single-bit flips (:1596-1614)
    for (int p = 0; p < nbits; ++p) {
        if ((magic0 ^ magic_after(p)) != 0x554C) continue;
        if ((base_syn ^ hsyn(p)) != 0) continue;
        d[p / 8] ^= static_cast<uint8_t>(1u << (p % 8));
        if (frame_valid(f, s.trial)) return true;
        d[p / 8] ^= static_cast<uint8_t>(1u << (p % 8));
    }
    // two-bit flips (:1617-1643)
    for (int b1 = 0; b1 < nbits; ++b1) {
        const uint16_t m1 = static_cast<uint16_t>(magic0 ^ magic_after(b1));
        const uint16_t s1 = static_cast<uint16_t>(base_syn ^ hsyn(b1));
        for (int b2 = b1 + 1; b2 < nbits; ++b2) {
            if ((m1 ^ magic_after(b2)) != 0x554C) continue;
            if ((s1 ^ hsyn(b2)) != 0) continue;
            d[b1 / 8] ^= static_cast<uint8_t>(1u << (b1 % 8));
            d[b2 / 8] ^= static_cast<uint8_t>(1u << (b2 % 8));
            if (frame_valid(f, s.trial)) return true;
            d[b2 / 8] ^= static_cast<uint8_t>(1u << (b2 % 8));
            d[b1 / 8] ^= static_cast<uint8_t>(1u << (b1 % 8));
        }
    }
    return false;
}

RC_HD bool repair_payload(const Frame& f, const float* soft, const Scratch& s, int frame_len) {
    // Case 2 (:1646-1843): header parses, frame CRC (or size) fails.  s.frame holds reassemble().
    uint8_t* fd = s.frame;
    const Header h = parse_header(fd, frame_len);
    if (!h.valid || h.is_control) return false;
    const int expected = 17 + h.payload_len + 2;
    if (frame_len < expected) return false;
    const uint16_t stored = static_cast<uint16_t>((fd[expected - 2] << 8) | fd[expected - 1]);
    const int data_bytes = expected - 2;
    const uint16_t orig = static_cast<uint16_t>(crc16(fd, data_bytes));
    const uint16_t syndrome = static_cast<uint16_t>(stored ^ orig);
    const int data_bits = data_bytes * 8;
    // deltas[p] = orig ^ crc(frame with bit p flipped) (:1664-1669): linear in the flipped bit
    {
        uint16_t st[8];
        for (int b = 0; b < 8; ++b) st[b] = crc_shift8(static_cast<uint16_t>((1u << b) << 8));
        for (int B = data_bytes - 1; B >= 0; --B)
            for (int b = 0; b < 8; ++b) { s.deltas[B * 8 + b] = st[b]; st[b] = crc_shift8(st[b]); }
    }
    // single data bit (:1672-1686): applied without verification
    for (int p = 0; p < data_bits; ++p) {
        if (s.deltas[p] == syndrome) {
            const int fb = p / 8, cw = fb / f.bpc;
            if (cw < 4) { f.cw[cw][fb % f.bpc] ^= static_cast<uint8_t>(1u << (p % 8)); return true; }
        }
    }
    // single bit in the stored CRC (:1689-1704)
    for (int bit = 0; bit < 16; ++bit) {
        if (syndrome == (1u << bit)) {
            const int fb = (bit >= 8) ? (expected - 2) : (expected - 1);
            const int cw = fb / f.bpc;
            if (cw < 4) { f.cw[cw][fb % f.bpc] ^= static_cast<uint8_t>(1u << (bit % 8)); return true; }
        }
    }
    // suspects (:1729-1750): info bits whose channel sign disagrees with bit (i % 8) of the decoded byte
    int ns_all = 0;
    for (int c = 0; c < 4; ++c) {
        for (int i = 0; i < f.bpc * 8 && i < 648; ++i) {
            const int frame_bit = c * f.bpc * 8 + i;
            if (frame_bit / 8 >= data_bytes) continue;
            const float v = soft[c * 648 + i];
            const int ch_bit = (v < 0) ? 1 : 0;
            const int dec_bit = (f.cw[c][i / 8] >> (i % 8)) & 1;
            if (ch_bit != dec_bit) { s.key[ns_all] = v < 0 ? -v : v; s.val[ns_all] = static_cast<uint16_t>(frame_bit); ++ns_all; }
        }
    }
    libstdcxx_sort(s.key, s.val, ns_all);
    const int ns = ns_all < 30 ? ns_all : 30;
    uint16_t sd[30];
    for (int i = 0; i < ns; ++i) sd[i] = s.deltas[s.val[i]];
    // 2-bit (:1761-1779)
    for (int a = 0; a < ns; ++a)
        for (int b = a + 1; b < ns; ++b)
            if ((sd[a] ^ sd[b]) == syndrome) {
                flip_frame_bit(f, s.val[a]); flip_frame_bit(f, s.val[b]);
                if (frame_valid(f, s.trial)) return true;
                flip_frame_bit(f, s.val[a]); flip_frame_bit(f, s.val[b]);
            }
    // 3-bit (:1782-1806)
    for (int a = 0; a < ns; ++a)
        for (int b = a + 1; b < ns; ++b) {
            const uint16_t dab = static_cast<uint16_t>(sd[a] ^ sd[b]);
            for (int c = b + 1; c < ns; ++c)
                if ((dab ^ sd[c]) == syndrome) {
                    flip_frame_bit(f, s.val[a]); flip_frame_bit(f, s.val[b]); flip_frame_bit(f, s.val[c]);
                    if (frame_valid(f, s.trial)) return true;
                    flip_frame_bit(f, s.val[a]); flip_frame_bit(f, s.val[b]); flip_frame_bit(f, s.val[c]);
                }
        }
    // 4-bit among the 15 weakest (:1812-1842)
    const int ns4 = ns < 15 ? ns : 15;
    for (int a = 0; a < ns4; ++a)
        for (int b = a + 1; b < ns4; ++b) {
            const uint16_t dab = static_cast<uint16_t>(sd[a] ^ sd[b]);
            for (int c = b + 1; c < ns4; ++c) {
                const uint16_t dabc = static_cast<uint16_t>(dab ^ sd[c]);
                for (int e = c + 1; e < ns4; ++e)
                    if ((dabc ^ sd[e]) == syndrome) {
                        flip_frame_bit(f, s.val[a]); flip_frame_bit(f, s.val[b]); flip_frame_bit(f, s.val[c]); flip_frame_bit(f, s.val[e]);
                        if (frame_valid(f, s.trial)) return true;
                        flip_frame_bit(f, s.val[a]); flip_frame_bit(f, s.val[b]); flip_frame_bit(f, s.val[c]); flip_frame_bit(f, s.val[e]);
                    }
            }
        }
    return false;
}

// The searches that precede the re-decode fallback.  Returns true when recovered.
RC_HD bool repair_bitflips(const Frame& f, const float* soft, const Scratch& s) {
    const int len = reassemble(f, s.frame);
    if (len == 0) return repair_header(f, s);
    return repair_payload(f, soft, s, len);
}

}  // namespace ria_repair
