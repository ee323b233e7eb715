// Host-side generation of the n=648 parity-check structure and the device table layouts.
//
// The reference derives H = [H_data | I] at construction time from a seeded std::mt19937
// (LDPCDecoder::Impl::buildMatrix, src/fec/ldpc_decoder.cpp:65-138; the encoder builds the same
// H_data, src/fec/ldpc_encoder.cpp:70-129).  There is no stored matrix to load, so a drop-in
// decoder has to regenerate the identical structure: same seed (0x12345678 + rate), same
// draw order (Fisher-Yates with rng() % i over the checks that still have room, then one draw
// per empty row).  std::mt19937 is specified bit-exactly by the C++ standard.

#include "ria_internal.h"

#include <algorithm>
#include <mutex>
#include <random>
#include <stdexcept>

namespace ria {

namespace {

struct Params { int k, m; };

// getCodeParams, src/fec/ldpc_decoder.cpp:21-36 (unknown rates fall back to R1/2 dimensions,
// but still seed the generator with their own enum value).
Params params_for(int rate) {
    switch (rate) {
        case RIA_R1_4: return {162, 486};
        case RIA_R1_2: return {324, 324};
        case RIA_R2_3: return {432, 216};
        case RIA_R3_4: return {486, 162};
        case RIA_R5_6: return {540, 108};
        default:       return {324, 324};
    }
}

void build(int rate, LdpcCodeHost& c) {
    const Params p = params_for(rate);
    const int k = p.k, m = p.m, n = k + m;
    c.rate = rate; c.k = k; c.m = m; c.n = n;

    std::mt19937 rng(static_cast<uint32_t>(0x12345678 + rate));
    std::vector<std::vector<int>> rows(m);
    std::vector<int> room(m, 0);

    const int check_target = 4;
    const int check_cap = check_target + 2;
    int var_degree = std::max(3, (check_target * m) / k);
    var_degree = std::min(var_degree, m / 2);

    std::vector<int> open;
    open.reserve(m);
    for (int j = 0; j < k; ++j) {
        open.clear();
        for (int i = 0; i < m; ++i)
            if (room[i] < check_cap) open.push_back(i);
        for (size_t i = open.size(); i > 1; --i) {
            size_t pick = rng() % i;
            std::swap(open[i - 1], open[pick]);
        }
        const int take = std::min<int>(var_degree, static_cast<int>(open.size()));
        for (int d = 0; d < take; ++d) {
            rows[open[d]].push_back(j);
            room[open[d]]++;
        }
    }
    for (int i = 0; i < m; ++i)
        if (rows[i].empty()) rows[i].push_back(static_cast<int>(rng() % k));

    // CSR in H_rows order; the identity column k+i closes every row.
    c.row_ptr.assign(m + 1, 0);
    c.edge_var.clear();
    for (int i = 0; i < m; ++i) {
        c.row_ptr[i] = static_cast<int32_t>(c.edge_var.size());
        for (int j : rows[i]) c.edge_var.push_back(j);
        c.edge_var.push_back(k + i);
    }
    c.row_ptr[m] = static_cast<int32_t>(c.edge_var.size());
    c.n_edges = static_cast<int>(c.edge_var.size());

    // ---- device layouts -------------------------------------------------------------------
    // Per check: 8 message slots; 0..5 = info edges, 6 = the identity edge, 7 = spare (the kernel
    // keeps the parity variable's running total there).  chk_var mirrors the slots with variable
    // indices and stores the info-edge count in slot 7.
    c.chk_var.assign(static_cast<size_t>(m) * 8, 0xFFFF);
    for (int i = 0; i < m; ++i) {
        if (rows[i].size() > 6) throw std::runtime_error("ldpc: check degree exceeds slot layout");
        for (size_t d = 0; d < rows[i].size(); ++d) c.chk_var[i * 8 + d] = static_cast<uint16_t>(rows[i][d]);
        c.chk_var[i * 8 + 7] = static_cast<uint16_t>(rows[i].size());
    }
    // Per info variable: its message slots ordered by ascending check index -- the reference
    // accumulates llr_total[j] over checks i = 0..m-1 in that order (ldpc_decoder.cpp:207-214),
    // and fp32 addition order is part of the bit-exact contract.
    std::vector<std::vector<uint16_t>> per_var(k);
    for (int i = 0; i < m; ++i)
        for (size_t d = 0; d < rows[i].size(); ++d) {
            const int slot = static_cast<int>((d >> 2) * (m * 4) + i * 4 + (d & 3));
            per_var[rows[i][d]].push_back(static_cast<uint16_t>(slot));
        }
    c.dv_max = 0;
    for (auto& v : per_var) c.dv_max = std::max<int>(c.dv_max, static_cast<int>(v.size()));
    c.var_slot.assign(static_cast<size_t>(c.dv_max) * k, 0xFFFF);
    for (int j = 0; j < k; ++j)
        for (size_t d = 0; d < per_var[j].size(); ++d) c.var_slot[d * k + j] = per_var[j][d];
}

}  // namespace

bool ldpc_rate_valid(int rate) { return rate >= RIA_R1_4 && rate <= RIA_R7_8; }

const LdpcCodeHost& ldpc_code_host(int rate) {
    static std::mutex mu;
    static LdpcCodeHost cache[8];
    if (!ldpc_rate_valid(rate)) throw std::invalid_argument("ldpc: bad rate");
    std::lock_guard<std::mutex> lock(mu);
    if (cache[rate].rate != rate) build(rate, cache[rate]);
    return cache[rate];
}

}  // namespace ria

extern "C" int ria_ldpc_params(int rate, int* k_info, int* m_parity, int* n_edges) {
    if (!ria::ldpc_rate_valid(rate)) return RIA_E_INVAL;
    try {
        const auto& c = ria::ldpc_code_host(rate);
        if (k_info) *k_info = c.k;
        if (m_parity) *m_parity = c.m;
        if (n_edges) *n_edges = c.n_edges;
    } catch (...) { return RIA_E_INVAL; }
    return RIA_OK;
}

extern "C" int ria_ldpc_get_matrix(int rate, int32_t* row_ptr, int32_t* edge_var) {
    if (!ria::ldpc_rate_valid(rate) || !row_ptr || !edge_var) return RIA_E_INVAL;
    try {
        const auto& c = ria::ldpc_code_host(rate);
        std::copy(c.row_ptr.begin(), c.row_ptr.end(), row_ptr);
        std::copy(c.edge_var.begin(), c.edge_var.end(), edge_var);
    } catch (...) { return RIA_E_INVAL; }
    return RIA_OK;
}
