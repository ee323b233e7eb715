/* TEST INFRASTRUCTURE ONLY -- see ria_oracle.h.
 *
 * LDPC (n = 648) restatement: H construction, systematic encoder, flooding normalised
 * min-sum decoder, plus the CRC-16 used by the frame header.  Float order follows the reference
 * exactly (no FMA: built with -ffp-contract=off).
 */
#include "ria_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------
 * std::mt19937 -- ISO C++ [rand.predef]: mersenne_twister_engine<uint32,32,624,397,31,
 * 0x9908b0df,11,0xffffffff,7,0x9d2c5680,15,0xefc60000,18,1812433253>.
 * The reference seeds it with 0x12345678 + rate (src/fec/ldpc_decoder.cpp:73).
 * ------------------------------------------------------------------------------------------- */
void orc_mt_seed(orc_mt19937* g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i)
        g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

uint32_t orc_mt_next(orc_mt19937* g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* src/fec/ldpc_decoder.cpp:21-36 (getCodeParams) */
static void code_params(int rate, int* k, int* m) {
    switch (rate) {
        case ORC_R1_4: *k = 162; *m = 486; break;
        case ORC_R1_2: *k = 324; *m = 324; break;
        case ORC_R2_3: *k = 432; *m = 216; break;
        case ORC_R3_4: *k = 486; *m = 162; break;
        case ORC_R5_6: *k = 540; *m = 108; break;
        default:       *k = 324; *m = 324; break;
    }
}

/* src/fec/ldpc_decoder.cpp:65-138 (buildMatrix; twin at ldpc_encoder.cpp:70-129).
 * H = [H_data | I].  Rows keep insertion order: info bits in ascending j as they were
 * connected, then any "empty row" fix-up, then the identity column k+i. */
void orc_ldpc_build(int rate, orc_ldpc_code* code) {
    int k, m;
    code_params(rate, &k, &m);
    int n = k + m;
    code->rate = rate; code->k = k; code->m = m; code->n = n;

    orc_mt19937 rng;
    orc_mt_seed(&rng, (uint32_t)(0x12345678 + rate));

    /* rows[i][..] up to 8 entries (max_check_degree 6 + identity; fix-up only hits empty rows) */
    int (*rows)[8] = calloc((size_t)m, sizeof *rows);
    int* deg = calloc((size_t)m, sizeof *deg);
    int* check_degrees = calloc((size_t)m, sizeof *check_degrees);
    int* avail = malloc((size_t)m * sizeof *avail);

    int target_check_degree = 4;
    int target_var_degree = (target_check_degree * m) / k;
    if (target_var_degree < 3) target_var_degree = 3;
    if (target_var_degree > m / 2) target_var_degree = m / 2;
    int max_check_degree = target_check_degree + 2;

    for (int j = 0; j < k; ++j) {
        int na = 0;
        for (int i = 0; i < m; ++i)
            if (check_degrees[i] < max_check_degree) avail[na++] = i;
        /* Fisher-Yates with rng() % i (ldpc_decoder.cpp:101-104) */
        for (int i = na; i > 1; --i) {
            int r = (int)(orc_mt_next(&rng) % (uint32_t)i);
            int t = avail[i - 1]; avail[i - 1] = avail[r]; avail[r] = t;
        }
        int connections = target_var_degree < na ? target_var_degree : na;
        for (int d = 0; d < connections; ++d) {
            int c = avail[d];
            rows[c][deg[c]++] = j;
            check_degrees[c]++;
        }
    }
    for (int i = 0; i < m; ++i) {
        if (deg[i] == 0) {
            int j = (int)(orc_mt_next(&rng) % (uint32_t)k);
            rows[i][deg[i]++] = j;
        }
    }
    for (int i = 0; i < m; ++i) rows[i][deg[i]++] = k + i;

    int e = 0;
    for (int i = 0; i < m; ++i) {
        code->row_ptr[i] = e;
        for (int d = 0; d < deg[i]; ++d) code->edge_var[e++] = rows[i][d];
    }
    code->row_ptr[m] = e;
    code->n_edges = e;
    free(rows); free(deg); free(check_degrees); free(avail);
}

/* src/fec/ldpc_encoder.cpp:193-257.  Bit-level blocking: k info bits per block, coded blocks are
 * byte-packed one after another, each padded to a whole byte (81 bytes for n=648). */
int orc_ldpc_encode(const orc_ldpc_code* code, const uint8_t* data, int len, uint8_t* out, int out_cap) {
    int k = code->k, m = code->m, n = code->n;
    int total_bits = len * 8, o = 0;
    uint8_t cw[ORC_LDPC_N];
    for (int off = 0; off < total_bits; off += k) {
        for (int j = 0; j < k; ++j) {
            int b = off + j;
            cw[j] = (b < total_bits) ? (uint8_t)((data[b >> 3] >> (7 - (b & 7))) & 1) : 0;
        }
        for (int i = 0; i < m; ++i) {
            uint8_t s = 0;
            /* identity edge is last in each row; parity uses the H_data part only */
            for (int e = code->row_ptr[i]; e < code->row_ptr[i + 1] - 1; ++e) s ^= cw[code->edge_var[e]];
            cw[k + i] = s;
        }
        uint8_t byte = 0; int cnt = 0;
        for (int j = 0; j < n; ++j) {
            byte = (uint8_t)((byte << 1) | cw[j]);
            if (++cnt == 8) { if (o >= out_cap) return -1; out[o++] = byte; byte = 0; cnt = 0; }
        }
        if (cnt > 0) { if (o >= out_cap) return -1; out[o++] = (uint8_t)(byte << (8 - cnt)); }
    }
    return o;
}

/* src/fec/ldpc_decoder.cpp:154-260 (decodeBP). */
int orc_ldpc_decode(const orc_ldpc_code* code, const float* llr, int n_llr, int max_iter,
                    float factor, uint8_t* out, int* iters) {
    const int n = code->n, k = code->k, m = code->m, E = code->n_edges;
    static _Thread_local float v2c[ORC_LDPC_MAX_EDGES], c2v[ORC_LDPC_MAX_EDGES];
    float llr_in[ORC_LDPC_N], total[ORC_LDPC_N];
    uint8_t hard[ORC_LDPC_N];

    for (int j = 0; j < n; ++j) { llr_in[j] = (j < n_llr) ? llr[j] : 0.0f; total[j] = llr_in[j]; }
    for (int e = 0; e < E; ++e) { v2c[e] = llr_in[code->edge_var[e]]; c2v[e] = 0.0f; }

    int success = 0, it;
    for (it = 0; it < max_iter; ++it) {
        /* :182-203 check update -- product of signs (msg < 0) and min |msg| over the other edges */
        for (int i = 0; i < m; ++i) {
            int b = code->row_ptr[i], d = code->row_ptr[i + 1] - b;
            for (int e = 0; e < d; ++e) {
                float sign = 1.0f, min_abs = FLT_MAX;
                for (int e2 = 0; e2 < d; ++e2) {
                    if (e2 == e) continue;
                    float msg = v2c[b + e2];
                    if (msg < 0) sign = -sign;
                    float a = fabsf(msg);
                    if (a < min_abs) min_abs = a;
                }
                c2v[b + e] = sign * min_abs * factor;
            }
        }
        /* :207-214 totals: channel LLR plus c2v in ascending (check, edge) order */
        for (int j = 0; j < n; ++j) total[j] = llr_in[j];
        for (int e = 0; e < E; ++e) total[code->edge_var[e]] += c2v[e];
        /* :217-225 v2c with clamp std::max(-50, std::min(50, x)) */
        for (int e = 0; e < E; ++e) {
            float x = total[code->edge_var[e]] - c2v[e];
            float y = (x < 50.0f) ? x : 50.0f;
            v2c[e] = (-50.0f < y) ? y : -50.0f;
        }
        /* :228-236 hard decision + parity */
        for (int j = 0; j < n; ++j) hard[j] = (total[j] < 0) ? 1 : 0;
        int ok = 1;
        for (int i = 0; i < m && ok; ++i) {
            uint8_t s = 0;
            for (int e = code->row_ptr[i]; e < code->row_ptr[i + 1]; ++e) s ^= hard[code->edge_var[e]];
            if (s) ok = 0;
        }
        if (ok) { success = 1; break; }
    }
    *iters = it;

    /* :240-257 pack first k bits MSB-first, last byte left-aligned */
    int o = 0, cnt = 0; uint8_t byte = 0;
    for (int j = 0; j < k; ++j) {
        byte = (uint8_t)((byte << 1) | ((total[j] < 0) ? 1 : 0));
        if (++cnt == 8) { out[o++] = byte; byte = 0; cnt = 0; }
    }
    if (cnt > 0) out[o++] = (uint8_t)(byte << (8 - cnt));
    return success;
}

void orc_ldpc_decode_batch(const orc_ldpc_code* code, const float* llr, int n_cw, int max_iter,
                           float factor, uint8_t* out, int out_stride, uint8_t* ok, int32_t* iters) {
    uint8_t tmp[ORC_LDPC_N / 8 + 1];
    int nb = (code->k + 7) / 8;
    for (int c = 0; c < n_cw; ++c) {
        int it = 0;
        ok[c] = (uint8_t)orc_ldpc_decode(code, llr + (size_t)c * ORC_LDPC_N, ORC_LDPC_N, max_iter, factor, tmp, &it);
        iters[c] = it;
        memset(out + (size_t)c * out_stride, 0, (size_t)out_stride);
        memcpy(out + (size_t)c * out_stride, tmp, (size_t)(nb < out_stride ? nb : out_stride));
    }
}

/* src/protocol/frame_v2.cpp:115-128 */
uint16_t orc_crc16(const uint8_t* data, int len) {
    uint16_t crc = 0xFFFF;
    for (int i = 0; i < len; ++i) {
        crc ^= (uint16_t)((uint16_t)data[i] << 8);
        for (int j = 0; j < 8; ++j)
            crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ 0x1021) : (uint16_t)(crc << 1);
    }
    return crc;
}
