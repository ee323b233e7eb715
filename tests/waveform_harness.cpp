// TEST INFRASTRUCTURE: the IWaveform drop-ins of include/ria_b200_adapters.hpp next to the reference's own
// waveform objects on identical samples, in the style of the reference's tools/test_waveform_simple.cpp.
//
// Built HERE (where /root/reference exists) by oracle/Makefile into oracle/_ref/waveform_harness from this file,
// the reference's own headers / objects and libria_b200.so; tests/test_waveform_dropin_gpu.py runs the binary
// on the B200.  Every waveform is created twice -- ultra::OFDMChirpWaveform / MCDPSKWaveform / OFDMNvisWaveform and
// ria::createWaveform(mode) -- and driven through the same IWaveform calls a StreamingDecoder makes:
//   connect frame:   generatePreamble + modulate -> channel -> detectSync -> setFrequencyOffset -> process -> getSoftBits
//   connected frame: generateDataPreamble + modulate -> channel -> detectDataSync -> process -> getSoftBits
// Sync positions, CFO and the soft bits must be identical; correlation values agree to 1e-4.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#define RIA_WITH_ULTRA 1
#include "ria_b200_adapters.hpp"
#include "ultra/logging.hpp"

using namespace ultra;

static int g_checks = 0, g_fail = 0;
#define CHECK(cond, ...)                                                      \
    do {                                                                      \
        ++g_checks;                                                           \
        if (!(cond)) { ++g_fail; printf("FAIL %s:%d ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } \
    } while (0)

static Samples channel(const Samples& tx, size_t lead, size_t tail, float snr_db, uint32_t seed) {
    double p = 0.0;
    for (float v : tx) p += static_cast<double>(v) * v;
    p /= tx.size();
    const float sigma = static_cast<float>(std::sqrt(p / std::pow(10.0, snr_db / 10.0)));
    std::mt19937 rng(seed);
    std::normal_distribution<float> g(0.0f, 1.0f);
    Samples rx(lead + tx.size() + tail, 0.0f);
    std::memcpy(rx.data() + lead, tx.data(), tx.size() * sizeof(float));
    for (float& v : rx) v += sigma * g(rng);
    return rx;
}

static bool same_bits(const std::vector<float>& a, const std::vector<float>& b) {
    return a.size() == b.size() && (a.empty() || std::memcmp(a.data(), b.data(), a.size() * sizeof(float)) == 0);
}

static void compare_sync(const char* what, bool da, const SyncResult& a, bool db, const SyncResult& b) {
    CHECK(da == db && a.detected == b.detected, "%s detected %d vs %d", what, a.detected, b.detected);
    if (!a.detected || !b.detected) return;
    CHECK(a.start_sample == b.start_sample, "%s start %d vs %d", what, a.start_sample, b.start_sample);
    CHECK(a.cfo_hz == b.cfo_hz, "%s cfo %.6f vs %.6f", what, a.cfo_hz, b.cfo_hz);
    CHECK(std::fabs(a.correlation - b.correlation) <= 1e-4f * std::max(1.0f, std::fabs(a.correlation)),
          "%s correlation %.6f vs %.6f", what, a.correlation, b.correlation);
    CHECK(a.has_training == b.has_training, "%s has_training", what);
}

static void run_frame(const char* what, IWaveform& ref, IWaveform& dut, const Samples& rx, bool data_preamble, int frame_samples,
                      float known_cfo, float sync_threshold = 0.15f, size_t sync_window = 120000) {
    SyncResult ra, rb;
    bool da, db;
    if (data_preamble) {
        da = ref.detectDataSync(SampleSpan(rx.data(), rx.size()), ra, known_cfo, 0.2f);
        db = dut.detectDataSync(SampleSpan(rx.data(), rx.size()), rb, known_cfo, 0.2f);
    } else {
        const size_t window = std::min<size_t>(rx.size(), sync_window);      // default: StreamingDecoder's chirp search window
        da = ref.detectSync(SampleSpan(rx.data(), window), ra, sync_threshold);
        db = dut.detectSync(SampleSpan(rx.data(), window), rb, sync_threshold);
    }
    compare_sync(what, da, ra, db, rb);
    if (!da || !db) { CHECK(false, "%s: not detected (ref %d, dut %d)", what, da, db); return; }
    if (!data_preamble) { ref.setFrequencyOffset(ra.cfo_hz); dut.setFrequencyOffset(rb.cfo_hz); }
    const size_t start = static_cast<size_t>(ra.start_sample);
    const size_t n = std::min<size_t>(frame_samples, rx.size() - start);
    ref.setAbsoluteTrainingPosition(start);
    dut.setAbsoluteTrainingPosition(start);
    const bool pa = ref.process(SampleSpan(rx.data() + start, n));
    const bool pb = dut.process(SampleSpan(rx.data() + start, n));
    CHECK(pa == pb, "%s process %d vs %d", what, pa, pb);
    CHECK(ref.hasData() == dut.hasData(), "%s hasData", what);
    const std::vector<float> sa = ref.getSoftBits(), sb = dut.getSoftBits();
    CHECK(!sa.empty(), "%s: reference produced no soft bits", what);
    CHECK(same_bits(sa, sb), "%s soft bits differ (%zu vs %zu values)", what, sa.size(), sb.size());
    CHECK(ref.getSoftBits().empty() && dut.getSoftBits().empty(), "%s getSoftBits must move out", what);
    CHECK(std::fabs(ref.estimatedSNR() - dut.estimatedSNR()) <= 1e-4f * std::max(1.0f, std::fabs(ref.estimatedSNR())),
          "%s snr %.5f vs %.5f", what, ref.estimatedSNR(), dut.estimatedSNR());
    CHECK(ref.estimatedCFO() == dut.estimatedCFO(), "%s estimatedCFO %.6f vs %.6f", what, ref.estimatedCFO(), dut.estimatedCFO());
    CHECK(ref.getFrequencyOffset() == dut.getFrequencyOffset(), "%s getFrequencyOffset", what);
    CHECK(std::fabs(ref.getFadingIndex() - dut.getFadingIndex()) <= 1e-5f, "%s fading %.6f vs %.6f", what, ref.getFadingIndex(),
          dut.getFadingIndex());
    CHECK(ref.wasBurstInterleaved() == dut.wasBurstInterleaved(), "%s burst marker", what);
    ref.reset();
    dut.reset();
    CHECK(ref.getFrequencyOffset() == dut.getFrequencyOffset(), "%s CFO must survive reset", what);
}

int main() {
    setLogLevel(LogLevel::ERROR);
    std::mt19937 rng(2026);
    // ---- OFDM-Chirp ----
    struct { Modulation mod; CodeRate rate; float snr; } ofdm_cases[] = {
        {Modulation::DQPSK, CodeRate::R1_2, 18.0f}, {Modulation::QAM64, CodeRate::R3_4, 30.0f}, {Modulation::QAM16, CodeRate::R2_3, 24.0f}};
    for (auto& oc : ofdm_cases) {
        // the modem's configuration (ModemConfig defaults: 1024-point FFT, 59 carriers), as WaveformFactory::create(mode,
        // config) builds it; the default-constructed waveform is a 512-point / 30-carrier variant the chain does not cover
        ModemConfig mcfg;
        ultra::OFDMChirpWaveform ref(mcfg);
#ifdef HARNESS_SELFTEST      // reference against itself: validates the scenario without a GPU
        WaveformPtr dut = std::make_unique<ultra::OFDMChirpWaveform>(mcfg);
#else
        WaveformPtr dut = ria::createWaveform(protocol::WaveformMode::OFDM_CHIRP, mcfg);
#endif
        CHECK(dut != nullptr && dut->getMode() == protocol::WaveformMode::OFDM_CHIRP, "factory");
        ref.configure(oc.mod, oc.rate);
        dut->configure(oc.mod, oc.rate);
        CHECK(ref.getSamplesPerSymbol() == dut->getSamplesPerSymbol() && ref.getPreambleSamples() == dut->getPreambleSamples() &&
              ref.getMinSamplesForFrame() == dut->getMinSamplesForFrame() && ref.getPilotSpacing() == dut->getPilotSpacing(), "sizing");
        Bytes coded(324);
        for (auto& b : coded) b = static_cast<uint8_t>(rng());
        char name[96];
        for (int rep = 0; rep < 3; ++rep) {
            // connect-style frame behind the dual chirp
            Samples pre = ref.generatePreamble(), body = ref.modulate(coded);
            Samples tx = pre;
            tx.insert(tx.end(), body.begin(), body.end());
            const int frame_samples = static_cast<int>(tx.size()) - (2 * 24000 + 2 * 4800);     // training + data
            Samples rx = channel(tx, 700 + 97 * rep, 2000, oc.snr, 100 + rep);
            snprintf(name, sizeof name, "ofdm_chirp mod %d chirp frame %d", static_cast<int>(oc.mod), rep);
            run_frame(name, ref, *dut, rx, false, frame_samples, 0.0f);
            // connected-mode frame: training-only preamble (the modulator's mixer restarts with the training symbols,
            // so the data symbols are modulated again behind it)
            Samples dpre = ref.generateDataPreamble();
            Samples dbody = ref.modulate(coded);
            Samples dtx = dpre;
            dtx.insert(dtx.end(), dbody.begin(), dbody.end());
            Samples drx = channel(dtx, 300 + 41 * rep, 1500, oc.snr, 200 + rep);
            snprintf(name, sizeof name, "ofdm_chirp mod %d data frame %d", static_cast<int>(oc.mod), rep);
            run_frame(name, ref, *dut, drx, true, static_cast<int>(dtx.size()), 0.0f);
        }
    }
    // ---- MC-DPSK ----
    struct { Modulation mod; SpreadingMode spread; float snr; } mc_cases[] = {
        {Modulation::DBPSK, SpreadingMode::TIME_4X, -6.0f}, {Modulation::DQPSK, SpreadingMode::NONE, 8.0f}};
    for (auto& mc : mc_cases) {
        MultiCarrierDPSKConfig cfg;
        cfg.num_carriers = 10;
        cfg.bits_per_symbol = mc.mod == Modulation::DBPSK ? 1 : 2;
        cfg.spreading_mode = mc.spread;
        ultra::MCDPSKWaveform ref(cfg);
#ifdef HARNESS_SELFTEST
        ultra::MCDPSKWaveform dut(cfg);
#else
        ria::MCDPSKWaveform dut(cfg);
#endif
        IWaveform& d = dut;
        CHECK(ref.getSamplesPerSymbol() == d.getSamplesPerSymbol() && ref.getPreambleSamples() == d.getPreambleSamples() &&
              ref.getDataPreambleSamples() == d.getDataPreambleSamples() && ref.getMinSamplesForFrame() == d.getMinSamplesForFrame(), "mc sizing");
        Bytes coded(81);
        for (auto& b : coded) b = static_cast<uint8_t>(rng());
        char name[96];
        for (int rep = 0; rep < 3; ++rep) {
            Samples pre = ref.generatePreamble(), body = ref.modulate(coded);
            Samples tx = pre;
            tx.insert(tx.end(), body.begin(), body.end());
            Samples rx = channel(tx, 900 + 53 * rep, 1500, mc.snr, 300 + rep);
            // process() is handed [training][reference][data]: everything behind the 57 600-sample dual chirp
            snprintf(name, sizeof name, "mc_dpsk bits %d chirp frame %d", cfg.bits_per_symbol, rep);
            run_frame(name, ref, d, rx, false, static_cast<int>(tx.size()) - 57600, 0.0f);
            // connected mode: ZC preamble (2512 samples) + training + reference, then the data
            Samples dpre = ref.generateDataPreamble();
            Samples dbody = ref.modulate(coded);
            Samples dtx = dpre;
            dtx.insert(dtx.end(), dbody.begin(), dbody.end());
            Samples drx = channel(dtx, 400 + 29 * rep, 1200, mc.snr + 4.0f, 400 + rep);
            snprintf(name, sizeof name, "mc_dpsk bits %d zc frame %d", cfg.bits_per_symbol, rep);
            run_frame(name, ref, d, drx, true, static_cast<int>(dtx.size()) - 2512, 0.0f);
        }
    }
    // ---- OFDM-COX (Schmidl-Cox acquisition) ----
    struct { Modulation mod; CodeRate rate; float snr; } cox_cases[] = {
        {Modulation::DQPSK, CodeRate::R1_2, 20.0f}, {Modulation::QPSK, CodeRate::R1_2, 22.0f}, {Modulation::QAM16, CodeRate::R3_4, 26.0f}};
    for (auto& oc : cox_cases) {
        ModemConfig mcfg;
        ultra::OFDMNvisWaveform ref(mcfg);
#ifdef HARNESS_SELFTEST
        WaveformPtr dut = std::make_unique<ultra::OFDMNvisWaveform>(mcfg);
#else
        WaveformPtr dut = ria::createWaveform(protocol::WaveformMode::OFDM_COX, mcfg);
#endif
        CHECK(dut != nullptr && dut->getMode() == protocol::WaveformMode::OFDM_COX, "cox factory");
        ref.configure(oc.mod, oc.rate);
        dut->configure(oc.mod, oc.rate);
        CHECK(ref.getSamplesPerSymbol() == dut->getSamplesPerSymbol() && ref.getPreambleSamples() == dut->getPreambleSamples() &&
              ref.getMinSamplesForFrame() == dut->getMinSamplesForFrame() && ref.getPilotSpacing() == dut->getPilotSpacing(), "cox sizing");
        Bytes coded(324);
        for (auto& b : coded) b = static_cast<uint8_t>(rng());
        char name[96];
        for (int rep = 0; rep < 4; ++rep) {
            Samples pre = ref.generatePreamble(), body = ref.modulate(coded);
            Samples tx = pre;
            tx.insert(tx.end(), body.begin(), body.end());
            // LTS + data = everything behind the silent symbol and the four STS
            const int frame_samples = static_cast<int>(tx.size()) - 5 * ref.getSamplesPerSymbol();
            Samples rx = channel(tx, 900 + 1337 * rep, 3000, oc.snr, 500 + rep);
            snprintf(name, sizeof name, "ofdm_cox mod %d frame %d", static_cast<int>(oc.mod), rep);
            // the same objects see all four frames, so the noise-floor tracker of the search is carried along
            run_frame(name, ref, *dut, rx, false, frame_samples, 0.0f, 0.8f, 48000);
        }
        // noise only: nothing to find, result untouched; noise + a steady tone: whatever the reference makes of it
        for (int tone = 0; tone < 2; ++tone) {
            std::mt19937 nrng(77 + tone);
            std::normal_distribution<float> g(0.0f, 0.1f);
            Samples noise(30000);
            for (size_t i = 0; i < noise.size(); ++i) noise[i] = g(nrng) + (tone ? 0.5f * std::sin(0.37f * static_cast<float>(i)) : 0.0f);
            SyncResult ra, rb;
            ra.start_sample = rb.start_sample = -7;
            const bool da = ref.detectSync(SampleSpan(noise.data(), noise.size()), ra, 0.8f);
            const bool db = dut->detectSync(SampleSpan(noise.data(), noise.size()), rb, 0.8f);
            CHECK(da == db && ra.start_sample == rb.start_sample && (tone || !da), "cox %s window: %d %d, %d %d",
                  tone ? "tone" : "noise-only", da, db, ra.start_sample, rb.start_sample);
            if (da && db) CHECK(ra.cfo_hz == rb.cfo_hz, "cox tone cfo %.6f vs %.6f", ra.cfo_hz, rb.cfo_hz);
            ref.reset();
            dut->reset();
        }
    }
#ifndef HARNESS_SELFTEST
    CHECK(ria::createWaveform(protocol::WaveformMode::OTFS_EQ) == nullptr, "modes off the hot path are not ours");
#endif
    printf("%s: %d checks, %d failed\n", g_fail ? "FAIL" : "PASS", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
