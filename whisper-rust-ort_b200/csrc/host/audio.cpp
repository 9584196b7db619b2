// audio.cpp — host-side audio ingest (replaces load_audio_16k_mono + resample_linear,
// /root/reference/src/main.rs:207-316, which sit on the symphonia crate).  RIFF/WAVE in the sample
// formats the reference accepts (U8 / S16 / F32; its match bails on S24 / S32 / F64, and so on every FLAC
// file, which symphonia decodes to S32).  MP3 is the one container the reference reads and this does
// not (no decoder offline; SURVEY.md §8f3 ranks that "next"): reported as unsupported.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../common.h"

extern "C" {

// main.rs:207-226 — linear interpolation, positions in f64, out-of-range samples are 0.
int64_t wb_host_resample_linear(const float* x, int64_t n, uint32_t sr_in, uint32_t sr_out, float* out, int64_t cap) {
    if (sr_in == sr_out) {
        if (out) std::memcpy(out, x, sizeof(float) * (size_t)(n < cap ? n : cap));
        return n;
    }
    const double ratio = (double)sr_out / (double)sr_in;
    const int64_t n_out = (int64_t)std::llround((double)n * ratio);       // f64::round: half away from zero
    if (!out) return n_out;
    for (int64_t i = 0; i < n_out && i < cap; ++i) {
        const double t = (double)i / ratio;
        const int64_t i0 = (int64_t)std::floor(t), i1 = i0 + 1;
        const double a = t - (double)i0;
        const float s0 = (i0 < 0 || i0 >= n) ? 0.0f : x[i0];
        const float s1 = (i1 < 0 || i1 >= n) ? 0.0f : x[i1];
        out[i] = (float)(1.0 - a) * s0 + (float)a * s1;
    }
    return n_out;
}

void wb_host_free(void* p) { std::free(p); }

}  // extern "C"

namespace {

uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

void load_wav(const char* path, std::vector<float>& mono, uint32_t& sr) {
    FILE* f = std::fopen(path, "rb");
    WB_REQUIRE(f != nullptr, WB_EIO, "Failed to open audio: %s", path);                 // main.rs:237-238
    std::vector<unsigned char> buf;
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    buf.resize(sz > 0 ? (size_t)sz : 0);
    size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), f);
    std::fclose(f);
    WB_REQUIRE(got == buf.size() && buf.size() >= 12, WB_EIO, "short read: %s", path);
    // FLAC: symphonia's FLAC decoder hands back S32 buffers, which the reference's match rejects (main.rs:265-303,
    // `_ => bail!`), so a faithful drop-in fails on .flac too -- with the reference's own message.
    if (std::memcmp(buf.data(), "fLaC", 4) == 0) WB_THROW(WB_EINVAL, "Unsupported decoded sample format");
    if (std::memcmp(buf.data(), "RIFF", 4) != 0 || std::memcmp(buf.data() + 8, "WAVE", 4) != 0)
        WB_THROW(WB_EINVAL, "unsupported audio container (only RIFF/WAVE is decodable offline): %s", path);
    int fmt_tag = 0, channels = 0, bits = 0;
    sr = 0;
    const unsigned char* data = nullptr;
    size_t data_len = 0, pos = 12;
    while (pos + 8 <= buf.size()) {
        const unsigned char* ck = buf.data() + pos;
        size_t len = rd32(ck + 4);
        const size_t body = pos + 8;
        if (std::memcmp(ck, "fmt ", 4) == 0 && body + 16 <= buf.size()) {
            fmt_tag = rd16(ck + 8); channels = rd16(ck + 10); sr = rd32(ck + 12); bits = rd16(ck + 22);
            if (fmt_tag == 0xFFFE && len >= 26 && body + 26 <= buf.size()) fmt_tag = rd16(ck + 8 + 24);   // WAVE_FORMAT_EXTENSIBLE
        } else if (std::memcmp(ck, "data", 4) == 0) {
            if (body + len > buf.size()) len = buf.size() - body;      // streaming writers leave bogus sizes
            data = buf.data() + body;
            data_len = len;
            break;
        }
        pos = body + len + (len & 1);
    }
    WB_REQUIRE(sr > 0, WB_EINVAL, "Unknown sample rate");                                // main.rs:252
    WB_REQUIRE(channels > 0, WB_EINVAL, "Unknown channels");                             // main.rs:253
    WB_REQUIRE(data != nullptr, WB_EINVAL, "No default track");                          // main.rs:249
    const int bps = bits / 8;
    // symphonia hands back U8 / S16 / F32 for these; S24/S32/F64 hit `_ => bail!` (main.rs:303)
    const bool ok = (fmt_tag == 1 && (bits == 8 || bits == 16)) || (fmt_tag == 3 && bits == 32);
    WB_REQUIRE(ok, WB_EINVAL, "Unsupported decoded sample format");
    const size_t frames = data_len / ((size_t)bps * channels);
    mono.resize(frames);
    const float fc = (float)channels;
    for (size_t i = 0; i < frames; ++i) {
        const unsigned char* p = data + i * (size_t)bps * channels;
        float acc = 0.0f;
        for (int c = 0; c < channels; ++c, p += bps) {
            if (bits == 8) acc += ((float)p[0] - 128.0f) / 128.0f;                       // main.rs:280
            else if (bits == 16) acc += (float)(int16_t)rd16(p) / 32768.0f;              // main.rs:298
            else { float v; std::memcpy(&v, p, 4); acc += v; }                           // main.rs:271
        }
        mono[i] = acc / fc;
    }
}

}  // namespace

extern "C" int wb_host_load_audio_16k_mono(const char* path, float** pcm_out, int64_t* n_out, double* dur_out) {
    try {
        WB_REQUIRE(path && pcm_out && n_out, WB_EINVAL, "null argument");
        std::vector<float> mono;
        uint32_t sr = 0;
        load_wav(path, mono, sr);
        int64_t n = (int64_t)mono.size();
        float* out;
        if (sr != 16000) {                                                               // main.rs:309-312
            int64_t m = wb_host_resample_linear(mono.data(), n, sr, 16000, nullptr, 0);
            out = (float*)std::malloc(sizeof(float) * (size_t)(m > 0 ? m : 1));
            WB_REQUIRE(out, WB_EINVAL, "out of memory");
            wb_host_resample_linear(mono.data(), n, sr, 16000, out, m);
            n = m;
        } else {
            out = (float*)std::malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
            WB_REQUIRE(out, WB_EINVAL, "out of memory");
            std::memcpy(out, mono.data(), sizeof(float) * (size_t)n);
        }
        *pcm_out = out;
        *n_out = n;
        if (dur_out) *dur_out = (double)n / 16000.0;                                     // main.rs:314
        return WB_OK;
    } catch (const WbError& e) {
        wb_set_error(e.what());
        return e.code;
    } catch (const std::exception& e) {
        wb_set_error(e.what());
        return WB_EINVAL;
    }
}
