// audio.cpp — host-side audio ingest (replaces load_audio_16k_mono + resample_linear,
// /root/reference/src/main.rs:207-316, which sit on the symphonia crate).  RIFF/WAVE in the sample
// formats the reference accepts (U8 / S16 / F32; its match bails on S24 / S32 / F64, and so on every FLAC
// file, which symphonia decodes to S32), and MPEG Layer III streams (mp3.cpp; symphonia hands the reference F32
// planes for them, main.rs:266-275).  Like symphonia's probe, the container is recognised by content, not by
// the file extension.
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

namespace wbmp3 {                                  // mp3.cpp
bool probe(const uint8_t* p, size_t n, size_t* first);
void decode_to_mono(const uint8_t* p, size_t n, size_t first, std::vector<float>& mono, uint32_t& sr);
}

namespace {

uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

// A-law / mu-law expansion (ITU-T G.711), as symphonia-codec-pcm does for CODEC_TYPE_PCM_ALAW / _MULAW: S16 out.
int16_t alaw_to_s16(uint8_t a) {
    a ^= 0x55;
    int t = (a & 0x0f) << 4;
    const int seg = (a & 0x70) >> 4;
    if (seg == 0) t += 8; else if (seg == 1) t += 0x108; else { t += 0x108; t <<= seg - 1; }
    return (int16_t)((a & 0x80) ? t : -t);
}
int16_t mulaw_to_s16(uint8_t u) {
    u = ~u;
    int t = ((u & 0x0f) << 3) + 0x84;
    t <<= (u & 0x70) >> 4;
    return (int16_t)((u & 0x80) ? (0x84 - t) : (t - 0x84));
}

enum Codec { C_NONE, C_U8, C_S16, C_F32, C_ALAW, C_MULAW, C_OTHER /* decodes to S24/S32/F64: main.rs:303 bails */ };

// The RIFF/WAVE reader of symphonia-format-riff 0.5.5 (Cargo.lock:1184-1309; sources are not vendored in the
// reference, so this restates the published crate: wav/mod.rs WavReader::try_new, wav/chunks.rs, common.rs
// ChunksReader::next / next_packet) as far as load_audio_16k_mono (main.rs:228-316) can observe it:
//  * the RIFF header's length bounds the chunk walk; a chunk longer than what is left of its parent is a
//    decode error unless both lengths are 0xFFFFFFFF (what ffmpeg writes to a pipe); chunks are 2-byte aligned;
//    unknown chunks are skipped; the walk stops at the first `data` chunk; no `data` chunk = unsupported.
//  * `fmt `: PCM (len 16/18/40, 8/16/24/32 bit, 1-2 channels), IEEE float (len 16/18, 32/64 bit, 1-2 channels),
//    WAVE_FORMAT_EXTENSIBLE (len 40, channel-mask popcount == channel count, PCM / IEEE sub-formats), A-law / mu-law
//    (8 bit, 1-2 channels); everything else is rejected.
//  * packets are at most 1152 blocks of `block_align` bytes, cut from the DECLARED data length; a packet that runs
//    past the end of the file is an IoError, which the reference's loop turns into `break` (main.rs:258-262): the
//    whole partial packet is dropped, so an over-declared stream loses its last (n mod 1152) frames.  The pad byte
//    after an odd-length data chunk is never decoded.
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
    if (std::memcmp(buf.data(), "RIFF", 4) != 0) {
        size_t first = 0;
        if (wbmp3::probe(buf.data(), buf.size(), &first)) { wbmp3::decode_to_mono(buf.data(), buf.size(), first, mono, sr); return; }
        WB_THROW(WB_EINVAL, "unsupported audio container (RIFF/WAVE and MPEG Layer III are the decodable ones): %s", path);
    }
    if (std::memcmp(buf.data() + 8, "WAVE", 4) != 0) WB_THROW(WB_EINVAL, "wav: riff form is not wave");
    const uint64_t riff_len = rd32(buf.data() + 4);
    uint64_t consumed = 0;                       // ChunksReader::consumed (counts from after the form id, like the crate)
    size_t pos = 12;
    int channels = 0, bits = 0, block_align = 0;
    Codec codec = C_NONE;
    bool have_fmt = false;
    sr = 0;
    const unsigned char* data = nullptr;
    uint64_t data_decl = 0;
    for (;;) {
        if (consumed & 1) { WB_REQUIRE(pos < buf.size(), WB_EIO, "end of stream"); ++pos; ++consumed; }
        if (consumed + 8 > riff_len) WB_THROW(WB_EINVAL, "wav: missing data chunk");
        WB_REQUIRE(pos + 8 <= buf.size(), WB_EIO, "end of stream");
        const unsigned char* ck = buf.data() + pos;
        const uint64_t len = rd32(ck + 4);
        pos += 8; consumed += 8;
        if (riff_len - consumed < len && !(riff_len == len && len == 0xFFFFFFFFull))
            WB_THROW(WB_EINVAL, "riff: chunk length exceeds parent (list) chunk length");
        consumed = consumed + len > 0xFFFFFFFFull ? 0xFFFFFFFFull : consumed + len;       // u32 saturating_add
        if (std::memcmp(ck, "fmt ", 4) == 0) {
            WB_REQUIRE(len >= 16, WB_EINVAL, "wav: malformed fmt chunk");
            WB_REQUIRE(pos + len <= buf.size(), WB_EIO, "end of stream");
            const unsigned char* p = buf.data() + pos;
            int tag = rd16(p);
            channels = rd16(p + 2); sr = rd32(p + 4); block_align = rd16(p + 12); bits = rd16(p + 14);
            auto mono_or_stereo = [&](const char* what) {
                WB_REQUIRE(channels == 1 || channels == 2, WB_EINVAL, "wav: channel layout is not stereo or mono for %s", what);
            };
            if (tag == 1) {                                        // WAVE_FORMAT_PCM
                WB_REQUIRE(len == 16 || len == 18 || len == 40, WB_EINVAL, "wav: malformed fmt_pcm chunk");
                WB_REQUIRE(bits == 8 || bits == 16 || bits == 24 || bits == 32, WB_EINVAL, "wav: bits per sample for fmt_pcm must be 8, 16, 24 or 32 bits");
                mono_or_stereo("fmt_pcm");
                codec = bits == 8 ? C_U8 : bits == 16 ? C_S16 : C_OTHER;
            } else if (tag == 3) {                                 // WAVE_FORMAT_IEEE_FLOAT
                WB_REQUIRE(len == 16 || len == 18, WB_EINVAL, "wav: malformed fmt_ieee chunk");
                if (len == 18) WB_REQUIRE(rd16(p + 16) == 0, WB_EINVAL, "wav: extension length not 0 for fmt_ieee chunk");
                WB_REQUIRE(bits == 32 || bits == 64, WB_EINVAL, "wav: bits per sample for fmt_ieee must be 32 or 64 bits");
                mono_or_stereo("fmt_ieee");
                codec = bits == 32 ? C_F32 : C_OTHER;
            } else if (tag == 0xFFFE) {                            // WAVE_FORMAT_EXTENSIBLE
                WB_REQUIRE(len == 40, WB_EINVAL, "wav: malformed fmt_ext chunk");
                WB_REQUIRE(rd16(p + 16) == 22, WB_EINVAL, "wav: extension length not 22 for fmt_ext chunk");
                WB_REQUIRE((bits & 7) == 0, WB_EINVAL, "wav: bits per sample for fmt_ext must be a multiple of 8");
                WB_REQUIRE(rd16(p + 18) <= bits, WB_EINVAL, "wav: bits per sample exceeds coded bits per sample for fmt_ext");
                const uint32_t mask = rd32(p + 20);
                WB_REQUIRE(__builtin_popcount(mask) == channels, WB_EINVAL, "wav: channel mask mismatch for fmt_ext");
                WB_REQUIRE((mask >> 26) == 0, WB_EINVAL, "wav: too many channel masks");     // Channels::from_bits: 26 positions
                static const unsigned char kTail[14] = {0x00, 0x00, 0x00, 0x00, 0x10, 0x00, 0x80, 0x00, 0x00, 0xaa, 0x00, 0x38, 0x9b, 0x71};
                WB_REQUIRE(std::memcmp(p + 26, kTail, 14) == 0, WB_EINVAL, "wav: unsupported fmt_ext sub-type");
                const int sub = rd16(p + 24);
                if (sub == 1) {
                    WB_REQUIRE(bits == 8 || bits == 16 || bits == 24 || bits == 32, WB_EINVAL, "wav: bits per sample for fmt_ext PCM sub-type must be 8, 16, 24 or 32 bits");
                    codec = bits == 8 ? C_U8 : bits == 16 ? C_S16 : C_OTHER;
                } else if (sub == 3) {
                    WB_REQUIRE(bits == 32 || bits == 64, WB_EINVAL, "wav: bits per sample for fmt_ext IEEE sub-type must be 32 or 64 bits");
                    codec = bits == 32 ? C_F32 : C_OTHER;
                } else {
                    WB_THROW(WB_EINVAL, "wav: unsupported fmt_ext sub-type");
                }
            } else if (tag == 6 || tag == 7) {                     // WAVE_FORMAT_ALAW / MULAW -> S16 (accepted, main.rs:293-300)
                WB_REQUIRE(len == 18, WB_EINVAL, "wav: malformed fmt_alaw/fmt_mulaw chunk");
                WB_REQUIRE(bits == 8, WB_EINVAL, "wav: bits per sample for fmt_alaw/fmt_mulaw must be 8 bits");
                mono_or_stereo(tag == 6 ? "fmt_alaw" : "fmt_mulaw");
                codec = tag == 6 ? C_ALAW : C_MULAW;
            } else if (tag == 2 || tag == 0x11) {                  // MS / IMA ADPCM decode to S32 buffers -> main.rs:303
                codec = C_OTHER;
            } else {
                WB_THROW(WB_EINVAL, "wav: unsupported wave format");
            }
            have_fmt = true;
        } else if (std::memcmp(ck, "data", 4) == 0) {
            data = buf.data() + pos;
            data_decl = len;
            break;
        }
        pos += (size_t)len;                                         // known-but-unused and unknown chunks are skipped
        WB_REQUIRE(pos <= buf.size(), WB_EIO, "end of stream");
    }
    WB_REQUIRE(have_fmt, WB_EINVAL, "No default track");                                 // main.rs:249 (no codec parameters yet)
    WB_REQUIRE(sr > 0, WB_EINVAL, "Unknown sample rate");                                // main.rs:252
    WB_REQUIRE(channels > 0, WB_EINVAL, "Unknown channels");                             // main.rs:253
    // symphonia hands back U8 / S16 / F32 for these; S24/S32/F64 hit `_ => bail!` (main.rs:303)
    WB_REQUIRE(codec != C_OTHER, WB_EINVAL, "Unsupported decoded sample format");
    const int bps = bits / 8;
    WB_REQUIRE(block_align > 0, WB_EINVAL, "riff: block size is 0");
    WB_REQUIRE(block_align == bps * channels, WB_EINVAL, "wav: block_align %d does not match %d channels x %d bytes", block_align, channels, bps);
    // packet walk over the declared length (common.rs next_packet): 1152 blocks per packet, partial packets at EOF dropped
    const uint64_t avail = buf.size() - (size_t)(data - buf.data());
    const uint64_t blocks_decl = data_decl / (uint64_t)block_align;
    uint64_t frames = 0;
    while (frames < blocks_decl) {
        const uint64_t n = blocks_decl - frames < 1152 ? blocks_decl - frames : 1152;
        if ((frames + n) * (uint64_t)block_align > avail) break;                         // read_boxed_slice_exact -> UnexpectedEof
        frames += n;
    }
    mono.resize((size_t)frames);
    const float fc = (float)channels;
    for (size_t i = 0; i < (size_t)frames; ++i) {
        const unsigned char* p = data + i * (size_t)block_align;
        float acc = 0.0f;
        for (int c = 0; c < channels; ++c, p += bps) {
            switch (codec) {
                case C_U8: acc += ((float)p[0] - 128.0f) / 128.0f; break;                // main.rs:280
                case C_S16: acc += (float)(int16_t)rd16(p) / 32768.0f; break;            // main.rs:298
                case C_ALAW: acc += (float)alaw_to_s16(p[0]) / 32768.0f; break;
                case C_MULAW: acc += (float)mulaw_to_s16(p[0]) / 32768.0f; break;
                default: { float v; std::memcpy(&v, p, 4); acc += v; }                   // main.rs:271
            }
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
