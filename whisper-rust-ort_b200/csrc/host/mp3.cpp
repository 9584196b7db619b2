// mp3.cpp — MPEG-1 / MPEG-2 LSF / MPEG-2.5 Layer III ingest for wb_host_load_audio_16k_mono: the one compressed
// container the reference both lists (main.rs:1116, Cargo.toml:19 features = [... "mp3" ...]) and can use (symphonia's
// MPA decoder hands back F32 planes -> main.rs:266-275).  Written from the published algorithm (ISO/IEC 11172-3
// clause 2.4, 13818-3 for the low-sampling-frequency extension) in its plain textbook form: this is host-side file
// ingest, a few hundred MFLOP per file, not a hot path.  Constant tables: mp3_tables.h (generated, see its header).
//
// What the reference's loop (main.rs:228-316) can observe of symphonia-bundle-mp3 0.5.x is kept:
//  * an ID3v2 tag in front is skipped; the stream starts at the first frame header that is followed by a second one;
//  * a first frame carrying a Xing / Info / VBRI tag is metadata, not audio (the demuxer does not hand it out);
//    FormatOptions::default() has gapless off, so encoder delay and padding are NOT trimmed;
//  * every frame yields 1152 (MPEG-1) or 576 (LSF) samples per channel; the channel count is the first frame's;
//  * a frame cut off by the end of the file is an IoError -> `break` (main.rs:258-262): dropped;
//  * Layer I / II (cargo feature "mp3" builds Layer III only) and free-format streams are errors.
// Known limit: mixed blocks at 8 kHz (MPEG-2.5), where the long-window part spans four subbands instead of two, are
// windowed as if it spanned two (libavcodec refuses that combination as well; no encoder is known to produce it).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../common.h"
#include "mp3_tables.h"

#pragma GCC optimize("O3")        // lets the y[i] += x[k] * table[k][i] loops vectorise (no reassociation: -ffp-contract=off, no fast-math)
// the two filterbank routines are also built for AVX2 and picked at load time (same operations in the same order, 8 lanes wide)
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define WB_MP3_CLONES __attribute__((target_clones("avx2", "default")))
#else
#define WB_MP3_CLONES
#endif

namespace wbmp3 {
namespace {

constexpr double kPi = 3.14159265358979323846;

struct Bits {                                   // MSB-first reader over a byte range; reads past the end give zeros
    const uint8_t* p; size_t n_bits; size_t pos;
    uint32_t get(int n) {
        uint32_t v = 0;
        for (int i = 0; i < n; ++i, ++pos) v = (v << 1) | (pos < n_bits ? (p[pos >> 3] >> (7 - (pos & 7))) & 1u : 0u);
        return v;
    }
};

struct Header {
    int version;            // 0 = MPEG-1, 1 = MPEG-2 (LSF), 2 = MPEG-2.5
    int layer;              // 1, 2, 3
    bool crc;
    int bitrate;            // bit/s, 0 = free format
    int sr, band_row;
    int mode, mode_ext, channels;
    size_t frame_size, side_size;
    int granules;
};

bool parse_header(const uint8_t* b, Header& h) {
    const uint32_t w = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
    if ((w >> 21) != 0x7FFu) return false;
    const int vbits = (w >> 19) & 3, lbits = (w >> 17) & 3, br = (w >> 12) & 15, sri = (w >> 10) & 3;
    if (vbits == 1 || lbits == 0 || br == 15 || sri == 3) return false;
    h.version = vbits == 3 ? 0 : vbits == 2 ? 1 : 2;
    h.layer = 4 - lbits;
    h.crc = ((w >> 16) & 1) == 0;
    static const int kSr[3][3] = {{44100, 48000, 32000}, {22050, 24000, 16000}, {11025, 12000, 8000}};
    h.sr = kSr[h.version][sri];
    h.band_row = h.version * 3 + sri;
    static const int kBrV1L3[15] = {0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320};
    static const int kBrV2L3[15] = {0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160};
    h.bitrate = 1000 * (h.version == 0 ? kBrV1L3[br] : kBrV2L3[br]);           // Layer III columns; other layers are rejected by the caller
    const int pad = (w >> 9) & 1;
    h.mode = (w >> 6) & 3;
    h.mode_ext = (w >> 4) & 3;
    h.channels = h.mode == 3 ? 1 : 2;
    h.granules = h.version == 0 ? 2 : 1;
    h.side_size = h.version == 0 ? (h.channels == 1 ? 17 : 32) : (h.channels == 1 ? 9 : 17);
    h.frame_size = h.bitrate ? (size_t)((h.version == 0 ? 144 : 72) * (int64_t)h.bitrate / h.sr + pad) : 0;
    return true;
}

struct Granule {
    int part2_3_length, big_values, global_gain, scalefac_compress;
    bool window_switching, mixed;
    int block_type, table_select[3], subblock_gain[3], region0_count, region1_count;
    bool preflag, scalefac_scale, count1_table;
    int sf_l[23], sf_s[13][3];
    bool is_bad_l[23], is_bad_s[13][3];     // right channel in intensity stereo: the scalefactor is an ILLEGAL intensity position
};

struct Trie { std::vector<int16_t> next; };      // node i: next[2i + bit] >= 0 -> child node, < 0 -> ~symbol

struct Tables {
    Trie book[15];
    int book_of[32], linbits[32];
    Trie quad_a;
    // cosine tables are stored input-major (table[k][i]) so that the sums below run as vectorisable y[i] += x[k] * table[k][i]
    // updates that keep the textbook summation order
    float win[4][36], cs[8], ca[8], synth_cos[32][64], synth_win[512], imdct36[18][36], imdct12[6][12];
    Tables() {
        for (int b = 0; b < 15; ++b) {
            const int n = kBookDim[b] * kBookDim[b];
            Trie& t = book[b];
            t.next.assign(2, 0);
            uint64_t acc = 0;
            for (int i = 0; i < n; ++i) {
                const int len = kHuffLen[kBookOff[b] + i];
                insert(t, (uint32_t)(acc >> (32 - len)), len, kHuffSym[kBookOff[b] + i]);
                acc += 1ull << (32 - len);
            }
        }
        quad_a.next.assign(2, 0);
        for (int s = 0; s < 16; ++s) insert(quad_a, kQuadACode[s], kQuadALen[s], s);
        static const int kLin[16] = {1, 2, 3, 4, 6, 8, 10, 13, 4, 5, 6, 7, 8, 9, 11, 13};
        for (int t = 0; t < 32; ++t) {
            book_of[t] = -1; linbits[t] = 0;
            if (t >= 16) { book_of[t] = t < 24 ? 13 : 14; linbits[t] = kLin[t - 16]; }
            else for (int b = 0; b < 13; ++b) if (kBookIds[b] == t) book_of[t] = b;
        }
        for (int i = 0; i < 36; ++i) {
            const double s36 = std::sin(kPi / 36.0 * (i + 0.5));
            win[0][i] = (float)s36;
            win[1][i] = (float)(i < 18 ? s36 : i < 24 ? 1.0 : i < 30 ? std::sin(kPi / 12.0 * (i - 18 + 0.5)) : 0.0);
            win[3][i] = (float)(i < 6 ? 0.0 : i < 12 ? std::sin(kPi / 12.0 * (i - 6 + 0.5)) : i < 18 ? 1.0 : s36);
            win[2][i] = (float)(i < 12 ? std::sin(kPi / 12.0 * (i + 0.5)) : 0.0);
            for (int k = 0; k < 18; ++k) imdct36[k][i] = (float)std::cos(kPi / 72.0 * (2 * i + 1 + 18) * (2 * k + 1));
        }
        for (int i = 0; i < 12; ++i)
            for (int k = 0; k < 6; ++k) imdct12[k][i] = (float)std::cos(kPi / 24.0 * (2 * i + 1 + 6) * (2 * k + 1));
        static const double kC[8] = {-0.6, -0.535, -0.33, -0.185, -0.095, -0.041, -0.0142, -0.0037};
        for (int i = 0; i < 8; ++i) { cs[i] = (float)(1.0 / std::sqrt(1.0 + kC[i] * kC[i])); ca[i] = (float)(kC[i] / std::sqrt(1.0 + kC[i] * kC[i])); }
        for (int i = 0; i < 64; ++i)
            for (int k = 0; k < 32; ++k) synth_cos[k][i] = (float)std::cos((16 + i) * (2 * k + 1) * kPi / 64.0);
        for (int i = 0; i < 512; ++i) synth_win[i] = (float)kSynthWindowQ16[i] * (1.0f / 65536.0f);
    }
    static void insert(Trie& t, uint32_t code, int len, int sym) {
        int node = 0;
        for (int i = len - 1; i >= 0; --i) {
            const int bit = (code >> i) & 1;
            if (i == 0) { t.next[2 * node + bit] = (int16_t)~sym; break; }
            int child = t.next[2 * node + bit];
            if (child == 0) {
                child = (int)(t.next.size() / 2);
                t.next.push_back(0); t.next.push_back(0);
                t.next[2 * node + bit] = (int16_t)child;
            }
            node = child;
        }
    }
    static int decode(const Trie& t, Bits& b) {
        int node = 0;
        for (;;) {
            const int16_t v = t.next[2 * node + (int)b.get(1)];
            if (v < 0) return ~v;
            if (v == 0) return 0;               // unreachable for a complete code
            node = v;
        }
    }
};

const Tables& tables() { static const Tables t; return t; }

struct ChannelState { float overlap[32][18]; float v[1024]; int v_off; };

struct Decoder {
    std::vector<uint8_t> reservoir;
    ChannelState ch[2];
    Decoder() { std::memset(ch, 0, sizeof ch); }

    void read_side_info(const Header& h, Bits& b, int& main_data_begin, int scfsi[2][4], Granule g[2][2]) {
        const bool v1 = h.version == 0;
        main_data_begin = (int)b.get(v1 ? 9 : 8);
        b.get(v1 ? (h.channels == 1 ? 5 : 3) : (h.channels == 1 ? 1 : 2));
        for (int c = 0; c < h.channels; ++c) for (int k = 0; k < 4; ++k) scfsi[c][k] = v1 ? (int)b.get(1) : 0;
        for (int gr = 0; gr < h.granules; ++gr)
            for (int c = 0; c < h.channels; ++c) {
                Granule& q = g[gr][c];
                std::memset(&q, 0, sizeof q);
                q.part2_3_length = (int)b.get(12);
                q.big_values = (int)b.get(9);
                q.global_gain = (int)b.get(8);
                q.scalefac_compress = (int)b.get(v1 ? 4 : 9);
                q.window_switching = b.get(1) != 0;
                if (q.window_switching) {
                    q.block_type = (int)b.get(2);
                    q.mixed = b.get(1) != 0;
                    WB_REQUIRE(q.block_type != 0, WB_EINVAL, "mpa: invalid block_type");
                    for (int r = 0; r < 2; ++r) q.table_select[r] = (int)b.get(5);
                    for (int w = 0; w < 3; ++w) q.subblock_gain[w] = (int)b.get(3);
                    q.region0_count = (q.block_type == 2 && !q.mixed) ? 8 : 7;
                    q.region1_count = 36;
                } else {
                    for (int r = 0; r < 3; ++r) q.table_select[r] = (int)b.get(5);
                    q.region0_count = (int)b.get(4);
                    q.region1_count = (int)b.get(3);
                }
                q.preflag = v1 ? b.get(1) != 0 : false;
                q.scalefac_scale = b.get(1) != 0;
                q.count1_table = b.get(1) != 0;
                WB_REQUIRE(q.big_values <= 288, WB_EINVAL, "mpa: granule big_values > 288");
            }
    }

    // 11172-3 2.4.2.7 scalefactors (MPEG-1)
    void read_scalefactors_v1(Bits& b, Granule& q, const Granule* gr0, const int scfsi[4]) {
        static const int kSlen[2][16] = {{0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4}, {0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 2, 3}};
        const int s1 = kSlen[0][q.scalefac_compress], s2 = kSlen[1][q.scalefac_compress];
        if (q.window_switching && q.block_type == 2) {
            int sfb0 = 0;
            if (q.mixed) { for (int s = 0; s < 8; ++s) q.sf_l[s] = (int)b.get(s1); sfb0 = 3; }
            for (int s = sfb0; s < 12; ++s) for (int w = 0; w < 3; ++w) q.sf_s[s][w] = (int)b.get(s < 6 ? s1 : s2);
        } else {
            static const int kEdge[5] = {0, 6, 11, 16, 21};
            for (int k = 0; k < 4; ++k)
                for (int s = kEdge[k]; s < kEdge[k + 1]; ++s)
                    q.sf_l[s] = (gr0 && scfsi[k]) ? gr0->sf_l[s] : (int)b.get(k < 2 ? s1 : s2);
        }
        for (int s = 0; s < 23; ++s) q.is_bad_l[s] = q.sf_l[s] >= 7;                   // 2.4.3.4: is_pos = 7 means "not intensity coded"
        for (int s = 0; s < 13; ++s) for (int w = 0; w < 3; ++w) q.is_bad_s[s][w] = q.sf_s[s][w] >= 7;
    }

    // 13818-3 2.4.3.2 scalefactors (LSF): four partitions, widths from scalefac_compress
    void read_scalefactors_lsf(Bits& b, Granule& q, bool intensity_right) {
        int sfc = q.scalefac_compress, slen[4], row;
        if (!intensity_right) {
            if (sfc < 400) { slen[0] = (sfc >> 4) / 5; slen[1] = (sfc >> 4) % 5; slen[2] = (sfc % 16) >> 2; slen[3] = sfc % 4; row = 0; }
            else if (sfc < 500) { sfc -= 400; slen[0] = (sfc >> 2) / 5; slen[1] = (sfc >> 2) % 5; slen[2] = sfc % 4; slen[3] = 0; row = 1; }
            else { sfc -= 500; slen[0] = sfc / 3; slen[1] = sfc % 3; slen[2] = slen[3] = 0; row = 2; q.preflag = true; }
        } else {
            sfc >>= 1;
            if (sfc < 180) { slen[0] = sfc / 36; slen[1] = (sfc % 36) / 6; slen[2] = (sfc % 36) % 6; slen[3] = 0; row = 3; }
            else if (sfc < 244) { sfc -= 180; slen[0] = (sfc % 64) >> 4; slen[1] = (sfc % 16) >> 2; slen[2] = sfc % 4; slen[3] = 0; row = 4; }
            else { sfc -= 244; slen[0] = sfc / 3; slen[1] = sfc % 3; slen[2] = slen[3] = 0; row = 5; }
        }
        const int col = (q.window_switching && q.block_type == 2) ? (q.mixed ? 2 : 1) : 0;
        int flat[48] = {0}, n = 0;
        bool bad[48] = {false};                                      // 13818-3 2.4.3.2: the all-ones value of a field is the illegal position
        for (int k = 0; k < 4; ++k)
            for (int i = 0; i < kLsfPartitions[row][col][k]; ++i, ++n) {
                flat[n] = (int)b.get(slen[k]);
                bad[n] = slen[k] > 0 && flat[n] == (1 << slen[k]) - 1;
            }
        if (col == 0) { for (int s = 0; s < 21; ++s) { q.sf_l[s] = flat[s]; q.is_bad_l[s] = bad[s]; } }
        else {
            int i = 0, sfb0 = 0;
            if (q.mixed) { for (int s = 0; s < 6; ++s, ++i) { q.sf_l[s] = flat[i]; q.is_bad_l[s] = bad[i]; } sfb0 = 3; }
            for (int s = sfb0; s < 12; ++s) for (int w = 0; w < 3; ++w, ++i) { q.sf_s[s][w] = flat[i]; q.is_bad_s[s][w] = bad[i]; }
        }
    }

    // 2.4.2.7 Huffman code bits: big_values pairs in three regions, then count1 quadruples up to part2_3_length
    void read_spectrum(const Header& h, Bits& b, const Granule& q, size_t end, int is[576]) {
        const Tables& T = tables();
        std::memset(is, 0, sizeof(int) * 576);
        int edge[23]; edge[0] = 0;
        for (int s = 0; s < 22; ++s) edge[s + 1] = edge[s] + kBandLong[h.band_row][s];
        int r1, r2;
        if (q.window_switching) {
            // region0 ends where 2.4.2.7 puts it for window-switched granules: 36 lines (54 for long windows at LSF rates, doubled at 8 kHz)
            r1 = q.block_type == 2 ? (h.band_row == 8 ? 72 : 36) : (h.version == 0 ? 36 : h.band_row == 8 ? 108 : 54);
            r2 = 576;
        } else {
            const int a = q.region0_count + 1, c = a + q.region1_count + 1;
            r1 = edge[a > 22 ? 22 : a]; r2 = edge[c > 22 ? 22 : c];
        }
        const int bv = q.big_values * 2;
        const int lim[3] = {r1 < bv ? r1 : bv, r2 < bv ? r2 : bv, bv};
        int pos = 0;
        for (int r = 0; r < 3; ++r) {
            const int t = q.table_select[r];
            const int bk = T.book_of[t], lb = T.linbits[t];
            WB_REQUIRE(t == 0 || bk >= 0, WB_EINVAL, "mpa: invalid huffman table %d", t);
            for (; pos < lim[r]; pos += 2) {
                if (t == 0) continue;
                if (b.pos >= end) { pos = 576; break; }
                const int sym = Tables::decode(T.book[bk], b);
                int x = sym >> 4, y = sym & 15;
                if (x == 15 && lb) x += (int)b.get(lb);
                if (x && b.get(1)) x = -x;
                if (y == 15 && lb) y += (int)b.get(lb);
                if (y && b.get(1)) y = -y;
                is[pos] = x; is[pos + 1] = y;
            }
        }
        for (pos = bv; pos <= 572 && b.pos < end; pos += 4) {
            const int sym = q.count1_table ? (int)(15 - b.get(4)) : Tables::decode(T.quad_a, b);
            int v[4] = {(sym >> 3) & 1, (sym >> 2) & 1, (sym >> 1) & 1, sym & 1};
            for (int k = 0; k < 4; ++k) if (v[k] && b.get(1)) v[k] = -1;
            if (b.pos > end) break;                                  // the quadruple ran past part2_3_length: not part of the granule
            for (int k = 0; k < 4; ++k) is[pos + k] = v[k];
        }
        b.pos = end;
    }

    // 2.4.3.4 requantisation, in bitstream order (short windows still band by band)
    void requantize(const Header& h, const Granule& q, const int is[576], float xr[576]) {
        static const int kPretab[22] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 2, 0};
        const double mult = q.scalefac_scale ? 1.0 : 0.5;
        static const std::vector<double> pow43 = [] {                // |is| <= 15 + 2^13 - 1
            std::vector<double> t(8207);
            for (int v = 0; v < 8207; ++v) t[(size_t)v] = std::pow((double)v, 4.0 / 3.0);
            return t;
        }();
        auto deq = [](int v, double gain) -> float {
            if (v == 0) return 0.0f;
            const double m = pow43[(size_t)(v < 0 ? -v : v)] * gain;
            return (float)(v < 0 ? -m : m);
        };
        const bool shortb = q.window_switching && q.block_type == 2;
        int i = 0;
        const int long_bands = !shortb ? 22 : q.mixed ? (h.version == 0 ? 8 : 6) : 0;
        for (int s = 0; s < long_bands; ++s) {
            const double g = std::exp2((q.global_gain - 210) / 4.0 - mult * ((s < 21 ? q.sf_l[s] : 0) + (q.preflag ? kPretab[s] : 0)));
            for (int k = 0; k < kBandLong[h.band_row][s]; ++k, ++i) xr[i] = deq(is[i], g);
        }
        if (!shortb) return;
        for (int s = q.mixed ? 3 : 0; s < 13; ++s) {
            const int w = kBandShort[h.band_row][s];
            for (int win = 0; win < 3; ++win) {
                const double g = std::exp2((q.global_gain - 210 - 8 * q.subblock_gain[win]) / 4.0 - mult * (s < 12 ? q.sf_s[s][win] : 0));
                for (int k = 0; k < w; ++k, ++i) xr[i] = deq(is[i], g);
            }
        }
    }

    // short windows: band-by-band order of the bitstream -> frequency-interleaved order (line 3 f + window)
    void reorder(const Header& h, const Granule& q, float xr[576]) {
        if (!(q.window_switching && q.block_type == 2)) return;
        float tmp[576];
        std::memcpy(tmp, xr, sizeof tmp);
        const int s0 = q.mixed ? 3 : 0;
        int start = 0;
        for (int s = 0; s < s0; ++s) start += kBandShort[h.band_row][s];
        int i = 3 * start;
        for (int s = s0; s < 13; ++s) {
            const int w = kBandShort[h.band_row][s];
            for (int win = 0; win < 3; ++win)
                for (int k = 0; k < w; ++k, ++i) xr[3 * (start + k) + win] = tmp[i];
            start += w;
        }
    }

    // 2.4.3.4 stereo processing on the two channels of a granule, in bitstream order.  MS: (m + s, m - s) / sqrt 2.  Intensity:
    // scanning down from the top, bands (per window for short blocks) in which the right channel is still all zero take
    // the left channel's lines split by the ratio the right channel's scalefactor encodes (MPEG-1: tan(is_pos pi / 12);
    // LSF: powers of 2^-1/4 or 2^-1/2, 13818-3 2.4.3.2); the band above the last scalefactor band reuses the last position;
    // an illegal position, and everything from the right channel's highest non-zero band down, is plain or MS stereo.
    void stereo(const Header& h, const Granule& gr, float* l, float* r) {
        const bool ms = (h.mode_ext & 2) != 0;
        auto mid_side = [&](int at, int n) {
            for (int j = at; j < at + n; ++j) {
                const float m = l[j], s = r[j];
                l[j] = (m + s) * 0.70710678118654752f;
                r[j] = (m - s) * 0.70710678118654752f;
            }
        };
        if (!(h.mode_ext & 1)) { if (ms) mid_side(0, 576); return; }
        auto intensity = [&](int at, int n, int pos) {
            float kl, kr;
            if (h.version == 0) {
                const double t = std::tan(pos * kPi / 12.0);
                kl = pos == 6 ? 1.0f : (float)(t / (1.0 + t));
                kr = pos == 6 ? 0.0f : (float)(1.0 / (1.0 + t));
            } else {
                const double f = std::exp2(-(double)((gr.scalefac_compress & 1) + 1) * ((pos + 1) >> 1) / 4.0);
                kl = (pos & 1) ? (float)f : 1.0f;
                kr = (pos & 1) ? 1.0f : (float)f;
            }
            for (int j = at; j < at + n; ++j) { const float x = l[j]; l[j] = x * kl; r[j] = x * kr; }
        };
        auto silent = [&](int at, int n) { for (int j = at; j < at + n; ++j) if (r[j] != 0.0f) return false; return true; };
        const bool shortb = gr.window_switching && gr.block_type == 2;
        const int long_end = !shortb ? 22 : gr.mixed ? (h.version == 0 ? 8 : 6) : 0;
        const int short_start = !shortb ? 13 : gr.mixed ? 3 : 0;
        int at = 576;
        bool found_w[3] = {false, false, false};
        for (int s = 12; s >= short_start; --s) {
            const int n = kBandShort[h.band_row][s], sf = s == 12 ? 11 : s;
            for (int w = 2; w >= 0; --w) {
                at -= n;
                if (!found_w[w] && !silent(at, n)) found_w[w] = true;
                if (!found_w[w] && !gr.is_bad_s[sf][w]) intensity(at, n, gr.sf_s[sf][w]);
                else if (ms) mid_side(at, n);
            }
        }
        bool found = found_w[0] || found_w[1] || found_w[2];
        for (int s = long_end - 1; s >= 0; --s) {
            const int n = kBandLong[h.band_row][s], sf = s == 21 ? 20 : s;
            at -= n;
            if (!found && !silent(at, n)) found = true;
            if (!found && !gr.is_bad_l[sf]) intensity(at, n, gr.sf_l[sf]);
            else if (ms) mid_side(at, n);
        }
    }

    // 2.4.3.4 alias reduction, IMDCT with windowing and overlap-add, frequency inversion
    WB_MP3_CLONES void hybrid(const Granule& q, float xr[576], ChannelState& st, float sb_out[18][32]) {
        const Tables& T = tables();
        const bool shortb = q.window_switching && q.block_type == 2;
        const int alias_sb = shortb ? (q.mixed ? 2 : 0) : 32;
        for (int sb = 1; sb < alias_sb; ++sb)
            for (int i = 0; i < 8; ++i) {
                const float a = xr[18 * sb - 1 - i], b = xr[18 * sb + i];
                xr[18 * sb - 1 - i] = a * T.cs[i] - b * T.ca[i];
                xr[18 * sb + i] = b * T.cs[i] + a * T.ca[i];
            }
        for (int sb = 0; sb < 32; ++sb) {
            float out[36];
            const float* X = xr + 18 * sb;
            bool any = false;
            for (int k = 0; k < 18; ++k) any |= X[k] != 0.0f;
            if (!any) {                                                // an empty subband (most of the upper ones at speech bit rates)
                std::memset(out, 0, sizeof out);
            } else if (shortb && !(q.mixed && sb < 2)) {
                std::memset(out, 0, sizeof out);
                for (int win = 0; win < 3; ++win) {
                    float y[12] = {0};
                    for (int k = 0; k < 6; ++k) for (int i = 0; i < 12; ++i) y[i] += X[3 * k + win] * T.imdct12[k][i];
                    for (int i = 0; i < 12; ++i) out[6 + 6 * win + i] += y[i] * T.win[2][i];
                }
            } else {
                const float* w = T.win[(q.window_switching && !(q.mixed && sb < 2)) ? q.block_type : 0];
                float y[36] = {0};
                for (int k = 0; k < 18; ++k) for (int i = 0; i < 36; ++i) y[i] += X[k] * T.imdct36[k][i];
                for (int i = 0; i < 36; ++i) out[i] = y[i] * w[i];
            }
            for (int i = 0; i < 18; ++i) {
                float v = out[i] + st.overlap[sb][i];
                st.overlap[sb][i] = out[18 + i];
                if ((sb & 1) && (i & 1)) v = -v;
                sb_out[i][sb] = v;
            }
        }
    }

    // 2.4.3.2 synthesis subband filter: 32 subband samples -> 32 PCM samples
    WB_MP3_CLONES void synth(ChannelState& st, const float s[32], float* pcm) {
        const Tables& T = tables();
        st.v_off = (st.v_off - 64) & 1023;                           // a multiple of 64: the 64- and 32-sample runs below never wrap
        float acc[64] = {0};                                         // local accumulators: no aliasing with the tables, stay in registers
        for (int k = 0; k < 32; ++k) {
            const float sk = s[k];
            const float* c = T.synth_cos[k];
            for (int i = 0; i < 64; ++i) acc[i] += c[i] * sk;
        }
        std::memcpy(st.v + st.v_off, acc, sizeof acc);
        float out[32] = {0};
        for (int i = 0; i < 8; ++i) {
            const float* a = st.v + ((st.v_off + 128 * i) & 1023);
            const float* b = st.v + ((st.v_off + 128 * i + 96) & 1023);
            const float* wa = T.synth_win + 64 * i;
            for (int j = 0; j < 32; ++j) { out[j] += a[j] * wa[j]; out[j] += b[j] * wa[32 + j]; }
        }
        std::memcpy(pcm, out, sizeof out);
    }

    // one frame -> granules * 576 samples per channel, appended to out[c]
    void decode_frame(const Header& h, const uint8_t* frame, std::vector<float> out[2]) {
        const size_t head = 4 + (h.crc ? 2 : 0);
        WB_REQUIRE(h.frame_size > head + h.side_size, WB_EINVAL, "mpa: frame too small");
        Bits sb{frame + head, h.side_size * 8, 0};
        int main_data_begin, scfsi[2][4];
        Granule g[2][2];
        read_side_info(h, sb, main_data_begin, scfsi, g);
        const uint8_t* main_data = frame + head + h.side_size;
        const size_t main_len = h.frame_size - head - h.side_size;
        const bool underflow = (size_t)main_data_begin > reservoir.size();
        std::vector<uint8_t> data;
        if (!underflow) data.assign(reservoir.end() - main_data_begin, reservoir.end());
        data.insert(data.end(), main_data, main_data + main_len);
        reservoir.insert(reservoir.end(), main_data, main_data + main_len);
        if (reservoir.size() > 4096) reservoir.erase(reservoir.begin(), reservoir.end() - 2048);
        const size_t n = (size_t)h.granules * 576;
        if (underflow) {                                             // the referenced bytes were never seen: the frame decodes to silence
            for (int c = 0; c < h.channels; ++c) out[c].insert(out[c].end(), n, 0.0f);
            return;
        }
        const bool joint = h.mode == 1;
        Bits b{data.data(), data.size() * 8, 0};
        for (int gr = 0; gr < h.granules; ++gr) {
            static thread_local float xr[2][576];
            static thread_local int is[576];
            for (int c = 0; c < h.channels; ++c) {
                Granule& q = g[gr][c];
                const size_t end = b.pos + (size_t)q.part2_3_length;
                if (h.version == 0) read_scalefactors_v1(b, q, gr == 1 ? &g[0][c] : nullptr, scfsi[c]);
                else read_scalefactors_lsf(b, q, joint && (h.mode_ext & 1) && c == 1);
                WB_REQUIRE(b.pos <= end, WB_EINVAL, "mpa: part2_3_length shorter than the scalefactors");
                read_spectrum(h, b, q, end, is);
                std::memset(xr[c], 0, sizeof xr[c]);
                requantize(h, q, is, xr[c]);
            }
            if (joint) stereo(h, g[gr][1], xr[0], xr[1]);
            for (int c = 0; c < h.channels; ++c) reorder(h, g[gr][c], xr[c]);
            for (int c = 0; c < h.channels; ++c) {
                float sbs[18][32];
                hybrid(g[gr][c], xr[c], ch[c], sbs);
                const size_t at = out[c].size();
                out[c].resize(at + 576);
                for (int t = 0; t < 18; ++t) synth(ch[c], sbs[t], out[c].data() + at + 32 * t);
            }
        }
    }
};

size_t id3v2_size(const uint8_t* p, size_t n) {
    if (n < 10 || std::memcmp(p, "ID3", 3) != 0) return 0;
    const size_t body = ((size_t)(p[6] & 0x7f) << 21) | ((size_t)(p[7] & 0x7f) << 14) | ((size_t)(p[8] & 0x7f) << 7) | (size_t)(p[9] & 0x7f);
    return 10 + body + ((p[5] & 0x10) ? 10 : 0);
}

bool same_stream(const Header& a, const Header& b) { return a.version == b.version && a.layer == b.layer && a.sr == b.sr; }

}  // namespace

// True when the bytes look like an MPEG audio stream (optionally behind an ID3v2 tag): two consecutive frame headers of
// one stream within the first 1 MiB, the depth symphonia's probe searches.  *first = offset of the first frame.
bool probe(const uint8_t* p, size_t n, size_t* first) {
    size_t pos = 0;
    for (size_t t; (t = id3v2_size(p + pos, n - pos)) != 0 && pos + t <= n;) pos += t;
    const size_t limit = pos + (1u << 20);
    for (; pos + 4 <= n && pos < limit; ++pos) {
        Header h, h2;
        if (p[pos] != 0xFF || !parse_header(p + pos, h) || h.frame_size < 4) continue;
        if (pos + h.frame_size + 4 <= n) { if (!parse_header(p + pos + h.frame_size, h2) || !same_stream(h, h2)) continue; }
        else if (pos + h.frame_size != n) continue;
        *first = pos;
        return true;
    }
    return false;
}

// Whole-file decode to the reference's mono mix: mean over channels in f32, left to right (main.rs:266-275).
void decode_to_mono(const uint8_t* p, size_t n, size_t first, std::vector<float>& mono, uint32_t& sr) {
    Header h0;
    WB_REQUIRE(parse_header(p + first, h0), WB_EINVAL, "mpa: no frame header");
    WB_REQUIRE(h0.layer == 3, WB_EINVAL, "unsupported codec: MPEG audio layer %d (the reference builds symphonia with Layer III only)", h0.layer);
    WB_REQUIRE(h0.bitrate != 0, WB_EINVAL, "mpa: free bit-rate is not supported");
    sr = (uint32_t)h0.sr;
    const int channels = h0.channels;
    size_t pos = first;
    {   // Xing / Info / VBRI frame: read by the demuxer for its metadata, never decoded.  The tag sits where the main data
        // of an audio frame would start, behind an all-zero side information block (VBRI: at a fixed offset).
        const size_t off = 4 + h0.side_size;
        bool tag = false;
        if (pos + h0.frame_size <= n && off + 4 <= h0.frame_size &&
            (std::memcmp(p + pos + off, "Xing", 4) == 0 || std::memcmp(p + pos + off, "Info", 4) == 0)) {
            tag = true;
            for (size_t i = 4; i < off; ++i) tag = tag && p[pos + i] == 0;
        }
        if (pos + h0.frame_size <= n && 36 + 4 <= h0.frame_size && std::memcmp(p + pos + 36, "VBRI", 4) == 0) tag = true;
        if (tag) pos += h0.frame_size;
    }
    Decoder dec;
    std::vector<float> out[2];
    while (pos + 4 <= n) {
        Header h;
        if (p[pos] != 0xFF || !parse_header(p + pos, h) || !same_stream(h, h0) || h.bitrate == 0) { ++pos; continue; }   // resync byte by byte
        if (pos + h.frame_size > n) break;                          // truncated last frame: IoError -> break (main.rs:258-262)
        std::vector<float> got[2];
        dec.decode_frame(h, p + pos, got);
        for (int c = 0; c < channels; ++c) {
            const std::vector<float>& src = got[c < h.channels ? c : 0];
            out[c].insert(out[c].end(), src.begin(), src.end());
        }
        pos += h.frame_size;
    }
    mono.resize(out[0].size());
    const float fc = (float)channels;
    for (size_t i = 0; i < mono.size(); ++i) {
        float acc = 0.0f;
        for (int c = 0; c < channels; ++c) acc += out[c][i];
        mono[i] = acc / fc;
    }
}

}  // namespace wbmp3

// Test hook: an in-memory Layer III stream -> the mono mix load_audio_16k_mono builds before resampling (malloc'd, free with
// wb_host_free), the stream's channel count and sample rate.
extern "C" int wb_host_mp3_decode_mono(const uint8_t* bytes, int64_t n, float** out, int64_t* n_out, int* channels, uint32_t* sr) {
    try {
        WB_REQUIRE(bytes && out && n_out && channels && sr && n >= 4, WB_EINVAL, "null or empty argument");
        size_t first = 0;
        WB_REQUIRE(wbmp3::probe(bytes, (size_t)n, &first), WB_EINVAL, "mpa: no MPEG audio stream found");
        wbmp3::Header h0;
        wbmp3::parse_header(bytes + first, h0);
        std::vector<float> mono;
        uint32_t rate = 0;
        wbmp3::decode_to_mono(bytes, (size_t)n, first, mono, rate);
        float* buf = (float*)std::malloc(sizeof(float) * (mono.size() ? mono.size() : 1));
        WB_REQUIRE(buf, WB_EINVAL, "out of memory");
        std::memcpy(buf, mono.data(), sizeof(float) * mono.size());
        *out = buf; *n_out = (int64_t)mono.size(); *channels = h0.channels; *sr = rate;
        return WB_OK;
    } catch (const WbError& e) {
        wb_set_error(e.what());
        return e.code;
    } catch (const std::exception& e) {
        wb_set_error(e.what());
        return WB_EINVAL;
    }
}
