// text.cpp — host-side text + statistics helpers of the hot path's tail:
//   stitch_texts / word_overlap      /root/reference/src/main.rs:659-696
//   percentile / stat_block          :1021-1048
//   special_tokens                   :528-569
//   chunk list                       :875-882
//   decode_tokens (no-tokenizer arm) :644-647   (tokenizer arm lives in tokenizer.cpp)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../common.h"
#include "json.h"
#include "unicode_tables.h"
#include "utf8.h"

namespace wbtext {

// Rust str::split_whitespace: split on Unicode White_Space, drop empties.
std::vector<std::string> split_whitespace(const std::string& s) {
    std::vector<std::string> out;
    size_t i = 0, start = std::string::npos;
    while (i < s.size()) {
        size_t j = i;
        uint32_t cp = wbutf8::decode(s, j);
        if (wbutf8::is_whitespace(cp)) {
            if (start != std::string::npos) out.push_back(s.substr(start, i - start));
            start = std::string::npos;
        } else if (start == std::string::npos) {
            start = i;
        }
        i = j;
    }
    if (start != std::string::npos) out.push_back(s.substr(start));
    return out;
}

std::string trim(const std::string& s) {
    size_t b = 0, e = s.size();
    while (b < e) {
        size_t j = b;
        if (!wbutf8::is_whitespace(wbutf8::decode(s, j))) break;
        b = j;
    }
    while (e > b) {
        size_t k = e - 1;
        while (k > b && ((unsigned char)s[k] & 0xC0) == 0x80) --k;
        size_t j = k;
        if (!wbutf8::is_whitespace(wbutf8::decode(s, j))) break;
        e = k;
    }
    return s.substr(b, e - b);
}

namespace {
bool in_ranges(const uint32_t (*r)[2], int n, uint32_t c) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) / 2;
        if (c < r[mid][0]) hi = mid - 1;
        else if (c > r[mid][1]) lo = mid + 1;
        else return true;
    }
    return false;
}
bool is_cased(uint32_t c) { return in_ranges(wbuni::CASED, wbuni::N_CASED, c); }
bool is_case_ignorable(uint32_t c) { return in_ranges(wbuni::CASE_IGNORABLE, wbuni::N_CASE_IGNORABLE, c); }
const wbuni::LowerEntry* lower_entry(uint32_t c) {
    int lo = 0, hi = wbuni::N_LOWER - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) / 2;
        if (c < wbuni::LOWER[mid].cp) hi = mid - 1;
        else if (c > wbuni::LOWER[mid].cp) lo = mid + 1;
        else return &wbuni::LOWER[mid];
    }
    return nullptr;
}
}  // namespace

// Rust's str::to_lowercase (alloc::str): the full Unicode mapping per scalar value (one char may become up to three,
// U+0130 -> "i" + U+0307) and the one context rule, Final_Sigma: U+03A3 becomes U+03C2 when it is preceded by a cased
// letter (skipping case-ignorable characters) and not followed by one.  Tables: unicode_tables.h (generated).
std::string to_lowercase(const std::string& s) {
    std::vector<uint32_t> cps;
    for (size_t i = 0; i < s.size();) cps.push_back(wbutf8::decode(s, i));
    std::string out;
    for (size_t i = 0; i < cps.size(); ++i) {
        const uint32_t c = cps[i];
        if (c == 0x3A3) {
            bool before = false, after = false;
            for (size_t k = i; k-- > 0;) {
                if (is_case_ignorable(cps[k])) continue;
                before = is_cased(cps[k]);
                break;
            }
            for (size_t k = i + 1; k < cps.size(); ++k) {
                if (is_case_ignorable(cps[k])) continue;
                after = is_cased(cps[k]);
                break;
            }
            wbutf8::encode(out, (before && !after) ? 0x3C2u : 0x3C3u);
            continue;
        }
        if (const wbuni::LowerEntry* e = lower_entry(c)) {
            for (int k = 0; k < 3 && e->to[k]; ++k) wbutf8::encode(out, e->to[k]);
        } else {
            wbutf8::encode(out, c);
        }
    }
    return out;
}

// csv crate, QuoteStyle::Necessary (main.rs:1216-1229): a field is quoted only when it holds the delimiter, a quote,
// CR or LF; quotes inside are doubled.
std::string csv_field(const std::string& s) {
    bool q = false;
    for (char c : s) if (c == ',' || c == '"' || c == '\n' || c == '\r') { q = true; break; }
    if (!q) return s;
    std::string o = "\"";
    for (char c : s) { if (c == '"') o += '"'; o += c; }
    return o + "\"";
}

int word_overlap(const std::string& a, const std::string& b, int max_words) {      // main.rs:686-696
    std::vector<std::string> aw = split_whitespace(a), bw = split_whitespace(b);
    for (auto& w : aw) w = to_lowercase(w);
    for (auto& w : bw) w = to_lowercase(w);
    int mx = std::min<int>({max_words, (int)aw.size(), (int)bw.size()});
    for (int k = mx; k >= 1; --k) {
        bool eq = true;
        for (int i = 0; i < k && eq; ++i) eq = aw[aw.size() - k + i] == bw[i];
        if (eq) return k;
    }
    return 0;
}

std::string stitch_texts(const std::vector<std::string>& chunks) {                  // main.rs:659-684
    std::string out;
    for (const auto& chunk : chunks) {
        std::string t = trim(chunk);
        if (t.empty()) continue;
        if (out.empty()) { out = t; continue; }
        int ov = word_overlap(out, t, 16);
        if (ov > 0) {
            std::vector<std::string> words = split_whitespace(t);
            std::string rem;
            for (size_t i = (size_t)ov; i < words.size(); ++i) {
                if (!rem.empty()) rem += ' ';
                rem += words[i];
            }
            if (!rem.empty()) { out += ' '; out += rem; }
        } else {
            out += ' ';
            out += t;
        }
    }
    return out;
}

double percentile(std::vector<double> xs, double p) {                               // main.rs:1021-1031
    if (xs.empty()) return NAN;
    std::sort(xs.begin(), xs.end());
    double k = ((double)xs.size() - 1.0) * (p / 100.0);
    size_t f = (size_t)std::floor(k), c = (size_t)std::ceil(k);
    if (f == c) return xs[f];
    return xs[f] + (xs[c] - xs[f]) * (k - (double)f);
}

}  // namespace wbtext

extern "C" {

int wb_host_chunk_starts(int64_t n_samples, int64_t chunk_len, int64_t step, int64_t* out, int cap) {
    if (chunk_len <= 0) chunk_len = WB_CHUNK_SAMPLES;
    if (step <= 0) step = 400000;
    int n = 0;
    int64_t pos = 0;
    while (pos < n_samples) {
        int64_t end = std::min(pos + chunk_len, n_samples);
        if (out && n < cap) out[n] = pos;
        ++n;
        if (end == n_samples) break;
        pos += step;
    }
    return n;
}

int wb_host_word_overlap(const char* a, const char* b, int max_words) {
    return wbtext::word_overlap(a ? a : "", b ? b : "", max_words);
}

int64_t wb_host_stitch_texts(const char* const* chunks, int n, char* out, int64_t cap) {
    std::vector<std::string> v;
    for (int i = 0; i < n; ++i) v.emplace_back(chunks[i] ? chunks[i] : "");
    std::string s = wbtext::stitch_texts(v);
    if (out && cap > 0) {
        size_t m = std::min<size_t>(s.size(), (size_t)cap - 1);
        std::memcpy(out, s.data(), m);
        out[m] = '\0';
    }
    return (int64_t)s.size();
}

double wb_host_percentile(const double* xs, int n, double p) {
    return wbtext::percentile(std::vector<double>(xs, xs + (n > 0 ? n : 0)), p);
}

int wb_host_stat_block(const double* xs, int n, double* o) {                        // main.rs:1033-1048
    if (!o) return WB_EINVAL;
    std::vector<double> v(xs, xs + (n > 0 ? n : 0));
    std::sort(v.begin(), v.end());
    double sum = 0;
    for (double x : v) sum += x;
    o[0] = v.empty() ? NAN : v.front();
    o[1] = v.empty() ? NAN : v[v.size() / 2];          // upper median
    o[2] = wbtext::percentile(v, 90.0);
    o[3] = wbtext::percentile(v, 95.0);
    o[4] = v.empty() ? NAN : v.back();
    o[5] = v.empty() ? NAN : sum / (double)v.size();
    return WB_OK;
}

static int64_t copy_out(const std::string& r, char* out, int64_t cap) {
    if (out && cap > 0) {
        const size_t n = std::min<size_t>(r.size(), (size_t)cap - 1);
        std::memcpy(out, r.data(), n);
        out[n] = '\0';
    }
    return (int64_t)r.size();
}
int64_t wb_host_json_string(const char* s, char* out, int64_t cap) { return copy_out(wbjson::escape(s ? s : ""), out, cap); }
int64_t wb_host_csv_field(const char* s, char* out, int64_t cap) { return copy_out(wbtext::csv_field(s ? s : ""), out, cap); }

int64_t wb_host_to_lowercase(const char* s, char* out, int64_t cap) {
    const std::string r = wbtext::to_lowercase(s ? s : "");
    if (out && cap > 0) {
        const size_t n = std::min<size_t>(r.size(), (size_t)cap - 1);
        std::memcpy(out, r.data(), n);
        out[n] = '\0';
    }
    return (int64_t)r.size();
}

int64_t wb_host_format_f64(double v, char* out, int64_t cap) {
    const std::string s = wbjson::fmt_f64(v);
    if (out && cap > 0) {
        const size_t n = std::min<size_t>(s.size(), (size_t)cap - 1);
        std::memcpy(out, s.data(), n);
        out[n] = '\0';
    }
    return (int64_t)s.size();
}

}  // extern "C"
