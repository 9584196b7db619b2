// utf8.h — the few Unicode operations the reference's Rust std calls imply (char iteration, White_Space,
// from_utf8_lossy) without ICU.  Lowercasing lives in text.cpp on top of the generated unicode_tables.h.
#pragma once
#include <cstdint>
#include <string>

namespace wbutf8 {

// Decode one code point at s[i], advance i. Invalid bytes decode as U+FFFD and advance by one.
inline uint32_t decode(const std::string& s, size_t& i) {
    const unsigned char c = (unsigned char)s[i];
    auto cont = [&](size_t k) { return i + k < s.size() && ((unsigned char)s[i + k] & 0xC0) == 0x80; };
    if (c < 0x80) { i += 1; return c; }
    if ((c & 0xE0) == 0xC0 && cont(1)) {
        uint32_t cp = ((c & 0x1Fu) << 6) | ((unsigned char)s[i + 1] & 0x3Fu);
        i += 2;
        return cp;
    }
    if ((c & 0xF0) == 0xE0 && cont(1) && cont(2)) {
        uint32_t cp = ((c & 0x0Fu) << 12) | (((unsigned char)s[i + 1] & 0x3Fu) << 6) | ((unsigned char)s[i + 2] & 0x3Fu);
        i += 3;
        return cp;
    }
    if ((c & 0xF8) == 0xF0 && cont(1) && cont(2) && cont(3)) {
        uint32_t cp = ((c & 0x07u) << 18) | (((unsigned char)s[i + 1] & 0x3Fu) << 12) |
                      (((unsigned char)s[i + 2] & 0x3Fu) << 6) | ((unsigned char)s[i + 3] & 0x3Fu);
        i += 4;
        return cp;
    }
    i += 1;
    return 0xFFFD;
}

// String::from_utf8_lossy: well-formed sequences (Unicode Table 3-7: no overlongs, no surrogates, nothing above
// U+10FFFF) are copied; every ill-formed run becomes ONE U+FFFD per maximal prefix of a well-formed sequence
// (at least one byte) -- the policy Rust's Utf8Chunks and CPython's errors="replace" share.
inline std::string from_utf8_lossy(const std::string& b) {
    std::string out;
    out.reserve(b.size());
    size_t i = 0;
    const size_t n = b.size();
    auto at = [&](size_t k) -> int { return k < n ? (unsigned char)b[k] : -1; };
    auto in = [](int v, int lo, int hi) { return v >= lo && v <= hi; };
    while (i < n) {
        const int c = at(i);
        if (c < 0x80) { out += (char)c; ++i; continue; }
        int need = 0, lo = 0x80, hi = 0xBF;                 // range of the SECOND byte depends on the lead byte
        if (in(c, 0xC2, 0xDF)) need = 1;
        else if (c == 0xE0) { need = 2; lo = 0xA0; }
        else if (in(c, 0xE1, 0xEC) || in(c, 0xEE, 0xEF)) need = 2;
        else if (c == 0xED) { need = 2; hi = 0x9F; }
        else if (c == 0xF0) { need = 3; lo = 0x90; }
        else if (in(c, 0xF1, 0xF3)) need = 3;
        else if (c == 0xF4) { need = 3; hi = 0x8F; }
        size_t len = 1;                                     // bytes of the maximal well-formed prefix
        bool ok = need > 0;
        if (ok) {
            if (in(at(i + 1), lo, hi)) {
                len = 2;
                for (int k = 2; k <= need; ++k) {
                    if (in(at(i + (size_t)k), 0x80, 0xBF)) len = (size_t)k + 1; else { ok = false; break; }
                }
            } else {
                ok = false;
            }
        }
        if (ok) out.append(b, i, (size_t)need + 1);
        else out += "\xEF\xBF\xBD";
        i += ok ? (size_t)need + 1 : len;
    }
    return out;
}

inline void encode(std::string& out, uint32_t cp) {
    if (cp < 0x80) out += (char)cp;
    else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
}

inline bool is_whitespace(uint32_t c) {      // Unicode White_Space (what Rust's char::is_whitespace tests)
    return (c >= 0x09 && c <= 0x0D) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 ||
           (c >= 0x2000 && c <= 0x200A) || c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}

}  // namespace wbutf8
