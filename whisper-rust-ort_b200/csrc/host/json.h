// json.h — minimal JSON reader/writer (replaces serde_json uses in /root/reference/src/main.rs:
// discovery json :124-167, generation_config.json :650-657, tokenizer.json, output files :1232-1259).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace wbjson {

struct Value {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    bool b = false;
    double n = 0.0;
    bool is_int = false;
    std::string s;
    std::vector<Value> a;
    std::vector<std::pair<std::string, Value>> o;   // insertion order

    bool is_null() const { return type == Null; }
    double num() const { return n; }
    const std::string& str() const { return s; }
    const std::vector<Value>& arr() const { return a; }
    const Value* find(const std::string& k) const {
        for (const auto& kv : o) if (kv.first == k) return &kv.second;
        return nullptr;
    }
    const Value& operator[](const std::string& k) const {
        static const Value null_v;
        const Value* v = find(k);
        return v ? *v : null_v;
    }
};

struct Parser {
    const std::string& t;
    size_t i = 0;
    explicit Parser(const std::string& text) : t(text) {}
    [[noreturn]] void fail(const char* m) const {
        throw std::runtime_error(std::string("JSON parse error at byte ") + std::to_string(i) + ": " + m);
    }
    void ws() { while (i < t.size() && (t[i] == ' ' || t[i] == '\n' || t[i] == '\t' || t[i] == '\r')) ++i; }
    static void utf8(std::string& out, unsigned cp) {
        if (cp < 0x80) out += (char)cp;
        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
        else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    }
    unsigned hex4() {
        if (i + 4 > t.size()) fail("bad \\u escape");
        unsigned v = 0;
        for (int k = 0; k < 4; ++k) {
            char c = t[i++];
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
            else fail("bad hex digit");
        }
        return v;
    }
    std::string string() {
        if (t[i] != '"') fail("expected string");
        ++i;
        std::string out;
        while (true) {
            if (i >= t.size()) fail("unterminated string");
            char c = t[i++];
            if (c == '"') break;
            if (c != '\\') { out += c; continue; }
            if (i >= t.size()) fail("bad escape");
            char e = t[i++];
            switch (e) {
                case '"': out += '"'; break;
                case '\\': out += '\\'; break;
                case '/': out += '/'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'n': out += '\n'; break;
                case 'r': out += '\r'; break;
                case 't': out += '\t'; break;
                case 'u': {
                    unsigned cp = hex4();
                    if (cp >= 0xD800 && cp < 0xDC00 && i + 1 < t.size() && t[i] == '\\' && t[i + 1] == 'u') {
                        i += 2;
                        unsigned lo = hex4();
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    utf8(out, cp);
                    break;
                }
                default: fail("unknown escape");
            }
        }
        return out;
    }
    Value value() {
        ws();
        if (i >= t.size()) fail("unexpected end");
        Value v;
        char c = t[i];
        if (c == '{') {
            v.type = Value::Obj;
            ++i; ws();
            if (i < t.size() && t[i] == '}') { ++i; return v; }
            while (true) {
                ws();
                std::string k = string();
                ws();
                if (i >= t.size() || t[i] != ':') fail("expected ':'");
                ++i;
                v.o.emplace_back(std::move(k), value());
                ws();
                if (i < t.size() && t[i] == ',') { ++i; continue; }
                if (i < t.size() && t[i] == '}') { ++i; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.type = Value::Arr;
            ++i; ws();
            if (i < t.size() && t[i] == ']') { ++i; return v; }
            while (true) {
                v.a.push_back(value());
                ws();
                if (i < t.size() && t[i] == ',') { ++i; continue; }
                if (i < t.size() && t[i] == ']') { ++i; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v.type = Value::Str;
            v.s = string();
        } else if (t.compare(i, 4, "true") == 0) { v.type = Value::Bool; v.b = true; i += 4; }
        else if (t.compare(i, 5, "false") == 0) { v.type = Value::Bool; v.b = false; i += 5; }
        else if (t.compare(i, 4, "null") == 0) { i += 4; }
        else {
            size_t s0 = i;
            bool isint = true;
            if (t[i] == '-') ++i;
            while (i < t.size() && ((t[i] >= '0' && t[i] <= '9') || t[i] == '.' || t[i] == 'e' || t[i] == 'E' || t[i] == '+' || t[i] == '-')) {
                if (t[i] == '.' || t[i] == 'e' || t[i] == 'E') isint = false;
                ++i;
            }
            if (s0 == i) fail("unexpected character");
            v.type = Value::Num;
            v.n = std::strtod(t.substr(s0, i - s0).c_str(), nullptr);
            v.is_int = isint;
        }
        return v;
    }
};

inline Value parse(const std::string& text) {
    Parser p(text);
    Value v = p.value();
    p.ws();
    if (p.i != text.size()) p.fail("trailing characters");
    return v;
}

// ---- writer: serde_json-compatible escaping and pretty printing (2-space indent) ----
inline std::string escape(const std::string& s) {
    std::string o = "\"";
    for (unsigned char c : s) {
        switch (c) {
            case '"': o += "\\\""; break;
            case '\\': o += "\\\\"; break;
            case '\n': o += "\\n"; break;
            case '\r': o += "\\r"; break;
            case '\t': o += "\\t"; break;
            case '\b': o += "\\b"; break;
            case '\f': o += "\\f"; break;
            default:
                if (c < 0x20) { char b[8]; snprintf(b, sizeof(b), "\\u%04x", c); o += b; }
                else o += (char)c;
        }
    }
    return o + "\"";
}

// Shortest round-trip f64 formatting with ryu's "pretty" layout, which is what serde_json emits
// ("14.884440201999999", "1.0", "0.0000482", "4.82e-6", "1e16").
inline std::string fmt_f64(double v) {
    if (std::isnan(v) || std::isinf(v)) return "null";       // serde_json writes null
    if (v == 0.0) return std::signbit(v) ? "-0.0" : "0.0";
    // Shortest digit string that reads back as v, and among those the one closest to v (ties -> even): what ryu /
    // Grisu / David Gay's algorithm all define.  The correctly rounded n-digit decimal is NOT always it: at an exact
    // decimal tie (v = 2^-24 = 5.9604644775390625e-8) round-half-even gives ...062, which reads back as a different
    // double, while ...063 round-trips.  So both neighbours of the exact expansion are tried at every length.
    const bool neg = std::signbit(v);
    const double a = std::fabs(v);
    char big[832];
    snprintf(big, sizeof(big), "%.780e", a);                 // glibc prints the exact binary value (<= 767 significant digits)
    std::string exact;
    const char* ep = std::strchr(big, 'e');
    for (const char* q = big; q < ep; ++q) if (*q != '.') exact += *q;
    int e10 = std::atoi(ep + 1);
    std::string digits;
    auto reads_back = [&](const std::string& d, int e) {
        const std::string t = d.substr(0, 1) + "." + d.substr(1) + "e" + std::to_string(e);
        return std::strtod(t.c_str(), nullptr) == a;
    };
    for (size_t n = 1; n <= 17 && digits.empty(); ++n) {
        std::string down = exact.substr(0, n), up = down;
        int e_up = e10;
        int i = (int)n - 1;
        while (i >= 0 && up[(size_t)i] == '9') up[(size_t)i--] = '0';
        if (i >= 0) ++up[(size_t)i]; else { up = "1" + up.substr(0, n - 1); ++e_up; }
        const bool ok_d = reads_back(down, e10), ok_u = reads_back(up, e_up);
        if (!ok_d && !ok_u) continue;
        bool take_up = ok_u && !ok_d;
        if (ok_d && ok_u) {                                  // closer to the exact expansion; tie -> even last digit
            const std::string rest = exact.substr(n);
            const std::string half = "5" + std::string(rest.size() - 1, '0');
            take_up = rest > half || (rest == half && ((down[n - 1] - '0') & 1));
        }
        digits = take_up ? up : down;
        if (take_up) e10 = e_up;
    }
    if (digits.empty()) { digits = exact.substr(0, 17); }
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int len = (int)digits.size();
    const int k = e10 - (len - 1);          // value = digits * 10^k
    const int kk = len + k;                 // position of the decimal point
    std::string out;
    if (k >= 0 && kk <= 16) {
        out = digits + std::string((size_t)k, '0') + ".0";
    } else if (kk > 0 && kk <= 16) {
        out = digits.substr(0, (size_t)kk) + "." + digits.substr((size_t)kk);
    } else if (kk > -5 && kk <= 0) {
        out = "0." + std::string((size_t)(-kk), '0') + digits;
    } else if (len == 1) {
        out = digits + "e" + std::to_string(kk - 1);
    } else {
        out = digits.substr(0, 1) + "." + digits.substr(1) + "e" + std::to_string(kk - 1);
    }
    return neg ? "-" + out : out;
}

}  // namespace wbjson
